#!/bin/bash
# 2-GPU development pass: NCCL tests + pipelined halo exchange with 1 / 2 / 4 / 8 chunks
TAG=${1:-multi2}; N=2; WL=${2:-rmat22}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests/test_dist_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "nccl or tile_ranges" > $OUT/pytest_dist.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_dist.log; tail -n 4 $OUT/pytest_dist.log
for ch in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $WL --steps 5 --warmup 3 --exchange halo --chunks $ch > $OUT/bench_${WL}_n${N}_c${ch}.json 2> $OUT/bench_${WL}_n${N}_c${ch}.err
  echo "chunks $ch exit $?"
done
for f in $OUT/bench_*.json; do echo $f; python -c "
import json
try:
    l=json.loads(open('$f').read().strip().splitlines()[-1]); print('  %.2f Gedges/s  %.3f ms/step' % (l['value']/1e9, l['ms_per_step']))
except Exception as e: print('  failed', e)
"; done
grep -h -i "error\|Traceback" -A5 $OUT/*.err | tail -n 30

#!/bin/bash
# 2-GPU development pass: row-partition tests + peer/NCCL transports with 1 / 4 chunks
TAG=${1:-multi2}; N=2; WL=${2:-rmat22}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > $OUT/pytest_dist.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_dist.log; tail -n 6 $OUT/pytest_dist.log
for tr in peer nccl; do for ch in 1 4; do
  SGLB200_DIST_TRACE=${TRACE:-0} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $WL --steps 5 --warmup 3 --exchange halo --chunks $ch --transport $tr > $OUT/bench_${WL}_n${N}_${tr}_c${ch}.json 2> $OUT/bench_${WL}_n${N}_${tr}_c${ch}.err
  echo "$tr chunks $ch exit $?"
done; done
for f in $OUT/bench_*.json; do echo $f; python -c "
import json
try:
    l=json.loads(open('$f').read().strip().splitlines()[-1]); print('  %.2f Gedges/s  %.3f ms/step' % (l['value']/1e9, l['ms_per_step']), l['config'].get('transport'))
except Exception as e: print('  failed', e)
"; done
grep -h -i "error\|Traceback" -A8 $OUT/*.err | tail -n 40

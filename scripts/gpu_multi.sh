#!/bin/bash
# multi-GPU pass (run under gpurun --gpus N): NCCL tests + strong-scaling bench lines
TAG=${1:-multi}; N=${2:-2}; WL=${3:-rmat22}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
python -m pytest tests/test_dist_gpu.py -m gpu -x -q > $OUT/pytest_dist.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_dist.log; tail -n 3 $OUT/pytest_dist.log
for wl in $WL; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_${wl}_n1.json 2> $OUT/bench_${wl}_n1.err; echo "n1 $wl exit $?"
  for ex in ${EXCH:-halo allgather}; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $wl --steps 5 --warmup 3 --exchange $ex ${EXTRA} > $OUT/bench_${wl}_n${N}_${ex}.json 2> $OUT/bench_${wl}_n${N}_${ex}.err
    echo "n$N $wl $ex exit $?"
  done
done
for f in $OUT/bench_*.json; do echo $f; python -c "
import sys,json
try:
    l=json.loads(open('$f').read().strip().splitlines()[-1]); print('  %.2f Gedges/s  %.3f ms/step  frac %.3f' % (l['value']/1e9, l['ms_per_step'], l['roofline']['frac']), l['config'].get('halo_recv_bytes_per_hop_max_rank'))
except Exception as e: print('  failed', e)
"; done
tail -n 5 $OUT/*.err | tail -n 40

#!/bin/bash
# 8-GPU strong-scaling lines (run under gpurun --gpus 8)
TAG=${1:-multi8}; N=${2:-8}; WL=${3:-rmat24}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
for wl in $WL; do
  for ex in halo allgather; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $wl --steps 5 --warmup 3 --exchange $ex > $OUT/bench_${wl}_n${N}_${ex}.json 2> $OUT/bench_${wl}_n${N}_${ex}.err
    echo "n$N $wl $ex exit $?"
  done
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_arxiv_n${N}.json 2> $OUT/bench_arxiv_n${N}.err; echo "arxiv n$N exit $?"
for f in $OUT/bench_*.json; do echo $f; python -c "
import json
try:
    l=json.loads(open('$f').read().strip().splitlines()[-1]); print('  %.2f Gedges/s  %.3f ms/step  frac %.3f' % (l['value']/1e9, l['ms_per_step'], l['roofline']['frac']), l['config'].get('halo_recv_bytes_per_hop_max_rank'), l['setup'])
except Exception as e: print('  failed', e)
"; done
grep -h -i "error\|Traceback" -A3 $OUT/*.err | tail -n 20

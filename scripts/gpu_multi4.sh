#!/bin/bash
TAG=${1:-multi4}; N=4; WL=${2:-rmat22}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for cfg in "peer 4" "peer 1" "nccl 1"; do
  set -- $cfg
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $WL --steps 5 --warmup 3 --exchange halo --transport $1 --chunks $2 > $OUT/bench_${WL}_n${N}_$1_c$2.json 2> $OUT/bench_${WL}_n${N}_$1_c$2.err
  echo "n$N $WL $1 c$2 exit $?"
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_arxiv_n${N}.json 2> $OUT/bench_arxiv_n${N}.err; echo "arxiv n$N exit $?"
for f in $OUT/bench_*.json; do echo $f; python -c "
import json
try:
    l=json.loads(open('$f').read().strip().splitlines()[-1]); print('  %.2f Gedges/s  %.3f ms/step' % (l['value']/1e9, l['ms_per_step']), l['config'].get('transport'), l['config'].get('chunks'), l['config'].get('halo_recv_bytes_per_hop_max_rank'))
except Exception as e: print('  failed', e)
"; done
grep -h -i "error\|Traceback" -A6 $OUT/*.err | tail -n 20

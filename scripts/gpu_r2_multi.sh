#!/bin/bash
# N-GPU bench exactly as the driver launches it.  usage: gpu_r2_multi.sh N [extra bench args]
N=${1:-2}; shift
OUT=gpurun_out/r2_multi
mkdir -p $OUT
run() {
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 "$@" > $OUT/${tag}_n$N.json 2> $OUT/${tag}_n$N.err
  echo "$tag exit $?"
  python - <<PY
import json
try:
    l=json.loads(open('$OUT/${tag}_n$N.json').read().strip().splitlines()[-1])
    print('%-28s n=%d %8.2f ms/step %7.2f Gedges/s  e2e %s  %s' % ('$tag', l['n_gpus'], l['ms_per_step'], l['value']/1e9,
          ('%.2f' % (l['e2e']['value']/1e9)) if l.get('e2e') else None, json.dumps(l.get('timing') or l.get('parity') or '')[:200]))
    print('   partition:', l['partition'][:200]); print('   parity:', json.dumps(l.get('parity'))[:300])
except Exception as e:
    print('$tag FAILED', e, open('$OUT/${tag}_n$N.err').read()[-600:].replace(chr(10),' | '))
PY
}
run products_feature "$@"
run products_row --partition row --transport nccl "$@"
if [ "$N" -le 4 ]; then run rmat23_feature --workload rmat23 "$@"; fi

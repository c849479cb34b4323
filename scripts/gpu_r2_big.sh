#!/bin/bash
# large R-MAT instances on N GPUs, feature split (per-rank on-device generation, device-side row sampling for parity)
N=${1:-8}; shift
OUT=gpurun_out/r2_big
mkdir -p $OUT
for W in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 5 --warmup 3 --workload $W > $OUT/${W}_feature_n$N.json 2> $OUT/${W}_feature_n$N.err
  echo "$W exit $?"
  python - <<PY
import json
try:
    l=json.loads(open('$OUT/${W}_feature_n$N.json').read().strip().splitlines()[-1])
    print('%-10s n=%d N=%d nnz=%d %8.2f ms/step %7.2f Gedges/s  e2e %s  %s  setup %s' % ('$W', l['n_gpus'], l['config']['N'], l['config']['nnz'], l['ms_per_step'], l['value']/1e9,
          ('%.2f' % (l['e2e']['value']/1e9)) if l.get('e2e') else None, json.dumps(l.get('timing')), json.dumps(l.get('setup'))))
    print('   parity:', json.dumps(l.get('parity'))[:400])
except Exception as e:
    print('$W FAILED', e, open('$OUT/${W}_feature_n$N.err').read()[-900:].replace(chr(10),' | '))
PY
done
nvidia-smi --query-gpu=memory.used --format=csv | head -3

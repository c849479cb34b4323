#!/bin/bash
# 2 GPUs: the two-rank parity tests (row partition, feature split) and the products feature split with / without padded block slabs
OUT=gpurun_out/r2_pad2
mkdir -p $OUT
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
run() {
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 "$@" > $OUT/${tag}_n2.json 2> $OUT/${tag}_n2.err
  echo "$tag exit $?"
  python - <<PY
import json
try:
    l=json.loads(open('$OUT/${tag}_n2.json').read().strip().splitlines()[-1])
    print('%-28s n=%d %8.2f ms/step %7.2f Gedges/s  e2e %s  %s' % ('$tag', l['n_gpus'], l['ms_per_step'], l['value']/1e9,
          ('%.2f' % (l['e2e']['value']/1e9)) if l.get('e2e') else None, json.dumps(l.get('timing') or '')[:200]))
    print('   parity:', json.dumps(l.get('parity'))[:300])
except Exception as e:
    print('$tag FAILED', e, open('$OUT/${tag}_n2.err').read()[-600:].replace(chr(10),' | '))
PY
}
run products_feature_pad
run products_feature_nopad --no-pad --no-e2e

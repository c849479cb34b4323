#!/bin/bash
OUT=gpurun_out/r2_recheck
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300
run() {
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-comparators --traffic none "$@" > $OUT/$tag.json 2> $OUT/$tag.err
  python - <<PY
import json
try:
    l=json.loads(open('$OUT/$tag.json').read().strip().splitlines()[-1])
    print('%-40s %8.1f us/hop  frac %.3f  %.2f Gedges/s  fused %s' % ('$tag', l['roofline']['us_per_launch'], l['roofline']['frac'], l['value']/1e9, {k: round(v,2) for k,v in (l.get('fused') or {}).items() if k.endswith('_ms')}))
except Exception as e:
    print('$tag', 'FAILED', e, open('$OUT/$tag.err').read()[-300:].replace(chr(10),' '))
PY
}
run products_t256 -- --workload products
run products_t512 -- --workload products --tile-items 512
run products_t1024 -- --workload products --tile-items 1024
run rmat22_t256 -- --workload rmat22
run rmat22_t512 -- --workload rmat22 --tile-items 512
run products_d16_t512 -- --workload products --feat-dim 16 --tile-items 512
run products_d16_t128 -- --workload products --feat-dim 16 --tile-items 128

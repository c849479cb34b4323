#!/bin/bash
OUT=gpurun_out/r2_group2
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "narrow_rows or feature_widths or fused" 2>&1 | tail -4 | cut -c1-250
run() {
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/$tag.json 2> $OUT/$tag.err
  python - <<PY
import json
try:
    l=json.loads(open('$OUT/$tag.json').read().strip().splitlines()[-1])
    print('%-40s %8.1f us/hop  frac %.3f  %.2f Gedges/s' % ('$tag', l['roofline']['us_per_launch'], l['roofline']['frac'], l['value']/1e9))
except Exception as e:
    print('$tag', 'FAILED', e, open('$OUT/$tag.err').read()[-200:].replace(chr(10),' '))
PY
}
for D in 16 32 64; do
  run rmat22_d${D}_u8 -- --workload rmat22 --feat-dim $D
  run rmat22_d${D}_u4 SGLB200_GROUP_U=4 -- --workload rmat22 --feat-dim $D
done
run products_d12_u8 -- --workload products --feat-dim 12
run products_d12_u4 SGLB200_GROUP_U=4 -- --workload products --feat-dim 12
run products_d52_u8 -- --workload products --feat-dim 52
run products_d52_u4 SGLB200_GROUP_U=4 -- --workload products --feat-dim 52

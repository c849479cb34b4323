#!/bin/bash
OUT=gpurun_out/r2_group3
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "narrow_rows or feature_widths or fused or cut_rows or osd or nafs or golden_combiners" 2>&1 | tail -4 | cut -c1-250
run() {
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-comparators --traffic none "$@" > $OUT/$tag.json 2> $OUT/$tag.err
  python - <<PY
import json
try:
    l=json.loads(open('$OUT/$tag.json').read().strip().splitlines()[-1])
    print('%-40s %8.1f us/hop  frac %.3f  %.2f Gedges/s' % ('$tag', l['roofline']['us_per_launch'], l['roofline']['frac'], l['value']/1e9))
except Exception as e:
    print('$tag', 'FAILED', e, open('$OUT/$tag.err').read()[-300:].replace(chr(10),' '))
PY
}
for D in 12 16 24 28 52; do
  run products_d${D}_coop -- --workload products --feat-dim $D
  run products_d${D}_nocoop SGLB200_GROUP_COOP=0 -- --workload products --feat-dim $D
done
run rmat22_d16_coop -- --workload rmat22 --feat-dim 16
run rmat22_d16_nocoop SGLB200_GROUP_COOP=0 -- --workload rmat22 --feat-dim 16
run rmat22_d64_coop -- --workload rmat22 --feat-dim 64
run rmat22_d64_nocoop SGLB200_GROUP_COOP=0 -- --workload rmat22 --feat-dim 64
python scripts/fused_probe.py products 2>&1 | grep -E "separate|plain" | cut -c1-120

#!/bin/bash
OUT=gpurun_out/r2_tiles
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
run() {
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-comparators --traffic none "$@" > $OUT/$tag.json 2> $OUT/$tag.err
  python - <<PY
import json
try:
    l=json.loads(open('$OUT/$tag.json').read().strip().splitlines()[-1])
    print('%-40s %8.1f us/hop  frac %.3f  %.2f Gedges/s  tile_items %s' % ('$tag', l['roofline']['us_per_launch'], l['roofline']['frac'], l['value']/1e9, l['setup']['tile_items']))
except Exception as e:
    print('$tag', 'FAILED', e, open('$OUT/$tag.err').read()[-300:].replace(chr(10),' '))
PY
}
run products_default -- --workload products
run products_t2048 -- --workload products --tile-items 2048
run products_t4096 -- --workload products --tile-items 4096
run rmat22_default -- --workload rmat22
run rmat22_t2048 -- --workload rmat22 --tile-items 2048
run arxiv_default -- --workload arxiv
run products_d16_default -- --workload products --feat-dim 16
run products_d16_t2048 -- --workload products --feat-dim 16 --tile-items 2048
run products_d52_default -- --workload products --feat-dim 52

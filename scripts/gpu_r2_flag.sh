#!/bin/bash
# round 2: flagged-stream walk in the warp kernel + the fused driver tests
OUT=gpurun_out/r2_flag
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12 | cut -c1-250
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
for W in arxiv products rmat22; do
  run ${W}_noflag SGLB200_FLAGS=0 -- --workload $W
  run ${W}_flag -- --workload $W
  run ${W}_flag_minb4 SGLB200_FLAG_MINB4=1 -- --workload $W
done
run products_d12 -- --workload products --feat-dim 12
run products_d16 -- --workload products --feat-dim 16
run products_d24 -- --workload products --feat-dim 24
run products_d52 -- --workload products --feat-dim 52
run rmat24_d16 -- --workload rmat24 --feat-dim 16
run rmat24_d32 -- --workload rmat24 --feat-dim 32
run rmat24_d64 -- --workload rmat24 --feat-dim 64

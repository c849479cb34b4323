#!/bin/bash
# round 2: TMA-staged hop kernel -- parity + throughput against the register-staged kernel
OUT=gpurun_out/r2_tma
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma or narrow_rows or feature_widths or cut_rows or tile_ranges" 2>&1 | tail -15
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
  run ${W}_ldg -- --workload $W
  run ${W}_tma SGLB200_TMA=1 -- --workload $W
  run ${W}_tma_t512 SGLB200_TMA=1 -- --workload $W --tile-items 512
  run ${W}_tma_t1024 SGLB200_TMA=1 -- --workload $W --tile-items 1024
done
run rmat22_d16_group -- --workload rmat22 --feat-dim 16

#!/bin/bash
# round 2: lane-group kernel for narrow rows -- parity tests + per-width throughput on rmat22 (feature-split shapes)
OUT=gpurun_out/r2_group
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "feature_widths or narrow_rows or cut_rows or tile_ranges or empty_rows" 2>&1 | tail -5
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
    print('$tag', 'FAILED', e); print(open('$OUT/$tag.err').read()[-800:])
PY
}
for D in 16 32 64; do
  true
  run rmat22_d${D}_group -- --workload rmat22 --feat-dim $D
  run rmat22_d${D}_warp SGLB200_GROUP=0 -- --workload rmat22 --feat-dim $D
done
run rmat24_d16_group -- --workload rmat24 --feat-dim 16
run rmat24_d128 -- --workload rmat24
run products_d12_group -- --workload products --feat-dim 12
run products_d52_group -- --workload products --feat-dim 52
for RF in 100 128; do ./scripts/bin/l2_policy_bench $RF 2449029 131072 0.74; done
./scripts/bin/l2_policy_bench 100 2449029 65536 0.62
./scripts/bin/l2_policy_bench 100 2449029 262144 0.85
python scripts/host_path_probe.py arxiv
python scripts/host_path_probe.py products

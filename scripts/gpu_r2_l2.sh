#!/bin/bash
# round 2 experiment: L2 residency levers on the HBM-resident workloads (relabel, cache hints, policy window, column tiles)
OUT=gpurun_out/r2_l2
mkdir -p $OUT
run() { # tag, env..., -- bench args
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/$tag.json 2> $OUT/$tag.err
  python - <<PY
import json
try:
    l=json.loads(open('$OUT/$tag.json').read().strip().splitlines()[-1])
    print('%-40s %8.1f us/hop  frac %.3f  %.2f Gedges/s' % ('$tag', l['roofline']['us_per_launch'], l['roofline']['frac'], l['value']/1e9))
except Exception as e:
    print('$tag', 'FAILED', e)
PY
}
python - <<'PY'
import ctypes, torch
torch.cuda.init()
rt = ctypes.CDLL('libcudart.so.12') if False else None
print('L2', torch.cuda.get_device_properties(0).L2_cache_size)
try:
    from cuda.bindings import runtime as cr
    err, v = cr.cudaDeviceGetAttribute(cr.cudaDeviceAttr.cudaDevAttrMaxPersistingL2CacheSize, 0); print('maxPersistingL2', v)
    err, v = cr.cudaDeviceGetAttribute(cr.cudaDeviceAttr.cudaDevAttrMaxAccessPolicyWindowSize, 0); print('maxWindow', v)
except Exception as e:
    print('cuda-python query failed', e)
PY
for W in products rmat22; do
  run ${W}_base -- --workload $W
  run ${W}_relabel -- --workload $W --relabel degree
  for H in 32768 65536 131072 262144; do
    run ${W}_relabel_hint${H}_cold1 SGLB200_HUB_COLS=$H SGLB200_COLD_POLICY=1 -- --workload $W --relabel degree
  done
  run ${W}_relabel_hint131072_cold0 SGLB200_HUB_COLS=131072 SGLB200_COLD_POLICY=0 -- --workload $W --relabel degree
  for MB in 32 64 96; do
    run ${W}_relabel_window${MB} SGLB200_L2_WINDOW_MB=$MB -- --workload $W --relabel degree
  done
done
run rmat22_relabel_tile64 SGLB200_COL_TILE=64 -- --workload rmat22 --relabel degree
run rmat22_relabel_tile64_hint262144 SGLB200_COL_TILE=64 SGLB200_HUB_COLS=262144 -- --workload rmat22 --relabel degree
run rmat22_relabel_tile32_hint524288 SGLB200_COL_TILE=32 SGLB200_HUB_COLS=524288 -- --workload rmat22 --relabel degree
run products_relabel_tile50_hint262144 SGLB200_COL_TILE=50 SGLB200_HUB_COLS=262144 -- --workload products --relabel degree

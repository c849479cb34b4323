#!/bin/bash
OUT=gpurun_out/r2_pairs
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "narrow_rows or feature_widths or learnable or fused or sgap or row_partition or device_built" 2>&1 | tail -3 | cut -c1-300
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
for D in 12 16 24 52; do
  run products_d${D}_pairs -- --workload products --feat-dim $D
  run products_d${D}_nopairs SGLB200_GROUP_PAIRS=0 -- --workload products --feat-dim $D
done
run rmat22_d16_pairs -- --workload rmat22 --feat-dim 16
run rmat22_d16_nopairs SGLB200_GROUP_PAIRS=0 -- --workload rmat22 --feat-dim 16
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"lw_back_score" -c 2 --csv python scripts/aux_kernels_probe.py 2>&1 | grep lw_back | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | cut -c1-200

#!/bin/bash
# round 2 experiment: TMA gather4 microbenchmark + rows-in-flight variants of the hop kernel on HBM-resident workloads
OUT=gpurun_out/r2_g4
mkdir -p $OUT
./scripts/bin/gather_bench 169343 16777216 0 > $OUT/gather_l2.txt 2>&1
./scripts/bin/gather_bench 2449029 16777216 0 > $OUT/gather_hbm.txt 2>&1
./scripts/bin/gather_bench 2449029 16777216 1 > $OUT/gather_hbm_skew.txt 2>&1
grep -E "^ldg    U=8  warps/cta=8 per_warp=512|^gather4|^bulk   stages=2 warps/cta=1" $OUT/gather_l2.txt $OUT/gather_hbm.txt $OUT/gather_hbm_skew.txt
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
    print('$tag', 'FAILED', e)
PY
}
for V in 2 3 4 5 7 8 10 11 12; do
  run products_v$V SGLB200_SPMM_VARIANT=$V SGLB200_FOLD=fixup -- --workload products
done
run products_default_fixup SGLB200_FOLD=fixup -- --workload products
for V in 3 7 10 11; do
  run rmat22_v$V SGLB200_SPMM_VARIANT=$V SGLB200_FOLD=fixup -- --workload rmat22
done
# hit rates with hints: ncu metrics only
for H in 0 131072; do
SGLB200_HUB_COLS=$H timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:spmm_flat_kernel -s 8 -c 1 --csv --log-file $OUT/ncu_hint$H.csv python bench.py --workload products --relabel degree --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -E "spmm_flat" $OUT/ncu_hint$H.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done

#!/bin/bash
OUT=gpurun_out/r2_ncu_group2
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_group_kernel -s 8 -c 1 -o $OUT/group_products_d16 \
    python bench.py --workload products --feat-dim 16 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-comparators --traffic none > $OUT/ncu.log 2>&1
echo "ncu exit $?"

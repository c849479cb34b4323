#!/bin/bash
# round 2, first call: where does the round-1 hop kernel stand on the HBM-resident workloads?
OUT=gpurun_out/r2_base
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.txt
for W in products rmat22; do
  timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_$W.json 2> $OUT/bench_$W.err
  echo "bench $W exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_flat_kernel -s 8 -c 1 -o $OUT/spmm_full_$W \
      python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_$W.log 2>&1
  echo "ncu $W exit $?"
done
tail -c 600 $OUT/bench_products.json; tail -c 600 $OUT/bench_rmat22.json

#!/bin/bash
# final pass of the round: sanitizer smoke, all GPU tests, smoke(), default bench + launch list
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_smoke.py > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $OUT/sanitizer_memcheck.log; tail -n 4 $OUT/sanitizer_memcheck.log
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -n 3 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -n 2 $OUT/smoke.log
python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 600 $OUT/bench.json; tail -n 3 $OUT/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sglb200 -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_bench.log 2>&1; echo "ncu exit $?"

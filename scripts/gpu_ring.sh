#!/bin/bash
TAG=${1:-ring}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SGLB200_SPMM_VARIANT=10 python -m pytest tests -m gpu -x -q > $OUT/pytest_ring.log 2>&1; echo "pytest(ring) exit $?" >> $OUT/pytest_ring.log; tail -n 4 $OUT/pytest_ring.log
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -n 4 $OUT/pytest_gpu.log
for v in -1 10 11 12; do
 for cfg in "256 64" "128 64" "512 64" "256 256"; do
  set -- $cfg
  echo -n "variant=$v tile_items=$1 split=$2 : " >> $OUT/sweep.log
  SGLB200_SPMM_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --tile-items $1 --split-threshold $2 2>&1 | \
    python -c "import sys,json; l=json.loads(sys.stdin.readlines()[-1]); print('%.1f us/hop  %.2f Gedges/s  frac %.3f cut_rows %d' % (l['roofline']['us_per_launch'], l['value']/1e9, l['roofline']['frac'], l['setup']['cut_rows']))" >> $OUT/sweep.log 2>&1
 done
done
cat $OUT/sweep.log
python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 1500 $OUT/bench.json; tail -n 5 $OUT/bench.err

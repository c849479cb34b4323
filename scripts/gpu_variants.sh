#!/bin/bash
TAG=${1:-var}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -n 3 $OUT/pytest_gpu.log
for v in ${VARIANTS:--1 1 2 3 4 6 7}; do
 for cfg in "256 64" "512 64" "128 64"; do
  set -- $cfg
  echo -n "variant=$v tile_items=$1 split=$2 : " >> $OUT/sweep.log
  SGLB200_SPMM_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --tile-items $1 --split-threshold $2 ${EXTRA} 2>&1 | \
    python -c "import sys,json; l=json.loads(sys.stdin.readlines()[-1]); print('%.1f us/hop  %.2f Gedges/s  frac %.3f cut_rows %d' % (l['roofline']['us_per_launch'], l['value']/1e9, l['roofline']['frac'], l['setup']['cut_rows']))" >> $OUT/sweep.log 2>&1
 done
done
cat $OUT/sweep.log

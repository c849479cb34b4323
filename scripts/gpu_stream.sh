#!/bin/bash
TAG=${1:-stream}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -n 3 $OUT/pytest_gpu.log
for sy in 0 1; do
for v in -1 1 3; do
 for cfg in "256 64"; do
  set -- $cfg
  echo -n "stream_y=$sy variant=$v tile_items=$1 split=$2 : " >> $OUT/sweep.log
  SGLB200_STREAM_Y=$sy SGLB200_SPMM_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --tile-items $1 --split-threshold $2 ${EXTRA} 2>&1 | \
    python -c "import sys,json; l=json.loads(sys.stdin.readlines()[-1]); print('%.1f us/hop  %.2f Gedges/s  frac %.3f cut_rows %d' % (l['roofline']['us_per_launch'], l['value']/1e9, l['roofline']['frac'], l['setup']['cut_rows']))" >> $OUT/sweep.log 2>&1
 done
done
done
cat $OUT/sweep.log
SGLB200_SPMM_VARIANT=1 bash scripts/gpu_prof.sh $TAG/prof --tile-items 256 --split-threshold 64

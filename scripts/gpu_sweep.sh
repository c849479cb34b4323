#!/bin/bash
# parity tests + a schedule sweep of the bench workload (device-timed only)
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -n 4 $OUT/pytest_gpu.log
for ti in ${TILES:-64 128 256 512}; do
  for st in ${SPLITS:-128 256 1024}; do
    echo -n "tile_items=$ti split=$st : " >> $OUT/sweep.log
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --tile-items $ti --split-threshold $st ${EXTRA} 2>&1 | \
      python -c "import sys,json; l=json.loads(sys.stdin.readlines()[-1]); print('%.1f us/hop  %.2f Gedges/s  frac %.3f cut_rows %d' % (l['roofline']['us_per_launch'], l['value']/1e9, l['roofline']['frac'], l['setup']['cut_rows']))" >> $OUT/sweep.log
  done
done
cat $OUT/sweep.log

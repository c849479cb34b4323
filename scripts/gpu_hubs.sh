#!/bin/bash
TAG=${1:-hubs}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -n 3 $OUT/pytest_gpu.log
for h in 0 128 256 448 1024 4096; do
 for v in -1 8; do
  echo -n "hubs=$h variant=$v : " >> $OUT/sweep.log
  SGLB200_HUBS=$h SGLB200_SPMM_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | \
    python -c "import sys,json; l=json.loads(sys.stdin.readlines()[-1]); print('%.1f us/hop  %.2f Gedges/s  frac %.3f' % (l['roofline']['us_per_launch'], l['value']/1e9, l['roofline']['frac']))" >> $OUT/sweep.log 2>&1
 done
done
for wl in products rmat22; do
 for h in 0 448; do
  echo -n "$wl hubs=$h : " >> $OUT/sweep.log
  SGLB200_HUBS=$h python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | \
    python -c "import sys,json; l=json.loads(sys.stdin.readlines()[-1]); print('%.1f us/hop  %.2f Gedges/s  frac %.3f' % (l['roofline']['us_per_launch'], l['value']/1e9, l['roofline']['frac']))" >> $OUT/sweep.log 2>&1
 done
done
cat $OUT/sweep.log
SGLB200_HUBS=448 bash scripts/gpu_prof.sh $TAG/prof

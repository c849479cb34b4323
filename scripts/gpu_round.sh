#!/bin/bash
# full pass: all GPU tests, smoke, default bench, other workloads
TAG=${1:-round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -n 4 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -n 2 $OUT/smoke.log
python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -n 3 $OUT/bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"
for wl in ${WORKLOADS:-products pubmed}; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "$wl exit $?"; tail -n 2 $OUT/bench_$wl.err
done
for f in $OUT/bench*.json; do echo $f; python -c "
import json
try:
    l=json.loads(open('$f').read().strip().splitlines()[-1]); print('  %.2f Gedges/s  %.3f ms/step' % (l['value']/1e9, l['ms_per_step']), 'frac', (l.get('roofline') or {}).get('frac'), 'e2e', (l.get('e2e') or {}).get('value'), 'cpu', (l.get('cpu_baseline') or {}).get('value'))
except Exception as e: print('  failed', e)
"; done

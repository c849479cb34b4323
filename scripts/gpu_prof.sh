#!/bin/bash
# ncu launch list + --set full capture of the hop kernel on the bench workload.  usage: gpu_prof.sh tag [bench args]
TAG=${1:-prof}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spmm_\|agg_\|normalize_values\|build_tiles\|carry_runs -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/ncu_bench.log 2>&1
echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:spmm_flat_kernel -s 20 -c 2 -o $OUT/spmm_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/ncu_full.log 2>&1
echo "ncu full exit $?"
python - <<PY
import csv
from collections import defaultdict
lines=[l for l in open('$OUT/launches.csv') if not l.startswith('==')]
agg=defaultdict(list)
for row in csv.DictReader(lines):
    if row.get('Metric Name')=='gpu__time_duration.sum' and 'sglb200' in row['Kernel Name']:
        agg[row['Kernel Name'][:60]].append(float(row['Metric Value'].replace(',','')))
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print(f"{k:60s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.1f}us")
PY

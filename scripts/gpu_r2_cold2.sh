#!/bin/bash
OUT=gpurun_out/r2_cold2
mkdir -p $OUT
run() {
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-comparators --traffic none "$@" > $OUT/$tag.json 2> $OUT/$tag.err
  python - <<PY
import json
try:
    l=json.loads(open('$OUT/$tag.json').read().strip().splitlines()[-1])
    print('%-40s %8.1f us/hop  frac %.3f  %.2f Gedges/s' % ('$tag', l['roofline']['us_per_launch'], l['roofline']['frac'], l['value']/1e9))
except Exception as e:
    print('$tag', 'FAILED', e, open('$OUT/$tag.err').read()[-300:].replace(chr(10),' '))
PY
}
run products_default X=1 -- --workload products
run products_cold48 SGLB200_COLD_HINT=1 -- --workload products
run products_cold64 SGLB200_COLD_HINT=1 SGLB200_HUB_MB=64 -- --workload products
run rmat22_default X=1 -- --workload rmat22
run rmat22_cold48 SGLB200_COLD_HINT=1 -- --workload rmat22
M=gpu__time_duration.sum,dram__bytes_read.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum
for v in default cold48; do
  if [ $v = cold48 ]; then export SGLB200_COLD_HINT=1; fi
  timeout 280 ncu --metrics $M --clock-control none -k regex:spmm_flat -s 6 -c 1 --csv --log-file $OUT/$v.csv \
    python bench.py --workload products --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-comparators --traffic none > $OUT/$v.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('$OUT/$v.csv')) if len(r)>10]
hdr=rows[0]
print('$v', ' | '.join('%s=%s' % (dict(zip(hdr,r))['Metric Name'].split('.')[0][-40:], dict(zip(hdr,r))['Metric Value']) for r in rows[1:]))
PY
done

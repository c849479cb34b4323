#!/bin/bash
# DRAM bytes / hit rates of the default hop vs the cold-tagged hop on products-shape (metrics pass only, no --set full)
OUT=gpurun_out/r2_cold_ncu
mkdir -p $OUT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__m_xbar2l1tex_read_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed
for v in default cold48; do
  if [ $v = cold48 ]; then export SGLB200_COLD_HINT=1; fi
  timeout 280 ncu --metrics $M --clock-control none -k regex:spmm_flat -s 6 -c 2 --csv --log-file $OUT/$v.csv \
    python bench.py --workload products --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-comparators --traffic none > $OUT/$v.log 2>&1
  echo "$v ncu exit $?"
  python - <<PY
import csv
rows=[r for r in csv.reader(open('$OUT/$v.csv')) if len(r)>10]
hdr=rows[0]; 
for r in rows[1:]:
    d=dict(zip(hdr,r))
    print('$v', d['ID'], d['Metric Name'][:60], d['Metric Value'], d['Metric Unit'])
PY
done

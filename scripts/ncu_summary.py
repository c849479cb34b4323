"""Prints the headline metrics of an .ncu-rep (run here, no GPU needed):  python scripts/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'launch__grid_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active', 'launch__waves_per_multiprocessor']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:72s} {units[i]:10s}", [r[i][:40] for r in data])
d = data[0]
st = [(h, float(d[i].replace(',', ''))) for i, h in enumerate(hdr)
      if h.startswith('smsp__pcsamp_warps_issue_stalled') and 'not_issued' not in h and d[i] not in ('', 'n/a')]
tot = sum(v for _, v in st) or 1
print("stall reasons (first launch):")
for h, v in sorted(st, key=lambda x: -x[1])[:8]:
    print(f"   {h[33:]:30s} {100 * v / tot:5.1f}%")

#!/bin/bash
# round 2, closing call on one GPU: full test suite, smoke, the driver's bench commands, ncu evidence, sanitizers
OUT=gpurun_out/r2_final
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300 | tee $OUT/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
( time timeout 900 python bench.py ) > $OUT/bench_products_n1.json 2> $OUT/bench_products_n1.err; echo "bench exit $?"; tail -4 $OUT/bench_products_n1.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err; echo "ref exit $?"
timeout 300 python bench.py --workload arxiv --steps 20 > $OUT/bench_arxiv_n1.json 2> $OUT/bench_arxiv_n1.err; echo "arxiv exit $?"
timeout 300 python bench.py --workload pubmed --steps 20 > $OUT/bench_pubmed_n1.json 2> $OUT/bench_pubmed_n1.err; echo "pubmed exit $?"
python - <<PY
import json
for f in ("bench_products_n1", "bench_arxiv_n1", "bench_pubmed_n1", "bench_reference_arm"):
    try:
        l = json.loads(open("$OUT/%s.json" % f).read().strip().splitlines()[-1])
        r = l.get("roofline") or {}
        e = l.get("e2e") or {}
        print("%-22s %7.2f Gedges/s  %8.2f ms/step  frac %s  traffic %s  e2e %s  cpu %s" % (f, l["value"] / 1e9, l["ms_per_step"],
              ("%.3f" % r["frac"]) if r else None, r.get("traffic"), ("%.2f" % (e["value"] / 1e9)) if e.get("value") else None,
              ("%.3f" % (l["cpu_baseline"]["value"] / 1e9)) if l.get("cpu_baseline") else None))
    except Exception as exc:
        print(f, "FAILED", exc)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sglb200 -c 200 --csv --log-file $OUT/launches_products.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-comparators --traffic none > $OUT/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_flat_kernel -s 20 -c 1 -o $OUT/spmm_flat_products \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-comparators --traffic none > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"agg_|gather_rows|lw_|it_" -s 14 -c 14 --csv \
    --log-file $OUT/aux_kernels.csv python scripts/aux_kernels_probe.py > $OUT/aux.log 2>&1; echo "ncu aux exit $?"
for TOOL in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $TOOL python scripts/sanitize_smoke.py > $OUT/compute_sanitizer_$TOOL.txt 2>&1
  echo "$TOOL exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize smoke ok" $OUT/compute_sanitizer_$TOOL.txt | head -3
done

#!/bin/bash
OUT=gpurun_out/r2_final2
mkdir -p $OUT
( time timeout 900 python bench.py ) > $OUT/bench_products_n1.json 2> $OUT/bench_products_n1.err; echo "bench exit $?"; tail -4 $OUT/bench_products_n1.err
timeout 300 python bench.py --workload arxiv --steps 20 > $OUT/bench_arxiv_n1.json 2> $OUT/bench_arxiv_n1.err; echo "arxiv exit $?"
timeout 300 python bench.py --workload rmat24 --steps 5 --no-e2e --no-cpu-baseline > $OUT/bench_rmat24_n1.json 2> $OUT/bench_rmat24_n1.err; echo "rmat24 exit $?"
python - <<PY
import json
for f in ("bench_products_n1", "bench_arxiv_n1", "bench_rmat24_n1"):
    try:
        l = json.loads(open("$OUT/%s.json" % f).read().strip().splitlines()[-1])
        r = l.get("roofline") or {}
        e = l.get("e2e") or {}
        print("%-22s %7.2f Gedges/s %8.2f ms/step frac %.3f dramfrac %s gather %.3f e2e %s cpu %s fused %s" % (f, l["value"] / 1e9, l["ms_per_step"],
              r["frac"], r.get("dram_measured_frac"), r["gather_path"]["frac"], ("%.2f" % (e["value"] / 1e9)) if e.get("value") else None,
              ("%.3f" % (l["cpu_baseline"]["value"] / 1e9)) if l.get("cpu_baseline") else None,
              {k: round(v, 1) for k, v in (l.get("fused") or {}).items() if k.endswith("_ms")}))
    except Exception as exc:
        print(f, "FAILED", exc)
PY
for V in 3 7; do
SGLB200_SPMM_VARIANT=$V timeout 300 python bench.py --steps 5 --no-e2e --no-cpu-baseline --no-comparators --traffic none 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $V', round(l['roofline']['us_per_launch'],1), 'us/hop')"
done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | cut -c1-200

#!/bin/bash
# cold-column evict_first tags (SGLB200_COLD_HINT): products / rmat22 per-hop time against the default kernel
OUT=gpurun_out/r2_cold
mkdir -p $OUT
run() {
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-comparators --traffic none "$@" > $OUT/$tag.json 2> $OUT/$tag.err
  python - <<PY
import json
try:
    l=json.loads(open('$OUT/$tag.json').read().strip().splitlines()[-1])
    print('%-40s %8.1f us/hop  frac %.3f  %.2f Gedges/s parity %s' % ('$tag', l['roofline']['us_per_launch'], l['roofline']['frac'], l['value']/1e9, str(l.get('parity'))[:80]))
except Exception as e:
    print('$tag', 'FAILED', e, open('$OUT/$tag.err').read()[-300:].replace(chr(10),' '))
PY
}
run products_default X=1 -- --workload products
run products_cold48 SGLB200_COLD_HINT=1 -- --workload products
run products_cold32 SGLB200_COLD_HINT=1 SGLB200_HUB_MB=32 -- --workload products
run products_cold64 SGLB200_COLD_HINT=1 SGLB200_HUB_MB=64 -- --workload products
run products_cold80 SGLB200_COLD_HINT=1 SGLB200_HUB_MB=80 -- --workload products
run products_cold16 SGLB200_COLD_HINT=1 SGLB200_HUB_MB=16 -- --workload products
run rmat22_default X=1 -- --workload rmat22
run rmat22_cold48 SGLB200_COLD_HINT=1 -- --workload rmat22
run arxiv_cold48 SGLB200_COLD_HINT=1 -- --workload arxiv
run arxiv_default X=1 -- --workload arxiv

#!/bin/bash
# the driver's own commands at N=1: our arm and the reference arm on the default workload
OUT=gpurun_out/r2_bench1
mkdir -p $OUT
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err
echo "bench exit $?"; tail -3 $OUT/bench_n1.err
python - <<PY
import json
l=json.loads(open('$OUT/bench_n1.json').read().strip().splitlines()[-1])
print(json.dumps({k: l[k] for k in ('value','ms_per_step','gpu_launches','clocks','parity')}))
print('roofline', json.dumps(l['roofline']))
print('cpu', json.dumps(l['cpu_baseline']))
e=l['e2e']; print('e2e', {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='api'}) for k,v in e.items() if k!='api'})
print('fused', {k:v for k,v in l['fused'].items() if k!='what'})
print('comparators', {k:{kk:vv for kk,vv in v.items() if kk!='note'} for k,v in l['comparators'].items()})
PY
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "ref exit $?"; tail -3 $OUT/bench_ref.err; cut -c1-400 $OUT/bench_ref.json
timeout 300 python bench.py --workload arxiv --steps 10 --traffic none > $OUT/bench_arxiv.json 2> $OUT/bench_arxiv.err
python - <<PY
import json
l=json.loads(open('$OUT/bench_arxiv.json').read().strip().splitlines()[-1])
print('arxiv', l['value']/1e9, l['roofline']['frac'], 'cpu', l['cpu_baseline']['value']/1e9)
e=l['e2e']; print('e2e', {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='api'}) for k,v in e.items() if k!='api'})
PY

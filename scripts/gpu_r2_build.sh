#!/bin/bash
OUT=gpurun_out/r2_build
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "adjacency or device_built or device_builder or nafs or label_prop" 2>&1 | tail -15 | cut -c1-300
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "native_adjacency_builder_vs_scipy and 300-4000" > $OUT/memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 $OUT/memcheck.log | cut -c1-200
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "native_adjacency_builder_vs_scipy and 300-4000" > $OUT/racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 $OUT/racecheck.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k products 2>&1 | tail -5 | cut -c1-300
python - <<'PY'
import time, torch, bench
from sgl_b200.graph_build import normalized_adjacency_device
dev = torch.device("cuda", 0)
rows, cols, n, d, K = bench.device_graph("products", dev)
for eng in ("native", "torch", "native", "torch"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = normalized_adjacency_device(rows, cols, n, r=0.5, engine=eng)
    torch.cuda.synchronize(); print(eng, "build of products-shape A^ structure: %.1f ms" % (1e3 * (time.perf_counter() - t0)), "nnz", p["indices"].numel())
    del p
PY

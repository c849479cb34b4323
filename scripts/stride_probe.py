"""One hop at the workload's own width with different row strides of the X / Y slabs (alignment of the gathered rows to
32-byte sectors / 128-byte lines).  Usage: python scripts/stride_probe.py [workload] [width]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sgl_b200.graph_build import build_operator_device  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "products"
dev = torch.device("cuda", 0)
rows, cols, n, d, K = bench.device_graph(name, dev)
if len(sys.argv) > 2:
    d = int(sys.argv[2])
op = build_operator_device(rows, cols, n, r=0.5)
del rows, cols
nnz = int(op.nnz)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def time_hop(x, y, reps=5):
    for _ in range(2):
        op.spmm(x, out=y)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        op.spmm(x, out=y)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


xd = torch.randn(n, d, device=dev)
ref = None
for ld in sorted({d, ((d + 7) // 8) * 8, ((d + 15) // 16) * 16, ((d + 31) // 32) * 32}):
    x = torch.empty((n, ld), device=dev)[:, :d]
    y = torch.empty((n, ld), device=dev)[:, :d]
    x.copy_(xd)
    t = time_hop(x, y)
    if ref is None:
        ref = y.clone()
    print(f"{name} d={d} row stride {ld:4d} floats ({ld * 4:4d} B): {t:8.1f} us/hop  {nnz / t / 1e3:6.2f} G edges/s  bit-equal {bool(torch.equal(y, ref))}")
    del x, y

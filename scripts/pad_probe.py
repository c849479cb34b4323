"""Narrow column blocks (the feature split's per-rank work): one hop at width w with the dense row stride vs a row
stride rounded up to 16 floats (FeatureSplitOperator.block_slab).  Usage: python scripts/pad_probe.py [workload]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sgl_b200.dist import FeatureSplitOperator  # noqa: E402
from sgl_b200.graph_build import build_operator_device  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "products"
dev = torch.device("cuda", 0)
rows, cols, n, d, K = bench.device_graph(name, dev)
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


for w in (12, 13, 16, 24, 28, 48, 52):
    xd = torch.randn(n, w, device=dev)
    yd = torch.empty_like(xd)
    t_dense = time_hop(xd, yd)
    xp, yp = FeatureSplitOperator.block_slab(n, w, dev), FeatureSplitOperator.block_slab(n, w, dev)
    xp.copy_(xd)
    t_pad = time_hop(xp, yp)
    same = bool(torch.equal(yp, yd))
    print(f"{name} w={w:3d}  dense stride {t_dense:8.1f} us/hop ({nnz / t_dense / 1e3:6.1f} G nnz/s)   "
          f"stride {xp.stride(0):3d}: {t_pad:8.1f} us/hop ({nnz / t_pad / 1e3:6.1f} G nnz/s)   bit-equal {same}")
    del xd, yd, xp, yp

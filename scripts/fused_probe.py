"""Where does the fused driver's time go?  (run on the GPU box)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sgl_b200.graph_build import build_operator_device

name = sys.argv[1] if len(sys.argv) > 1 else "products"
dev = torch.device("cuda", 0)
rows, cols, n, d, K = bench.device_graph(name, dev)
op = build_operator_device(rows, cols, n, r=0.5)
del rows, cols
x = torch.randn(n, d, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def t(f, reps=3):
    f(); torch.cuda.synchronize()
    ms = 0.0
    for i in range(reps):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / reps

hops = [x] + [torch.empty_like(x) for _ in range(K)]
def plain():
    for k in range(1, K + 1):
        op.spmm(hops[k - 1], out=hops[k])
print(f"{name}: plain K hops                          {t(plain):8.2f} ms")
for label, kw in [("fused norm, keep all, no agg", dict(keep="all", fuse_norm=True)),
                  ("fused norm, keep none, no agg", dict(keep="none", fuse_norm=True)),
                  ("vals stream, keep all, no agg", dict(keep="all", fuse_norm=False)),
                  ("vals stream, keep none, agg mean", dict(keep="none", agg="mean", fuse_norm=False)),
                  ("fused norm, keep none, agg mean", dict(keep="none", agg="mean", fuse_norm=True)),
                  ("fused norm, keep none, agg last", dict(keep="none", agg="last", fuse_norm=True)),
                  ("fused norm, keep none, agg osd", dict(keep="none", agg="osd", fuse_norm=True)),
                  ("fused norm, keep none, agg concat", dict(keep="none", agg="concat", fuse_norm=True))]:
    print(f"{name}: {label:38s} {t(lambda: op.propagate_fused(x, K, **kw)):8.2f} ms")
from sgl_b200.runtime import aggregate
from sgl_b200 import _lib
print(f"{name}: separate mean aggregation pass         {t(lambda: aggregate(_lib.AGG_MEAN, hops)):8.2f} ms")
print(f"{name}: separate osd aggregation pass          {t(lambda: aggregate(_lib.AGG_OSD, hops)):8.2f} ms")

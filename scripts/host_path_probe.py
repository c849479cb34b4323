"""Where does the time of the host-in / host-out propagate go?  (run on the GPU box)"""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sgl_b200.graph_build import build_operator_device

name = sys.argv[1] if len(sys.argv) > 1 else "arxiv"
dev = torch.device("cuda", 0)
rows, cols, n, d, K = bench.device_graph(name, dev)
op = build_operator_device(rows, cols, n, r=0.5)
del rows, cols
x = torch.randn(n, d).pin_memory()
slab = n * d * 4

def t(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

keep = []
print(f"{name}: n={n} d={d} K={K} slab={slab/1e6:.1f} MB")
print("pinned alloc of K slabs, fresh   : %.2f ms" % (1e3 * t(lambda: keep.append([torch.empty((n, d), pin_memory=True) for _ in range(K)]), 2)))
keep.clear()
print("pinned alloc of K slabs, recycled: %.2f ms" % (1e3 * t(lambda: [torch.empty((n, d), pin_memory=True) for _ in range(K)])))
xd = x.to(dev)
outs = [torch.empty((n, d), pin_memory=True) for _ in range(K)]
def d2h():
    for o in outs:
        o.copy_(xd, non_blocking=True)
print("D2H of K slabs into pinned       : %.2f ms  (%.1f GB/s)" % (1e3 * t(d2h), K * slab / t(d2h) / 1e9))
pag = [torch.empty((n, d)) for _ in range(K)]
def d2h_pag():
    for o in pag:
        o.copy_(xd)
print("D2H of K slabs into pageable     : %.2f ms  (%.1f GB/s)" % (1e3 * t(d2h_pag, 2), K * slab / t(d2h_pag, 2) / 1e9))
print("H2D of X from pinned             : %.2f ms" % (1e3 * t(lambda: x.to(dev, non_blocking=True))))
def hops_only():
    op.propagate(xd, K, mode="fast")
print("K hops device resident           : %.2f ms" % (1e3 * t(hops_only)))
print("propagate_host keep=all          : %.2f ms" % (1e3 * t(lambda: op.propagate_host(x, K, mode="fast", keep="all"))))
print("propagate_host keep=last         : %.2f ms" % (1e3 * t(lambda: op.propagate_host(x, K, mode="fast", keep="last"))))

"""Tiny pass over every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
Sizes are small on purpose: the sanitizer serialises and instruments every launch."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgl_b200 import _lib  # noqa: E402
from sgl_b200.operators.message_op import LearnableWeightedMessageOp  # noqa: E402
from sgl_b200.operators.utils import adj_to_symmetric_norm  # noqa: E402
from sgl_b200.runtime import CsrOperator, aggregate, gather_rows  # noqa: E402

rng = np.random.default_rng(0)
n, m = 300, 4000
rows = rng.integers(0, n, m)
cols = (rng.zipf(1.2, m) - 1) % n
adj = sp.csr_matrix((np.ones(2 * m, dtype=np.float32), (np.concatenate([rows, cols]), np.concatenate([cols, rows]))), shape=(n, n))
norm = adj_to_symmetric_norm(adj, 0.5).tocsr()
for d in (128, 100, 47):
    op = CsrOperator.from_scipy(norm, tile_items=32, split_threshold=8)      # many cut rows: exercises the carry path
    x = torch.randn(n, d, device="cuda")
    y_fast = op.spmm(x, mode="fast")
    y_exact = op.spmm(x, mode="exact")
    assert torch.allclose(y_fast, y_exact, rtol=1e-4, atol=1e-5)
    tiles, rows_b = op.chunks(3)
    out = torch.empty_like(y_fast)
    for c in range(3):
        op.spmm_tiles(x, out, tiles[c], tiles[c + 1])
    assert torch.equal(out, y_fast)
    hops = op.propagate(x, 3)
    for code in (_lib.AGG_SUM, _lib.AGG_MEAN, _lib.AGG_MAX, _lib.AGG_MIN, _lib.AGG_CONCAT, _lib.AGG_OSD):
        aggregate(code, hops)
    aggregate(_lib.AGG_WEIGHTED, hops, [0.5, 0.25, 0.125, 0.0625])
    gather_rows(hops, torch.arange(0, n, 7, device="cuda"))
    # round 2 kernels: fused driver (in-kernel fold + fused flush), lane-group kernel, TMA-staged kernel
    for agg in ("mean", "max", "osd", "concat", "last"):
        op.propagate_fused(x, 3, mode="fast", keep="none", agg=agg)
    if d % 4 == 0:
        xn = x[:, :16].contiguous()
        assert torch.allclose(op.spmm(xn, mode="fast"), op.spmm(xn, mode="exact"), rtol=1e-4, atol=1e-5)
        os.environ["SGLB200_TMA"] = "1"
        y_tma = op.spmm(x, mode="fast")
        os.environ.pop("SGLB200_TMA")
        assert torch.allclose(y_tma, y_exact, rtol=1e-4, atol=1e-5)
    op.close()
feats = [torch.randn(64, 16, device="cuda", requires_grad=True) for _ in range(4)]
for kind, args in (("gate", (16,)), ("ori_ref", (16,)), ("jk", (3, 16))):
    lw = LearnableWeightedMessageOp(0, 4, kind, *args).cuda()
    lw.aggregate(feats).sum().backward()
from sgl_b200.graph_build import operator_from_scipy_device  # noqa: E402
from sgl_b200.operators.message_op import IterateLearnableWeightedMessageOp  # noqa: E402
opd = operator_from_scipy_device(adj, r=0.5, alpha=0.15, tile_items=32, split_threshold=8)   # fused normalisation + PPR
opd.propagate_fused(torch.randn(n, 100, device="cuda"), 3, mode="fast", keep="all", agg="mean", fuse_norm=True)
opd.propagate_fused(torch.randn(n, 100, device="cuda"), 3, mode="fast", keep="none", agg="mean")
opd.close()
it = IterateLearnableWeightedMessageOp(0, 4, "recursive", 16).cuda()
it.aggregate(feats).sum().backward()
torch.cuda.synchronize()
print("sanitize smoke ok")

"""One launch of every auxiliary kernel family at a training-batch size, for an ncu line each (run under ncu):
aggregation (mean / NAFS), gather, learnable-weighted (jk) forward+backward, iterate forward+backward."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgl_b200 import _lib
from sgl_b200.runtime import aggregate, gather_rows
from sgl_b200.operators.message_op import LearnableWeightedMessageOp, IterateLearnableWeightedMessageOp

B, d, K = 500_000, 128, 6
feats = [torch.randn(B, d, device="cuda") for _ in range(K + 1)]
for _ in range(2):
    aggregate(_lib.AGG_MEAN, feats)
    aggregate(_lib.AGG_OSD, feats)
    gather_rows(feats, torch.randint(0, B, (100_000,), device="cuda"))
fr = [f.clone().requires_grad_(True) for f in feats]
lw = LearnableWeightedMessageOp(0, K + 1, "jk", K, d).cuda()
it = IterateLearnableWeightedMessageOp(0, K + 1, "recursive", d).cuda()
for _ in range(2):
    lw.aggregate(fr).sum().backward()
    it.aggregate(fr).sum().backward()
torch.cuda.synchronize()
print("bytes per pass: (K'+1) * 4 * B * d =", (K + 2) * 4 * B * d)

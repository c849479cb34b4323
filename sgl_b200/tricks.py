"""tricks.py -- adjacent consumers of the hop kernel, routed through the same handle.

  label_propagation        reference sgl/tricks/utils.py:40-58          out = alpha * A^ out + (1-alpha) * H0, clamped
  nafs_smoothed_features   reference sgl/tasks/node_clustering.py:205-251 and sgl/tasks/link_prediction.py:233-284
                           (the feature construction of the NAFS tasks, which bypasses sgl.operators and re-normalises
                           with scipy + torch.spmm for each of six r values and each hop count)

The reference runs these with torch.spmm on CPU COO tensors; here the hops are sglb200_spmm launches, the NAFS weights
are the OverSmoothDistance kernel, and for the r sweep only the float32 values of ONE resident CSR are rewritten
(sglb200_normalize_values) -- the structure of A^ does not depend on r.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .graph_build import degree_powers, normalized_adjacency_device
from .runtime import CsrOperator, aggregate, require_cuda


def _default_clamp(x):
    return x.clamp_(0., 1.)


@torch.no_grad()
def label_propagation(labels, adj, num_layers, alpha, post_process: Callable = _default_clamp, mask=None,
                      mode: str = "fast"):
    """Same signature and semantics as the reference's label_propagation: `adj` is the already normalised scipy matrix
    (cast to float32 like sparse_mx_to_torch_sparse_tensor does), labels a long vector or a float matrix."""
    require_cuda()
    if labels.dtype == torch.long:
        labels = F.one_hot(labels.reshape(-1)).to(torch.float)
    on_cpu = not labels.is_cuda
    dev = torch.device("cuda", torch.cuda.current_device()) if on_cpu else labels.device
    lab = labels.to(dev, dtype=torch.float32)
    out = lab.clone()
    if mask is not None:
        out = torch.zeros_like(lab)
        m = mask.to(dev) if isinstance(mask, torch.Tensor) else mask
        out[m] = lab[m]
    op = CsrOperator.from_scipy(adj.tocsr())
    try:
        res = (1 - alpha) * out
        fused = post_process is _default_clamp and alpha != 0 and out.shape[1] <= 512
        for _ in range(num_layers):
            if fused:   # scale, residual add and clamp inside the hop kernel's row flush: one launch per layer
                out = op.spmm_axpby(out, alpha, res, clamp=(0.0, 1.0), mode=mode)
            else:
                out = alpha * op.spmm(out.contiguous(), mode=mode) + res
                out = post_process(out)
    finally:
        op.close()
    return out.cpu() if on_cpu else out


@torch.no_grad()
def nafs_smoothed_features(adj, features, hops: int, r_list: Sequence[float] = (0.5, 0.4, 0.3, 0.2, 0.1, 0.0),
                           method: str = "mean", mode: str = "fast") -> torch.Tensor:
    """Node features smoothed the NAFS way: for every r, `hops` propagation steps of D^(r-1) (A+I)^T D^(-r), combined per
    node with softmax-over-hops of the cosine to the raw features, then mean / max / concat over r ('simple': the last
    hop of the first r only).  adj: scipy sparse adjacency (raw), features: [N, d] tensor or array."""
    if method not in ("mean", "max", "concat", "simple"):
        raise ValueError("method must be 'mean', 'max', 'concat' or 'simple'")
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    x = torch.as_tensor(np.asarray(features) if not isinstance(features, torch.Tensor) else features,
                        dtype=torch.float32).to(dev).contiguous()
    coo = adj.tocoo()
    parts = normalized_adjacency_device(torch.from_numpy(coo.row.astype(np.int64)).to(dev),
                                        torch.from_numpy(coo.col.astype(np.int64)).to(dev), adj.shape[0],
                                        torch.from_numpy(np.asarray(coo.data, dtype=np.float32)).to(dev), r=float(r_list[0]))
    op = CsrOperator(parts["indptr"], parts["indices"], None, adj.shape)
    deg = parts["deg"]
    per_r = []
    try:
        for r in r_list:
            dl, dr = degree_powers(deg, float(r))        # numpy pow on the table of distinct integer degrees
            op.normalize_values(parts["raw_w"], dl, dr)
            hop_list = op.propagate(x, hops, mode=mode)
            if method == "simple":
                per_r.append(hop_list[-1])
                break
            per_r.append(aggregate(_lib.AGG_OSD, hop_list))
    finally:
        op.close()
    if method == "mean":
        out = per_r[0].clone()
        for t in per_r[1:]:
            out = out + t
        return (out / len(per_r)).cpu()
    if method == "max":
        return torch.stack(per_r, dim=0).max(0)[0].cpu()
    if method == "concat":
        return torch.cat(per_r, dim=1).cpu()
    return per_r[-1].cpu()


@torch.no_grad()
def nafs_smoothed_features_sweep(adj, features, max_hops: int, r_list: Sequence[float] = (0.5, 0.4, 0.3, 0.2, 0.1, 0.0),
                                 method: str = "mean", mode: str = "fast"):
    """[nafs_smoothed_features(adj, features, h, ...) for h in 1..max_hops] with every hop computed ONCE per r.
    The reference's NAFS tasks call the feature construction from scratch for every hop count
    (sgl/tasks/node_clustering.py:177-179, link_prediction.py:163-165): O(max_hops^2 * len(r_list)) hops.  The hop list of
    h+1 is the hop list of h plus one more hop, so one propagation of max_hops hops per r serves the whole sweep; only the
    over-smoothing-distance combination (a streaming kernel over h+1 slabs) is repeated per hop count."""
    if method not in ("mean", "max", "concat", "simple"):
        raise ValueError("method must be 'mean', 'max', 'concat' or 'simple'")
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    x = torch.as_tensor(np.asarray(features) if not isinstance(features, torch.Tensor) else features,
                        dtype=torch.float32).to(dev).contiguous()
    coo = adj.tocoo()
    parts = normalized_adjacency_device(torch.from_numpy(coo.row.astype(np.int64)).to(dev),
                                        torch.from_numpy(coo.col.astype(np.int64)).to(dev), adj.shape[0],
                                        torch.from_numpy(np.asarray(coo.data, dtype=np.float32)).to(dev), r=float(r_list[0]))
    op = CsrOperator(parts["indptr"], parts["indices"], None, adj.shape)
    deg = parts["deg"]
    per_hop = [[] for _ in range(max_hops)]          # per_hop[h-1] = one tensor per r
    try:
        for r in (r_list[:1] if method == "simple" else r_list):
            dl, dr = degree_powers(deg, float(r))        # numpy pow on the table of distinct integer degrees
            op.normalize_values(parts["raw_w"], dl, dr)
            hop_list = op.propagate(x, max_hops, mode=mode)
            for h in range(1, max_hops + 1):
                per_hop[h - 1].append(hop_list[h] if method == "simple" else aggregate(_lib.AGG_OSD, hop_list[:h + 1]))
    finally:
        op.close()
    outs = []
    for per_r in per_hop:
        if method == "mean":
            acc = per_r[0].clone()
            for t in per_r[1:]:
                acc = acc + t
            outs.append((acc / len(per_r)).cpu())
        elif method == "max":
            outs.append(torch.stack(per_r, dim=0).max(0)[0].cpu())
        elif method == "concat":
            outs.append(torch.cat(per_r, dim=1).cpu())
        else:
            outs.append(per_r[-1].cpu())
    return outs

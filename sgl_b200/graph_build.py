"""graph_build.py -- construction of the normalised adjacency A^ on the device.

Restates what the reference does with scipy on one CPU core
    A~ = A + I                          sgl/operators/utils.py:77      (duplicates of A merged in A's dtype first)
    deg = rowsum(A~)                    utils.py:78                    (float64)
    dL = deg^(r-1), dR = deg^(-r)       utils.py:79-85                 (inf -> 0)
    A^ = (A~ diag(dL))^T diag(dR)       utils.py:87, .tocsr() at graph_op/laplacian_graph_op.py:19
    (1-alpha) A^ + alpha I              graph_op/ppr_graph_op.py:19
as sort / segment-reduce passes over the COO entries with torch on the GPU (memory + plumbing), and the value pass
in libsglb200 (sglb200_normalize_values: IEEE float64 products in the reference's order).  The CSR structure
(indptr, sorted indices) is bit-identical to the reference's; the float32 values are bit-identical when the degree
powers are evaluated on the host (`pow_on="host"`: numpy/glibc pow, O(N)) and the weights are integers (every
dataset of the reference: 1 or 2); `pow_on="device"` may differ by one float64 ulp before the
float32 rounding (CUDA pow is <= 2 ulp).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .runtime import CsrOperator, require_cuda


def _segment_sum_sorted(values: torch.Tensor, first: torch.Tensor) -> torch.Tensor:
    """Sum of `values` over runs of a sorted key; `first` marks run starts.  Deterministic (lengths-based)."""
    starts = torch.nonzero(first, as_tuple=False).reshape(-1)
    lengths = torch.diff(torch.cat([starts, starts.new_tensor([values.numel()])]))
    return torch.segment_reduce(values, "sum", lengths=lengths, unsafe=True)


def merge_duplicates(keys: torch.Tensor, w: torch.Tensor):
    """Sort (key, w) by key and add the weights of equal keys (what csr_matrix((w,(row,col))) does)."""
    keys, order = torch.sort(keys, stable=True)
    w = w[order]
    first = torch.ones_like(keys, dtype=torch.bool)
    first[1:] = keys[1:] != keys[:-1]
    if bool(first.all()):
        return keys, w
    return keys[first], _segment_sum_sorted(w, first)


def normalized_adjacency_device(rows: torch.Tensor, cols: torch.Tensor, n: int, weights: Optional[torch.Tensor] = None,
                                r: float = 0.5, alpha: Optional[float] = None, pow_on: str = "host"):
    """COO of A on the device (int64 ids, duplicates allowed, float32 weights or None = 1) ->
    dict(indptr int64 [n+1], indices int32 [nnz], raw_w float64 [nnz], d_left, d_right float64 [n], deg) on the
    device, describing A^ = diag(dL) (A+I)^T diag(dR) before the value pass."""
    dev = rows.device
    rows = rows.to(torch.int64)
    cols = cols.to(torch.int64)
    w = torch.ones(rows.numel(), dtype=torch.float32, device=dev) if weights is None else weights.to(torch.float32)
    keys, w = merge_duplicates(rows * n + cols, w)                       # A, canonical, float32 sums
    diag = torch.arange(n, dtype=torch.int64, device=dev) * (n + 1)
    keys2 = torch.cat([keys, diag])
    w2 = torch.cat([w.to(torch.float64), torch.ones(n, dtype=torch.float64, device=dev)])
    keys2, w2 = merge_duplicates(keys2, w2)                              # A + I in float64 (w_ii + 1)
    keep = w2 != 0                                                       # scipy's sparse add drops exact zeros
    if not bool(keep.all()):
        keys2, w2 = keys2[keep], w2[keep]
    r_t = torch.div(keys2, n, rounding_mode="floor")                     # row of A~
    c_t = keys2 - r_t * n
    first = torch.ones_like(r_t, dtype=torch.bool)
    first[1:] = r_t[1:] != r_t[:-1]
    deg = torch.zeros(n, dtype=torch.float64, device=dev)
    deg[r_t[first]] = _segment_sum_sorted(w2, first)                     # weighted degrees, column order
    if pow_on == "host":
        d = deg.cpu().numpy()
        with np.errstate(divide="ignore", invalid="ignore"):
            dl, dr = np.power(d, r - 1), np.power(d, -r)
        dl[np.isinf(dl)] = 0.0
        dr[np.isinf(dr)] = 0.0
        d_left, d_right = torch.from_numpy(dl).to(dev), torch.from_numpy(dr).to(dev)
    else:
        d_left, d_right = torch.pow(deg, r - 1), torch.pow(deg, -r)
        d_left[torch.isinf(d_left)] = 0.0
        d_right[torch.isinf(d_right)] = 0.0
    # transpose: entry (j, i) of A~ lives at (i, j) of A^
    keys_t, order = torch.sort(c_t * n + r_t, stable=True)
    out_rows = torch.div(keys_t, n, rounding_mode="floor")
    out_cols = (keys_t - out_rows * n).to(torch.int32)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(torch.bincount(out_rows, minlength=n), 0)
    return {"indptr": indptr, "indices": out_cols, "raw_w": w2[order].contiguous(), "d_left": d_left,
            "d_right": d_right, "deg": deg, "alpha": alpha}


def values_from_parts(parts) -> torch.Tensor:
    """float64 values of A^ from the builder's parts with torch (same IEEE products as sglb200_normalize_values);
    used where the values are needed outside a handle: the lazy scipy view of GraphOp._adj and the row partitioner."""
    indptr, cols = parts["indptr"], parts["indices"].to(torch.int64)
    n = indptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n, device=indptr.device), torch.diff(indptr))
    v = (parts["raw_w"] * parts["d_left"][rows]) * parts["d_right"][cols]
    if parts.get("alpha") is not None:
        a = float(parts["alpha"])
        v = (1 - a) * v
        diag = rows == cols
        v[diag] = v[diag] + a
    return v


def parts_to_scipy(parts):
    """scipy CSR (float64, int32 indices) of A^ -- what the reference keeps in GraphOp._adj."""
    import scipy.sparse as sp
    n = parts["indptr"].numel() - 1
    return sp.csr_matrix((values_from_parts(parts).cpu().numpy(), parts["indices"].cpu().numpy(),
                          parts["indptr"].cpu().numpy().astype(np.int32 if parts["indices"].numel() < 2 ** 31 else np.int64)),
                         shape=(n, n))


def build_operator_device(rows, cols, n, weights=None, r=0.5, alpha=None, pow_on="host", **op_kw) -> CsrOperator:
    """CsrOperator of A^ (LaplacianGraphOp semantics, or PprGraphOp when alpha is given) without any host scipy."""
    require_cuda()
    parts = normalized_adjacency_device(rows, cols, n, weights, r, alpha, pow_on)
    op = CsrOperator(parts["indptr"], parts["indices"], None, (n, n), **op_kw)
    op.normalize_values(parts["raw_w"], parts["d_left"], parts["d_right"], alpha=alpha or 0.0,
                        apply_ppr=alpha is not None)
    op.parts = parts
    return op


def operator_from_scipy_device(adj, r=0.5, alpha=None, pow_on="host", **op_kw) -> CsrOperator:
    """Same, starting from the scipy CSR/COO adjacency the reference API receives (uploaded as COO)."""
    coo = adj.tocoo()
    dev = torch.device("cuda", torch.cuda.current_device())
    rows = torch.from_numpy(coo.row.astype(np.int64)).to(dev)
    cols = torch.from_numpy(coo.col.astype(np.int64)).to(dev)
    w = torch.from_numpy(np.asarray(coo.data, dtype=np.float32)).to(dev)
    return build_operator_device(rows, cols, adj.shape[0], w, r, alpha, pow_on, **op_kw)

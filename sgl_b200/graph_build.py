"""graph_build.py -- construction of the normalised adjacency A^ on the device.

Restates what the reference does with scipy on one CPU core
    A~ = A + I                          sgl/operators/utils.py:77      (duplicates of A merged in A's dtype first)
    deg = rowsum(A~)                    utils.py:78                    (float64)
    dL = deg^(r-1), dR = deg^(-r)       utils.py:79-85                 (inf -> 0)
    A^ = (A~ diag(dL))^T diag(dR)       utils.py:87, .tocsr() at graph_op/laplacian_graph_op.py:19
    (1-alpha) A^ + alpha I              graph_op/ppr_graph_op.py:19
with our own device kernels: sglb200_adjacency_build (csrc/build_adj.cu: LSD radix sort of (row | col | identity) keys,
run fold, per-row degree chains, transposing sort by column) for the structure, degrees and merged weights, and
sglb200_normalize_values for the value pass (IEEE float64 products in the reference's order).  `engine="torch"` keeps the
first version (torch.sort / segment_reduce) as an independent cross-check for the tests.  The CSR structure (indptr,
sorted indices) and the degrees are bit-identical to the reference's; the float32 values are bit-identical when the
degree powers come from numpy's pow (`pow_on="host"`, the default: evaluated on a table of the distinct integer degrees
-- max_deg + 1 entries -- when every degree is an integer, which holds for every dataset of the reference (weights 1 or 2,
SURVEY.md section 9.3); only non-integer degree vectors make the O(N) round trip); `pow_on="device"` may differ by one
float64 ulp before the float32 rounding (CUDA pow is <= 2 ulp).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ctypes import byref, c_int64, c_void_p

from . import _lib
from ._lib import check
from .runtime import CsrOperator, _stream_ptr, require_cuda


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


def adjacency_structure_native(rows: torch.Tensor, cols: torch.Tensor, n: int, weights: Optional[torch.Tensor] = None):
    """(indptr int64 [n+1], indices int32 [nnz], raw_w float64 [nnz], deg float64 [n]) of (A + I)^T through
    sglb200_adjacency_build / _export -- hand-written sort, fold and degree kernels, nothing leaves the device."""
    require_cuda()
    dev = rows.device
    rows = rows.to(torch.int64).contiguous()
    cols = cols.to(torch.int64).contiguous()
    w = None if weights is None else weights.to(device=dev, dtype=torch.float32).contiguous()
    lib = _lib.load()
    h, nnz = c_void_p(), c_int64(0)
    with torch.cuda.device(dev):
        # the builder allocates with cudaMalloc, outside torch's pool: hand cached blocks back when it would not fit otherwise
        need = 56 * (int(rows.numel()) + int(n))
        if torch.cuda.mem_get_info()[0] < need:
            torch.cuda.empty_cache()
        check(lib.sglb200_adjacency_build(byref(h), int(n), int(rows.numel()), c_void_p(rows.data_ptr()),
                                          c_void_p(cols.data_ptr()), None if w is None else c_void_p(w.data_ptr()), 1,
                                          byref(nnz), _stream_ptr()), "adjacency_build")
        try:
            indptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
            indices = torch.empty(nnz.value, dtype=torch.int32, device=dev)
            raw_w = torch.empty(nnz.value, dtype=torch.float64, device=dev)
            deg = torch.empty(n, dtype=torch.float64, device=dev)
            check(lib.sglb200_adjacency_export(h, c_void_p(indptr.data_ptr()), c_void_p(indices.data_ptr()),
                                               c_void_p(raw_w.data_ptr()), c_void_p(deg.data_ptr()), _stream_ptr()),
                  "adjacency_export")
            torch.cuda.current_stream().synchronize()
        finally:
            lib.sglb200_adjacency_free(h)
    return indptr, indices, raw_w, deg


_POW_TABLE_MAX = 1 << 22


def degree_powers(deg: torch.Tensor, r: float, pow_on: str = "host"):
    """(deg^(r-1), deg^(-r)) as float64 device vectors, inf -> 0 (utils.py:79-85).  pow_on="host": numpy's pow, like the
    reference -- on the table 0..max_deg when the degrees are integers (the vector itself never leaves the device)."""
    dev = deg.device
    if pow_on != "host":
        d_left, d_right = torch.pow(deg, r - 1), torch.pow(deg, -r)
        d_left[torch.isinf(d_left)] = 0.0
        d_right[torch.isinf(d_right)] = 0.0
        return d_left, d_right
    top = float(deg.max()) if deg.numel() else 0.0
    integral = deg.numel() > 0 and 0 <= top < _POW_TABLE_MAX and bool((deg == torch.floor(deg)).all()) and float(deg.min()) >= 0
    src = np.arange(int(top) + 1, dtype=np.float64) if integral else deg.cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        dl, dr = np.power(src, r - 1), np.power(src, -r)
    dl[np.isinf(dl)] = 0.0
    dr[np.isinf(dr)] = 0.0
    d_left, d_right = torch.from_numpy(dl).to(dev), torch.from_numpy(dr).to(dev)
    if integral:
        idx = deg.to(torch.int64)
        d_left, d_right = d_left[idx], d_right[idx]
    return d_left, d_right


def normalized_adjacency_device(rows: torch.Tensor, cols: torch.Tensor, n: int, weights: Optional[torch.Tensor] = None,
                                r: float = 0.5, alpha: Optional[float] = None, pow_on: str = "host", engine: str = "native"):
    """COO of A on the device (int64 ids, duplicates allowed, float32 weights or None = 1) ->
    dict(indptr int64 [n+1], indices int32 [nnz], raw_w float64 [nnz], d_left, d_right float64 [n], deg) on the
    device, describing A^ = diag(dL) (A+I)^T diag(dR) before the value pass."""
    if engine == "native":
        indptr, indices, raw_w, deg = adjacency_structure_native(rows, cols, n, weights)
        d_left, d_right = degree_powers(deg, r, pow_on)
        return {"indptr": indptr, "indices": indices, "raw_w": raw_w, "d_left": d_left, "d_right": d_right, "deg": deg,
                "alpha": alpha}
    if engine != "torch":
        raise ValueError("engine must be 'native' or 'torch'")
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
    d_left, d_right = degree_powers(deg, r, pow_on)
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


def build_operator_device(rows, cols, n, weights=None, r=0.5, alpha=None, pow_on="host", engine="native", **op_kw) -> CsrOperator:
    """CsrOperator of A^ (LaplacianGraphOp semantics, or PprGraphOp when alpha is given) without any host scipy."""
    require_cuda()
    parts = normalized_adjacency_device(rows, cols, n, weights, r, alpha, pow_on, engine)
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

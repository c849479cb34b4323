"""runtime.py -- device-resident CSR operator and launchers over the C ABI (include/sglb200.h).

PyTorch is used for device memory, streams and pinned host buffers only; all arithmetic runs in libsglb200.so.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int64, c_void_p
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import SglB200Error, check

_MODES = {"fast": _lib.MODE_FAST, "exact": _lib.MODE_EXACT, 0: 0, 1: 1}


def require_cuda() -> None:
    """Fail loudly when there is no GPU: this package never computes on the CPU."""
    if not torch.cuda.is_available():
        raise SglB200Error("no CUDA device visible: sgl_b200 runs on B200 (sm_100a) only and has no CPU fallback")
    _lib.load()


def _stream_ptr() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _as_f32_host(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


class CsrOperator:
    """A sparse operator A (n_rows x n_cols, CSR, float32 values) resident in HBM, plus its warp schedules.

    Created once per normalised adjacency; every hop reuses it (the reference re-casts and re-copies per hop,
    sgl/operators/utils.py:31-35).
    """

    def __init__(self, indptr, indices, vals, shape, *, tile_items: int = 0, split_threshold: int = 0,
                 device: Optional[int] = None):
        require_cuda()
        lib = _lib.load()
        if device is not None:
            torch.cuda.set_device(device)
        torch.cuda.init()
        torch.cuda.current_stream()  # make sure the primary context is current on this thread
        self.shape = (int(shape[0]), int(shape[1]))
        self._keep = []
        on_device = isinstance(indptr, torch.Tensor)
        if on_device:
            if not (indptr.is_cuda and indices.is_cuda and (vals is None or vals.is_cuda)):
                raise TypeError("CsrOperator: tensor inputs must all be CUDA tensors")
            if indptr.dtype not in (torch.int32, torch.int64):
                raise TypeError("indptr must be int32 or int64")
            indptr = indptr.contiguous()
            indices = indices.to(torch.int32).contiguous()
            vals = None if vals is None else vals.to(torch.float32).contiguous()
            is64 = indptr.dtype == torch.int64
            nnz = int(indices.numel())
            p_indptr, p_indices = indptr.data_ptr(), indices.data_ptr()
            p_vals = None if vals is None else vals.data_ptr()
            loc = _lib.DEVICE
            self._keep = [indptr, indices, vals]
        else:
            indptr = np.ascontiguousarray(indptr)
            if indptr.dtype not in (np.int32, np.int64):
                indptr = indptr.astype(np.int64)
            indices = np.ascontiguousarray(indices, dtype=np.int32)
            vals = None if vals is None else np.ascontiguousarray(vals, dtype=np.float32)
            is64 = indptr.dtype == np.int64
            nnz = int(indices.shape[0])
            p_indptr, p_indices = indptr.ctypes.data, indices.ctypes.data
            p_vals = None if vals is None else vals.ctypes.data
            loc = _lib.HOST
            self._keep = [indptr, indices, vals]
        if indptr.shape[0] != self.shape[0] + 1:
            raise ValueError("indptr length does not match the number of rows")
        self.nnz = nnz
        handle = c_void_p()
        check(lib.sglb200_graph_create(ctypes.byref(handle), self.shape[0], self.shape[1], nnz, c_void_p(p_indptr),
                                       int(is64), c_void_p(p_indices), c_void_p(p_vals) if p_vals else None, loc,
                                       int(tile_items), int(split_threshold), _stream_ptr()), "graph_create")
        torch.cuda.current_stream().synchronize()
        self._keep = []
        self._h = handle
        self.device = torch.device("cuda", torch.cuda.current_device())

    # ---- construction helpers ---------------------------------------------------------------------------------
    @classmethod
    def from_scipy(cls, csr, **kw) -> "CsrOperator":
        """From a scipy CSR matrix (or any object with indptr/indices/data/shape).  float64 values are cast to
        float32 once -- the same rounding the reference applies every hop (sgl/operators/utils.py:32)."""
        return cls(csr.indptr, csr.indices, np.asarray(csr.data, dtype=np.float32), csr.shape, **kw)

    # ---- bookkeeping ------------------------------------------------------------------------------------------
    def info(self) -> dict:
        buf = (c_int64 * 9)()
        check(_lib.load().sglb200_graph_info(self._h, buf), "graph_info")
        keys = ["n_rows", "n_cols", "nnz", "tiles_fast", "carry_runs", "tiles_exact", "tile_items", "split_threshold",
                "bytes_resident"]
        return dict(zip(keys, [int(v) for v in buf]))

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.load().sglb200_graph_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_values(self, vals) -> None:
        if isinstance(vals, torch.Tensor) and vals.is_cuda:
            v = vals.to(torch.float32).contiguous()
            check(_lib.load().sglb200_graph_set_values(self._h, c_void_p(v.data_ptr()), _lib.DEVICE, _stream_ptr()))
            torch.cuda.current_stream().synchronize()
        else:
            v = np.ascontiguousarray(vals, dtype=np.float32)
            check(_lib.load().sglb200_graph_set_values(self._h, c_void_p(v.ctypes.data), _lib.HOST, _stream_ptr()))

    def normalize_values(self, raw_w, d_left, d_right, alpha: float = 0.0, apply_ppr: bool = False) -> None:
        """vals[i,j] = fl32((1-alpha) * ((w*dL[i])*dR[j]) + alpha*[i==j]) in float64 on the device (a4)."""
        if isinstance(raw_w, torch.Tensor) and raw_w.is_cuda:
            ts = [t.to(device=raw_w.device, dtype=torch.float64).contiguous() for t in (raw_w, d_left, d_right)]
            with torch.cuda.device(self.device):
                check(_lib.load().sglb200_normalize_values(self._h, *[c_void_p(t.data_ptr()) for t in ts], float(alpha),
                                                           int(apply_ppr), _lib.DEVICE, _stream_ptr()), "normalize_values")
                torch.cuda.current_stream().synchronize()
            return
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (raw_w, d_left, d_right)]
        check(_lib.load().sglb200_normalize_values(self._h, *[c_void_p(a.ctypes.data) for a in arrs], float(alpha),
                                                   int(apply_ppr), _lib.HOST, _stream_ptr()), "normalize_values")

    # ---- one hop ----------------------------------------------------------------------------------------------
    def spmm(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, mode="fast", accumulate: bool = False):
        """Y = A x (+ Y when accumulate).  x: CUDA float32 [n_cols, d] (row stride may exceed d)."""
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
            raise TypeError("spmm: x must be a 2-D CUDA float32 tensor")
        if x.shape[0] != self.shape[1]:
            raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
        if x.stride(1) != 1:
            x = x.contiguous()
        d = int(x.shape[1])
        if out is None:
            if accumulate:
                raise ValueError("spmm: accumulate needs an existing out tensor")
            out = torch.empty((self.shape[0], d), dtype=torch.float32, device=x.device)
        if out.shape != (self.shape[0], d) or out.stride(1) != 1 or not out.is_cuda:
            raise ValueError("spmm: bad out tensor")
        ldx = int(x.stride(0)) if x.shape[0] > 1 else max(d, int(x.stride(0)))
        ldy = int(out.stride(0)) if out.shape[0] > 1 else max(d, int(out.stride(0)))
        with torch.cuda.device(self.device):      # the handle lives on self.device; launch there, on its current stream
            check(_lib.load().sglb200_spmm(self._h, c_void_p(x.data_ptr()), ldx, c_void_p(out.data_ptr()), ldy, d,
                                           _MODES[mode], int(accumulate), _stream_ptr()), "spmm")
        return out

    def spmm_axpby(self, x: torch.Tensor, alpha: float, res: Optional[torch.Tensor] = None, clamp=None, mode="fast",
                   out: Optional[torch.Tensor] = None):
        """out = clamp(alpha * (A x) + res, *clamp) in one launch (label propagation layer, sglb200_spmm_axpby)."""
        x = x.contiguous()
        d = int(x.shape[1])
        if out is None:
            out = torch.empty((self.shape[0], d), dtype=torch.float32, device=x.device)
        if res is not None:
            res = res.contiguous()
        lo, hi = (float(clamp[0]), float(clamp[1])) if clamp is not None else (0.0, 0.0)
        with torch.cuda.device(self.device):
            check(_lib.load().sglb200_spmm_axpby(self._h, c_void_p(x.data_ptr()), d, c_void_p(out.data_ptr()), d, d, _MODES[mode],
                                                 float(alpha), None if res is None else c_void_p(res.data_ptr()), d,
                                                 int(clamp is not None), lo, hi, _stream_ptr()), "spmm_axpby")
        return out

    def chunks(self, n_chunks: int, mode="fast"):
        """(tile_bounds, row_bounds) of n_chunks consecutive tile ranges of the schedule (sglb200_graph_chunks)."""
        tb = (c_int64 * (n_chunks + 1))()
        rb = (c_int64 * (n_chunks + 1))()
        check(_lib.load().sglb200_graph_chunks(self._h, _MODES[mode], int(n_chunks), tb, rb), "graph_chunks")
        return [int(v) for v in tb], [int(v) for v in rb]

    def spmm_tiles(self, x: torch.Tensor, out: torch.Tensor, tile_begin: int, tile_end: int, mode="fast"):
        """The hop restricted to the tiles [tile_begin, tile_end): fills the rows that range finishes."""
        d = int(x.shape[1])
        ldx = int(x.stride(0)) if x.shape[0] > 1 else d
        ldy = int(out.stride(0)) if out.shape[0] > 1 else d
        with torch.cuda.device(self.device):
            check(_lib.load().sglb200_spmm_tiles(self._h, c_void_p(x.data_ptr()), ldx, c_void_p(out.data_ptr()), ldy, d,
                                                 _MODES[mode], int(tile_begin), int(tile_end), _stream_ptr()), "spmm_tiles")
        return out

    # ---- K hops, device resident ------------------------------------------------------------------------------
    def propagate(self, x: torch.Tensor, prop_steps: int, mode="fast", concat: bool = False) -> List[torch.Tensor]:
        """[x, A x, ..., A^K x] as CUDA tensors.  concat=True lays the K+1 slabs out as column blocks of one
        [n, (K+1)*d] buffer (the ConcatMessageOp result, written by the hop kernels directly)."""
        if self.shape[0] != self.shape[1]:
            raise ValueError("propagate needs a square operator")
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
            raise TypeError("propagate: x must be a 2-D CUDA float32 tensor")
        n, d = int(x.shape[0]), int(x.shape[1])
        if n != self.shape[1]:
            raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
        K = int(prop_steps)
        if concat:
            slab = torch.empty((n, (K + 1) * d), dtype=torch.float32, device=x.device)
            hops = [slab[:, k * d:(k + 1) * d] for k in range(K + 1)]
            hops[0].copy_(x)
            ld = (K + 1) * d
        else:
            x = x.contiguous()
            hops = [x] + [torch.empty_like(x) for _ in range(K)]
            ld = d
        if n and d and K:
            with torch.cuda.device(self.device):
                check(_lib.load().sglb200_propagate(self._h, _lib.ptr_array([h.data_ptr() for h in hops]), ld, d, K,
                                                    _MODES[mode], _stream_ptr()), "propagate")
        return hops

    # ---- K hops + degree normalisation + cross-hop aggregation in one pass per hop -----------------------------
    def propagate_fused(self, x: torch.Tensor, prop_steps: int, mode="fast", keep="none", agg: Optional[str] = None,
                        start: int = 0, end: Optional[int] = None, weights=None, fuse_norm: bool = False):
        """K hops through sglb200_propagate_fused.  keep: "none" | "last" | "all" -- which hops are stored;
        agg: None | "sum" | "mean" | "max" | "min" | "weighted" | "concat" | "osd" | "last" over hops [start, end)
        (reference message_op/*.py; weights: one float per hop 0..K for "weighted").  Returns (hops, out): hops is a list
        of K+1 entries (None where a hop was not stored; entry 0 is x), out the aggregate or None.
        With fuse_norm=True and FAST mode on an operator built by sgl_b200.graph_build the normalised values are never
        read: the kernel streams the raw weights and applies deg^(r-1) / deg^(-r) in the row flush.  Off by default: it
        saves the value array (4 bytes per edge of HBM) but measures 35 % slower per hop than the exact-values stream on
        products-shape (profiles/r02_fused_driver.txt)."""
        if self.shape[0] != self.shape[1]:
            raise ValueError("propagate needs a square operator")
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
            raise TypeError("propagate_fused: x must be a 2-D CUDA float32 tensor")
        n, d = int(x.shape[0]), int(x.shape[1])
        if n != self.shape[1]:
            raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
        if x.stride(1) != 1:
            x = x.contiguous()
        K = int(prop_steps)
        end = K + 1 if end is None else int(end)
        ops = {None: -1, "sum": _lib.AGG_SUM, "mean": _lib.AGG_MEAN, "max": _lib.AGG_MAX, "min": _lib.AGG_MIN,
               "weighted": _lib.AGG_WEIGHTED, "concat": _lib.AGG_CONCAT, "osd": _lib.AGG_OSD, "last": _lib.AGG_LAST}
        if agg not in ops:
            raise ValueError(f"unknown aggregation {agg!r}")
        hops: List[Optional[torch.Tensor]] = [x] + [None] * K
        if keep == "all":
            hops = [x] + [torch.empty((n, d), dtype=torch.float32, device=x.device) for _ in range(K)]
        elif keep == "last" and K > 0 and agg != "last":
            hops[K] = torch.empty((n, d), dtype=torch.float32, device=x.device)
        out = None
        if agg is not None:
            out = torch.empty((n, (end - start) * d if agg == "concat" else d), dtype=torch.float32, device=x.device)
        w = None
        if weights is not None:
            wl = [float(v) for v in weights]
            if len(wl) != K + 1:
                raise ValueError("propagate_fused: one weight per hop 0..K")
            w = (c_float * (K + 1))(*wl)
        ptrs = _lib.ptr_array([None] + [None if h is None else h.data_ptr() for h in hops[1:]])
        ldx = int(x.stride(0)) if n > 1 else max(d, int(x.stride(0)))
        if n and d:
            with torch.cuda.device(self.device):
                check(_lib.load().sglb200_propagate_fused(
                    self._h, c_void_p(x.data_ptr()), ldx, ptrs, d, d, K, _MODES[mode], ops[agg], int(start), int(end), w,
                    None if out is None else c_void_p(out.data_ptr()), 0 if out is None else int(out.shape[1]),
                    int(bool(fuse_norm)), _stream_ptr()), "propagate_fused")
        if agg == "last" and keep in ("last", "all"):
            hops[K] = out
        return hops, out

    # ---- K hops, host in / host out ---------------------------------------------------------------------------
    def propagate_host(self, x, prop_steps: int, mode="fast", keep: str = "all", pin: bool = True) -> List[torch.Tensor]:
        """Host [n,d] features in, K host tensors out (hop 1..K; keep='last' downloads only hop K).  Uploads, the
        K hops and the downloads overlap on two streams inside the library (sglb200_propagate_host)."""
        x = _as_f32_host(x)
        if x.ndim != 2 or x.shape[0] != self.shape[1]:
            raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
        n, d = x.shape
        K = int(prop_steps)
        outs: List[Optional[torch.Tensor]] = []
        for k in range(1, K + 1):
            if keep == "all" or k == K:
                outs.append(torch.empty((n, d), dtype=torch.float32, pin_memory=pin))
            else:
                outs.append(None)
        if K and n and d:
            ptrs = _lib.ptr_array([None if o is None else o.data_ptr() for o in outs])
            with torch.cuda.device(self.device):
                check(_lib.load().sglb200_propagate_host(self._h, c_void_p(x.ctypes.data), ptrs, int(d), K, _MODES[mode]),
                      "propagate_host")
        return outs


# ---------------------------------------------------------------------------------------------------------------
# aggregation launchers (device tensors in, device tensor out)
# ---------------------------------------------------------------------------------------------------------------
def _check_feats(feats: Sequence[torch.Tensor]):
    f0 = feats[0]
    if not all(isinstance(f, torch.Tensor) and f.is_cuda and f.dtype == torch.float32 and f.dim() == 2 for f in feats):
        raise TypeError("aggregate: feature matrices must be 2-D CUDA float32 tensors")
    if not all(f.shape == f0.shape for f in feats):
        raise ValueError("aggregate: feature matrices differ in shape")
    n, d = int(f0.shape[0]), int(f0.shape[1])
    ld = int(f0.stride(0)) if n > 1 else max(d, int(f0.stride(0)))
    fixed = []
    for f in feats:
        fld = int(f.stride(0)) if n > 1 else ld
        if f.stride(1) != 1 or fld != ld:
            f = f.contiguous()
            fld = d
        fixed.append((f, fld))
    if any(fld != fixed[0][1] for _, fld in fixed):
        fixed = [(f.contiguous(), d) for f, _ in fixed]
    return [f for f, _ in fixed], n, d, fixed[0][1]


def aggregate(op: int, feats: Sequence[torch.Tensor], weights=None) -> torch.Tensor:
    require_cuda()
    feats, n, d, ld = _check_feats(list(feats))
    k = len(feats)
    out_cols = k * d if op == _lib.AGG_CONCAT else d
    out = torch.empty((n, out_cols), dtype=torch.float32, device=feats[0].device)
    w = None
    if weights is not None:
        wl = [float(v) for v in weights]
        if len(wl) != k:
            raise ValueError("The feature list and the weight list have different lengths!")
        w = (c_float * k)(*wl)
    if n and d:
        with torch.cuda.device(feats[0].device):
            check(_lib.load().sglb200_aggregate(int(op), _lib.ptr_array([f.data_ptr() for f in feats]), k, n, d, ld, w,
                                                c_void_p(out.data_ptr()), out_cols, _stream_ptr()), "aggregate")
    return out


def gather_rows(feats: Sequence[torch.Tensor], idx: torch.Tensor) -> List[torch.Tensor]:
    """outs[k] = feats[k][idx] for all hops in one launch (f1: device-resident feature store for forward())."""
    require_cuda()
    feats, n, d, ld = _check_feats(list(feats))
    idx = idx.to(device=feats[0].device, dtype=torch.int64).contiguous()
    B = int(idx.numel())
    outs = [torch.empty((B, d), dtype=torch.float32, device=feats[0].device) for _ in feats]
    if B and d:
        with torch.cuda.device(feats[0].device):
            check(_lib.load().sglb200_gather_rows(_lib.ptr_array([f.data_ptr() for f in feats]), len(feats), ld,
                                                  c_void_p(idx.data_ptr()), B, d,
                                                  _lib.ptr_array([o.data_ptr() for o in outs]), d, _stream_ptr()),
                  "gather_rows")
    return outs

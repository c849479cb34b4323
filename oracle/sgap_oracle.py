"""oracle/sgap_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy + the C file next to it) of the reference's SGAP pre-processing path:

  adj_to_symmetric_norm      reference sgl/operators/utils.py:76-88
  LaplacianGraphOp adjacency reference sgl/operators/graph_op/laplacian_graph_op.py:12-19
  PprGraphOp adjacency       reference sgl/operators/graph_op/ppr_graph_op.py:13-21
  GraphOp.propagate          reference sgl/operators/base_op.py:19-36
  csr_sparse_dense_matmul    reference sgl/operators/utils.py:10-40 -> csrc/matmul.c:23-40
  MessageOp combiners        reference sgl/operators/message_op/*.py (cited per function below)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker.  The product package sgl_b200 never imports it.

Parity pinning.  The reference has NO tests and NO golden vectors for this path, so the
oracle is pinned against outputs of the reference itself: oracle/gen_golden.py imports the unmodified
sgl.operators from /root/reference in the build container, runs it on small seeded graphs and stores inputs
and outputs under tests/golden/*.npz (committed together with the generator).  tests/test_oracle.py checks
every function here against those fixtures (bit-exact for CSR structure and the fma hop, exact for the
combiners) and, when oracle/_ref/libmatmul_ref.so (the reference's own matmul.c compiled in place) is
present, against that library on random inputs.

Third-party pieces of the algorithm: scipy.sparse (unpinned in the reference's requirements.txt:5; 1.18.1 in
this image) supplies sparse add / diag products / transpose / tocsr and csr row sums.  Their published
behaviour is restated explicitly here (duplicate summation in the storage dtype, canonical sorted CSR,
row sums through numpy.add.reduceat in float64) instead of being called.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle` (or __graft_entry__.build())")
        lib = ctypes.CDLL(path)
        f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
        f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
        i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
        for name, ip in (("oracle_spmm_f32_fma_i32", i32p), ("oracle_spmm_f32_fma_i64", i64p),
                         ("oracle_spmm_f32_muladd_i32", i32p), ("oracle_spmm_f32_muladd_i64", i64p)):
            fn = getattr(lib, name)
            fn.argtypes = [f32p, f32p, i32p, ip, f32p, ctypes.c_int64, ctypes.c_int64]
            fn.restype = None
        lib.oracle_spmm_f64_i64.argtypes = [f64p, f64p, i32p, i64p, f64p, ctypes.c_int64, ctypes.c_int64]
        lib.oracle_spmm_f64_i64.restype = None
        lib.oracle_num_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def num_threads() -> int:
    return int(_lib().oracle_num_threads())


# ----------------------------------------------------------------------------------------------------------
# CSR container (plain numpy; no scipy objects inside the oracle)
# ----------------------------------------------------------------------------------------------------------
class Csr:
    """Minimal CSR triple.  indptr int64, indices int32, data float64 or float32."""

    def __init__(self, indptr, indices, data, shape):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data)
        self.shape = (int(shape[0]), int(shape[1]))

    @property
    def nnz(self):
        return int(self.indptr[-1])


def _coo_from_any(adj):
    """Accept a scipy csr/coo matrix (only attribute access, no scipy calls) or a Csr; return row, col, data."""
    if isinstance(adj, Csr):
        n = adj.shape[0]
        row = np.repeat(np.arange(n, dtype=np.int64), np.diff(adj.indptr))
        return row, adj.indices.astype(np.int64), adj.data, adj.shape
    fmt = getattr(adj, "format", None)
    if fmt == "csr":
        n = adj.shape[0]
        row = np.repeat(np.arange(n, dtype=np.int64), np.diff(adj.indptr).astype(np.int64))
        return row, adj.indices.astype(np.int64), adj.data, adj.shape
    if fmt == "coo":
        return adj.row.astype(np.int64), adj.col.astype(np.int64), adj.data, adj.shape
    raise TypeError("The adjacency matrix must be a scipy.sparse.coo_matrix/csr_matrix!")


def _canonical_csr(row, col, data, shape) -> Csr:
    """COO -> canonical CSR: sort by (row, col), add duplicates left to right IN THE STORAGE DTYPE
    (what scipy's coo.tocsr()/csr_sum_duplicates do before anything is upcast)."""
    order = np.lexsort((col, row))
    row, col, data = row[order], col[order], data[order]
    if row.size:
        first = np.empty(row.size, dtype=bool)
        first[0] = True
        first[1:] = (row[1:] != row[:-1]) | (col[1:] != col[:-1])
        starts = np.flatnonzero(first)
        if starts.size != row.size:
            # sequential left-to-right addition in the storage dtype, like csr_sum_duplicates
            merged = data[starts].copy()
            seg = np.cumsum(first) - 1
            dup = np.flatnonzero(~first)
            for j in dup:  # duplicates are rare; plain loop keeps the addition order explicit
                merged[seg[j]] = merged[seg[j]] + data[j]
            data = merged
            row, col = row[starts], col[starts]
    indptr = np.zeros(shape[0] + 1, dtype=np.int64)
    np.add.at(indptr, row + 1, 1)
    np.cumsum(indptr, out=indptr)
    return Csr(indptr, col, data, shape)


# ----------------------------------------------------------------------------------------------------------
# a4: adj_to_symmetric_norm   (reference sgl/operators/utils.py:76-88)
# ----------------------------------------------------------------------------------------------------------
def add_self_loops(adj) -> Csr:
    """A + I in float64 (utils.py:77).  Duplicates of A are merged first in A's dtype, the identity is float64, so
    the sum is float64; an existing diagonal entry w becomes w + 1."""
    row, col, data, shape = _coo_from_any(adj)
    a = _canonical_csr(row, col, np.asarray(data), shape)
    n = shape[0]
    arow = np.repeat(np.arange(n, dtype=np.int64), np.diff(a.indptr))
    acol = a.indices.astype(np.int64)
    adat = a.data.astype(np.float64)
    is_diag = arow == acol
    has_diag = np.zeros(n, dtype=bool)
    has_diag[arow[is_diag]] = True
    adat = adat.copy()
    adat[is_diag] = adat[is_diag] + 1.0
    miss = np.flatnonzero(~has_diag)
    row2 = np.concatenate([arow, miss])
    col2 = np.concatenate([acol, miss])
    dat2 = np.concatenate([adat, np.ones(miss.size, dtype=np.float64)])
    order = np.lexsort((col2, row2))
    row2, col2, dat2 = row2[order], col2[order], dat2[order]
    # scipy's csr + csr drops entries whose sum is exactly zero (csr_binop_csr keeps result != 0)
    keep = dat2 != 0.0
    row2, col2, dat2 = row2[keep], col2[keep], dat2[keep]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, row2 + 1, 1)
    np.cumsum(indptr, out=indptr)
    return Csr(indptr, col2, dat2, shape)


def weighted_degrees(a_tilde: Csr) -> np.ndarray:
    """Row sums of A+I in float64 (utils.py:78).  scipy 1.18 reduces the minor axis with numpy.add.reduceat over
    the float64 data (scipy/sparse/_compressed.py sum -> _minor_reduce), so the same numpy primitive is used."""
    n = a_tilde.shape[0]
    deg = np.zeros(n, dtype=np.float64)
    nonempty = np.flatnonzero(np.diff(a_tilde.indptr))
    if nonempty.size:
        deg[nonempty] = np.add.reduceat(a_tilde.data, a_tilde.indptr[:-1][nonempty])
    return deg


def norm_vectors(deg: np.ndarray, r: float):
    """dL = deg^(r-1), dR = deg^(-r) with inf -> 0 (utils.py:79-85)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        d_left = np.power(deg, r - 1)
        d_right = np.power(deg, -r)
    d_left[np.isinf(d_left)] = 0.0
    d_right[np.isinf(d_right)] = 0.0
    return d_left, d_right


def symmetric_norm_csr(adj, r: float) -> Csr:
    """CSR (float64 data, canonical) of  (A~ . diag(dL))^T . diag(dR)  (utils.py:87):
    entry (i, j) = fl64(fl64(A~[j, i] * dL[i]) * dR[j])   (probed)."""
    at = add_self_loops(adj)
    deg = weighted_degrees(at)
    d_left, d_right = norm_vectors(deg, r)
    n = at.shape[0]
    src = np.repeat(np.arange(n, dtype=np.int64), np.diff(at.indptr))  # j : row of A~
    dst = at.indices.astype(np.int64)                                  # i : col of A~
    # first product of the reference: (A~ diag(dL))[j, i] = A~[j, i] * dL[i]
    v = at.data * d_left[dst]
    # transpose: entry lives at (i, j); second product: * dR[j]
    v = v * d_right[src]
    order = np.lexsort((src, dst))
    out_row, out_col, v = dst[order], src[order], v[order]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, out_row + 1, 1)
    np.cumsum(indptr, out=indptr)
    return Csr(indptr, out_col, v, at.shape)


def laplacian_adj(adj, r: float = 0.5) -> Csr:
    """LaplacianGraphOp._construct_adj (graph_op/laplacian_graph_op.py:12-19)."""
    return symmetric_norm_csr(adj, r)


def ppr_adj(adj, r: float = 0.5, alpha: float = 0.15) -> Csr:
    """PprGraphOp._construct_adj (graph_op/ppr_graph_op.py:13-21): (1-alpha) * A^ + alpha * I, float64.
    Every row of A^ already holds its diagonal (self loops were added), so the pattern is unchanged unless a
    diagonal sum cancels to exactly zero, which scipy would prune; that case is kept as is and flagged by tests."""
    a = symmetric_norm_csr(adj, r)
    n = a.shape[0]
    row = np.repeat(np.arange(n, dtype=np.int64), np.diff(a.indptr))
    data = (1 - alpha) * a.data
    diag = row == a.indices
    if int(diag.sum()) != n:
        raise NotImplementedError("oracle: a row without a stored diagonal (cancelled self loop) is not restated")
    data[diag] = data[diag] + alpha * 1.0
    return Csr(a.indptr, a.indices, data, a.shape)


# ----------------------------------------------------------------------------------------------------------
# a5/a6: one hop, and a1: the K-hop loop
# ----------------------------------------------------------------------------------------------------------
def spmm_hop(adj: Csr, x: np.ndarray, flavour: str = "fma") -> np.ndarray:
    """One hop Y = A^ X.
    flavour 'fma'    : float32 values (cast per hop, utils.py:32), float32 fused chain == shipped libmatmul.so
    flavour 'muladd' : same with separate multiply/add                    == scipy csr(float32).dot(X)
    flavour 'f64'    : float64 values and accumulation on float64 X       == base_op.py:34 (non-Linux branch)"""
    n, d = x.shape
    if adj.shape[1] != n:
        raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
    lib = _lib()
    if flavour == "f64":
        y = np.zeros((adj.shape[0], d), dtype=np.float64)
        lib.oracle_spmm_f64_i64(y, adj.data.astype(np.float64), adj.indices, adj.indptr,
                                np.ascontiguousarray(x, dtype=np.float64), adj.shape[0], d)
        return y
    y = np.zeros((adj.shape[0], d), dtype=np.float32)
    a32 = adj.data.astype(np.float32)
    x32 = np.ascontiguousarray(x, dtype=np.float32)
    fn = lib.oracle_spmm_f32_fma_i64 if flavour == "fma" else lib.oracle_spmm_f32_muladd_i64
    fn(y, a32, adj.indices, adj.indptr, x32, adj.shape[0], d)
    return y


def propagate(adj_norm: Csr, x: np.ndarray, prop_steps: int, flavour: str = "fma") -> List[np.ndarray]:
    """GraphOp.propagate (base_op.py:19-36): [X, A^X, ..., A^^K X] as float32 arrays.
    'fma'/'muladd' round to float32 after every hop (Linux branch, :31-32); 'f64' keeps float64 across hops and
    casts once at the end (:34,:36)."""
    feats = [np.asarray(x)]
    for _ in range(prop_steps):
        feats.append(spmm_hop(adj_norm, feats[-1], flavour))
    return [np.asarray(f, dtype=np.float32) for f in feats]


# ----------------------------------------------------------------------------------------------------------
# a9/a10/a13: non-learnable combiners (float32, numpy; addition order explicit)
# ----------------------------------------------------------------------------------------------------------
def combine_last(feats):   # message_op/last_message_op.py:9-10
    return feats[-1]


def combine_sum(feats, start, end):   # sum_message_op.py:9-10 : python sum() = 0 + f0 + f1 + ... left to right
    sel = feats[start:end]
    acc = sel[0].astype(np.float32).copy()  # 0 + f0 == f0 exactly (except -0.0 -> +0.0, handled below)
    acc = acc + np.float32(0.0)
    for f in sel[1:]:
        acc = acc + f
    return acc


def combine_mean(feats, start, end):  # mean_message_op.py:9-10
    return combine_sum(feats, start, end) / np.float32(end - start)


def combine_max(feats, start, end):   # max_message_op.py:11-12
    return np.stack(feats[start:end], axis=0).max(axis=0)


def combine_min(feats, start, end):   # min_message_op.py:11-12
    return np.stack(feats[start:end], axis=0).min(axis=0)


def combine_concat(feats, start, end):  # concat_message_op.py:11-12
    return np.hstack(feats[start:end])


def alpha_weights(n_feats: int, alpha: float, start, end) -> np.ndarray:
    """simple_weighted_message_op.py:42-47: w0 = alpha, wk = (1-alpha) * w(k-1) in python float64, then float32."""
    w = [alpha]
    for _ in range(n_feats - 1):
        w.append((1 - alpha) * w[-1])
    return np.asarray(w[start:end], dtype=np.float32)


def combine_weighted(feats, weights: np.ndarray, start, end):
    """one_dim_weighted_add (utils.py:91-102): (stack * w[:, None]).sum(dim=0) in float32.
    torch's sum over the short leading dim of a [K', N*d] tensor adds rows in order, so this is the sequential
    acc = f0*w0; acc += fk*wk  (products rounded separately, no fma) -- verified against the reference goldens."""
    sel = feats[start:end]
    w = np.asarray(weights, dtype=np.float32)
    acc = sel[0] * w[0]
    for f, wk in zip(sel[1:], w[1:]):
        acc = acc + f * wk
    return acc.astype(np.float32)


def osd_weights(feats) -> np.ndarray:
    """over_smooth_distance_op.py:13-22: w = softmax_k( <x, y_k> / (|y_k|+1e-10) / (|x|+1e-10) ), float32.
    Row reductions are evaluated in float64 then rounded, so the comparison tolerance for this op is 1e-6."""
    x = feats[0].astype(np.float64)
    nx = np.sqrt((x * x).sum(1)).astype(np.float32) + np.float32(1e-10)
    cols = []
    for f in feats:
        y = f.astype(np.float64)
        ny = np.sqrt((y * y).sum(1)).astype(np.float32) + np.float32(1e-10)
        dot = (x * y).sum(1).astype(np.float32)
        cols.append((dot / ny) / nx)
    c = np.stack(cols, axis=1).astype(np.float32)
    c = c - c.max(axis=1, keepdims=True)
    e = np.exp(c.astype(np.float64))
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def combine_osd(feats):
    """over_smooth_distance_op.py:24-33: out_i = sum_k w_ik * y_k,i  (hop order)."""
    w = osd_weights(feats)
    out = np.zeros_like(feats[0], dtype=np.float32)
    for k, f in enumerate(feats):
        out = out + w[:, k:k + 1] * f
    return out


# ----------------------------------------------------------------------------------------------------------
# a11: learnable weighted op, forward only, explicit parameters (float32 semantics, float64 internals noted)
# ----------------------------------------------------------------------------------------------------------
def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def _softmax_rows(z):
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def learnable_weights(feats, start, end, kind: str, weight=None, bias=None) -> np.ndarray:
    """learnable_weighted_messahe_op.py:59-86.  Returns the [B, end-start] weight matrix (or [end-start] vector
    for simple/simple_allow_neg) in float64 for tolerance comparison.  The 'ori_ref' and 'jk' reshapes follow the
    reference AS WRITTEN: a hop-major score vector of length K'*B is viewed as [B, K'] (:78,:86), i.e. weight
    (i, j) = flat[i*K' + j]."""
    sel = feats[start:end]
    kp = end - start
    if kind == "simple":
        s = _sigmoid(np.asarray(weight, dtype=np.float64)[start:end])
        e = np.exp(s - s.max())
        return e / e.sum()
    if kind == "simple_allow_neg":
        return np.asarray(weight, dtype=np.float64)[start:end]
    w = np.asarray(weight, dtype=np.float64).reshape(-1)
    b = float(np.asarray(bias).reshape(-1)[0])
    stacked = np.vstack(sel).astype(np.float64)  # [K'*B, d], hop-major
    if kind == "gate":
        score = stacked @ w + b                      # :68-71
        return _softmax_rows(_sigmoid(score.reshape(kp, -1).T))
    if kind == "ori_ref":
        ref = np.tile(feats[0].astype(np.float64), (kp, 1))           # :74
        score = np.hstack([ref, stacked]) @ w + b
        return _softmax_rows(_sigmoid(score.reshape(-1, kp)))          # :78 (as written)
    if kind == "jk":
        ref = np.tile(np.hstack(feats).astype(np.float64), (kp, 1))    # :81 all hops, regardless of start/end
        score = np.hstack([ref, stacked]) @ w + b
        return _softmax_rows(_sigmoid(score.reshape(-1, kp)))          # :86 (as written)
    raise NotImplementedError(kind)


def combine_learnable(feats, start, end, kind: str, weight=None, bias=None) -> np.ndarray:
    """learnable_weighted_messahe_op.py:91-101 (+ utils.py:91-116)."""
    sel = feats[start:end]
    w = learnable_weights(feats, start, end, kind, weight, bias)
    if w.ndim == 1:
        out = sum(f.astype(np.float64) * wk for f, wk in zip(sel, w))
    else:
        out = sum(f.astype(np.float64) * w[:, k:k + 1] for k, f in enumerate(sel))
    return out.astype(np.float32)


# ----------------------------------------------------------------------------------------------------------
# reference-library timing helper (bench.py --impl reference / cpu_baseline only)
# ----------------------------------------------------------------------------------------------------------
def load_reference_kernel():
    """ctypes handle on oracle/_ref/libmatmul_ref.so::FloatCSRMulDenseOMP, the reference's own kernel compiled from
    its own source by oracle/Makefile; None when that build is absent."""
    path = os.path.join(_HERE, "_ref", "libmatmul_ref.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    f32p = np.ctypeslib.ndpointer(dtype=np.float32, ndim=1, flags="C_CONTIGUOUS")
    i32p = np.ctypeslib.ndpointer(dtype=np.int32, ndim=1, flags="C_CONTIGUOUS")
    lib.FloatCSRMulDenseOMP.argtypes = [f32p, f32p, i32p, i32p, f32p, ctypes.c_int, ctypes.c_int]
    lib.FloatCSRMulDenseOMP.restype = None
    return lib


def reference_kernel_hop(lib, adj: Csr, x: np.ndarray, a32: Optional[np.ndarray] = None,
                         indptr32: Optional[np.ndarray] = None) -> np.ndarray:
    """One hop through the reference's compiled kernel with the reference wrapper's argument preparation
    (utils.py:31-38) minus its redundant copies.  Valid only while N*d < 2^31 and nnz < 2^31 (matmul.c:33)."""
    n, d = x.shape
    if n * d >= 2 ** 31 or adj.nnz >= 2 ** 31:
        raise OverflowError("reference kernel uses 32-bit offsets (matmul.c:33)")
    y = np.zeros(n * d, dtype=np.float32)
    a32 = adj.data.astype(np.float32) if a32 is None else a32
    indptr32 = adj.indptr.astype(np.int32) if indptr32 is None else indptr32
    lib.FloatCSRMulDenseOMP(y, a32, adj.indices, indptr32, np.ascontiguousarray(x, dtype=np.float32).reshape(-1),
                            n, d)
    return y.reshape(n, d)


# ----------------------------------------------------------------------------------------------------------
# 8f-3: adjacent consumers (restated for tests/test_gpu_parity.py; tolerance parity -- the reference uses torch.spmm
# on CPU COO tensors, whose accumulation order is unspecified)
# ----------------------------------------------------------------------------------------------------------
def label_propagation(labels: np.ndarray, adj_norm: Csr, num_layers: int, alpha: float, mask=None,
                      clamp=(0.0, 1.0)) -> np.ndarray:
    """sgl/tricks/utils.py:40-58 with the default post_process (clamp to [0, 1]); labels already one-hot float32."""
    lab = np.asarray(labels, dtype=np.float32)
    out = lab.copy()
    if mask is not None:
        out = np.zeros_like(lab)
        out[mask] = lab[mask]
    res = np.float32(1 - alpha) * out
    a32 = Csr(adj_norm.indptr, adj_norm.indices, adj_norm.data.astype(np.float32), adj_norm.shape)
    for _ in range(num_layers):
        out = np.float32(alpha) * spmm_hop(a32, out, "muladd") + res
        if clamp is not None:
            out = np.clip(out, clamp[0], clamp[1])
    return out.astype(np.float32)


def nafs_smoothed_features(adj, x: np.ndarray, hops: int, r_list=(0.5, 0.4, 0.3, 0.2, 0.1, 0.0), method="mean"):
    """Feature construction of NodeClusteringNAFS._k_hop_cluster (sgl/tasks/node_clustering.py:205-251): per r the
    hops of adj_to_symmetric_norm(adj, r) (tasks/utils.py:412-424, same formula as operators/utils.py:76-88), the
    over-smoothing-distance weights (:226-243, identical to over_smooth_distance_op.py:13-31) and the combination
    over r (:246-253)."""
    per_r = []
    for r in r_list:
        a = symmetric_norm_csr(adj, r)
        a32 = Csr(a.indptr, a.indices, a.data.astype(np.float32), a.shape)
        feats = [np.asarray(x, dtype=np.float32)]
        for _ in range(hops):
            feats.append(spmm_hop(a32, feats[-1], "muladd"))
        if method == "simple":
            per_r.append(feats[-1])
            break
        per_r.append(combine_osd(feats))
    if method == "mean":
        return (sum(per_r) / np.float32(len(per_r))).astype(np.float32)
    if method == "max":
        return np.stack(per_r, axis=0).max(axis=0)
    if method == "concat":
        return np.concatenate(per_r, axis=1)
    return per_r[-1]

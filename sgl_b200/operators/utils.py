"""Mirror of the reference's sgl/operators/utils.py on top of libsglb200.

  csr_sparse_dense_matmul        reference utils.py:10-40   (CPU, ctypes -> libmatmul.so)
  cuda_csr_sparse_dense_matmul   reference utils.py:43-73   (cuSPARSE wrapper, never called there)
  adj_to_symmetric_norm          reference utils.py:76-88
  one_dim_weighted_add           reference utils.py:91-102
  two_dim_weighted_add           reference utils.py:105-116
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
from torch import Tensor

from .. import _lib
from ..runtime import CsrOperator, aggregate, require_cuda


def csr_sparse_dense_matmul(adj, feature, mode: str = "exact"):
    """One hop ``adj @ feature`` on the GPU; numpy in, fresh float32 numpy out (reference utils.py:10-40).

    The default EXACT mode reproduces the reference library's float32 result bit for bit.  Callers that run several
    hops should build one :class:`~sgl_b200.runtime.CsrOperator` instead of paying the CSR upload per call."""
    require_cuda()
    feat = np.ascontiguousarray(feature, dtype=np.float32)
    if feat.ndim != 2 or adj.shape[1] != feat.shape[0]:
        raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
    op = CsrOperator.from_scipy(adj)
    try:
        x = torch.from_numpy(feat).cuda(non_blocking=False)
        y = op.spmm(x, mode=mode)
        return y.cpu().numpy()
    finally:
        op.close()


def cuda_csr_sparse_dense_matmul(adj, feature):
    """Same hop; kept under the name of the reference's dormant cuSPARSE wrapper (utils.py:43-73)."""
    return csr_sparse_dense_matmul(adj, feature, mode="fast")


def normalisation_parts(adj, r: float):
    """Pieces of  A^ = diag(deg^(r-1)) (A+I)^T diag(deg^-r)  (reference utils.py:76-88):
    the CSR structure of (A+I)^T with its raw float64 weights, and the two float64 scaling vectors."""
    n = adj.shape[0]
    with_loops = (adj + sp.identity(n, format="csr")).tocsr()      # float64; an existing diagonal w becomes w + 1
    deg = np.asarray(with_loops.sum(axis=1)).reshape(-1)           # weighted degrees (row sums), float64
    with np.errstate(divide="ignore", invalid="ignore"):
        d_left = np.power(deg, r - 1)
        d_right = np.power(deg, -r)
    d_left[np.isinf(d_left)] = 0.0
    d_right[np.isinf(d_right)] = 0.0
    transposed = with_loops.T.tocsr()
    transposed.sort_indices()
    return transposed, d_left, d_right


def adj_to_symmetric_norm(adj, r):
    """Normalised adjacency with self loops as a scipy sparse matrix in float64 (reference utils.py:76-88).
    Entry (i, j) = fl64(fl64((A+I)[j, i] * deg_i^(r-1)) * deg_j^(-r)), the reference's product order."""
    transposed, d_left, d_right = normalisation_parts(adj, r)
    rows = np.repeat(np.arange(transposed.shape[0]), np.diff(transposed.indptr))
    data = (transposed.data * d_left[rows]) * d_right[transposed.indices]
    return sp.csr_matrix((data, transposed.indices, transposed.indptr), shape=transposed.shape)


def _to_cuda(feat_list):
    dev = next((f.device for f in feat_list if f.is_cuda), torch.device("cuda", torch.cuda.current_device()))
    return [f.detach().to(device=dev, dtype=torch.float32) for f in feat_list], dev


def one_dim_weighted_add(feat_list, weight_list):
    """sum_k weight[k] * feat[k] with one scalar weight per hop (reference utils.py:91-102)."""
    if not isinstance(feat_list, list) or not isinstance(weight_list, Tensor):
        raise TypeError("This function is designed for list(feature) and tensor(weight)!")
    elif len(feat_list) != weight_list.shape[0]:
        raise ValueError("The feature list and the weight list have different lengths!")
    elif len(weight_list.shape) != 1:
        raise ValueError("The weight list should be a 1d tensor!")
    if weight_list.requires_grad or any(f.requires_grad for f in feat_list):
        # autograd path (learnable scalar weights): plain tensor algebra on whatever device the batch lives on
        return sum(f * w for f, w in zip(feat_list, weight_list))
    require_cuda()
    on_cpu = not feat_list[0].is_cuda
    feats, _ = _to_cuda(feat_list)
    out = aggregate(_lib.AGG_WEIGHTED, feats, weight_list.detach().cpu().tolist())
    return out.cpu() if on_cpu else out


def two_dim_weighted_add(feat_list, weight_list):
    """out[i] = sum_k weight[i, k] * feat[k][i] with per-node hop weights (reference utils.py:105-116)."""
    if not isinstance(feat_list, list) or not isinstance(weight_list, Tensor):
        raise TypeError("This function is designed for list(feature) and tensor(weight)!")
    elif len(feat_list) != weight_list.shape[1]:
        raise ValueError("The feature list and the weight list have different lengths!")
    elif len(weight_list.shape) != 2:
        raise ValueError("The weight list should be a 2d tensor!")
    out = feat_list[0] * weight_list[:, 0:1]
    for k in range(1, len(feat_list)):
        out = out + feat_list[k] * weight_list[:, k:k + 1]
    return out

"""Writes a minimal stand-in for the reference's `sgl.operators` package into a directory (test scaffolding).

The GPU box has no /root/reference, so the two drop-in routes of INTEGRATION.md (swapping csrc/libmatmul.so; calling
sgl_b200.patch.install()) are exercised there against this stand-in.  It restates only the INTERFACE the routes bind
to -- module paths, names, argument types and the ctypes call with its numpy.ctypeslib argtypes
(reference sgl/operators/utils.py:10-40, sgl/operators/base_op.py:11-60, graph_op/laplacian_graph_op.py:7-19,
message_op/{sum,mean,max,min,concat}_message_op.py, over_smooth_distance_op.py) -- in a few lines each.
"""
import os
import textwrap

FILES = {
    "sgl/__init__.py": "",
    "sgl/operators/__init__.py": "",
    "sgl/operators/utils.py": '''
        import os.path as osp
        from ctypes import c_int
        import numpy as np
        import numpy.ctypeslib as ctl

        def csr_sparse_dense_matmul(adj, feature):
            # same binding as the reference wrapper: library next to this file, 1-D contiguous ndpointers, void return
            lib = ctl.load_library("./csrc/libmatmul.so", osp.split(osp.abspath(__file__))[0])
            ints = ctl.ndpointer(dtype=np.int32, ndim=1, flags="CONTIGUOUS")
            floats = ctl.ndpointer(dtype=np.float32, ndim=1, flags="CONTIGUOUS")
            lib.FloatCSRMulDenseOMP.argtypes = [floats, floats, ints, ints, floats, c_int, c_int]
            lib.FloatCSRMulDenseOMP.restypes = None
            answer = np.zeros(feature.shape).astype(np.float32).flatten()
            rows, cols = feature.shape
            lib.FloatCSRMulDenseOMP(answer, adj.data.astype(np.float32), adj.indices, adj.indptr, feature.flatten(), rows, cols)
            return answer.reshape(feature.shape)

        def cuda_csr_sparse_dense_matmul(adj, feature):
            raise RuntimeError("stand-in: the dormant cuSPARSE wrapper is never called")

        def adj_to_symmetric_norm(adj, r):
            import scipy.sparse as sp
            adj = adj + sp.eye(adj.shape[0])
            deg = np.array(adj.sum(1)).flatten()
            left, right = np.power(deg, r - 1), np.power(deg, -r)
            left[np.isinf(left)] = 0.0
            right[np.isinf(right)] = 0.0
            return adj.dot(sp.diags(left)).transpose().dot(sp.diags(right))
    ''',
    "sgl/operators/base_op.py": '''
        import numpy as np
        import scipy.sparse as sp
        import torch
        import torch.nn as nn
        from torch import Tensor
        from sgl.operators.utils import csr_sparse_dense_matmul, cuda_csr_sparse_dense_matmul

        class GraphOp:
            def __init__(self, prop_steps):
                self._prop_steps = prop_steps
                self._adj = None

            def _construct_adj(self, adj):
                raise NotImplementedError

            def propagate(self, adj, feature):
                self._adj = self._construct_adj(adj)
                if not isinstance(adj, sp.csr_matrix):
                    raise TypeError("The adjacency matrix must be a scipy csr sparse matrix!")
                elif not isinstance(feature, np.ndarray):
                    raise TypeError("The feature matrix must be a numpy.ndarray!")
                elif self._adj.shape[1] != feature.shape[0]:
                    raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
                feats = [feature]
                for _ in range(self._prop_steps):
                    feats.append(csr_sparse_dense_matmul(self._adj, feats[-1]))
                return [torch.FloatTensor(f) for f in feats]

        class MessageOp(nn.Module):
            def __init__(self, start=None, end=None):
                super().__init__()
                self._aggr_type = None
                self._start, self._end = start, end

            @property
            def aggr_type(self):
                return self._aggr_type

            def _combine(self, feat_list):
                return NotImplementedError

            def aggregate(self, feat_list):
                for feat in feat_list:
                    if not isinstance(feat, Tensor):
                        raise TypeError("The feature matrices must be tensors!")
                return self._combine(feat_list)
    ''',
    "sgl/operators/graph_op/__init__.py": '''
        import scipy.sparse as sp
        from sgl.operators.base_op import GraphOp
        from sgl.operators.utils import adj_to_symmetric_norm

        class LaplacianGraphOp(GraphOp):
            def __init__(self, prop_steps, r=0.5):
                super().__init__(prop_steps)
                self._r = r

            def _construct_adj(self, adj):
                if isinstance(adj, sp.csr_matrix):
                    adj = adj.tocoo()
                elif not isinstance(adj, sp.coo_matrix):
                    raise TypeError("The adjacency matrix must be a scipy.sparse.coo_matrix/csr_matrix!")
                return adj_to_symmetric_norm(adj, self._r).tocsr()
    ''',
    "sgl/operators/message_op/__init__.py": '''
        import torch
        from sgl.operators.base_op import MessageOp

        class _Sliced(MessageOp):
            def __init__(self, start, end):
                super().__init__(start, end)

        class SumMessageOp(_Sliced):
            def _combine(self, feat_list):
                return sum(feat_list[self._start:self._end])

        class MeanMessageOp(_Sliced):
            def _combine(self, feat_list):
                return sum(feat_list[self._start:self._end]) / (self._end - self._start)

        class MaxMessageOp(_Sliced):
            def _combine(self, feat_list):
                return torch.stack(feat_list[self._start:self._end], dim=0).max(dim=0)[0]

        class MinMessageOp(_Sliced):
            def _combine(self, feat_list):
                return torch.stack(feat_list[self._start:self._end], dim=0).min(dim=0)[0]

        class ConcatMessageOp(_Sliced):
            def _combine(self, feat_list):
                return torch.hstack(feat_list[self._start:self._end])

        class OverSmoothDistanceWeightedOp(MessageOp):
            def __init__(self):
                super().__init__()

            def _combine(self, feat_list):
                x = feat_list[0]
                cos = [(x * f).sum(1) / (f.norm(dim=1) + 1e-10) / (x.norm(dim=1) + 1e-10) for f in feat_list]
                w = torch.softmax(torch.stack(cos, dim=1), dim=1)
                return sum(f * w[:, k:k + 1] for k, f in enumerate(feat_list))
    ''',
}


def write(root):
    for rel, body in FILES.items():
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(textwrap.dedent(body).lstrip("\n"))
    os.makedirs(os.path.join(root, "sgl", "operators", "csrc"), exist_ok=True)
    return root

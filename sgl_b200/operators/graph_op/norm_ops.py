"""The two graph operators of the reference, built on one degree-normalisation recipe.

  LaplacianGraphOp   reference sgl/operators/graph_op/laplacian_graph_op.py:7-19   A^ = D^(r-1) (A+I)^T D^(-r)
  PprGraphOp         reference sgl/operators/graph_op/ppr_graph_op.py:7-21         (1 - alpha) A^ + alpha I

`_construct_adj` returns the float64 scipy CSR exactly as the reference would (host pass, sgl_b200.operators.utils);
`_norm_spec` publishes (r, alpha) so that GraphOp can build the same operator on the GPU instead (build_on="device").
"""
import numpy as np
import scipy.sparse as sp

from ..base_op import GraphOp
from ..utils import adj_to_symmetric_norm


class _DegreeNormalisedOp(GraphOp):
    def __init__(self, prop_steps, r, alpha=None):
        super().__init__(prop_steps)
        self._r = r
        self._alpha = alpha

    def _norm_spec(self):
        return (self._r, self._alpha)

    def _construct_adj(self, adj):
        if not isinstance(adj, (sp.csr_matrix, sp.coo_matrix)):
            raise TypeError("The adjacency matrix must be a scipy.sparse.coo_matrix/csr_matrix!")
        base = adj_to_symmetric_norm(adj.tocsr(), self._r)
        if self._alpha is None:
            return base
        # teleport term: every row of A^ stores its diagonal (self loops), so only existing entries change
        rows = np.repeat(np.arange(base.shape[0]), np.diff(base.indptr))
        on_diagonal = rows == base.indices
        if int(on_diagonal.sum()) != base.shape[0]:   # a cancelled self loop: fall back to the sparse sum
            return ((1 - self._alpha) * base + self._alpha * sp.identity(base.shape[0], format="csr")).tocsr()
        data = (1 - self._alpha) * base.data
        data[on_diagonal] = data[on_diagonal] + self._alpha
        return sp.csr_matrix((data, base.indices, base.indptr), shape=base.shape)


class LaplacianGraphOp(_DegreeNormalisedOp):
    """r = 0.5 is the symmetric GCN normalisation, r = 0 the random-walk one."""

    def __init__(self, prop_steps, r=0.5):
        super().__init__(prop_steps, r, None)


class PprGraphOp(_DegreeNormalisedOp):
    """Personalised-PageRank style propagation with teleport probability alpha."""

    def __init__(self, prop_steps, r=0.5, alpha=0.15):
        super().__init__(prop_steps, r, alpha)

"""PprGraphOp -- mirror of the reference's sgl/operators/graph_op/ppr_graph_op.py:7-21."""
import numpy as np
import scipy.sparse as sp

from ..base_op import GraphOp
from ..utils import adj_to_symmetric_norm


class PprGraphOp(GraphOp):
    """(1 - alpha) * A^ + alpha * I   with A^ as in LaplacianGraphOp (personalised-PageRank style propagation)."""

    def __init__(self, prop_steps, r=0.5, alpha=0.15):
        super(PprGraphOp, self).__init__(prop_steps)
        self._r = r
        self._alpha = alpha

    def _norm_spec(self):
        return (self._r, self._alpha)

    def _construct_adj(self, adj):
        if not isinstance(adj, (sp.csr_matrix, sp.coo_matrix)):
            raise TypeError("The adjacency matrix must be a scipy.sparse.coo_matrix/csr_matrix!")
        base = adj_to_symmetric_norm(adj.tocsr(), self._r)
        # every row of A^ stores its diagonal (self loops), so the teleport term only touches existing entries
        rows = np.repeat(np.arange(base.shape[0]), np.diff(base.indptr))
        data = (1 - self._alpha) * base.data
        diagonal = rows == base.indices
        if int(diagonal.sum()) != base.shape[0]:
            mixed = (1 - self._alpha) * base + self._alpha * sp.identity(base.shape[0], format="csr")
            return mixed.tocsr()
        data[diagonal] = data[diagonal] + self._alpha
        return sp.csr_matrix((data, base.indices, base.indptr), shape=base.shape)

"""LaplacianGraphOp -- mirror of the reference's sgl/operators/graph_op/laplacian_graph_op.py:7-19."""
import scipy.sparse as sp

from ..base_op import GraphOp
from ..utils import adj_to_symmetric_norm


class LaplacianGraphOp(GraphOp):
    """A^ = D^(r-1) (A+I)^T D^(-r); r = 0.5 is the symmetric GCN normalisation."""

    def __init__(self, prop_steps, r=0.5):
        super(LaplacianGraphOp, self).__init__(prop_steps)
        self._r = r

    def _norm_spec(self):
        return (self._r, None)

    def _construct_adj(self, adj):
        if not isinstance(adj, (sp.csr_matrix, sp.coo_matrix)):
            raise TypeError("The adjacency matrix must be a scipy.sparse.coo_matrix/csr_matrix!")
        return adj_to_symmetric_norm(adj.tocsr(), self._r)

"""Graph operators of the SGAP propagation step (mirror of the reference package sgl.operators.graph_op)."""
from . import laplacian_graph_op as _lap, ppr_graph_op as _ppr

LaplacianGraphOp = _lap.LaplacianGraphOp
PprGraphOp = _ppr.PprGraphOp

__all__ = sorted(name for name in dir() if name.endswith("GraphOp"))

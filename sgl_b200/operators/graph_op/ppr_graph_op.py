"""Module path kept for callers that import `...graph_op.ppr_graph_op` like in the reference."""
from .norm_ops import PprGraphOp  # noqa: F401

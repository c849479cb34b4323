"""patch.py -- route an importable reference ``sgl.operators`` through libsglb200 without editing SGL.

    import sgl_b200.patch as p; p.install()

After install():
  * ``sgl.operators.base_op.GraphOp.propagate`` runs the K hops on the B200 (every LaplacianGraphOp / PprGraphOp
    instance, hence SGC / GAMLP / NAFS / NARS ... unchanged), keeping the subclass' own ``_construct_adj``;
  * ``sgl.operators.utils.csr_sparse_dense_matmul`` (and the name imported into base_op) is the GPU hop;
  * the non-learnable ``MessageOp._combine`` implementations are the fused kernels.
uninstall() restores the originals.  See INTEGRATION.md for the equivalent three-line edit inside SGL.
"""
from __future__ import annotations

import importlib

_saved = {}

_COMBINE_MAP = {
    "SumMessageOp": "SumMessageOp", "MeanMessageOp": "MeanMessageOp", "MaxMessageOp": "MaxMessageOp",
    "MinMessageOp": "MinMessageOp", "ConcatMessageOp": "ConcatMessageOp",
    "OverSmoothDistanceWeightedOp": "OverSmoothDistanceWeightedOp",
}


def install():
    from .operators import base_op as ours_base, utils as ours_utils
    from .operators.message_op import simple_ops as ours_ops

    ref_base = importlib.import_module("sgl.operators.base_op")
    ref_utils = importlib.import_module("sgl.operators.utils")
    ref_msg = importlib.import_module("sgl.operators.message_op")
    if _saved:
        return
    _saved["propagate"] = ref_base.GraphOp.propagate
    _saved["matmul_utils"] = ref_utils.csr_sparse_dense_matmul
    _saved["matmul_base"] = ref_base.csr_sparse_dense_matmul

    def propagate(self, adj, feature):
        # the reference contract, bound to the reference class: self._construct_adj stays the reference's own
        return ours_base.propagate_reference_contract(self, adj, feature)

    ref_base.GraphOp.propagate = propagate
    ref_base.GraphOp.mode = ours_base.GraphOp.mode
    ref_base.GraphOp.output_device = ours_base.GraphOp.output_device
    ref_utils.csr_sparse_dense_matmul = ours_utils.csr_sparse_dense_matmul
    ref_base.csr_sparse_dense_matmul = ours_utils.csr_sparse_dense_matmul
    for ref_name, our_name in _COMBINE_MAP.items():
        ref_cls = getattr(ref_msg, ref_name)
        _saved["combine_" + ref_name] = ref_cls._combine
        ref_cls._combine = getattr(ours_ops, our_name)._combine


def uninstall():
    if not _saved:
        return
    ref_base = importlib.import_module("sgl.operators.base_op")
    ref_utils = importlib.import_module("sgl.operators.utils")
    ref_msg = importlib.import_module("sgl.operators.message_op")
    ref_base.GraphOp.propagate = _saved.pop("propagate")
    ref_utils.csr_sparse_dense_matmul = _saved.pop("matmul_utils")
    ref_base.csr_sparse_dense_matmul = _saved.pop("matmul_base")
    for ref_name in _COMBINE_MAP:
        getattr(ref_msg, ref_name)._combine = _saved.pop("combine_" + ref_name)

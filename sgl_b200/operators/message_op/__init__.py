"""Cross-hop message operators (mirror of the reference package sgl.operators.message_op: the same eleven names)."""
from . import learnable_ops as _learnable, simple_ops as _simple

_EXPORTS = {
    _simple: ("LastMessageOp", "SumMessageOp", "MeanMessageOp", "MaxMessageOp", "MinMessageOp", "ConcatMessageOp",
              "SimpleWeightedMessageOp", "OverSmoothDistanceWeightedOp"),
    _learnable: ("LearnableWeightedMessageOp", "IterateLearnableWeightedMessageOp", "ProjectedConcatMessageOp"),
}
for _module, _names in _EXPORTS.items():
    for _name in _names:
        globals()[_name] = getattr(_module, _name)

__all__ = [n for names in _EXPORTS.values() for n in names]

"""Mirror of the reference's sgl/operators/message_op package (same eleven class names)."""
from .simple_ops import (ConcatMessageOp, LastMessageOp, MaxMessageOp, MeanMessageOp, MinMessageOp,
                         OverSmoothDistanceWeightedOp, SimpleWeightedMessageOp, SumMessageOp)
from .learnable_ops import (IterateLearnableWeightedMessageOp, LearnableWeightedMessageOp,
                            ProjectedConcatMessageOp)

__all__ = [
    "ConcatMessageOp",
    "IterateLearnableWeightedMessageOp",
    "LastMessageOp",
    "LearnableWeightedMessageOp",
    "MaxMessageOp",
    "MeanMessageOp",
    "MinMessageOp",
    "ProjectedConcatMessageOp",
    "SimpleWeightedMessageOp",
    "SumMessageOp",
    "OverSmoothDistanceWeightedOp",
]

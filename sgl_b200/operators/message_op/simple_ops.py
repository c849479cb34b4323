"""Non-learnable cross-hop combiners on the B200 (reference sgl/operators/message_op/{last,sum,mean,max,min,concat,
simple_weighted}_message_op.py and over_smooth_distance_op.py).

Inputs are the K+1 hop tensors.  CUDA tensors are combined in place on their device; CPU tensors (what the
reference's preprocess hands over) are uploaded, combined by the kernel and the result is returned on the CPU,
so callers see the reference's types.  There is no CPU implementation behind these classes.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .. import utils as _u
from ... import _lib
from ...runtime import aggregate, require_cuda
from ..base_op import MessageOp


def _needs_autograd(feats) -> bool:
    """The kernels are forward-only: inputs that carry gradients (a learnable aggregator upstream, reference
    models/base_model.py:119-222) are combined with the reference's own differentiable torch expressions instead."""
    return torch.is_grad_enabled() and any(isinstance(f, Tensor) and f.requires_grad for f in feats)


def _torch_combine(op: int, feats, weights=None) -> Tensor:
    if op == _lib.AGG_SUM:
        return sum(feats)
    if op == _lib.AGG_MEAN:
        return sum(feats) / len(feats)
    if op == _lib.AGG_MAX:
        return torch.stack(feats, dim=0).max(dim=0)[0]
    if op == _lib.AGG_MIN:
        return torch.stack(feats, dim=0).min(dim=0)[0]
    if op == _lib.AGG_CONCAT:
        return torch.hstack(feats)
    if op == _lib.AGG_WEIGHTED:
        return sum(f * float(w) for f, w in zip(feats, weights))
    x = feats[0]
    cos = [(x * f).sum(1) / (f.norm(dim=1) + 1e-10) / (x.norm(dim=1) + 1e-10) for f in feats]
    w = torch.softmax(torch.stack(cos, dim=1), dim=1)
    return sum(f * w[:, k:k + 1] for k, f in enumerate(feats))


def _run(op: int, feats, weights=None) -> Tensor:
    if _needs_autograd(feats):
        return _torch_combine(op, list(feats), weights)
    require_cuda()
    on_cpu = not feats[0].is_cuda
    dev_feats, _ = _u._to_cuda(list(feats))
    out = aggregate(op, dev_feats, weights)
    return out.cpu() if on_cpu else out


class LastMessageOp(MessageOp):
    """Keeps the last hop only (reference last_message_op.py:4-10)."""

    def __init__(self):
        super(LastMessageOp, self).__init__()
        self._aggr_type = "last"

    def _combine(self, feat_list):
        return feat_list[-1]

    def fused_spec(self, prop_steps):
        """How sglb200_propagate_fused folds this combiner into the hop kernels (None: not fusable)."""
        return {"agg": "last"}


class SumMessageOp(MessageOp):
    """Left-to-right float32 sum of hops [start, end) (reference sum_message_op.py:4-10)."""

    def __init__(self, start, end):
        super(SumMessageOp, self).__init__(start, end)
        self._aggr_type = "sum"

    def _combine(self, feat_list):
        return _run(_lib.AGG_SUM, feat_list[self._start:self._end])

    def fused_spec(self, prop_steps):
        if not (0 <= self._start < self._end <= prop_steps + 1):
            return None
        return {"agg": "sum", "start": self._start, "end": self._end}


class MeanMessageOp(MessageOp):
    """Sum of hops [start, end) divided once by end-start (reference mean_message_op.py:4-10)."""

    def __init__(self, start, end):
        super(MeanMessageOp, self).__init__(start, end)
        self._aggr_type = "mean"

    def _combine(self, feat_list):
        sel = feat_list[self._start:self._end]
        if len(sel) != self._end - self._start:
            # the reference divides by end-start even when the slice is shorter; keep that by scaling afterwards
            return _run(_lib.AGG_SUM, sel) / (self._end - self._start)
        return _run(_lib.AGG_MEAN, sel)

    def fused_spec(self, prop_steps):
        if not (0 <= self._start < self._end <= prop_steps + 1):
            return None
        return {"agg": "mean", "start": self._start, "end": self._end}


class MaxMessageOp(MessageOp):
    """Element-wise maximum over hops [start, end) (reference max_message_op.py:6-12)."""

    def __init__(self, start, end):
        super(MaxMessageOp, self).__init__(start, end)
        self._aggr_type = "max"

    def _combine(self, feat_list):
        return _run(_lib.AGG_MAX, feat_list[self._start:self._end])

    def fused_spec(self, prop_steps):
        if not (0 <= self._start < self._end <= prop_steps + 1):
            return None
        return {"agg": "max", "start": self._start, "end": self._end}


class MinMessageOp(MessageOp):
    """Element-wise minimum over hops [start, end) (reference min_message_op.py:6-12)."""

    def __init__(self, start, end):
        super(MinMessageOp, self).__init__(start, end)
        self._aggr_type = "min"

    def _combine(self, feat_list):
        return _run(_lib.AGG_MIN, feat_list[self._start:self._end])

    def fused_spec(self, prop_steps):
        if not (0 <= self._start < self._end <= prop_steps + 1):
            return None
        return {"agg": "min", "start": self._start, "end": self._end}


class ConcatMessageOp(MessageOp):
    """[N, (end-start)*d] horizontal stack of hops [start, end) (reference concat_message_op.py:6-12)."""

    def __init__(self, start, end):
        super(ConcatMessageOp, self).__init__(start, end)
        self._aggr_type = "concat"

    def _combine(self, feat_list):
        return _run(_lib.AGG_CONCAT, feat_list[self._start:self._end])

    def fused_spec(self, prop_steps):
        if not (0 <= self._start < self._end <= prop_steps + 1):
            return None
        return {"agg": "concat", "start": self._start, "end": self._end}


class SimpleWeightedMessageOp(MessageOp):
    """Fixed scalar hop weights (reference simple_weighted_message_op.py:8-56).

    'alpha'        one extra argument alpha (float in [0, 1]): w_0 = alpha, w_k = (1 - alpha) * w_{k-1}
    'hand_crafted' one extra argument: the weight list (list or tensor)
    """

    def __init__(self, start, end, combination_type, *args):
        super(SimpleWeightedMessageOp, self).__init__(start, end)
        self._aggr_type = "simple_weighted"

        if combination_type not in ["alpha", "hand_crafted"]:
            raise ValueError("Invalid weighted combination type! Type must be 'alpha' or 'hand_crafted'.")
        self._combination_type = combination_type

        if len(args) != 1:
            raise ValueError("Invalid parameter numbers for the simple weighted aggregator!")
        self._alpha, self._weight_list = None, None
        if combination_type == "alpha":
            self._alpha = args[0]
            if not isinstance(self._alpha, float):
                raise TypeError("The alpha must be a float!")
            elif self._alpha > 1 or self._alpha < 0:
                raise ValueError("The alpha must be a float in [0,1]!")
        else:
            self._weight_list = args[0]
            if isinstance(self._weight_list, list):
                self._weight_list = torch.FloatTensor(self._weight_list)
            elif not isinstance(self._weight_list, (list, Tensor)):
                raise TypeError("The input weight list must be a list or a tensor!")

    def _combine(self, feat_list):
        if self._combination_type == "alpha":
            # geometric weights in python float64, rounded to float32 once (reference :42-47)
            weights = [self._alpha]
            for _ in range(len(feat_list) - 1):
                weights.append((1 - self._alpha) * weights[-1])
            self._weight_list = torch.FloatTensor(weights[self._start:self._end])
        return _u.one_dim_weighted_add(feat_list[self._start:self._end], weight_list=self._weight_list)

    def fused_spec(self, prop_steps):
        if not (0 <= self._start < self._end <= prop_steps + 1):
            return None
        if self._combination_type == "alpha":
            weights = [self._alpha]
            for _ in range(prop_steps):
                weights.append((1 - self._alpha) * weights[-1])
            per_hop = [float(v) for v in torch.FloatTensor(weights)]
        else:
            wl = [float(v) for v in self._weight_list]
            if len(wl) != self._end - self._start:
                return None
            per_hop = [0.0] * self._start + wl + [0.0] * (prop_steps + 1 - self._end)
        return {"agg": "weighted", "start": self._start, "end": self._end, "weights": per_hop}


class OverSmoothDistanceWeightedOp(MessageOp):
    """NAFS hop weights: softmax over hops of the cosine between a node's raw and smoothed features
    (reference over_smooth_distance_op.py:6-33, whose python loop over N x (K+1) becomes one warp per node)."""

    def __init__(self):
        super(OverSmoothDistanceWeightedOp, self).__init__()
        self._aggr_type = 'over_smooth_dis_weighted'

    def _combine(self, feat_list):
        return _run(_lib.AGG_OSD, feat_list)

    def fused_spec(self, prop_steps):
        return {"agg": "osd"}

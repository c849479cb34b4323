"""Learnable cross-hop combiners (reference sgl/operators/message_op/learnable_weighted_messahe_op.py,
iterate_learnable_weighted_message_op.py, projected_concat_message_op.py).

These run inside ``forward`` on mini-batches that already live on the training device, with autograd.  The weight
computation follows the reference expression by expression, including the as-written ``view(-1, end-start)`` of the
hop-major score vector for 'ori_ref' and 'jk' (SURVEY.md section 9 item 10).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import Linear, ModuleList, Parameter

from ctypes import c_void_p

from ... import _lib
from ..base_op import MessageOp
from ..utils import one_dim_weighted_add, two_dim_weighted_add


class _FusedHopWeights(torch.autograd.Function):
    """out = sum_j W[:, j] * feats[start+j] with W = softmax(sigmoid(Linear(...))) for gate / ori_ref / jk, forward and
    backward in libsglb200 (sglb200_lw_forward / sglb200_lw_backward): the reference row is dotted once per node
    instead of being repeated K' times, nothing of size [(K'*B), (K+2)*d] is materialised."""

    @staticmethod
    def forward(ctx, kind, start, end, weight, bias, *feats):
        lib = _lib.load()
        feats = [f.detach().contiguous() for f in feats]
        B, d = int(feats[0].shape[0]), int(feats[0].shape[1])
        kp = end - start
        dev = feats[0].device
        w = weight.detach().reshape(-1).contiguous()
        b = bias.detach().reshape(-1).contiguous()
        scores = torch.empty(kp * B, dtype=torch.float32, device=dev)
        hop_w = torch.empty((B, kp), dtype=torch.float32, device=dev)
        out = torch.empty((B, d), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.sglb200_lw_forward(kind, _lib.ptr_array([f.data_ptr() for f in feats]), len(feats), start,
                                              end, B, d, c_void_p(w.data_ptr()), c_void_p(b.data_ptr()),
                                              c_void_p(scores.data_ptr()), c_void_p(hop_w.data_ptr()),
                                              c_void_p(out.data_ptr()), stream), "lw_forward")
        ctx.meta = (kind, start, end, weight.shape, bias.shape)
        ctx.save_for_backward(w, b, scores, hop_w, *feats)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        kind, start, end, w_shape, b_shape = ctx.meta
        w, b, scores, hop_w, *feats = ctx.saved_tensors
        B, d = int(feats[0].shape[0]), int(feats[0].shape[1])
        dev = feats[0].device
        grad_out = grad_out.contiguous().float()
        grads = [torch.zeros_like(f) for f in feats]
        gw = torch.zeros_like(w)
        gb = torch.zeros_like(b)
        scratch = torch.empty((end - start) * B, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.sglb200_lw_backward(kind, _lib.ptr_array([f.data_ptr() for f in feats]), len(feats), start,
                                               end, B, d, c_void_p(w.data_ptr()), c_void_p(b.data_ptr()),
                                               c_void_p(scores.data_ptr()), c_void_p(hop_w.data_ptr()),
                                               c_void_p(grad_out.data_ptr()),
                                               _lib.ptr_array([g.data_ptr() for g in grads]), c_void_p(gw.data_ptr()),
                                               c_void_p(gb.data_ptr()), c_void_p(scratch.data_ptr()), stream),
                       "lw_backward")
        return (None, None, None, gw.view(w_shape), gb.view(b_shape), *grads)


class LearnableWeightedMessageOp(MessageOp):
    """Per-node (or global) learnable hop weights.

    'simple' / 'simple_allow_neg'  extra argument prop_steps        -> prop_steps+1 scalars
    'gate'                         extra argument feat_dim          -> Linear(feat_dim, 1) per hop row
    'ori_ref'                      extra argument feat_dim          -> Linear(2*feat_dim, 1) on [hop0 | hop_k]
    'jk'                           extra arguments prop_steps, feat_dim -> Linear((prop_steps+2)*feat_dim, 1)
    """

    def __init__(self, start, end, combination_type, *args):
        super(LearnableWeightedMessageOp, self).__init__(start, end)
        self._aggr_type = "learnable_weighted"

        if combination_type not in ["simple", "simple_allow_neg", "gate", "ori_ref", "jk"]:
            raise ValueError(
                "Invalid weighted combination type! Type must be 'simple', 'simple_allow_neg', 'gate', 'ori_ref' or 'jk'.")
        self._combination_type = combination_type

        expected = 2 if combination_type == "jk" else 1
        if len(args) != expected:
            raise ValueError(f"Invalid parameter numbers for the {combination_type} learnable weighted aggregator!")
        if combination_type in ("simple", "simple_allow_neg"):
            seed = torch.FloatTensor(1, args[0] + 1)  # xavier needs a 2-d tensor
            nn.init.xavier_normal_(seed)
            self._learnable_weight = Parameter(seed.view(-1))
        elif combination_type == "gate":
            self._learnable_weight = Linear(args[0], 1)
        elif combination_type == "ori_ref":
            self._learnable_weight = Linear(2 * args[0], 1)
        else:
            prop_steps, feat_dim = args
            self._learnable_weight = Linear(feat_dim + (prop_steps + 1) * feat_dim, 1)

    def hop_weights(self, feat_list):
        kind, s, e = self._combination_type, self._start, self._end
        if kind == "simple":
            return F.softmax(torch.sigmoid(self._learnable_weight[s:e]), dim=0)
        if kind == "simple_allow_neg":
            return self._learnable_weight[s:e]
        stacked = torch.vstack(feat_list[s:e])  # [(e-s)*B, d], hop-major
        if kind == "gate":
            score = self._learnable_weight(stacked).view(e - s, -1).T
        else:
            ref = feat_list[0] if kind == "ori_ref" else torch.hstack(feat_list)
            score = self._learnable_weight(torch.hstack((ref.repeat(e - s, 1), stacked))).view(-1, e - s)
        return F.softmax(torch.sigmoid(score), dim=1)

    fused = True  # per-node kinds on CUDA batches run the fused kernels; False keeps the torch expressions

    def _combine(self, feat_list):
        kind = self._combination_type
        if (self.fused and kind in ("gate", "ori_ref", "jk") and len(feat_list) <= 64 and feat_list[0].dim() == 2
                and all(f.is_cuda and f.dtype == torch.float32 and f.shape == feat_list[0].shape for f in feat_list)):
            lin = self._learnable_weight
            return _FusedHopWeights.apply(_lib.LW_KINDS[kind], self._start, self._end, lin.weight, lin.bias,
                                          *feat_list)
        weights = self.hop_weights(feat_list)
        sel = feat_list[self._start:self._end]
        if weights.dim() == 1:
            return one_dim_weighted_add(sel, weight_list=weights)
        return two_dim_weighted_add(sel, weight_list=weights)


class _FusedIterate(torch.autograd.Function):
    """The recursive gated combination of IterateLearnableWeightedMessageOp, forward and backward in libsglb200
    (sglb200_it_forward / sglb200_it_backward, csrc/iterate.cu): 2 K' dot products per node replace the K' hstacks of
    [B, 2d] and the O(K'^2) [B, d] products of the torch expression."""

    @staticmethod
    def forward(ctx, weight, bias, *feats):
        lib = _lib.load()
        feats = [f.detach().float().contiguous() for f in feats]
        B, d = int(feats[0].shape[0]), int(feats[0].shape[1])
        kp, dev = len(feats), feats[0].device
        w = weight.detach().reshape(-1).float().contiguous()
        b = bias.detach().reshape(-1).float().contiguous()
        dots = torch.empty((B, 2 * kp), dtype=torch.float32, device=dev)
        hop_w = torch.empty((B, kp), dtype=torch.float32, device=dev)
        out = torch.empty((B, d), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.sglb200_it_forward(_lib.ptr_array([f.data_ptr() for f in feats]), kp, B, d, c_void_p(w.data_ptr()),
                                              c_void_p(b.data_ptr()), c_void_p(dots.data_ptr()), c_void_p(hop_w.data_ptr()),
                                              c_void_p(out.data_ptr()), stream), "it_forward")
        ctx.meta = (weight.shape, bias.shape)
        ctx.save_for_backward(w, b, dots, *feats)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        w_shape, b_shape = ctx.meta
        w, b, dots, *feats = ctx.saved_tensors
        B, d = int(feats[0].shape[0]), int(feats[0].shape[1])
        dev = feats[0].device
        grad_out = grad_out.contiguous().float()
        grads = [torch.zeros_like(f) for f in feats]
        gw, gb = torch.zeros_like(w), torch.zeros_like(b)
        with torch.cuda.device(dev):
            stream = c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.sglb200_it_backward(_lib.ptr_array([f.data_ptr() for f in feats]), len(feats), B, d,
                                               c_void_p(w.data_ptr()), c_void_p(b.data_ptr()), c_void_p(dots.data_ptr()),
                                               c_void_p(grad_out.data_ptr()), _lib.ptr_array([g.data_ptr() for g in grads]),
                                               c_void_p(gw.data_ptr()), c_void_p(gb.data_ptr()), stream), "it_backward")
        return (gw.view(w_shape), gb.view(b_shape), *grads)


class _FusedReluConcat(torch.autograd.Function):
    """hstack(y_0, relu(y_1), ..., relu(y_{K'-1})) in one pass (sglb200_relu_concat), gradient masked on the way back."""

    @staticmethod
    def forward(ctx, *ys):
        lib = _lib.load()
        ys = [y.detach().float().contiguous() for y in ys]
        B, h = int(ys[0].shape[0]), int(ys[0].shape[1])
        out = torch.empty((B, len(ys) * h), dtype=torch.float32, device=ys[0].device)
        with torch.cuda.device(ys[0].device):
            _lib.check(lib.sglb200_relu_concat(_lib.ptr_array([y.data_ptr() for y in ys]), len(ys), B, h, c_void_p(out.data_ptr()),
                                               c_void_p(torch.cuda.current_stream().cuda_stream)), "relu_concat")
        ctx.save_for_backward(*ys)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        ys = ctx.saved_tensors
        B, h = int(ys[0].shape[0]), int(ys[0].shape[1])
        grad_out = grad_out.contiguous().float()
        grads = [torch.empty_like(y) for y in ys]
        with torch.cuda.device(ys[0].device):
            _lib.check(lib.sglb200_relu_concat_backward(_lib.ptr_array([y.data_ptr() for y in ys]), len(ys), B, h,
                                                        c_void_p(grad_out.data_ptr()), _lib.ptr_array([g.data_ptr() for g in grads]),
                                                        c_void_p(torch.cuda.current_stream().cuda_stream)), "relu_concat_backward")
        return tuple(grads)


class IterateLearnableWeightedMessageOp(MessageOp):
    """Recursive gated combination (reference iterate_learnable_weighted_message_op.py:8-51)."""

    def __init__(self, start, end, combination_type, *args):
        super(IterateLearnableWeightedMessageOp, self).__init__(start, end)
        self._aggr_type = "iterate_learnable_weighted"
        if combination_type not in ["recursive"]:
            raise ValueError("Invalid weighted combination type! Type must be 'recursive'.")
        self._combination_type = combination_type
        if len(args) != 1:
            raise ValueError("Invalid parameter numbers for the recursive iterate weighted aggregator!")
        self._learnable_weight = Linear(2 * args[0], 1)

    fused = True   # CUDA batches go through the fused kernels (csrc/iterate.cu); False keeps the torch expression

    def _combine(self, feat_list):
        s, e = self._start, self._end
        if s > 0 and e > s:
            # the reference indexes the running weight matrix with ABSOLUTE hop numbers (:44-46: `for j in range(1, i + 1)`
            # with i starting at `start`), so start > 0 reads column 1 of a one-column matrix: same failure here
            raise IndexError("index 1 is out of bounds for dimension 1 with size 1")
        sel = feat_list[s:e]
        if self.fused and sel and all(f.is_cuda for f in sel) and 1 <= len(sel) <= 16 \
                and len({tuple(f.shape) for f in sel}) == 1:
            return _FusedIterate.apply(self._learnable_weight.weight, self._learnable_weight.bias, *sel)
        combined = feat_list[s]
        scores = None
        for i in range(s, e):
            gate = torch.sigmoid(self._learnable_weight(torch.hstack((feat_list[i], combined))))
            scores = gate if scores is None else torch.hstack((scores, gate))
            scores = F.softmax(scores, dim=1)  # the reference re-normalises the running matrix every step (:37)
            combined = feat_list[s] * scores[:, 0:1]
            for j in range(1, i + 1):
                combined = combined + feat_list[s + j] * scores[:, j:j + 1]
        return combined


class _Mlp(nn.Module):
    """Dense head used per hop: Linear -> shared PReLU -> Dropout(0.5) ... -> Linear, xavier(relu gain) weights and
    zero biases (the behaviour of the reference's MultiLayerPerceptron, sgl/models/simple_models.py:101-140, which is
    outside the hot path and therefore only restated as far as this op needs it)."""

    def __init__(self, feat_dim, hidden_dim, num_layers, output_dim, dropout=0.5):
        super().__init__()
        if num_layers < 2:
            raise ValueError("MLP must have at least two layers!")
        dims = [feat_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.fcs = ModuleList(Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.act = nn.PReLU()
        self.dropout = nn.Dropout(dropout)
        gain = nn.init.calculate_gain("relu")
        for fc in self.fcs:
            nn.init.xavier_uniform_(fc.weight, gain=gain)
            nn.init.zeros_(fc.bias)

    def forward(self, x):
        for fc in self.fcs[:-1]:
            x = self.dropout(self.act(fc(x)))
        return self.fcs[-1](x)


class ProjectedConcatMessageOp(MessageOp):
    """Per-hop MLP projection then concatenation (reference projected_concat_message_op.py:9-28)."""

    def __init__(self, start, end, feat_dim, hidden_dim, num_layers):
        super(ProjectedConcatMessageOp, self).__init__(start, end)
        self._aggr_type = "proj_concat"
        self._learnable_weight = ModuleList(
            _Mlp(feat_dim, hidden_dim, num_layers, hidden_dim) for _ in range(end - start))

    fused = True   # CUDA batches: ReLU + concat in one kernel (csrc/iterate.cu); the per-hop MLPs stay dense torch layers

    def _combine(self, feat_list):
        sel = feat_list[self._start:self._end]
        if self.fused and sel and all(f.is_cuda for f in sel) and len(sel) <= 16:
            return _FusedReluConcat.apply(*[mlp(f) for mlp, f in zip(self._learnable_weight, sel)])
        parts = [self._learnable_weight[0](sel[0])]
        parts += [F.relu(mlp(f)) for mlp, f in zip(list(self._learnable_weight)[1:], sel[1:])]
        return torch.hstack(parts)

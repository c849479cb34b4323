"""Mirror of the reference's sgl/operators/base_op.py: GraphOp (K-hop propagation) and MessageOp (aggregation).

GraphOp.propagate   reference base_op.py:19-36  -> CsrOperator (libsglb200: sglb200_graph_create + propagate)
MessageOp.aggregate reference base_op.py:53-60
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn as nn
from torch import Tensor

from ..runtime import CsrOperator, require_cuda


def adjacency_fingerprint(adj):
    """Cheap identity of a scipy CSR adjacency: shape, nnz, buffer addresses and a strided checksum of the three arrays.
    Lets repeated propagate() calls on the same matrix reuse the resident operator (the reference rebuilds A^ on every
    call, base_op.py:20; its label-use task calls preprocess every epoch) while an in-place edit of adj.data, which
    keeps the object identity, still invalidates it."""
    def probe(a):
        a = np.asarray(a)
        step = max(1, a.size // 4096)
        return (a.ctypes.data, a.size, a.dtype.str, float(np.asarray(a[::step], dtype=np.float64).sum()),
                float(a[-1]) if a.size else 0.0)
    return (tuple(adj.shape), int(adj.nnz), probe(adj.data), probe(adj.indices), probe(adj.indptr))


def propagate_reference_contract(op, adj, feature):
    """The body of GraphOp.propagate with the reference's exact contract (base_op.py:19-36): build A^ with the
    operator's own `_construct_adj`, validate, run the K hops on the GPU.  `op` may be one of our GraphOp objects or an
    instance of the reference's own classes (sgl_b200.patch binds this function to sgl.operators.base_op.GraphOp)."""
    fingerprint = adjacency_fingerprint(adj) if isinstance(adj, sp.csr_matrix) else None
    cached = getattr(op, "_sglb200_cached", None)
    reuse = fingerprint is not None and cached is not None and cached[0] == fingerprint \
        and getattr(op, "_operator", None) is not None and getattr(op._operator, "_h", None)
    if not reuse:
        op._adj = op._construct_adj(adj)

    if not isinstance(adj, sp.csr_matrix):
        raise TypeError("The adjacency matrix must be a scipy csr sparse matrix!")
    elif not isinstance(feature, (np.ndarray, Tensor)):
        raise TypeError("The feature matrix must be a numpy.ndarray!")
    elif op._adj.shape[1] != feature.shape[0]:
        raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")

    if isinstance(feature, Tensor):
        first = feature.detach()
        first = first if first.dtype == torch.float32 else first.float()
    else:
        if feature.dtype != np.float32:
            # the reference's ctypes signature rejects anything but float32 (utils.py:21-25)
            raise TypeError("The feature matrix must be a float32 numpy.ndarray!")
        first = torch.from_numpy(feature)  # shares memory, like torch.FloatTensor(ndarray) at base_op.py:36
    require_cuda()

    if not reuse:
        previous = getattr(op, "_operator", None)
        if previous is not None:
            previous.close()
        op._operator = CsrOperator.from_scipy(op._adj)
        op._sglb200_cached = (fingerprint,)
    mode = getattr(op, "mode", "fast")
    steps = op._prop_steps
    if getattr(op, "output_device", "cpu") == "cuda":
        return op._operator.propagate(first.cuda(), steps, mode=mode)
    rest = op._operator.propagate_host(first.cpu(), steps, mode=mode, keep="all")
    return [first.cpu()] + rest


class GraphOp:
    """K-hop propagation  [X, A^X, ..., A^^K X]  of a normalised adjacency A^ built by ``_construct_adj``.

    Same contract as the reference (base_op.py:11-36): ``adj`` is a scipy CSR matrix, ``feature`` a float32
    ``numpy.ndarray`` ([N, d]; a ``torch.Tensor`` is accepted too, which the reference's own callers pass,
    SURVEY.md section 9), the result is a list of K+1 ``torch.FloatTensor`` whose first element shares memory
    with ``feature``; ``self._adj`` holds the normalised CSR afterwards.

    B200 specifics, all optional: ``mode`` ("fast" | "exact", class attribute or env SGLB200_MODE) selects the
    accumulation-order contract of the hop kernel; ``output_device`` ("cpu" | "cuda", env SGLB200_OUTPUT) keeps the
    K+1 slabs resident in HBM and returns CUDA tensors (``feat[idx].to(device)`` in the caller's forward then is an
    on-device gather instead of a CPU gather + PCIe copy).
    """

    mode = os.environ.get("SGLB200_MODE", "fast")
    output_device = os.environ.get("SGLB200_OUTPUT", "cpu")
    # "host": A^ is built by _construct_adj with scipy exactly like the reference (default, any subclass);
    # "device": subclasses that publish _norm_spec() (r, alpha) get A^ built on the GPU (sgl_b200.graph_build) and
    #           self._adj becomes a lazily downloaded scipy view -- removes the single-core scipy pass from preprocess
    build_on = os.environ.get("SGLB200_BUILD", "host")
    # directory of the on-disk cache of propagated features (sgl_b200.cache); None = off
    cache_dir = os.environ.get("SGLB200_CACHE_DIR") or None

    def __init__(self, prop_steps):
        self._prop_steps = prop_steps
        self._adj_value = None
        self._adj_parts = None
        self._operator = None  # CsrOperator of the last propagate / prepare call
        self._prepared_for = None

    @property
    def _adj(self):
        if self._adj_value is None and self._adj_parts is not None:
            from ..graph_build import parts_to_scipy
            self._adj_value = parts_to_scipy(self._adj_parts)
        return self._adj_value

    @_adj.setter
    def _adj(self, value):
        self._adj_value = value
        self._adj_parts = None

    def _construct_adj(self, adj):
        raise NotImplementedError

    def _norm_spec(self):
        return None

    def prepare(self, adj):
        """Optional: build A^ and make it resident in HBM once; later propagate(adj, ...) calls with the SAME adjacency
        object skip normalisation and upload (the reference rebuilds A^ on every call, base_op.py:20 -- its label-use
        task calls preprocess every epoch, tasks/node_classification_with_label_use.py:79).  Returns self."""
        if not isinstance(adj, (sp.csr_matrix, sp.coo_matrix)):
            raise TypeError("The adjacency matrix must be a scipy.sparse.coo_matrix/csr_matrix!")
        require_cuda()
        if self._operator is not None:
            self._operator.close()
        spec = self._norm_spec() if self.build_on == "device" else None
        if spec is not None:
            from ..graph_build import operator_from_scipy_device
            self._operator = operator_from_scipy_device(adj, r=spec[0], alpha=spec[1])
            self._adj_value, self._adj_parts = None, self._operator.parts
        else:
            self._adj = self._construct_adj(adj)
            self._operator = CsrOperator.from_scipy(self._adj)
        self._prepared_for = adj
        self._prepared_print = adjacency_fingerprint(adj) if isinstance(adj, sp.csr_matrix) else None
        return self

    def propagate_device(self, adj, feature, concat=False):
        """Like propagate but returns the K+1 slabs as CUDA tensors regardless of `output_device` (used by the fused
        preprocess of sgl_b200.sgap so that only the aggregated result crosses PCIe)."""
        if getattr(self, "_prepared_for", None) is not adj or self._operator is None:
            self.prepare(adj)
        if feature.shape[0] != self._operator.shape[1]:
            raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
        if isinstance(feature, np.ndarray):
            if feature.dtype != np.float32:
                raise TypeError("The feature matrix must be a float32 numpy.ndarray!")
            feature = torch.from_numpy(feature)
        x = feature.detach().to(dtype=torch.float32).to(self._operator.device, non_blocking=True)
        return self._operator.propagate(x, self._prop_steps, mode=self.mode, concat=concat)

    def propagate_aggregate_device(self, adj, feature, spec, keep="none"):
        """propagate + aggregate in one pass per hop (sglb200_propagate_fused): `spec` is a MessageOp.fused_spec() dict.
        Returns (hops, out) as CUDA tensors; hops[k] is None where hop k was not stored."""
        if getattr(self, "_prepared_for", None) is not adj or self._operator is None:
            self.prepare(adj)
        if feature.shape[0] != self._operator.shape[1]:
            raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
        if isinstance(feature, np.ndarray):
            if feature.dtype != np.float32:
                raise TypeError("The feature matrix must be a float32 numpy.ndarray!")
            feature = torch.from_numpy(feature)
        x = feature.detach().to(dtype=torch.float32).to(self._operator.device, non_blocking=True)
        return self._operator.propagate_fused(x, self._prop_steps, mode=self.mode, keep=keep, agg=spec.get("agg"),
                                              start=spec.get("start", 0), end=spec.get("end"),
                                              weights=spec.get("weights"))

    def propagate(self, adj, feature):
        if self.cache_dir and isinstance(adj, sp.csr_matrix) and isinstance(feature, (np.ndarray, Tensor)) \
                and adj.shape[1] == feature.shape[0]:
            return self._propagate_cached(adj, feature)
        return self._propagate_uncached(adj, feature)

    def _propagate_cached(self, adj, feature):
        from ..cache import HopCache
        spec = self._norm_spec()
        if spec is not None:
            params = {"r": spec[0], "alpha": spec[1]}
        elif hasattr(self, "cache_params"):
            params = dict(self.cache_params())
        else:
            # a user-defined operator without declared hyper-parameters: its A^ cannot be keyed safely
            return self._propagate_uncached(adj, feature)
        cache = HopCache(self.cache_dir)
        key = cache.key(adj, feature, type(self).__name__ + ":" + self.mode, self._prop_steps, **params)
        hops = cache.load(key)
        if hops is not None:
            if self._adj_value is None and self._adj_parts is None:
                self._adj = self._construct_adj(adj)   # the reference contract: _adj is set after propagate (base_op.py:20)
            first = torch.from_numpy(feature) if isinstance(feature, np.ndarray) else feature.detach().float().cpu()
            hops = [first] + hops[1:]
            return [h.cuda() for h in hops] if self.output_device == "cuda" else hops
        hops = self._propagate_uncached(adj, feature)
        cache.save(key, hops)
        return hops

    def _propagate_uncached(self, adj, feature):
        if getattr(self, "_prepared_for", None) is adj and self._operator is not None \
                and getattr(self, "_prepared_print", None) == adjacency_fingerprint(adj) \
                and isinstance(feature, (np.ndarray, Tensor)) and feature.shape[0] == self._operator.shape[1]:
            hops = self.propagate_device(adj, feature)
            if self.output_device == "cuda":
                return hops
            first = torch.from_numpy(feature) if isinstance(feature, np.ndarray) else feature.detach().float().cpu()
            return [first] + [h.cpu() for h in hops[1:]]
        self._prepared_for = None
        spec = self._norm_spec() if self.build_on == "device" else None
        if spec is not None and isinstance(adj, sp.csr_matrix) and isinstance(feature, (np.ndarray, Tensor)) \
                and adj.shape[1] == feature.shape[0]:
            return self._propagate_device_built(adj, feature, spec)
        return propagate_reference_contract(self, adj, feature)

    def _propagate_device_built(self, adj, feature, spec):
        from ..graph_build import operator_from_scipy_device
        if isinstance(feature, np.ndarray) and feature.dtype != np.float32:
            raise TypeError("The feature matrix must be a float32 numpy.ndarray!")
        require_cuda()
        first = feature.detach().float() if isinstance(feature, Tensor) else torch.from_numpy(feature)
        if self._operator is not None:
            self._operator.close()
        r, alpha = spec
        self._operator = operator_from_scipy_device(adj, r=r, alpha=alpha)
        self._adj_value, self._adj_parts = None, self._operator.parts
        if self.output_device == "cuda":
            return self._operator.propagate(first.cuda(), self._prop_steps, mode=self.mode)
        return [first.cpu()] + self._operator.propagate_host(first.cpu(), self._prop_steps, mode=self.mode, keep="all")


class MessageOp(nn.Module):
    """Cross-hop combiner base class (reference base_op.py:40-60)."""

    def __init__(self, start=None, end=None):
        super(MessageOp, self).__init__()
        self._aggr_type = None
        self._start, self._end = start, end

    @property
    def aggr_type(self):
        return self._aggr_type

    def _combine(self, feat_list):
        raise NotImplementedError

    def aggregate(self, feat_list):
        if not isinstance(feat_list, list):
            # the reference *returns* the exception object here (base_op.py:55); raising is the evident intent
            raise TypeError("The input must be a list consists of feature matrices!")
        for feat in feat_list:
            if not isinstance(feat, Tensor):
                raise TypeError("The feature matrices must be tensors!")

        return self._combine(feat_list)

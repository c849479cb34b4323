"""-m gpu: the two drop-in routes of INTEGRATION.md on hardware, against a stand-in `sgl.operators` package
(tests/standin/make_standin.py; the GPU box has no /root/reference).

 1. swap the shared object: the stand-in's UNMODIFIED wrapper (the reference's ctypes binding with
    numpy.ctypeslib.ndpointer argtypes, sgl/operators/utils.py:10-40) loads libsglb200.so under the name
    csrc/libmatmul.so and calls FloatCSRMulDenseOMP -> every hop runs on the GPU, bit-exact vs the oracle's fma chain;
 2. sgl_b200.patch.install(): GraphOp.propagate / csr_sparse_dense_matmul / MessageOp._combine of the stand-in classes
    are re-routed; outputs bit-exact (EXACT mode) and the message ops equal the torch expressions they replace.
"""
import importlib
import os
import shutil
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import sgap_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture()
def standin(tmp_path):
    sys.path.insert(0, os.path.join(HERE, "standin"))
    import make_standin
    root = make_standin.write(str(tmp_path))
    for name in [m for m in sys.modules if m == "sgl" or m.startswith("sgl.")]:
        del sys.modules[name]
    sys.path.insert(0, root)
    yield root
    sys.path.remove(root)
    sys.path.remove(os.path.join(HERE, "standin"))
    for name in [m for m in sys.modules if m == "sgl" or m.startswith("sgl.")]:
        del sys.modules[name]


def _graph(rng, n=3000, m=24000):
    rows = rng.integers(0, n, m)
    cols = (rows + rng.zipf(1.5, m)) % n
    a = sp.csr_matrix((np.ones(2 * m, dtype=np.float32), (np.concatenate([rows, cols]), np.concatenate([cols, rows]))), shape=(n, n))
    a.sum_duplicates()
    return a


def test_swapped_shared_object_runs_the_unmodified_wrapper_on_the_gpu(standin):
    from sgl_b200 import _lib
    shutil.copy(_lib.LIB_PATH, os.path.join(standin, "sgl", "operators", "csrc", "libmatmul.so"))
    graph_op = importlib.import_module("sgl.operators.graph_op")
    rng = np.random.default_rng(5)
    adj = _graph(rng)
    x = rng.standard_normal((adj.shape[0], 64)).astype(np.float32)
    op = graph_op.LaplacianGraphOp(3, r=0.5)
    hops = op.propagate(adj, x)
    ref = O.propagate(O.laplacian_adj(adj, 0.5), x, 3, "fma")
    assert len(hops) == 4 and all(isinstance(h, torch.Tensor) and not h.is_cuda for h in hops)
    for k in range(4):
        assert np.array_equal(hops[k].numpy(), ref[k]), f"hop {k}"
    # the library really is ours: the legacy symbol and the handle API live in the same object
    import ctypes
    lib = ctypes.CDLL(os.path.join(standin, "sgl", "operators", "csrc", "libmatmul.so"))
    assert lib.sglb200_version() == _lib.load().sglb200_version()


def test_patch_install_reroutes_standin_classes(standin):
    import sgl_b200.patch as patch
    base = importlib.import_module("sgl.operators.base_op")
    graph_op = importlib.import_module("sgl.operators.graph_op")
    msg = importlib.import_module("sgl.operators.message_op")
    rng = np.random.default_rng(6)
    adj = _graph(rng)
    x = rng.standard_normal((adj.shape[0], 100)).astype(np.float32)
    K = 4
    ref = O.propagate(O.laplacian_adj(adj, 0.5), x, K, "fma")
    orig = base.GraphOp.propagate
    patch.install()
    try:
        assert base.GraphOp.propagate is not orig
        op = graph_op.LaplacianGraphOp(K, r=0.5)           # the stand-in's own class, its own _construct_adj
        op.mode = "exact"
        hops = op.propagate(adj, x)
        assert [tuple(h.shape) for h in hops] == [x.shape] * (K + 1)
        assert hops[0].numpy().ctypes.data == x.ctypes.data   # element 0 aliases the input like torch.FloatTensor(ndarray)
        for k in range(K + 1):
            assert np.array_equal(hops[k].numpy(), ref[k]), f"hop {k}"
        again = op.propagate(adj, x)                          # second call on the same adjacency: resident operator reused
        assert all(np.array_equal(a.numpy(), b.numpy()) for a, b in zip(hops, again))
        op.mode = "fast"
        fast = op.propagate(adj, x)
        for k in range(1, K + 1):
            err = np.abs(fast[k].numpy() - ref[k]).max() / np.abs(ref[k]).max()
            assert err <= 1e-5
        with pytest.raises(TypeError):
            op.propagate(adj.tocoo(), x)
        with pytest.raises(ValueError):
            op.propagate(adj, x[:-1])
        feats = [torch.from_numpy(r) for r in ref]
        assert np.array_equal(msg.SumMessageOp(0, K + 1).aggregate(feats).numpy(), O.combine_sum(ref, 0, K + 1))
        assert np.array_equal(msg.MeanMessageOp(1, K).aggregate(feats).numpy(), O.combine_mean(ref, 1, K))
        assert np.array_equal(msg.MaxMessageOp(0, K + 1).aggregate(feats).numpy(), np.stack(ref).max(0))
        assert np.array_equal(msg.ConcatMessageOp(0, 3).aggregate(feats).numpy(), np.hstack(ref[:3]))
        nafs = msg.OverSmoothDistanceWeightedOp().aggregate(feats).numpy()
        want = O.combine_osd(ref)
        assert np.abs(nafs - want).max() <= 3e-6 * max(1.0, np.abs(want).max())
    finally:
        patch.uninstall()
    assert base.GraphOp.propagate is orig

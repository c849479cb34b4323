"""CPU tests of the host-side mirror of sgl.operators: normalisation parity with the reference goldens, argument
checking and error behaviour, and the loud failure when no GPU is present."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import GRAPH_TAGS, golden_graph_files, load_graph
from sgl_b200 import SglB200Error
from sgl_b200.operators.graph_op import LaplacianGraphOp, PprGraphOp
from sgl_b200.operators.message_op import (ConcatMessageOp, LastMessageOp, LearnableWeightedMessageOp, MaxMessageOp,
                                           MeanMessageOp, OverSmoothDistanceWeightedOp, SimpleWeightedMessageOp,
                                           SumMessageOp, IterateLearnableWeightedMessageOp)
from sgl_b200.operators.utils import adj_to_symmetric_norm, one_dim_weighted_add, two_dim_weighted_add

HAS_GPU = torch.cuda.is_available()


def _scipy_adj(adj):
    return sp.csr_matrix((adj.data, adj.indices, adj.indptr), shape=adj.shape)


@pytest.mark.parametrize("path", golden_graph_files(), ids=lambda p: os.path.basename(p)[6:-4])
@pytest.mark.parametrize("tag,kind,r,alpha", GRAPH_TAGS)
def test_construct_adj_matches_reference(path, tag, kind, r, alpha):
    z, adj = load_graph(path)
    op = LaplacianGraphOp(3, r=r) if kind == "lap" else PprGraphOp(3, r=r, alpha=alpha)
    a = op._construct_adj(_scipy_adj(adj)).tocsr()
    assert np.array_equal(a.indptr.astype(np.int64), z[tag + "_norm_indptr"])
    assert np.array_equal(a.indices.astype(np.int32), z[tag + "_norm_indices"])
    assert a.data.dtype == np.float64
    assert np.array_equal(a.data, z[tag + "_norm_data"])
    # COO input is accepted like in the reference
    b = op._construct_adj(_scipy_adj(adj).tocoo()).tocsr()
    assert np.array_equal(b.data, a.data)


def test_symmetric_norm_formula_directed():
    # edges 0->1, 0->2: A^[1,0] = A^[2,0] = 1/sqrt(3), A^[0,0] = 1/3
    a = sp.csr_matrix((np.ones(2, dtype=np.float32), ([0, 0], [1, 2])), shape=(3, 3))
    n = adj_to_symmetric_norm(a, 0.5).toarray()
    assert n[1, 0] == pytest.approx(3 ** -0.5) and n[2, 0] == pytest.approx(3 ** -0.5)
    assert n[0, 0] == pytest.approx(1 / 3) and n[1, 1] == 1.0 and n[0, 1] == 0.0


def test_propagate_argument_errors():
    a = sp.csr_matrix(np.eye(3, dtype=np.float32))
    x = np.ones((3, 2), dtype=np.float32)
    with pytest.raises(TypeError):
        LaplacianGraphOp(1).propagate(a.todense(), x)            # not sparse: _construct_adj rejects
    with pytest.raises(TypeError):
        LaplacianGraphOp(1).propagate(a.tocoo(), x)              # COO passes _construct_adj, fails the CSR check
    with pytest.raises(TypeError):
        LaplacianGraphOp(1).propagate(a, [[1.0, 2.0]] * 3)       # feature must be an ndarray (or tensor)
    with pytest.raises(ValueError):
        LaplacianGraphOp(1).propagate(a, np.ones((4, 2), dtype=np.float32))
    if not HAS_GPU:
        with pytest.raises(SglB200Error):
            LaplacianGraphOp(1).propagate(a, x)                  # no silent CPU path
        with pytest.raises(TypeError):
            LaplacianGraphOp(1).propagate(a, x.astype(np.float64))


def test_message_op_validation_and_types():
    assert LastMessageOp().aggr_type == "last"
    assert SumMessageOp(0, 2).aggr_type == "sum" and MeanMessageOp(0, 2).aggr_type == "mean"
    assert MaxMessageOp(0, 2).aggr_type == "max" and ConcatMessageOp(0, 2).aggr_type == "concat"
    assert OverSmoothDistanceWeightedOp().aggr_type == "over_smooth_dis_weighted"
    assert LearnableWeightedMessageOp(0, 3, "gate", 8).aggr_type == "learnable_weighted"
    assert IterateLearnableWeightedMessageOp(0, 3, "recursive", 8).aggr_type == "iterate_learnable_weighted"
    with pytest.raises(ValueError):
        SimpleWeightedMessageOp(0, 2, "nope", 0.5)
    with pytest.raises(TypeError):
        SimpleWeightedMessageOp(0, 2, "alpha", 1)
    with pytest.raises(ValueError):
        SimpleWeightedMessageOp(0, 2, "alpha", 1.5)
    with pytest.raises(ValueError):
        LearnableWeightedMessageOp(0, 2, "jk", 3)
    with pytest.raises(TypeError):
        SumMessageOp(0, 2).aggregate([np.zeros((2, 2))])
    with pytest.raises(TypeError):
        SumMessageOp(0, 2).aggregate(torch.zeros(2, 2))
    f = [torch.ones(4, 3), torch.full((4, 3), 2.0)]
    assert torch.equal(LastMessageOp().aggregate(f), f[-1])
    if not HAS_GPU:
        with pytest.raises(SglB200Error):
            SumMessageOp(0, 2).aggregate(f)


def test_weighted_add_helpers_errors_and_autograd():
    f = [torch.randn(5, 3), torch.randn(5, 3)]
    with pytest.raises(TypeError):
        one_dim_weighted_add(f, [0.5, 0.5])
    with pytest.raises(ValueError):
        one_dim_weighted_add(f, torch.ones(3))
    with pytest.raises(ValueError):
        two_dim_weighted_add(f, torch.ones(5, 3))
    w = torch.tensor([0.25, 0.75], requires_grad=True)
    out = one_dim_weighted_add(f, w)          # autograd branch: plain tensor algebra, usable on CPU
    out.sum().backward()
    assert torch.allclose(w.grad, torch.stack([f[0].sum(), f[1].sum()]))
    w2 = torch.rand(5, 2)
    ref = torch.bmm(torch.stack(f, dim=2), w2.unsqueeze(2)).squeeze(2)
    assert torch.allclose(two_dim_weighted_add(f, w2), ref, atol=1e-6)


@pytest.mark.parametrize("kind", ["simple", "simple_allow_neg", "gate", "ori_ref", "jk"])
@pytest.mark.parametrize("se", [(0, 5), (1, 4)])
def test_learnable_forward_matches_reference(message_golden, kind, se):
    g = message_golden
    s, e = se
    K, d = 4, 16
    batch = [torch.from_numpy(h[g["batch_idx"]]) for h in g["hops"]]
    args = {"simple": (K,), "simple_allow_neg": (K,), "gate": (d,), "ori_ref": (d,), "jk": (K, d)}[kind]
    op = LearnableWeightedMessageOp(s, e, kind, *args)
    tag = f"lw_{kind}_{s}_{e}"
    with torch.no_grad():
        if kind in ("simple", "simple_allow_neg"):
            op._learnable_weight.copy_(torch.from_numpy(g[tag + "_w"]))
        else:
            op._learnable_weight.weight.copy_(torch.from_numpy(g[tag + "_w"]))
            op._learnable_weight.bias.copy_(torch.from_numpy(g[tag + "_b"]))
    feats = [b.clone().requires_grad_(True) for b in batch]
    out = op.aggregate(feats)
    np.testing.assert_allclose(out.detach().numpy(), g[tag + "_out"], rtol=1e-5, atol=1e-6)
    gout = torch.linspace(-1, 1, out.numel()).reshape(out.shape)
    out.backward(gout)
    gin = np.stack([f.grad.numpy() if f.grad is not None else np.zeros((len(g["batch_idx"]), d), np.float32)
                    for f in feats])
    np.testing.assert_allclose(gin, g[tag + "_gin"], rtol=1e-4, atol=1e-5)
    gw = op._learnable_weight.grad if kind.startswith("simple") else op._learnable_weight.weight.grad
    np.testing.assert_allclose(gw.numpy(), g[tag + "_gw"], rtol=1e-4, atol=1e-5)


def test_iterate_learnable_matches_reference(message_golden):
    g = message_golden
    batch = [torch.from_numpy(h[g["batch_idx"]]) for h in g["hops"]]
    op = IterateLearnableWeightedMessageOp(0, 5, "recursive", 16)
    with torch.no_grad():
        op._learnable_weight.weight.copy_(torch.from_numpy(g["iter_w"]))
        op._learnable_weight.bias.copy_(torch.from_numpy(g["iter_b"]))
    np.testing.assert_allclose(op.aggregate(batch).detach().numpy(), g["iter_out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("path", golden_graph_files(), ids=lambda p: os.path.basename(p)[6:-4])
@pytest.mark.parametrize("tag,kind,r,alpha", GRAPH_TAGS)
def test_device_builder_passes_match_reference(path, tag, kind, r, alpha):
    """graph_build.normalized_adjacency_device, engine="torch" (the sort / segment passes the CUDA builder is cross-checked
    against on the GPU; run here on CPU tensors, where it also exercises degree_powers' integer-degree table) must reproduce
    the reference's CSR structure bit-exactly and, with the float64 value formula of sglb200_normalize_values
    restated in numpy, its float64 values."""
    from sgl_b200.graph_build import normalized_adjacency_device
    z, adj = load_graph(path)
    coo = _scipy_adj(adj).tocoo()
    parts = normalized_adjacency_device(torch.from_numpy(coo.row.astype(np.int64)),
                                        torch.from_numpy(coo.col.astype(np.int64)), adj.shape[0],
                                        torch.from_numpy(coo.data.astype(np.float32)), r=r, alpha=alpha, engine="torch")
    assert np.array_equal(parts["indptr"].numpy(), z[tag + "_norm_indptr"])
    assert np.array_equal(parts["indices"].numpy(), z[tag + "_norm_indices"])
    rows = np.repeat(np.arange(adj.shape[0]), np.diff(parts["indptr"].numpy()))
    v = (parts["raw_w"].numpy() * parts["d_left"].numpy()[rows]) * parts["d_right"].numpy()[parts["indices"].numpy()]
    if alpha is not None:
        v = (1 - alpha) * v
        v[rows == parts["indices"].numpy()] += alpha
    ref = z[tag + "_norm_data"]
    if "weighted" in path:   # non-integer weights: the float64 degree sum order (numpy pairwise) is not reproduced
        np.testing.assert_allclose(v, ref, rtol=4e-16)
        assert np.array_equal(v.astype(np.float32), ref.astype(np.float32))
    else:
        assert np.array_equal(v, ref)


def test_hop_cache_round_trip_and_content_keys(tmp_path):
    from sgl_b200.cache import HopCache
    rng = np.random.default_rng(0)
    adj = sp.random(50, 50, density=0.1, format="csr", random_state=1, dtype=np.float32)
    x = rng.standard_normal((50, 8)).astype(np.float32)
    hops = [torch.from_numpy(x)] + [torch.from_numpy(rng.standard_normal((50, 8)).astype(np.float32)) for _ in range(3)]
    cache = HopCache(str(tmp_path))
    key = cache.key(adj, x, "LaplacianGraphOp:fast", 3, r=0.5, alpha=None)
    assert cache.load(key) is None
    cache.save(key, hops)
    back = cache.load(key)
    assert len(back) == 4 and all(torch.equal(a, b) for a, b in zip(back, hops))
    # any change of the inputs changes the key
    assert cache.key(adj, x, "LaplacianGraphOp:fast", 4, r=0.5, alpha=None) != key
    assert cache.key(adj, x, "LaplacianGraphOp:fast", 3, r=0.3, alpha=None) != key
    assert cache.key(adj, x + 1, "LaplacianGraphOp:fast", 3, r=0.5, alpha=None) != key
    adj2 = adj.copy()
    adj2.data[0] += 1
    assert cache.key(adj2, x, "LaplacianGraphOp:fast", 3, r=0.5, alpha=None) != key
    # a cache hit is served without a GPU (this test runs with -m "not gpu")
    op = LaplacianGraphOp(3, r=0.5)
    op.cache_dir = str(tmp_path)
    HopCache(str(tmp_path)).save(HopCache.key(adj, x, "LaplacianGraphOp:" + op.mode, 3, r=0.5, alpha=None), hops)
    got = op.propagate(adj, x)
    assert all(torch.equal(a, b) for a, b in zip(got, hops)) and got[0].data_ptr() == torch.from_numpy(x).data_ptr()


def test_custom_homo_layout_round_trip(tmp_path):
    """sgl_b200.io writes the raw files the reference's Custom_Homo reads (custom_dataset.py:38-85) and reads them back
    into the adjacency the hot path takes (duplicates summed like base_data.py:29-30)."""
    import scipy.sparse as sp
    from sgl_b200 import io
    rng = np.random.default_rng(0)
    n = 50
    row, col = rng.integers(0, n, 300), rng.integers(0, n, 300)
    w = np.ones(300, dtype=np.float32)
    x = rng.standard_normal((n, 6)).astype(np.float32)
    y = np.eye(4)[rng.integers(0, 4, n)]
    d = io.write_custom_homo(str(tmp_path), "toy", (row, col, w), x, y, train_idx=np.arange(10), test_idx=np.arange(10, 20))
    assert sorted(os.listdir(d)) == ["adj_matrix.npz", "indices.npz", "label.npy", "x.npy"]
    f = np.load(os.path.join(d, "adj_matrix.npz"))
    assert set(f.files) == {"row", "col", "data"}
    got = io.read_custom_homo(str(tmp_path), "toy")
    want = sp.csr_matrix((w, (row, col)), shape=(n, n))
    want.sum_duplicates()
    assert (got["adj"] != want).nnz == 0 and got["adj"].dtype == np.float32
    assert np.array_equal(got["x"], x) and np.array_equal(got["y"], y.argmax(1))
    assert np.array_equal(got["train_idx"], np.arange(10)) and got["val_idx"] is None
    with pytest.raises(ValueError):
        io.read_custom_homo(str(tmp_path), "missing", num_node=5)


def test_degree_powers_table_equals_numpy_pow_on_the_vector():
    """graph_build.degree_powers: integer degree vectors go through a numpy pow table of the distinct values (the vector never
    leaves the device); the result must be the same float64 bits as numpy's pow applied to the vector itself, inf -> 0
    (reference utils.py:79-85).  Non-integer vectors take the vector path."""
    from sgl_b200.graph_build import degree_powers
    rng = np.random.default_rng(3)
    for r in (0.5, 0.3, 0.0, 1.0):
        deg = rng.integers(0, 5000, 20000).astype(np.float64)
        deg[:5] = [0.0, 1.0, 2.0, 4999.0, 0.0]
        with np.errstate(divide="ignore", invalid="ignore"):
            dl, dr = np.power(deg, r - 1), np.power(deg, -r)
        dl[np.isinf(dl)] = 0.0
        dr[np.isinf(dr)] = 0.0
        got_l, got_r = degree_powers(torch.from_numpy(deg), r)
        assert np.array_equal(got_l.numpy(), dl) and np.array_equal(got_r.numpy(), dr)
        frac = deg + rng.random(deg.size)
        with np.errstate(divide="ignore", invalid="ignore"):
            fl, fr = np.power(frac, r - 1), np.power(frac, -r)
        got_l, got_r = degree_powers(torch.from_numpy(frac), r)
        assert np.array_equal(got_l.numpy(), fl) and np.array_equal(got_r.numpy(), fr)


def test_feature_split_column_blocks_and_padded_slabs():
    from sgl_b200.dist import FeatureSplitOperator as F
    assert F.column_bounds(100, 8).tolist() == [0, 12, 24, 36, 48, 60, 72, 84, 100]
    assert F.column_bounds(128, 8).tolist() == list(range(0, 129, 16))
    assert F.column_bounds(7, 4).tolist()[-1] == 7 and F.column_bounds(7, 4).tolist()[0] == 0
    for w, ld in ((12, 16), (16, 16), (24, 32), (52, 64), (64, 64), (100, 100), (128, 128)):
        slab = F.block_slab(10, w, "cpu")
        assert slab.shape == (10, w) and slab.stride(0) == ld and slab.stride(1) == 1
    # world = 1 operator on CPU tensors with a caller-supplied hop: propagate keeps dense slabs off the GPU
    a = sp.random(30, 30, 0.2, format="csr", dtype=np.float32, random_state=0)
    at = torch.from_numpy(a.toarray())
    fs = F(world=1, rank=0, local_hop=lambda x, out: out.copy_(at @ x))
    hops = fs.propagate(torch.ones(30, 12), 2)
    assert torch.allclose(hops[2], at @ (at @ torch.ones(30, 12)))
    assert F.row_bounds(10, 4).tolist() == [0, 2, 5, 7, 10]

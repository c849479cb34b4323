"""CPU tests (-m "not gpu"): pin the oracle against the golden vectors generated from the unmodified reference
(oracle/gen_golden.py) and, when present, against the reference's own kernel compiled in place (oracle/_ref)."""
import os

import numpy as np
import pytest

from conftest import GRAPH_TAGS, golden_graph_files, load_graph
from oracle import sgap_oracle as O


def _norm(adj, kind, r, alpha):
    return O.laplacian_adj(adj, r) if kind == "lap" else O.ppr_adj(adj, r, alpha)


@pytest.mark.parametrize("path", golden_graph_files(), ids=lambda p: os.path.basename(p)[6:-4])
@pytest.mark.parametrize("tag,kind,r,alpha", GRAPH_TAGS)
def test_normalised_adjacency_matches_reference(path, tag, kind, r, alpha):
    z, adj = load_graph(path)
    a = _norm(adj, kind, r, alpha)
    assert np.array_equal(a.indptr, z[tag + "_norm_indptr"])            # structure: bit-exact
    assert np.array_equal(a.indices, z[tag + "_norm_indices"])
    ref = z[tag + "_norm_data"]
    assert a.data.dtype == np.float64 and ref.dtype == np.float64
    assert np.array_equal(a.data, ref)                                   # float64 values: bit-exact


@pytest.mark.parametrize("path", golden_graph_files(), ids=lambda p: os.path.basename(p)[6:-4])
@pytest.mark.parametrize("tag,kind,r,alpha", GRAPH_TAGS)
def test_propagate_matches_reference(path, tag, kind, r, alpha):
    z, adj = load_graph(path)
    a = _norm(adj, kind, r, alpha)
    K = z[tag + "_hops_fma"].shape[0] - 1
    hops = O.propagate(a, z["x"], K, "fma")
    assert np.array_equal(np.stack(hops), z[tag + "_hops_fma"])          # == shipped libmatmul.so, bit-exact
    if tag + "_hops_f64" in z:
        hops = O.propagate(a, z["x"], K, "f64")
        assert np.array_equal(np.stack(hops), z[tag + "_hops_f64"])      # == scipy fp64 branch
        hops = O.propagate(a, z["x"], K, "muladd")
        assert np.array_equal(np.stack(hops), z[tag + "_hops_scipy32"])  # == scipy csr(f32).dot


@pytest.mark.parametrize("path", golden_graph_files(), ids=lambda p: os.path.basename(p)[6:-4])
def test_wrapper_hop(path):
    z, adj = load_graph(path)
    a = O.laplacian_adj(adj, 0.5)
    assert np.array_equal(O.spmm_hop(a, z["x"], "fma"), z["wrapper_hop"])


def test_combiners_match_reference(message_golden):
    g = message_golden
    hops = list(g["hops"])
    assert np.array_equal(O.combine_last(hops), g["last"])
    for (s, e) in [(0, 5), (1, 4)]:
        t = f"_{s}_{e}"
        assert np.array_equal(O.combine_sum(hops, s, e), g["sum" + t])
        assert np.array_equal(O.combine_mean(hops, s, e), g["mean" + t])
        assert np.array_equal(O.combine_max(hops, s, e), g["max" + t])
        assert np.array_equal(O.combine_min(hops, s, e), g["min" + t])
        assert np.array_equal(O.combine_concat(hops, s, e), g["concat" + t])
        for al in (0.85, 0.1):
            w = O.alpha_weights(len(hops), al, s, e)
            assert np.array_equal(O.combine_weighted(hops, w, s, e), g[f"alpha{al}" + t])
    assert np.array_equal(O.combine_weighted(hops, g["hand_weights"], 0, 5), g["hand_0_5"])
    np.testing.assert_allclose(O.combine_osd(hops), g["osd"], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("kind", ["simple", "simple_allow_neg", "gate", "ori_ref", "jk"])
@pytest.mark.parametrize("se", [(0, 5), (1, 4)])
def test_learnable_forward_matches_reference(message_golden, kind, se):
    g = message_golden
    s, e = se
    batch = [h[g["batch_idx"]] for h in g["hops"]]
    tag = f"lw_{kind}_{s}_{e}"
    out = O.combine_learnable(batch, s, e, kind, g[tag + "_w"], g.get(tag + "_b"))
    np.testing.assert_allclose(out, g[tag + "_out"], rtol=1e-5, atol=1e-6)


def test_against_reference_kernel_build():
    lib = O.load_reference_kernel()
    if lib is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(3)
    n, d, m = 500, 96, 6000
    rows, cols = rng.integers(0, n, m), rng.integers(0, n, m)
    adj = O._canonical_csr(rows.astype(np.int64), cols.astype(np.int64), rng.standard_normal(m).astype(np.float32), (n, n))
    x = rng.standard_normal((n, d)).astype(np.float32)
    ours = O.spmm_hop(adj, x, "fma")
    theirs = O.reference_kernel_hop(lib, adj, x)
    assert np.array_equal(ours, theirs)


def test_empty_and_ragged_rows():
    # rows 0 and 3 empty, row 2 long; zero-column feature matrix rejected by shape mismatch check
    indptr = np.array([0, 0, 1, 6, 6], dtype=np.int64)
    indices = np.array([1, 0, 1, 2, 3, 3], dtype=np.int32)[:6]
    adj = O.Csr(indptr, indices, np.arange(1, 7, dtype=np.float32), (4, 4))
    x = np.arange(8, dtype=np.float32).reshape(4, 2)
    y = O.spmm_hop(adj, x, "fma")
    assert np.array_equal(y[0], [0, 0]) and np.array_equal(y[3], [0, 0])
    dense = np.zeros((4, 4), dtype=np.float32)
    for i in range(4):
        for j in range(indptr[i], indptr[i + 1]):
            dense[i, indices[j]] += adj.data[j]
    np.testing.assert_allclose(y, dense @ x)
    with pytest.raises(ValueError):
        O.spmm_hop(adj, np.zeros((5, 2), dtype=np.float32))


def test_label_propagation_matches_reference():
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "tricks.npz")))
    z, adj = load_graph(os.path.join(GOLDEN, "graph_skewed200.npz"))
    a = O.laplacian_adj(adj, 0.5)
    onehot = np.eye(int(g["labels"].max()) + 1, dtype=np.float32)[g["labels"]]
    np.testing.assert_allclose(O.label_propagation(onehot, a, 4, 0.75, mask=g["mask"]), g["lp_masked"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(O.label_propagation(onehot, a, 3, 0.5), g["lp_full"], rtol=1e-6, atol=1e-7)

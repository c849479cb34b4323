"""CPU tests, build container only (skipped where /root/reference does not exist, e.g. on the GPU box): the oracle and
the host-side mirror against the UNMODIFIED reference on seeded random graphs -- beyond the committed golden vectors.
The reference is imported in place; nothing is copied."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import sgap_oracle as O

REF = os.environ.get("SGL_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "sgl", "operators")),
                                reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    try:
        from sgl.operators.graph_op import LaplacianGraphOp, PprGraphOp
        from sgl.operators.utils import adj_to_symmetric_norm
        yield {"lap": LaplacianGraphOp, "ppr": PprGraphOp, "norm": adj_to_symmetric_norm}
    finally:
        sys.path.remove(REF)


def _random_adj(rng, kind):
    n = int(rng.integers(3, 120))
    m = int(rng.integers(0, 6 * n))
    rows, cols = rng.integers(0, n, m), rng.integers(0, n, m)
    if kind == "undirected":
        rows, cols = np.concatenate([rows, cols]), np.concatenate([cols, rows])
    if kind == "weighted":
        vals = rng.uniform(0.1, 4.0, rows.size).astype(np.float32)
    else:
        vals = np.ones(rows.size, dtype=np.float32)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))          # duplicates summed, like the reference's Edge


@pytest.mark.parametrize("kind", ["undirected", "directed", "weighted"])
@pytest.mark.parametrize("seed", range(6))
def test_normalisation_and_hops_match_reference(ref, kind, seed):
    rng = np.random.default_rng(1000 * seed + len(kind))
    adj = _random_adj(rng, kind)
    n = adj.shape[0]
    r = float(rng.choice([0.0, 0.25, 0.5, 1.0]))
    alpha = float(rng.choice([0.1, 0.15, 0.5]))
    x = rng.standard_normal((n, int(rng.integers(1, 40)))).astype(np.float32)
    K = int(rng.integers(1, 4))
    for name, op_ref, ours in (("lap", ref["lap"](K, r=r), O.laplacian_adj(adj, r)),
                               ("ppr", ref["ppr"](K, r=r, alpha=alpha), O.ppr_adj(adj, r, alpha))):
        hops = op_ref.propagate(adj, x)                                # reference: scipy build + shipped libmatmul.so
        a = op_ref._adj.tocsr()
        assert np.array_equal(a.indptr, ours.indptr) and np.array_equal(a.indices, ours.indices), name
        if kind == "weighted":
            # non-integer weights: numpy's pairwise float64 row sums vs the oracle's same primitive -- still exact
            assert np.array_equal(a.data, ours.data), name
        else:
            assert np.array_equal(a.data, ours.data), name
        got = O.propagate(ours, x, K, "fma")
        for k in range(K + 1):
            assert np.array_equal(got[k], hops[k].numpy()), (name, k)


@pytest.mark.parametrize("seed", range(4))
def test_host_mirror_construct_adj_matches_reference(ref, seed):
    from sgl_b200.operators.graph_op import LaplacianGraphOp, PprGraphOp
    rng = np.random.default_rng(77 + seed)
    adj = _random_adj(rng, ["undirected", "directed", "weighted", "undirected"][seed])
    r = [0.5, 0.3, 0.0, 1.0][seed]
    for mine, theirs in ((LaplacianGraphOp(2, r=r), ref["lap"](2, r=r)), (PprGraphOp(2, r=r, alpha=0.2), ref["ppr"](2, r=r, alpha=0.2))):
        a = mine._construct_adj(adj).tocsr()
        b = theirs._construct_adj(adj).tocsr()
        a.sort_indices()
        b.sort_indices()
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
        assert np.array_equal(a.data, b.data)

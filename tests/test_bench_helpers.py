"""CPU tests of bench.py's workload generator and accounting helpers."""
import numpy as np
import torch

import bench


def test_rmat_edges_are_deterministic_and_in_range():
    a = bench.rmat_edges(1000, 5000, 12, seed=3, device="cpu")
    b = bench.rmat_edges(1000, 5000, 12, seed=3, device="cpu")
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert int(a[0].min()) >= 0 and int(a[0].max()) < 1000 and int(a[1].max()) < 1000
    c = bench.rmat_edges(1000, 5000, 12, seed=4, device="cpu")
    assert not torch.equal(a[0], c[0])


def test_build_adjacency_matches_reference_symmetrisation():
    adj, d, K = bench.build_adjacency("pubmed", "cpu")
    assert adj.shape == (19717, 19717) and d == 500 and K == 3
    assert adj.dtype == np.float32 and adj.indices.dtype == np.int32
    assert (abs(adj - adj.T) > 0).nnz == 0                       # concatenation of (src,dst) and (dst,src): symmetric
    assert adj.has_canonical_format and adj.data.min() >= 1.0      # duplicates summed, never dropped


def test_algorithmic_bytes_formula():
    # 8*nnz + 8*(N+1) + 8*N*d
    assert bench.algorithmic_bytes_per_hop(10, 100, 4) == 800 + 88 + 320

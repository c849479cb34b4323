"""GPU parity at BASELINE.json's sizes (-m gpu): products-shape (configs[2]), pubmed-shape (configs[0]) and a CSR whose
row pointers pass 2^31 (configs[3]/[4] need int64 indptr; the reference's kernel overflows there, matmul.c:33).

The oracle cannot run the full products-shape hop in seconds, so every hop is checked on a row sample: hop k of the
sampled rows is recomputed by the oracle's fma chain from the device's own hop k-1 (a one-hop check isolates the
kernel).  Tolerance (FAST mode): Frobenius-relative and max-abs relative to max|ref| <= 1e-5;
EXACT mode: bit-exact.  Nothing here reads /root/reference.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import sgap_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import bench
    from sgl_b200.graph_build import build_operator_device, values_from_parts
    from sgl_b200.operators.graph_op import LaplacianGraphOp
    from sgl_b200.runtime import CsrOperator


def _rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return (np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300),
            np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300))


def _sample_rows_csr(indptr, indices, vals, sample):
    """Host CSR (int64 indptr, int32 indices, float32 values) of the sampled rows of a device CSR."""
    sample_t = torch.from_numpy(sample).to(indptr.device)
    starts, ends = indptr[sample_t], indptr[sample_t + 1]
    lens = (ends - starts).cpu().numpy()
    sub_ptr = np.zeros(sample.size + 1, dtype=np.int64)
    np.cumsum(lens, out=sub_ptr[1:])
    pos = torch.repeat_interleave(starts - torch.from_numpy(sub_ptr[:-1]).to(indptr.device), torch.from_numpy(lens).to(indptr.device)) \
        + torch.arange(int(sub_ptr[-1]), device=indptr.device)
    return sub_ptr, indices[pos].to(torch.int32).cpu().numpy(), vals[pos].to(torch.float32).cpu().numpy()


def _oracle_rows(sub_ptr, sub_idx, sub_val, x_host, d):
    ref = np.zeros((sub_ptr.size - 1, d), dtype=np.float32)
    O._lib().oracle_spmm_f32_fma_i64(ref, sub_val, sub_idx, sub_ptr, np.ascontiguousarray(x_host), sub_ptr.size - 1, d)
    return ref


def test_products_shape_parity_per_hop():
    """configs[2] shape: N=2,449,029, nnz(A^) ~ 121M, d=100, K=6 -- every hop, FAST and EXACT, on 3000 sampled rows
    (the 64 longest rows included: they are the ones FAST mode cuts across warps)."""
    dev = torch.device("cuda", 0)
    rows, cols, n, d, K = bench.device_graph("products", dev)
    op = build_operator_device(rows, cols, n, r=0.5)
    parts = op.parts
    # the hand-written builder (7 + 3 radix passes over 124 M keys, three-level scan) against the torch.sort version
    from sgl_b200.graph_build import normalized_adjacency_device
    check = normalized_adjacency_device(rows, cols, n, r=0.5, engine="torch")
    for key in ("indptr", "indices", "raw_w", "deg", "d_left", "d_right"):
        assert torch.equal(check[key], parts[key]), key
    del rows, cols, check
    vals = values_from_parts(parts).to(torch.float32)
    indptr, indices = parts["indptr"], parts["indices"]
    assert int(indptr[-1]) == op.nnz and op.nnz > 100_000_000
    rng = np.random.default_rng(7)
    deg = torch.diff(indptr)
    longest = torch.topk(deg, 64).indices.cpu().numpy()
    sample = np.unique(np.concatenate([rng.choice(n, 3000, replace=False), longest])).astype(np.int64)
    sub_ptr, sub_idx, sub_val = _sample_rows_csr(indptr, indices, vals, sample)
    del vals
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(3)).to(dev)
    hops_fast = op.propagate(x, K, mode="fast")
    hops_exact = op.propagate(x, K, mode="exact")
    sample_t = torch.from_numpy(sample).to(dev)
    for k in range(1, K + 1):
        ref = _oracle_rows(sub_ptr, sub_idx, sub_val, hops_exact[k - 1].cpu().numpy(), d)
        assert np.array_equal(hops_exact[k][sample_t].cpu().numpy(), ref), f"exact hop {k} differs from the fma chain"
        ref_f = _oracle_rows(sub_ptr, sub_idx, sub_val, hops_fast[k - 1].cpu().numpy(), d)
        fro, mx = _rel(hops_fast[k][sample_t].cpu().numpy(), ref_f)
        assert fro <= 1e-5 and mx <= 1e-5, (k, fro, mx)
    # whole-matrix agreement of the two schedules on the last hop (north_star: 1e-5 on ogbn-products)
    fro, mx = _rel(hops_fast[K].cpu().numpy(), hops_exact[K].cpu().numpy())
    assert fro <= 1e-5 and mx <= 1e-5, (fro, mx)
    op.close()


def test_pubmed_shape_parity_full():
    """configs[0] shape: N=19,717, 44,324 undirected edges, d=500 (4 slices per lane), K=3, through GraphOp.propagate;
    the whole result against the oracle (bit-exact in EXACT mode), features row-normalised and non-negative like the
    reference's Planetoid loader (dataset/planetoid.py:40-47)."""
    rng = np.random.default_rng(11)
    n, m, d, K = 19_717, 44_324, 500, 3
    r0, c0 = rng.integers(0, n, m), rng.integers(0, n, m)
    adj = sp.csr_matrix((np.ones(2 * m, dtype=np.float32), (np.concatenate([r0, c0]), np.concatenate([c0, r0]))), shape=(n, n))
    x = rng.random((n, d)).astype(np.float32) * (rng.random((n, d)) < 0.1)
    x = (x / np.maximum(x.sum(1, keepdims=True), 1e-12)).astype(np.float32)
    ref = O.propagate(O.laplacian_adj(adj, 0.5), x, K, "fma")
    op = LaplacianGraphOp(K, r=0.5)
    op.mode = "exact"
    hops = op.propagate(adj, x)
    for k in range(K + 1):
        assert np.array_equal(hops[k].numpy(), ref[k]), f"hop {k}"
    op.mode = "fast"
    hops = op.propagate(adj, x)
    for k in range(1, K + 1):
        fro, mx = _rel(hops[k].numpy(), ref[k])
        assert fro <= 1e-5 and mx <= 1e-5, (k, fro, mx)


def test_row_pointers_beyond_2_31():
    """nnz = 2^31 + 2^27 (int64 indptr mandatory): rows of 4096 entries, tiny feature width.  Sampled rows -- in
    particular those whose entries straddle and follow offset 2^31 -- against numpy in float64 and, EXACT mode,
    against the oracle's fma chain bit for bit.  The reference's int32 `indptr` / `indices[j]*mat_col` cannot address
    this matrix (matmul.c:28-33)."""
    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~40 GB of device memory")
    row_len, d = 4096, 4
    n_rows = (2 ** 31 + 2 ** 27) // row_len
    nnz = n_rows * row_len
    n_cols = 1_000_003
    assert nnz > 2 ** 31
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    vals = torch.empty(nnz, dtype=torch.float32, device=dev)
    chunk = 1 << 27
    for j0 in range(0, nnz, chunk):
        j = torch.arange(j0, min(nnz, j0 + chunk), dtype=torch.int64, device=dev)
        indices[j0:j0 + j.numel()] = ((j * 2654435761 + (j >> 13)) % n_cols).to(torch.int32)
        vals[j0:j0 + j.numel()] = (((j * 40503) % 2001).to(torch.float32) - 1000.0) / 1024.0
        del j
    indptr = torch.arange(0, nnz + 1, row_len, dtype=torch.int64, device=dev)
    op = CsrOperator(indptr, indices, vals, (n_rows, n_cols))
    x = torch.randn(n_cols, d, generator=torch.Generator().manual_seed(5)).to(dev)
    boundary = (2 ** 31) // row_len
    sample = np.unique(np.concatenate([[0, 1, boundary - 1, boundary, boundary + 1, n_rows - 2, n_rows - 1],
                                       np.random.default_rng(2).choice(n_rows, 200, replace=False)])).astype(np.int64)
    sub_ptr, sub_idx, sub_val = _sample_rows_csr(indptr, indices, vals, sample)
    assert int(indptr[sample[-1] + 1]) > 2 ** 31
    xh = x.cpu().numpy()
    ref = _oracle_rows(sub_ptr, sub_idx, sub_val, xh, d)
    sample_t = torch.from_numpy(sample).to(dev)
    y_exact = op.spmm(x, mode="exact")
    assert np.array_equal(y_exact[sample_t].cpu().numpy(), ref)
    y_fast = op.spmm(x, mode="fast")
    ref64 = np.stack([(sub_val[sub_ptr[i]:sub_ptr[i + 1], None].astype(np.float64)
                       * xh[sub_idx[sub_ptr[i]:sub_ptr[i + 1]]].astype(np.float64)).sum(0) for i in range(sample.size)])
    fro, mx = _rel(y_fast[sample_t].cpu().numpy(), ref64)
    assert fro <= 1e-5 and mx <= 1e-5, (fro, mx)
    # checksum over ALL rows: column sums of Y against a float64 reduction of vals (x) X[indices] done in chunks
    total = torch.zeros(d, dtype=torch.float64, device=dev)
    for j0 in range(0, nnz, chunk):
        sl = slice(j0, min(nnz, j0 + chunk))
        total += (vals[sl].to(torch.float64)[:, None] * x[indices[sl].to(torch.int64)].to(torch.float64)).sum(0)
    got = y_fast.to(torch.float64).sum(0)
    scale = float((vals.abs().to(torch.float64).sum() * x.abs().max()).item())
    assert float((got - total).abs().max().item()) <= 1e-6 * scale
    op.close()

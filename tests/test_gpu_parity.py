"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle and the golden vectors.

Bars: CSR structure and EXACT-mode hops bit-exact against the reference goldens / the oracle's fma chain;
FAST-mode hops within 1e-5 (Frobenius-relative and max-abs relative to max|ref|, per hop: BASELINE.md section 4);
sum/mean/max/min/concat/weighted combiners bit-exact; NAFS weights within 2e-6.
Nothing here reads /root/reference.
"""
import ctypes
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import GRAPH_TAGS, ROOT, golden_graph_files, load_graph
from oracle import sgap_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from sgl_b200 import _lib
    from sgl_b200.operators.graph_op import LaplacianGraphOp, PprGraphOp
    from sgl_b200.operators.message_op import (ConcatMessageOp, LastMessageOp, MaxMessageOp, MeanMessageOp,
                                               MinMessageOp, OverSmoothDistanceWeightedOp, SimpleWeightedMessageOp,
                                               SumMessageOp)
    from sgl_b200.operators.utils import csr_sparse_dense_matmul, cuda_csr_sparse_dense_matmul
    from sgl_b200.runtime import CsrOperator, aggregate, gather_rows


def rel_errors(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    fro = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)
    mx = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300) if ref.size else 0.0
    return fro, mx


def assert_close_1e5(got, ref):
    fro, mx = rel_errors(got, ref)
    assert fro <= 1e-5 and mx <= 1e-5, (fro, mx)


def _scipy_adj(adj):
    return sp.csr_matrix((adj.data, adj.indices, adj.indptr), shape=adj.shape)


def random_graph(rng, n, m, skew=1.3, weights=False, undirected=True):
    rows = rng.integers(0, n, m)
    cols = (rng.zipf(skew, m) - 1) % n
    if undirected:
        rows, cols = np.concatenate([rows, cols]), np.concatenate([cols, rows])
    vals = rng.uniform(0.25, 2.0, rows.size).astype(np.float32) if weights else np.ones(rows.size, dtype=np.float32)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))


# --------------------------------------------------------------------------------------------------------------
# golden vectors from the unmodified reference
# --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", golden_graph_files(), ids=lambda p: os.path.basename(p)[6:-4])
@pytest.mark.parametrize("tag,kind,r,alpha", GRAPH_TAGS)
def test_golden_propagate(path, tag, kind, r, alpha):
    z, adj = load_graph(path)
    K = z[tag + "_hops_fma"].shape[0] - 1
    op = LaplacianGraphOp(K, r=r) if kind == "lap" else PprGraphOp(K, r=r, alpha=alpha)
    op.mode = "exact"
    hops = op.propagate(_scipy_adj(adj), z["x"].copy())
    assert len(hops) == K + 1 and all(isinstance(h, torch.Tensor) and h.dtype == torch.float32 and not h.is_cuda
                                      for h in hops)
    got = np.stack([h.numpy() for h in hops])
    assert np.array_equal(got, z[tag + "_hops_fma"])                      # == shipped libmatmul.so, bit for bit
    a = op._adj
    assert np.array_equal(a.indptr, z[tag + "_norm_indptr"]) and np.array_equal(a.indices, z[tag + "_norm_indices"])
    op.mode = "fast"
    got = np.stack([h.numpy() for h in op.propagate(_scipy_adj(adj), z["x"].copy())])
    for k in range(K + 1):
        assert_close_1e5(got[k], z[tag + "_hops_fma"][k])
        if tag + "_hops_f64" in z:
            assert_close_1e5(got[k], z[tag + "_hops_f64"][k])             # north_star: scipy CPU path
            assert_close_1e5(got[k], z[tag + "_hops_scipy32"][k])


@pytest.mark.parametrize("path", golden_graph_files(), ids=lambda p: os.path.basename(p)[6:-4])
def test_golden_wrapper_and_first_hop_aliases_input(path):
    z, adj = load_graph(path)
    op = LaplacianGraphOp(1, r=0.5)
    x = z["x"].copy()
    hops = op.propagate(_scipy_adj(adj), x)
    assert hops[0].data_ptr() == torch.from_numpy(x).data_ptr()          # element 0 shares memory
    assert np.array_equal(csr_sparse_dense_matmul(op._adj, x), z["wrapper_hop"])
    assert_close_1e5(cuda_csr_sparse_dense_matmul(op._adj, x), z["wrapper_hop"])


def test_golden_combiners(message_golden):
    g = message_golden
    hops = [torch.from_numpy(h) for h in g["hops"]]
    dev = [h.cuda() for h in hops]
    assert torch.equal(LastMessageOp().aggregate(hops), hops[-1])
    for (s, e) in [(0, 5), (1, 4)]:
        t = f"_{s}_{e}"
        for feats in (hops, dev):     # CPU tensors in -> CPU out; CUDA in -> CUDA out
            def run(op):
                out = op.aggregate(feats)
                assert out.is_cuda == feats[0].is_cuda
                return out.cpu().numpy()
            assert np.array_equal(run(SumMessageOp(s, e)), g["sum" + t])
            assert np.array_equal(run(MeanMessageOp(s, e)), g["mean" + t])
            assert np.array_equal(run(MaxMessageOp(s, e)), g["max" + t])
            assert np.array_equal(run(MinMessageOp(s, e)), g["min" + t])
            assert np.array_equal(run(ConcatMessageOp(s, e)), g["concat" + t])
            assert np.array_equal(run(SimpleWeightedMessageOp(s, e, "alpha", 0.85)), g["alpha0.85" + t])
            assert np.array_equal(run(SimpleWeightedMessageOp(s, e, "alpha", 0.1)), g["alpha0.1" + t])
    out = SimpleWeightedMessageOp(0, 5, "hand_crafted", [float(v) for v in g["hand_weights"]]).aggregate(hops)
    assert np.array_equal(out.numpy(), g["hand_0_5"])
    osd = OverSmoothDistanceWeightedOp().aggregate(hops).numpy()
    np.testing.assert_allclose(osd, g["osd"], rtol=2e-6, atol=2e-6)


# --------------------------------------------------------------------------------------------------------------
# oracle on seeded random inputs: every kernel shape, ragged rows, cut rows
# --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [1, 3, 4, 8, 12, 16, 32, 47, 48, 64, 68, 100, 128, 200, 256, 500, 1030])
def test_feature_widths_exact_and_fast(d):
    rng = np.random.default_rng(d)
    n = 700
    adj = random_graph(rng, n, 5000, weights=True)
    a = O.laplacian_adj(adj, 0.5)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.spmm_hop(a, x, "fma")
    op = CsrOperator(a.indptr, a.indices, a.data.astype(np.float32), a.shape)
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(op.spmm(xd, mode="exact").cpu().numpy(), ref)
    assert_close_1e5(op.spmm(xd, mode="fast").cpu().numpy(), ref)
    op.close()


@pytest.mark.parametrize("tile_items,split", [(32, 1), (32, 40), (128, 0), (512, 16)])
def test_cut_rows_are_folded_deterministically(tile_items, split):
    """tiny tiles + tiny split threshold force most rows across several warps: exercises the carry workspace."""
    rng = np.random.default_rng(tile_items + split)
    n, d = 400, 128
    adj = random_graph(rng, n, 30000, skew=1.15)           # a few rows with thousands of entries
    a = O.laplacian_adj(adj, 0.5)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.spmm_hop(a, x, "fma")
    op = CsrOperator(a.indptr, a.indices, a.data.astype(np.float32), a.shape, tile_items=tile_items,
                     split_threshold=split)
    info = op.info()
    if split in (1, 16, 40):
        assert info["carry_runs"] > 0
    xd = torch.from_numpy(x).cuda()
    y1 = op.spmm(xd, mode="fast").cpu().numpy()
    y2 = op.spmm(xd, mode="fast").cpu().numpy()
    assert np.array_equal(y1, y2)                           # run-to-run deterministic (no atomics)
    assert_close_1e5(y1, ref)
    assert np.array_equal(op.spmm(xd, mode="exact").cpu().numpy(), ref)
    op.close()


@pytest.mark.parametrize("d", [12, 16, 32, 64])
def test_narrow_rows_lane_group_kernel_cut_rows_and_tile_ranges(d):
    """d <= 64: several tiles per warp, one lane group each (spmm_group.cu) -- cut rows, tile ranges, accumulate."""
    rng = np.random.default_rng(100 + d)
    n = 900
    adj = random_graph(rng, n, 40000, skew=1.2)
    a = O.laplacian_adj(adj, 0.5)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.spmm_hop(a, x, "fma")
    op = CsrOperator(a.indptr, a.indices, a.data.astype(np.float32), a.shape, tile_items=64, split_threshold=24)
    assert op.info()["carry_runs"] > 0
    xd = torch.from_numpy(x).cuda()
    y1 = op.spmm(xd, mode="fast")
    assert torch.equal(y1, op.spmm(xd, mode="fast"))
    assert_close_1e5(y1.cpu().numpy(), ref)
    assert np.array_equal(op.spmm(xd, mode="exact").cpu().numpy(), ref)
    tb, rb = op.chunks(3, mode="fast")
    y3 = torch.full_like(y1, float("nan"))
    for c in range(3):
        op.spmm_tiles(xd, y3, tb[c], tb[c + 1], mode="fast")
    assert torch.equal(y3, y1)
    base = rng.standard_normal((n, d)).astype(np.float32)
    yacc = torch.from_numpy(base).cuda()
    op.spmm(xd, out=yacc, mode="exact", accumulate=True)
    want = base.copy()
    O._lib().oracle_spmm_f32_fma_i64(want, a.data.astype(np.float32), a.indices, a.indptr.astype(np.int64), x, n, d)
    assert np.array_equal(yacc.cpu().numpy(), want)
    op.close()


@pytest.fixture()
def tma_kernel(monkeypatch):
    monkeypatch.setenv("SGLB200_TMA", "1")
    yield


@pytest.mark.parametrize("d", [68, 100, 128, 200, 256])
@pytest.mark.parametrize("tile_items,split", [(0, 0), (64, 24), (32, 1)])
def test_tma_staged_kernel_matches_oracle(tma_kernel, d, tile_items, split):
    """spmm_tma.cu (gather4 feature rows + bulk-copied index stream): bit-exact in EXACT mode, 1e-5 in FAST mode with cut
    rows, equal to the register-staged kernel's tile ranges, deterministic run to run."""
    rng = np.random.default_rng(1000 + d + tile_items)
    n = 1500
    adj = random_graph(rng, n, 30000, skew=1.2, weights=True)
    a = O.laplacian_adj(adj, 0.5)                      # full diagonal: no empty rows
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.spmm_hop(a, x, "fma")
    op = CsrOperator(a.indptr, a.indices, a.data.astype(np.float32), a.shape, tile_items=tile_items, split_threshold=split)
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(op.spmm(xd, mode="exact").cpu().numpy(), ref)
    y1 = op.spmm(xd, mode="fast")
    assert torch.equal(y1, op.spmm(xd, mode="fast"))
    assert_close_1e5(y1.cpu().numpy(), ref)
    tb, rb = op.chunks(3, mode="fast")
    y3 = torch.full_like(y1, float("nan"))
    for c in range(3):
        op.spmm_tiles(xd, y3, tb[c], tb[c + 1], mode="fast")
    assert torch.equal(y3, y1)
    # strided input and output (column blocks of a concat slab)
    slab = torch.zeros((n, 3 * d), device="cuda")
    slab[:, d:2 * d] = xd
    op.spmm(slab[:, d:2 * d], out=slab[:, 2 * d:], mode="exact")
    assert np.array_equal(slab[:, 2 * d:].cpu().numpy(), ref)
    op.close()


def test_tma_kernel_falls_back_on_graphs_with_empty_rows(tma_kernel):
    rng = np.random.default_rng(77)
    n, d = 500, 128
    dense = (rng.random((n, n)) < 0.01).astype(np.float32)
    dense[::7] = 0.0                                   # empty rows: the row-end flags cannot describe them
    a = sp.csr_matrix(dense)
    x = rng.standard_normal((n, d)).astype(np.float32)
    op = CsrOperator(a.indptr, a.indices, a.data, a.shape)
    y = op.spmm(torch.from_numpy(x).cuda(), mode="exact").cpu().numpy()
    want = np.zeros((n, d), dtype=np.float32)
    O._lib().oracle_spmm_f32_fma_i64(want, a.data, a.indices, a.indptr.astype(np.int64), x, n, d)
    assert np.array_equal(y, want)
    op.close()


def test_empty_rows_empty_matrix_and_rectangular():
    rng = np.random.default_rng(5)
    # rows 0, 3 and the last 70 rows are empty; rectangular 200 x 90 operator (a row partition)
    n_rows, n_cols, d = 200, 90, 36
    rows = rng.integers(1, 130, 900)
    rows = rows[rows != 3]
    cols = rng.integers(0, n_cols, rows.size)
    m = sp.csr_matrix((rng.standard_normal(rows.size).astype(np.float32), (rows, cols)), shape=(n_rows, n_cols))
    m.sum_duplicates()
    a = O.Csr(m.indptr, m.indices, m.data, m.shape)
    x = rng.standard_normal((n_cols, d)).astype(np.float32)
    ref = np.zeros((n_rows, d), dtype=np.float32)
    O._lib().oracle_spmm_f32_fma_i64(ref, a.data.astype(np.float32), a.indices, a.indptr, x, n_rows, d)
    op = CsrOperator(a.indptr, a.indices, a.data, a.shape, tile_items=32)
    for mode in ("exact", "fast"):
        y = op.spmm(torch.from_numpy(x).cuda(), mode=mode).cpu().numpy()
        assert np.array_equal(y, ref) if mode == "exact" else np.allclose(y, ref, rtol=1e-5, atol=1e-6)
        assert not y[0].any() and not y[3].any() and not y[130:].any()
    op.close()
    # nnz == 0 and n == 0
    z = CsrOperator(np.zeros(6, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.float32), (5, 5))
    assert not z.spmm(torch.ones(5, 4, device="cuda")).cpu().numpy().any()
    z.close()
    e = CsrOperator(np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.float32), (0, 0))
    assert e.spmm(torch.ones(0, 4, device="cuda")).shape == (0, 4)
    e.close()


def test_int32_indptr_device_inputs_and_strided_output():
    rng = np.random.default_rng(9)
    n, d, K = 300, 64, 3
    adj = random_graph(rng, n, 2500)
    a = O.laplacian_adj(adj, 0.5)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.propagate(a, x, K, "fma")
    op = CsrOperator(torch.from_numpy(a.indptr.astype(np.int32)).cuda(), torch.from_numpy(a.indices).cuda(),
                     torch.from_numpy(a.data.astype(np.float32)).cuda(), a.shape)
    hops = op.propagate(torch.from_numpy(x).cuda(), K, mode="exact", concat=True)
    # the K+1 slabs are column blocks of one [n, (K+1) d] buffer == ConcatMessageOp(0, K+1) of the hop list
    base = hops[0]
    assert all(h.stride(0) == (K + 1) * d for h in hops)
    for k in range(K + 1):
        assert np.array_equal(hops[k].cpu().numpy(), ref[k])
    whole = torch.as_strided(base, (n, (K + 1) * d), ((K + 1) * d, 1)).cpu().numpy()
    assert np.array_equal(whole, O.combine_concat(ref, 0, K + 1))
    op.close()


def test_accumulate_matches_reference_answer_semantics():
    """matmul.c:36-37 accumulates into the caller's buffer: each chain starts from the value already there."""
    rng = np.random.default_rng(11)
    n, d = 256, 100
    adj = random_graph(rng, n, 3000, weights=True)
    a = O.laplacian_adj(adj, 0.3)
    x = rng.standard_normal((n, d)).astype(np.float32)
    y0 = rng.standard_normal((n, d)).astype(np.float32)
    ref = y0.copy()
    O._lib().oracle_spmm_f32_fma_i64(ref, a.data.astype(np.float32), a.indices, a.indptr, x, n, d)
    op = CsrOperator(a.indptr, a.indices, a.data.astype(np.float32), a.shape)
    y = torch.from_numpy(y0.copy()).cuda()
    op.spmm(torch.from_numpy(x).cuda(), out=y, accumulate=True)
    assert np.array_equal(y.cpu().numpy(), ref)
    op.close()


def test_legacy_abi_symbols_bit_exact():
    """The reference's two C entry points, same signatures, host pointers (matmul.h:5, cudamatmul.c:28)."""
    rng = np.random.default_rng(13)
    n, d = 500, 96
    adj = random_graph(rng, n, 6000, weights=True)
    a = O.laplacian_adj(adj, 0.5)
    data = a.data.astype(np.float32)
    indptr32 = a.indptr.astype(np.int32)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.spmm_hop(a, x, "fma")
    lib = _lib.load()
    ans = np.zeros(n * d, dtype=np.float32)
    lib.FloatCSRMulDenseOMP(ans.ctypes.data, data.ctypes.data, a.indices.ctypes.data, indptr32.ctypes.data,
                            x.reshape(-1).ctypes.data, n, d)
    assert np.array_equal(ans.reshape(n, d), ref)
    ans2 = np.full(n * d, 7.0, dtype=np.float32)          # the cuSPARSE-style entry overwrites (beta = 0)
    rc = lib.FloatCSRMulDense(ans2.ctypes.data, int(a.nnz), data.ctypes.data, a.indices.ctypes.data,
                              indptr32.ctypes.data, x.reshape(-1).ctypes.data, n, d)
    assert rc == 0 and np.array_equal(ans2.reshape(n, d), ref)
    ref_lib = O.load_reference_kernel()
    if ref_lib is not None:                                  # the reference's own kernel, compiled from its source
        assert np.array_equal(O.reference_kernel_hop(ref_lib, a, x), ans.reshape(n, d))


def test_propagate_host_keep_last_and_torch_feature():
    rng = np.random.default_rng(17)
    n, d, K = 1200, 128, 5
    adj = random_graph(rng, n, 9000)
    a = O.laplacian_adj(adj, 0.5)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.propagate(a, x, K, "fma")
    op = CsrOperator(a.indptr, a.indices, a.data.astype(np.float32), a.shape)
    outs = op.propagate_host(x, K, mode="exact", keep="last")
    assert outs[:-1] == [None] * (K - 1) and np.array_equal(outs[-1].numpy(), ref[K])
    outs = op.propagate_host(torch.from_numpy(x), K, mode="exact", keep="all")
    for k in range(1, K + 1):
        assert np.array_equal(outs[k - 1].numpy(), ref[k])
    op.close()
    # dataset.x is a torch.FloatTensor in the reference's own callers: accepted
    g = LaplacianGraphOp(K)
    g.mode = "exact"
    hops = g.propagate(adj, torch.from_numpy(x))
    assert np.array_equal(hops[K].numpy(), ref[K])


def test_cuda_resident_outputs_and_gather():
    rng = np.random.default_rng(19)
    n, d, K = 900, 100, 3
    adj = random_graph(rng, n, 7000)
    x = rng.standard_normal((n, d)).astype(np.float32)
    g = LaplacianGraphOp(K)
    g.mode = "exact"
    g.output_device = "cuda"
    hops = g.propagate(adj, x)
    ref = O.propagate(O.laplacian_adj(adj, 0.5), x, K, "fma")
    assert all(h.is_cuda for h in hops)
    for k in range(K + 1):
        assert np.array_equal(hops[k].cpu().numpy(), ref[k])
    idx = torch.from_numpy(rng.integers(0, n, 257))
    outs = gather_rows(hops, idx)
    for k in range(K + 1):
        assert np.array_equal(outs[k].cpu().numpy(), ref[k][idx.numpy()])


def test_nafs_weights_against_oracle_larger():
    rng = np.random.default_rng(23)
    n, d, K = 3000, 128, 6
    adj = random_graph(rng, n, 20000)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.propagate(O.laplacian_adj(adj, 0.5), x, K, "fma")
    out = aggregate(_lib.AGG_OSD, [torch.from_numpy(r).cuda() for r in ref]).cpu().numpy()
    np.testing.assert_allclose(out, O.combine_osd(ref), rtol=3e-6, atol=3e-6)


# --------------------------------------------------------------------------------------------------------------
# BASELINE.json sizes: size-independent properties (the oracle on a row sample, linearity, constant vectors)
# --------------------------------------------------------------------------------------------------------------
def _rmat_like(rng, n, m):
    rows = (rng.zipf(1.25, m) - 1) % n
    cols = rng.integers(0, n, m)
    r2, c2 = np.concatenate([rows, cols]), np.concatenate([cols, rows])
    return sp.csr_matrix((np.ones(r2.size, dtype=np.float32), (r2, c2)), shape=(n, n))


def test_arxiv_shape_properties():
    """configs[1] shape: N=169,343, ~2.3M undirected entries, d=128, K=5."""
    rng = np.random.default_rng(29)
    n, d, K = 169_343, 128, 5
    adj = _rmat_like(rng, n, 1_166_243)
    op = LaplacianGraphOp(K, r=0.5)
    x = rng.standard_normal((n, d)).astype(np.float32)
    hops = op.propagate(adj, x)
    a = op._adj
    a32 = a.astype(np.float32)
    # (1) oracle on a row sample of every hop: y_k[i] from y_{k-1} (one-hop check isolates the kernel)
    sample = rng.choice(n, 4000, replace=False)
    sub = a32[sample]
    sub_o = O.Csr(sub.indptr, sub.indices, sub.data, sub.shape)
    for k in range(1, K + 1):
        ref = np.zeros((sample.size, d), dtype=np.float32)
        O._lib().oracle_spmm_f32_fma_i64(ref, sub_o.data, sub_o.indices, sub_o.indptr, hops[k - 1].numpy(),
                                         sample.size, d)
        assert_close_1e5(hops[k].numpy()[sample], ref)
    # (2) r = 0 normalisation is row-stochastic: constant vectors are fixed points of every hop
    op0 = LaplacianGraphOp(3, r=0.0)
    ones = np.ones((n, 4), dtype=np.float32)
    for h in op0.propagate(adj, ones):
        np.testing.assert_allclose(h.numpy(), 1.0, rtol=0, atol=3e-5)   # deg terms of fp32 rounding on hub rows
    # (3) linearity: A(2x + y) == 2 Ax + Ay within fp32 rounding
    csr = CsrOperator.from_scipy(a)
    xd = torch.from_numpy(x).cuda()
    yd = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32)).cuda()
    lhs = csr.spmm(2 * xd + yd)
    rhs = 2 * csr.spmm(xd) + csr.spmm(yd)
    assert_close_1e5(lhs.cpu().numpy(), rhs.cpu().numpy())
    # (4) fast and exact schedules agree to tolerance on the full-size graph, exact is reproducible
    e1 = csr.spmm(xd, mode="exact")
    e2 = csr.spmm(xd, mode="exact")
    assert torch.equal(e1, e2)
    assert_close_1e5(csr.spmm(xd, mode="fast").cpu().numpy(), e1.cpu().numpy())
    csr.close()


# --------------------------------------------------------------------------------------------------------------
# f2: A^ built on the device; e: the row-partitioned operator on CUDA
# --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", golden_graph_files(), ids=lambda p: os.path.basename(p)[6:-4])
@pytest.mark.parametrize("tag,kind,r,alpha", GRAPH_TAGS)
def test_device_built_adjacency_matches_goldens(path, tag, kind, r, alpha):
    z, adj = load_graph(path)
    K = z[tag + "_hops_fma"].shape[0] - 1
    op = LaplacianGraphOp(K, r=r) if kind == "lap" else PprGraphOp(K, r=r, alpha=alpha)
    op.mode, op.build_on = "exact", "device"
    hops = op.propagate(_scipy_adj(adj), z["x"].copy())
    got = np.stack([h.numpy() for h in hops])
    a = op._adj                                                          # lazily downloaded scipy view
    assert np.array_equal(a.indptr, z[tag + "_norm_indptr"]) and np.array_equal(a.indices, z[tag + "_norm_indices"])
    if "weighted" in path:
        assert np.array_equal(a.data.astype(np.float32), z[tag + "_norm_data"].astype(np.float32))
    else:
        assert np.array_equal(a.data, z[tag + "_norm_data"])
    assert np.array_equal(got, z[tag + "_hops_fma"])


def test_device_builder_larger_random_vs_oracle():
    from sgl_b200.graph_build import operator_from_scipy_device, values_from_parts
    rng = np.random.default_rng(31)
    n = 5000
    adj = random_graph(rng, n, 60000, skew=1.2)
    for (r, alpha) in [(0.5, None), (0.3, 0.15)]:
        ref = O.laplacian_adj(adj, r) if alpha is None else O.ppr_adj(adj, r, alpha)
        op = operator_from_scipy_device(adj, r=r, alpha=alpha)
        assert np.array_equal(op.parts["indptr"].cpu().numpy(), ref.indptr)
        assert np.array_equal(op.parts["indices"].cpu().numpy(), ref.indices)
        assert np.array_equal(values_from_parts(op.parts).cpu().numpy(), ref.data)
        x = rng.standard_normal((n, 64)).astype(np.float32)
        y = op.spmm(torch.from_numpy(x).cuda(), mode="exact").cpu().numpy()     # values written by the CUDA value pass
        assert np.array_equal(y, O.spmm_hop(ref, x, "fma"))
        op.close()


def _scipy_structure(rows, cols, w, n):
    """What the reference computes on the host (utils.py:76-78,87): A canonical, A + I, row sums, transpose in CSR order."""
    import scipy.sparse as sp
    a = sp.csr_matrix((w, (rows, cols)), shape=(n, n))
    a.sum_duplicates()
    at = a + sp.eye(n)
    deg = np.asarray(at.sum(1)).reshape(-1)
    t = sp.csr_matrix(at.T)
    t.sort_indices()
    return t.indptr.astype(np.int64), t.indices.astype(np.int32), t.data.astype(np.float64), deg


@pytest.mark.parametrize("n,e,seed", [(1, 0, 0), (7, 0, 1), (300, 4000, 2), (5000, 70000, 3), (70001, 900000, 4)])
def test_native_adjacency_builder_vs_scipy(n, e, seed):
    """sglb200_adjacency_build (own radix sort / scan / fold kernels): unsorted COO with duplicates, self loops, a
    diagonal entry of -1 (A + I makes an exact zero, which scipy drops), asymmetric, integer weights 1..3 so that every
    float32 duplicate sum is exact whatever the order.  Structure, merged weights and degrees equal scipy's bit for bit;
    the torch engine (first version) agrees too."""
    from sgl_b200.graph_build import adjacency_structure_native, normalized_adjacency_device
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, n, e).astype(np.int64)
    cols = ((rng.zipf(1.3, e) - 1) % n).astype(np.int64)
    w = rng.integers(1, 4, e).astype(np.float32)
    if e:
        k = e // 10
        rows[:k], cols[:k] = rows[e - k:], cols[e - k:]          # duplicates far apart in the input
        rows[k:2 * k] = cols[k:2 * k]                            # self loops
        sel = rows == cols                                       # one diagonal entry summing to -1: dropped by A + I
        if sel.any():
            v = int(rows[sel][0])
            hit = (rows == v) & (cols == v)
            w[hit] = 0.0
            w[np.flatnonzero(hit)[0]] = -1.0
    for weights in ((w, None) if e else (None,)):
        ww = np.ones(e, dtype=np.float32) if weights is None else weights
        ref_ptr, ref_idx, ref_w, ref_deg = _scipy_structure(rows, cols, ww, n)
        rt, ct = torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda()
        wt = None if weights is None else torch.from_numpy(weights).cuda()
        indptr, indices, raw_w, deg = adjacency_structure_native(rt, ct, n, wt)
        assert np.array_equal(indptr.cpu().numpy(), ref_ptr)
        assert np.array_equal(indices.cpu().numpy(), ref_idx)
        assert np.array_equal(raw_w.cpu().numpy(), ref_w)
        assert np.array_equal(deg.cpu().numpy(), ref_deg)
        if e:
            parts_t = normalized_adjacency_device(rt, ct, n, wt, engine="torch")
            parts_n = normalized_adjacency_device(rt, ct, n, wt, engine="native")
            for key in ("indptr", "indices", "raw_w", "deg", "d_left", "d_right"):
                assert torch.equal(parts_t[key], parts_n[key]), key


def test_native_adjacency_builder_float_weights_input_order():
    """Non-integer weights: duplicates are added in float32 in INPUT order (the stable sort keeps it), then widened."""
    from sgl_b200.graph_build import adjacency_structure_native
    rng = np.random.default_rng(9)
    n, e = 50, 3000
    rows, cols = rng.integers(0, n, e).astype(np.int64), rng.integers(0, n, e).astype(np.int64)
    w = rng.standard_normal(e).astype(np.float32)
    acc = {}
    for r, c, v in zip(rows.tolist(), cols.tolist(), w):
        acc[(r, c)] = np.float32(acc[(r, c)] + v) if (r, c) in acc else np.float32(v)
    dense = np.zeros((n, n), dtype=np.float64)
    for (r, c), v in acc.items():
        dense[r, c] = float(v)
    dense += np.eye(n)
    indptr, indices, raw_w, deg = adjacency_structure_native(torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(), n,
                                                             torch.from_numpy(w).cuda())
    indptr, indices, raw_w = indptr.cpu().numpy(), indices.cpu().numpy(), raw_w.cpu().numpy()
    got = np.zeros((n, n), dtype=np.float64)
    for i in range(n):
        js = indices[indptr[i]:indptr[i + 1]]
        assert np.all(np.diff(js) > 0)
        got[js, i] = raw_w[indptr[i]:indptr[i + 1]]                 # entry (i, j) of the transpose holds A~[j, i]
    assert np.array_equal(got, dense)
    want_deg = np.zeros(n, dtype=np.float64)
    for i in range(n):
        acc_i = 0.0                                                 # one sequential float64 chain in column order
        for j in range(n):
            if dense[i, j] != 0.0:
                acc_i = acc_i + float(dense[i, j])
        want_deg[i] = acc_i
    assert np.array_equal(deg.cpu().numpy(), want_deg)


def test_native_adjacency_builder_rejects_bad_edges():
    from sgl_b200 import SglB200Error
    from sgl_b200.graph_build import adjacency_structure_native
    rows = torch.tensor([0, 5], dtype=torch.int64, device="cuda")
    cols = torch.tensor([1, 2], dtype=torch.int64, device="cuda")
    with pytest.raises(SglB200Error, match="outside"):
        adjacency_structure_native(rows, cols, 4, None)


def test_trace_mode_prints_per_hop_timers():
    """SGLB200_TRACE=1: CUDA-event timers around every hop of sglb200_propagate, one stderr line per hop (the NVTX ranges
    around the entry points need a profiler to be seen; this is the part a test can check)."""
    import subprocess
    import sys
    code = ("import numpy as np, scipy.sparse as sp, torch\n"
            "from sgl_b200.runtime import CsrOperator\n"
            "a = sp.random(500, 500, 0.02, format='csr', dtype=np.float32, random_state=1)\n"
            "op = CsrOperator.from_scipy(a)\n"
            "h = op.propagate(torch.ones(500, 8, device='cuda'), 3)\n"
            "torch.cuda.synchronize(); print(float(h[-1].sum()))\n")
    env = dict(os.environ, SGLB200_TRACE="1", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stderr.splitlines() if ln.startswith("sglb200 trace: propagate hop")]
    assert len(lines) == 3 and "sglb200 trace: propagate 3 hops" in r.stderr


def test_row_partition_single_rank_on_cuda():
    """world = 1 degenerates to the single-GPU operator; exercises DistOperator's CUDA path without NCCL."""
    from sgl_b200.dist import DistOperator, build_plan
    rng = np.random.default_rng(37)
    n, d, K = 2000, 128, 3
    a = O.laplacian_adj(random_graph(rng, n, 15000), 0.5)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.propagate(a, x, K, "fma")
    for mode in ("halo", "allgather"):
        plan = build_plan(a.indptr, a.indices, a.data, n, 1, 0, mode)
        op = DistOperator(plan, mode="exact")
        hops = op.propagate(torch.from_numpy(x).cuda(), K)
        for k in range(K + 1):
            assert np.array_equal(hops[k].cpu().numpy(), ref[k])
        op.close()


# --------------------------------------------------------------------------------------------------------------
# a14: the SGAP model glue (preprocess / forward contract) on top of the path
# --------------------------------------------------------------------------------------------------------------
def test_sgap_models_preprocess_and_forward():
    from sgl_b200 import sgap
    rng = np.random.default_rng(41)
    n, d, K = 1500, 32, 3
    adj = random_graph(rng, n, 12000)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.propagate(O.laplacian_adj(adj, 0.5), x, K, "fma")
    cases = [(sgap.SGC(K, d, 4).cuda(), ref[K]), (sgap.SSGC(K, d, 4).cuda(), O.combine_mean(ref, 0, K + 1)),
             (sgap.GBP(K, d, 4, 16, 2).cuda(),
              O.combine_weighted(ref, O.alpha_weights(K + 1, 0.85, 0, K + 1), 0, K + 1)),
             (sgap.SIGN(K, d, 4, 16, 2).cuda(), O.combine_concat(ref, 0, K + 1))]
    for fused, threshold in ((True, 0.5), (True, 0.0), (False, 0.5)):   # 0.0: always take the memory-lean fused aggregates
        for model, want in cases:
            model.fused_preprocess = fused
            model.fuse_when_slabs_exceed = threshold
            model._pre_graph_op.mode = "exact"
            model.preprocess(adj, x)
            feat = model._processed_feature
            assert isinstance(feat, torch.Tensor) and not feat.is_cuda and model._pre_msg_learnable is False
            assert np.array_equal(feat.numpy(), want), type(model).__name__
            out = model.model_forward(range(10, 50), "cuda")
            assert out.shape[0] == 40 and out.is_cuda
    nafs = sgap.NAFS(K, d, d)
    nafs._pre_graph_op.mode = "exact"
    nafs.preprocess(adj, x)
    np.testing.assert_allclose(nafs._processed_feature.numpy(), O.combine_osd(ref), rtol=3e-6, atol=3e-6)
    # learnable aggregator: slabs stay resident on the device, forward gathers there
    gamlp = sgap.GAMLP(K, d, 5, 16, 2).cuda()
    gamlp._pre_graph_op.mode = "exact"
    gamlp.preprocess(adj, x)
    assert gamlp._pre_msg_learnable and all(h.is_cuda for h in gamlp._processed_feat_list)
    for k in range(K + 1):
        assert np.array_equal(gamlp._processed_feat_list[k].cpu().numpy(), ref[k])
    idx = torch.arange(0, 300, 3)
    out = gamlp.model_forward(idx, "cuda")
    assert out.shape == (100, 5)
    out.sum().backward()
    # prepared adjacency is reused for the same object and rebuilt for another
    g = gamlp._pre_graph_op
    handle = g._operator
    gamlp.preprocess(adj, x)
    assert g._operator is handle
    gamlp.preprocess(adj.copy(), x)
    assert g._operator is not handle
    # postprocess path (softmax -> propagate -> aggregate) with a post graph op
    from sgl_b200.operators.graph_op import PprGraphOp
    from sgl_b200.operators.message_op import LastMessageOp
    sgc = sgap.SGC(K, d, 4)
    sgc._post_graph_op, sgc._post_msg_op = PprGraphOp(2, r=0.5, alpha=0.3), LastMessageOp()
    sgc._post_graph_op.mode = "exact"
    logits = torch.from_numpy(rng.standard_normal((n, 4)).astype(np.float32))
    got = sgc.postprocess(adj, logits)
    soft = torch.softmax(logits, dim=1).numpy()
    want = O.propagate(O.ppr_adj(adj, 0.5, 0.3), soft, 2, "fma")[-1]
    assert np.array_equal(got.numpy(), want)


# --------------------------------------------------------------------------------------------------------------
# a11: fused LearnableWeightedMessageOp kernels (forward + autograd) against the reference goldens
# --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["gate", "ori_ref", "jk"])
@pytest.mark.parametrize("se", [(0, 5), (1, 4)])
def test_fused_learnable_weighted_op(message_golden, kind, se):
    from sgl_b200.operators.message_op import LearnableWeightedMessageOp
    g = message_golden
    s, e = se
    K, d = 4, 16
    batch = [torch.from_numpy(h[g["batch_idx"]]).cuda() for h in g["hops"]]
    args = {"gate": (d,), "ori_ref": (d,), "jk": (K, d)}[kind]
    tag = f"lw_{kind}_{s}_{e}"
    results = {}
    for fused in (True, False):
        op = LearnableWeightedMessageOp(s, e, kind, *args).cuda()
        op.fused = fused
        with torch.no_grad():
            op._learnable_weight.weight.copy_(torch.from_numpy(g[tag + "_w"]))
            op._learnable_weight.bias.copy_(torch.from_numpy(g[tag + "_b"]))
        feats = [b.clone().requires_grad_(True) for b in batch]
        out = op.aggregate(feats)
        gout = torch.linspace(-1, 1, out.numel(), device="cuda").reshape(out.shape)
        out.backward(gout)
        gin = np.stack([f.grad.cpu().numpy() if f.grad is not None else np.zeros((len(g["batch_idx"]), d), np.float32)
                        for f in feats])
        results[fused] = (out.detach().cpu().numpy(), gin, op._learnable_weight.weight.grad.cpu().numpy(),
                          op._learnable_weight.bias.grad.cpu().numpy())
    for fused, (out, gin, gw, gb) in results.items():
        np.testing.assert_allclose(out, g[tag + "_out"], rtol=1e-5, atol=2e-6, err_msg=f"fused={fused}")
        np.testing.assert_allclose(gin, g[tag + "_gin"], rtol=1e-4, atol=1e-5, err_msg=f"fused={fused}")
        np.testing.assert_allclose(gw, g[tag + "_gw"], rtol=1e-4, atol=2e-5, err_msg=f"fused={fused}")
        np.testing.assert_allclose(gb, g[tag + "_gb"], rtol=1e-4, atol=2e-5, err_msg=f"fused={fused}")


def test_fused_learnable_larger_batch_matches_torch_expressions():
    from sgl_b200.operators.message_op import LearnableWeightedMessageOp
    torch.manual_seed(3)
    K, d, B = 6, 128, 5000
    feats = [torch.randn(B, d, device="cuda") for _ in range(K + 1)]
    for kind, args in [("gate", (d,)), ("ori_ref", (d,)), ("jk", (K, d))]:
        op = LearnableWeightedMessageOp(0, K + 1, kind, *args).cuda()
        outs, grads = [], []
        for fused in (True, False):
            op.fused = fused
            op.zero_grad()
            fs = [f.clone().requires_grad_(True) for f in feats]
            out = op.aggregate(fs)
            out.square().mean().backward()
            outs.append(out.detach())
            grads.append((torch.stack([f.grad for f in fs]), op._learnable_weight.weight.grad.clone(),
                          op._learnable_weight.bias.grad.clone()))
        assert torch.allclose(outs[0], outs[1], rtol=1e-5, atol=1e-6), kind
        for a, b in zip(grads[0], grads[1]):
            assert torch.allclose(a, b, rtol=2e-4, atol=1e-7), kind


def test_tile_ranges_compose_to_the_full_hop():
    """sglb200_spmm_tiles over consecutive chunks == sglb200_spmm, bit for bit (same tiles, same folds), and the device
    chunk bounds equal the host formula the row partitioner uses."""
    from sgl_b200.dist import chunk_row_bounds
    rng = np.random.default_rng(43)
    n, d = 6000, 128
    a = O.laplacian_adj(random_graph(rng, n, 90000, skew=1.15), 0.5)
    op = CsrOperator(a.indptr, a.indices, a.data.astype(np.float32), a.shape, tile_items=256)
    assert op.info()["carry_runs"] > 0
    x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32)).cuda()
    full = op.spmm(x, mode="fast")
    for n_chunks in (1, 3, 7):
        tiles, rows = op.chunks(n_chunks, "fast")
        assert rows == [int(v) for v in chunk_row_bounds(a.indptr, n_chunks, 256)]
        out = torch.full_like(full, float("nan"))
        for c in range(n_chunks):
            op.spmm_tiles(x, out, tiles[c], tiles[c + 1], mode="fast")
            torch.cuda.synchronize()
            done = out[:rows[c + 1]]
            assert torch.equal(done, full[:rows[c + 1]])       # rows finished so far are final
        assert torch.equal(out, full)
    op.close()


# --------------------------------------------------------------------------------------------------------------
# 8f-3: label propagation and the NAFS task-level feature construction on the same handle
# --------------------------------------------------------------------------------------------------------------
def test_label_propagation_and_nafs_features():
    from sgl_b200.operators.utils import adj_to_symmetric_norm
    from sgl_b200.tricks import label_propagation, nafs_smoothed_features
    rng = np.random.default_rng(47)
    n, classes = 3000, 7
    adj = random_graph(rng, n, 25000)
    y = torch.from_numpy(rng.integers(0, classes, n))
    mask = torch.from_numpy(rng.random(n) < 0.3)
    norm = adj_to_symmetric_norm(adj, 0.5)
    got = label_propagation(y, norm, num_layers=5, alpha=0.8, mask=mask)
    onehot = np.eye(classes, dtype=np.float32)[y.numpy()]
    want = O.label_propagation(onehot, O.laplacian_adj(adj, 0.5), 5, 0.8, mask=mask.numpy())
    assert got.shape == (n, classes) and not got.is_cuda
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-5, atol=1e-6)
    x = rng.standard_normal((n, 64)).astype(np.float32)
    for method in ("mean", "max", "concat", "simple"):
        got = nafs_smoothed_features(adj, x, hops=3, r_list=(0.5, 0.3, 0.0), method=method).numpy()
        want = O.nafs_smoothed_features(adj, x, 3, (0.5, 0.3, 0.0), method)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=4e-6, err_msg=method)
    # the hop sweep of the NAFS tasks: every hop computed once per r, equal to the per-hop-count calls
    from sgl_b200.tricks import nafs_smoothed_features_sweep
    sweep = nafs_smoothed_features_sweep(adj, x, 3, r_list=(0.5, 0.0), method="mean")
    for h in (1, 2, 3):
        single = nafs_smoothed_features(adj, x, hops=h, r_list=(0.5, 0.0), method="mean")
        assert torch.equal(sweep[h - 1], single), h
    # a user post_process keeps the unfused layer sequence; the default clamp takes the fused hop (same numbers)
    plain = label_propagation(y, norm, num_layers=5, alpha=0.8, mask=mask, post_process=lambda t: t.clamp_(0., 1.))
    want_lp = O.label_propagation(onehot, O.laplacian_adj(adj, 0.5), 5, 0.8, mask=mask.numpy())
    np.testing.assert_allclose(plain.numpy(), want_lp, rtol=1e-5, atol=1e-6)


# --------------------------------------------------------------------------------------------------------------
# g1: fused K-hop driver -- degree normalisation and cross-hop aggregation inside the hop kernel's row flush
# --------------------------------------------------------------------------------------------------------------
def _fused_reference(ref, agg, start, end, weights):
    if agg == "sum":
        return O.combine_sum(ref, start, end)
    if agg == "mean":
        return O.combine_mean(ref, start, end)
    if agg == "max":
        return O.combine_max(ref, start, end)
    if agg == "min":
        return O.combine_min(ref, start, end)
    if agg == "concat":
        return O.combine_concat(ref, start, end)
    if agg == "weighted":
        return O.combine_weighted(ref, np.asarray(weights[start:end], dtype=np.float32), start, end)
    if agg == "last":
        return ref[-1]
    return O.combine_osd(ref)


@pytest.mark.parametrize("d", [32, 100, 128])
@pytest.mark.parametrize("use_tma", [False, True])
def test_fused_driver_exact_mode_is_bit_exact(monkeypatch, d, use_tma):
    """sglb200_propagate_fused, EXACT mode: hops and sum / mean / max / min / weighted / concat / last aggregates equal the
    oracle's separate propagate + combine bit for bit (the running update keeps the reference's left-to-right order);
    NAFS weights within 3e-6.  d=32: lane-group kernel, d=100/128: warp kernel or the TMA-staged kernel."""
    if use_tma:
        monkeypatch.setenv("SGLB200_TMA", "1")
    rng = np.random.default_rng(500 + d)
    n, K = 1200, 4
    adj = random_graph(rng, n, 15000, skew=1.25, weights=True)
    a = O.laplacian_adj(adj, 0.5)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.propagate(a, x, K, "fma")
    op = CsrOperator(a.indptr, a.indices, a.data.astype(np.float32), a.shape)
    xd = torch.from_numpy(x).cuda()
    wts = [0.5, 0.25, -0.125, 2.0, 0.3]
    for agg, start, end in [("sum", 0, K + 1), ("mean", 1, K), ("max", 0, K + 1), ("min", 2, K + 1), ("weighted", 0, K + 1),
                            ("weighted", 1, 4), ("concat", 0, K + 1), ("concat", 1, 3), ("last", 0, K + 1), ("osd", 0, K + 1)]:
        hops, out = op.propagate_fused(xd, K, mode="exact", keep="none", agg=agg, start=start, end=end,
                                       weights=wts if agg == "weighted" else None)
        want = _fused_reference(ref, agg, start, end, wts)
        if agg == "osd":
            np.testing.assert_allclose(out.cpu().numpy(), want, rtol=3e-6, atol=3e-6)
        else:
            assert np.array_equal(out.cpu().numpy(), want), (agg, start, end)
        assert all(h is None for h in hops[1:])
    hops, out = op.propagate_fused(xd, K, mode="exact", keep="all", agg="mean")
    for k in range(K + 1):
        assert np.array_equal(hops[k].cpu().numpy(), ref[k])
    assert np.array_equal(out.cpu().numpy(), O.combine_mean(ref, 0, K + 1))
    hops, out = op.propagate_fused(xd, K, mode="exact", keep="last")
    assert out is None and np.array_equal(hops[K].cpu().numpy(), ref[K]) and hops[1] is None
    op.close()


@pytest.mark.parametrize("kind,r,alpha", [("lap", 0.5, None), ("lap", 0.0, None), ("ppr", 0.5, 0.15)])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("d", [16, 100, 128])
def test_fused_normalisation_fast_mode(kind, r, alpha, weighted, d):
    """FAST mode on a device-built operator: the hop streams raw weights (none when they are all 1) and applies
    deg^(r-1), deg^(-r) and the PPR teleport term in the row flush.  Every hop within 1e-5 of the oracle (Frobenius and
    max-abs relative), cut rows included; mean aggregate within 1e-5; equal to itself run to run."""
    from sgl_b200.graph_build import operator_from_scipy_device
    rng = np.random.default_rng(900 + d)
    n, K = 2500, 4
    rows = rng.integers(0, n, 30000)
    cols = (rng.zipf(1.2, 30000) - 1) % n
    if weighted:
        adj = sp.csr_matrix((np.ones(60000, dtype=np.float32), (np.concatenate([rows, cols]), np.concatenate([cols, rows]))), shape=(n, n))
    else:
        adj = sp.csr_matrix((np.ones(60000, dtype=np.float32), (np.concatenate([rows, cols]), np.concatenate([cols, rows]))), shape=(n, n))
        adj.data[:] = 1.0       # duplicates merged to weight 1: an unweighted graph
    a = O.laplacian_adj(adj, r) if kind == "lap" else O.ppr_adj(adj, r, alpha)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.propagate(a, x, K, "fma")
    op = operator_from_scipy_device(adj, r=r, alpha=alpha, tile_items=64, split_threshold=24)
    assert op.info()["carry_runs"] > 0
    xd = torch.from_numpy(x).cuda()
    hops, out = op.propagate_fused(xd, K, mode="fast", keep="all", agg="mean", fuse_norm=True)
    hops2, out2 = op.propagate_fused(xd, K, mode="fast", keep="all", agg="mean", fuse_norm=True)
    for k in range(1, K + 1):
        assert_close_1e5(hops[k].cpu().numpy(), ref[k])
        assert torch.equal(hops[k], hops2[k])
    assert_close_1e5(out.cpu().numpy(), O.combine_mean(ref, 0, K + 1))
    assert torch.equal(out, out2)
    # without keeping the hops (internal ping-pong slabs only) the aggregate is the same bits
    _, out3 = op.propagate_fused(xd, K, mode="fast", keep="none", agg="mean", fuse_norm=True)
    assert torch.equal(out, out3)
    # fuse_norm=False streams the materialised float32 values: same tolerance, and EXACT stays bit-exact
    hops4, _ = op.propagate_fused(xd, K, mode="fast", keep="all", fuse_norm=False)
    assert_close_1e5(hops4[K].cpu().numpy(), ref[K])
    # running mean on the LEAN flush (L2 reductions) with cut rows: the folded row, not its pieces, reaches the aggregate --
    # bit-equal to the left-to-right mean of the very hops this mode produces
    _, out4 = op.propagate_fused(xd, K, mode="fast", keep="none", agg="mean", fuse_norm=False)
    want4 = hops4[0].clone() + 0.0
    for k in range(1, K + 1):
        want4 = want4 + hops4[k]
    assert torch.equal(out4.cpu(), want4.cpu() / (K + 1))    # on the CPU: torch's CUDA division by a scalar multiplies by 1/b
    _, out5 = op.propagate_fused(xd, K, mode="fast", keep="none", agg="weighted", start=1, end=K, fuse_norm=False,
                                 weights=[0.0, 0.5, 0.25, 2.0, 0.0])
    want5 = hops4[1] * 0.5
    want5 = want5 + hops4[2] * 0.25
    want5 = want5 + hops4[3] * 2.0
    assert torch.equal(out5, want5)
    hops5, _ = op.propagate_fused(xd, K, mode="exact", keep="all")
    assert np.array_equal(hops5[K].cpu().numpy(), ref[K])
    op.close()


def test_message_ops_keep_the_autograd_graph():
    """combiners fed with tensors that require grad must stay differentiable (reference _combine is torch)."""
    feats = [torch.randn(50, 8, device="cuda", requires_grad=True) for _ in range(4)]
    for op in (SumMessageOp(0, 4), MeanMessageOp(1, 3), MaxMessageOp(0, 4), MinMessageOp(0, 4), ConcatMessageOp(0, 3),
               SimpleWeightedMessageOp(0, 4, "alpha", 0.85), OverSmoothDistanceWeightedOp()):
        out = op.aggregate(feats)
        assert out.requires_grad, type(op).__name__
        out.sum().backward()
    assert all(f.grad is not None for f in feats)
    with torch.no_grad():
        assert not SumMessageOp(0, 4).aggregate(feats).requires_grad


# --------------------------------------------------------------------------------------------------------------
# a12: IterateLearnableWeightedMessageOp / ProjectedConcatMessageOp fused kernels vs the reference's torch expressions
# --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("se", [(0, 5), (0, 3), (0, 1)])
@pytest.mark.parametrize("d", [16, 100])
def test_fused_iterate_learnable_forward_and_gradients(se, d):
    """csrc/iterate.cu against the torch expression restating iterate_learnable_weighted_message_op.py:28-51 (itself pinned
    to the reference's goldens on the CPU, tests/test_host_logic.py): outputs 1e-5, gradients 1e-4 (relative)."""
    from sgl_b200.operators.message_op import IterateLearnableWeightedMessageOp
    torch.manual_seed(3 + d)
    B, n_all = 700, 5
    s, e = se
    op = IterateLearnableWeightedMessageOp(s, e, "recursive", d).cuda()
    with torch.no_grad():
        op._learnable_weight.weight.mul_(3.0)       # gates away from 0.5 so that the recursion matters
    feats_a = [torch.randn(B, d, device="cuda", requires_grad=True) for _ in range(n_all)]
    feats_b = [f.detach().clone().requires_grad_(True) for f in feats_a]
    g = torch.randn(B, d, device="cuda")
    op.fused = True
    out_a = op.aggregate(feats_a)
    (out_a * g).sum().backward()
    gw_a, gb_a = op._learnable_weight.weight.grad.clone(), op._learnable_weight.bias.grad.clone()
    op.zero_grad()
    op.fused = False
    out_b = op.aggregate(feats_b)
    (out_b * g).sum().backward()
    gw_b, gb_b = op._learnable_weight.weight.grad, op._learnable_weight.bias.grad

    def close(x, y, tol):
        scale = max(float(y.abs().max()), 1e-12)
        assert float((x - y).abs().max()) <= tol * scale, (float((x - y).abs().max()), scale)

    close(out_a, out_b, 1e-5)
    for k in range(n_all):
        if s <= k < e:
            close(feats_a[k].grad, feats_b[k].grad, 1e-4)
        else:
            assert feats_a[k].grad is None and feats_b[k].grad is None
    close(gw_a, gw_b, 1e-4)
    close(gb_a, gb_b, 1e-4)
    with pytest.raises(IndexError):      # start > 0 fails in the reference (absolute hop index into the weight matrix)
        IterateLearnableWeightedMessageOp(1, 4, "recursive", d).cuda().aggregate(feats_a)


def test_fused_projected_concat_matches_torch():
    from sgl_b200.operators.message_op import ProjectedConcatMessageOp
    torch.manual_seed(11)
    B, d, h = 300, 24, 12
    op = ProjectedConcatMessageOp(0, 4, d, h, 2).cuda().eval()    # eval: dropout off, both paths deterministic
    feats_a = [torch.randn(B, d, device="cuda", requires_grad=True) for _ in range(4)]
    feats_b = [f.detach().clone().requires_grad_(True) for f in feats_a]
    g = torch.randn(B, 4 * h, device="cuda")
    op.fused = True
    out_a = op.aggregate(feats_a)
    (out_a * g).sum().backward()
    grads_a = [p.grad.clone() for p in op.parameters()]
    op.zero_grad()
    op.fused = False
    out_b = op.aggregate(feats_b)
    (out_b * g).sum().backward()
    assert torch.equal(out_a, out_b)
    for fa, fb in zip(feats_a, feats_b):
        assert torch.allclose(fa.grad, fb.grad, rtol=1e-5, atol=1e-6)
    for ga, pb in zip(grads_a, op.parameters()):
        assert torch.allclose(ga, pb.grad, rtol=1e-5, atol=1e-6)


def test_hop_cache_miss_compute_save_hit_on_gpu(tmp_path):
    """f4: GraphOp.cache_dir on hardware -- a miss runs the K hops on the GPU and stores them, a hit returns the same bits
    without touching the operator; a different graph op parameter misses again."""
    rng = np.random.default_rng(61)
    n, d, K = 800, 24, 3
    adj = random_graph(rng, n, 6000)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ref = O.propagate(O.laplacian_adj(adj, 0.5), x, K, "fma")
    op = LaplacianGraphOp(K, r=0.5)
    op.mode = "exact"
    op.cache_dir = str(tmp_path)
    first = op.propagate(adj, x)                      # miss -> GPU -> save
    assert len(os.listdir(tmp_path)) == 1
    for k in range(K + 1):
        assert np.array_equal(first[k].numpy(), ref[k])
    op2 = LaplacianGraphOp(K, r=0.5)
    op2.mode = "exact"
    op2.cache_dir = str(tmp_path)
    second = op2.propagate(adj, x)                    # hit: no operator is built
    assert op2._operator is None and op2._adj is not None
    assert all(np.array_equal(a.numpy(), b.numpy()) for a, b in zip(first, second))
    assert second[0].numpy().ctypes.data == x.ctypes.data     # element 0 still aliases the caller's array
    op3 = LaplacianGraphOp(K, r=0.3)
    op3.cache_dir = str(tmp_path)
    op3.propagate(adj, x)
    assert len(os.listdir(tmp_path)) == 2 and op3._operator is not None

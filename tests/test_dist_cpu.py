"""CPU tests (gloo, world_size 2 and 3): the 1-D row partition, the halo / all-gather exchange plans and the hop loop
of sgl_b200.dist, with the oracle's C hop injected as the local kernel (the CUDA kernel is covered by -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sgap_oracle as O
from sgl_b200.dist import DistOperator, build_plan, chunk_row_bounds, exchange_volume_bytes, partition_rows


def _graph(seed=0, n=600, m=5000):
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, n, m)
    cols = (rng.zipf(1.3, m) - 1) % n
    r2, c2 = np.concatenate([rows, cols]), np.concatenate([cols, rows])
    adj = sp.csr_matrix((np.ones(r2.size, dtype=np.float32), (r2, c2)), shape=(n, n))
    return O.laplacian_adj(adj, 0.5)


def test_partition_balances_work_and_covers_rows():
    a = _graph()
    for world in (1, 2, 3, 8):
        b = partition_rows(a.indptr, world)
        assert b[0] == 0 and b[-1] == a.shape[0] and np.all(np.diff(b) >= 0) and len(b) == world + 1
        work = np.diff(a.indptr[b]) + 4 * np.diff(b)
        assert work.max() <= work.mean() * 1.5 + a.indptr[1:].max()


def test_plans_renumber_columns_consistently():
    a = _graph(1)
    world = 4
    x = np.random.default_rng(2).standard_normal((a.shape[0], 8)).astype(np.float32)
    ref = O.spmm_hop(a, x, "fma")
    bounds = partition_rows(a.indptr, world)
    for mode, chunks in (("halo", 1), ("halo", 3), ("allgather", 1)):
        plans = [build_plan(a.indptr, a.indices, a.data, a.shape[1], world, r, mode, n_chunks=chunks)
                 for r in range(world)]
        for p in plans:
            lo, hi = bounds[p.rank], bounds[p.rank + 1]
            ext = np.zeros((p.n_ext, 8), dtype=np.float32)
            if mode == "halo":
                ext[:p.n_local] = x[lo:hi]
                pos = p.n_local
                for c in range(chunks):     # what all-to-all c would deliver: rows q sends to p, in q's send order
                    for q in range(world):
                        rows_q = plans[q].send_rows[c][p.rank] + bounds[q]
                        assert len(rows_q) == p.recv_counts[c][q]
                        # rows of chunk c are exactly those q's hop finishes in its c-th tile range
                        assert np.all((rows_q - bounds[q] >= plans[q].chunk_rows[c]) &
                                      (rows_q - bounds[q] < plans[q].chunk_rows[c + 1]))
                        ext[pos:pos + len(rows_q)] = x[rows_q]
                        pos += len(rows_q)
                assert pos == p.n_ext
            else:
                for q in range(world):
                    ext[q * p.max_rows:q * p.max_rows + (bounds[q + 1] - bounds[q])] = x[bounds[q]:bounds[q + 1]]
            y = np.zeros((p.n_local, 8), dtype=np.float32)
            O._lib().oracle_spmm_f32_fma_i64(y, p.data, p.indices, p.indptr, ext, p.n_local, 8)
            assert np.array_equal(y, ref[lo:hi])        # storage order is unchanged: bit-exact
        halo = sum(exchange_volume_bytes(p, 8) for p in plans) if mode == "halo" else None
        if halo is not None:
            assert halo <= (world - 1) * a.shape[0] * 8 * 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, mode, K, d, out_dir, chunks=1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a = _graph(3)
        x = np.random.default_rng(4).standard_normal((a.shape[0], d)).astype(np.float32)
        plan = build_plan(a.indptr, a.indices, a.data, a.shape[1], world, rank, mode, n_chunks=chunks)

        def oracle_hop(x_ext, out):          # checker kernel standing in for sglb200_spmm on the CPU
            y = np.zeros((plan.n_local, d), dtype=np.float32)
            O._lib().oracle_spmm_f32_fma_i64(y, plan.data, plan.indices, plan.indptr,
                                             np.ascontiguousarray(x_ext.numpy()), plan.n_local, d)
            out.copy_(torch.from_numpy(y))

        op = DistOperator(plan, local_hop=oracle_hop)
        lo, hi = plan.bounds[rank], plan.bounds[rank + 1]
        hops = op.propagate(torch.from_numpy(x[lo:hi].copy()), K)
        np.save(os.path.join(out_dir, f"hops_{mode}_{rank}.npy"), np.stack([h.numpy() for h in hops]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("mode,chunks", [("halo", 1), ("halo", 4), ("allgather", 1)])
def test_distributed_propagate_matches_single_process(tmp_path, world, mode, chunks):
    K, d = 3, 12
    port = _free_port()
    mp.spawn(_worker, args=(world, port, mode, K, d, str(tmp_path), chunks), nprocs=world, join=True)
    a = _graph(3)
    x = np.random.default_rng(4).standard_normal((a.shape[0], d)).astype(np.float32)
    ref = np.stack(O.propagate(a, x, K, "fma"))
    bounds = partition_rows(a.indptr, world)
    for r in range(world):
        got = np.load(tmp_path / f"hops_{mode}_{r}.npy")
        assert np.array_equal(got, ref[:, bounds[r]:bounds[r + 1]])


def _brute_force_rows_before(indptr, item_pos):
    """Walk the merged (row ends + non-zeros) stream item by item: rows whose end marker lies in the first item_pos items."""
    n = len(indptr) - 1
    done, pos, j = 0, 0, 0
    for r in range(n):
        # non-zeros of row r, then its end marker
        pos += indptr[r + 1] - indptr[r]
        if pos >= item_pos:
            return done
        pos += 1
        if pos > item_pos:
            return done
        done += 1
    return done


def test_chunk_row_bounds_against_item_walk():
    rng = np.random.default_rng(12)
    for trial in range(20):
        n = int(rng.integers(1, 400))
        deg = rng.integers(0, 40, n) * (rng.random(n) < 0.7)
        if trial % 4 == 0:
            deg[rng.integers(0, n)] = 3000            # a hub longer than several tiles
        indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
        for tile_items in (32, 256):
            for n_chunks in (1, 2, 5):
                got = chunk_row_bounds(indptr, n_chunks, tile_items)
                total = n + int(indptr[-1])
                n_tiles = (total + tile_items - 1) // tile_items
                want = [0] + [_brute_force_rows_before(indptr, min((n_tiles * c // n_chunks) * tile_items, total))
                              for c in range(1, n_chunks)] + [n]
                assert list(got) == want, (trial, tile_items, n_chunks)


def _collective_worker(rank, world, port, chunks, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sgl_b200.dist import build_plan_collective
        a = _graph(7, n=900, m=9000)
        bounds = partition_rows(a.indptr, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        loc_ptr = torch.from_numpy(a.indptr[lo:hi + 1] - a.indptr[lo])
        cols = torch.from_numpy(a.indices[a.indptr[lo]:a.indptr[hi]].astype(np.int64))
        vals = torch.from_numpy(a.data[a.indptr[lo]:a.indptr[hi]].astype(np.float32))
        got = build_plan_collective(loc_ptr, cols, vals, bounds, n_chunks=chunks)
        want = build_plan(a.indptr, a.indices, a.data, a.shape[1], world, rank, "halo", n_chunks=chunks)
        ok = (np.array_equal(got.indices.numpy(), want.indices) and got.n_ext == want.n_ext
              and got.recv_counts == want.recv_counts and list(got.chunk_rows) == list(want.chunk_rows)
              and all(np.array_equal(got.send_rows[c][q], want.send_rows[c][q])
                      for c in range(chunks) for q in range(world)))
        np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([ok]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,chunks", [(2, 1), (3, 4)])
def test_collective_plan_equals_full_matrix_plan(tmp_path, world, chunks):
    """Each rank sees only its own rows; the collectively built plan must equal the one derived from the full matrix."""
    port = _free_port()
    mp.spawn(_collective_worker, args=(world, port, chunks, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert bool(np.load(tmp_path / f"ok_{r}.npy")[0]), f"rank {r}"


def _feature_split_worker(rank, world, port, d, K, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sgl_b200.dist import FeatureSplitOperator
        a = _graph(9)
        x = np.random.default_rng(10).standard_normal((a.shape[0], d)).astype(np.float32)

        def oracle_hop(xb, out):
            out.copy_(torch.from_numpy(O.spmm_hop(a, xb.numpy(), "fma")))

        op = FeatureSplitOperator(world=world, rank=rank, local_hop=oracle_hop)
        cb = op.column_bounds(d, world)
        hops = op.propagate(torch.from_numpy(x[:, cb[rank]:cb[rank + 1]].copy()), K)
        full = op.gather_columns(hops[-1], d)
        np.save(os.path.join(out_dir, f"fs_{rank}.npy"), full.numpy())
        shard = op.rows_from_columns(hops[-1], d)          # the all-to-all form: this rank's rows, all columns
        np.save(os.path.join(out_dir, f"fs_rows_{rank}.npy"), shard.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,d", [(2, 12), (3, 10)])
def test_feature_split_needs_no_exchange_and_matches(tmp_path, world, d):
    from sgl_b200.dist import FeatureSplitOperator
    cb = FeatureSplitOperator.column_bounds(d, world)
    assert cb[0] == 0 and cb[-1] == d and np.all(np.diff(cb) >= 0)
    K = 3
    port = _free_port()
    mp.spawn(_feature_split_worker, args=(world, port, d, K, str(tmp_path)), nprocs=world, join=True)
    a = _graph(9)
    x = np.random.default_rng(10).standard_normal((a.shape[0], d)).astype(np.float32)
    ref = O.propagate(a, x, K, "fma")[-1]
    rb = FeatureSplitOperator.row_bounds(a.shape[0], world)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"fs_{r}.npy"), ref)   # column blocks are independent: bit-exact
        assert np.array_equal(np.load(tmp_path / f"fs_rows_{r}.npy"), ref[rb[r]:rb[r + 1]])

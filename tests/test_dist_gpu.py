"""GPU test (-m gpu, needs >= 2 devices, skipped otherwise): the row-partitioned operator on 2 ranks, halo rows moved
either by our own NVLink peer-memory kernels (CUDA IPC) or by NCCL."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import sgap_oracle as O

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, mode, tmp, chunks=1, transport="peer"):
    import scipy.sparse as sp
    import torch.distributed as dist
    from sgl_b200.dist import DistOperator, build_plan
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["SGLB200_DIST_TRANSPORT"] = transport
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rng = np.random.default_rng(5)
        n, d, K = 20000, 128, 3
        rows = rng.integers(0, n, 150000)
        cols = (rng.zipf(1.25, rows.size) - 1) % n
        adj = sp.csr_matrix((np.ones(2 * rows.size, dtype=np.float32),
                             (np.concatenate([rows, cols]), np.concatenate([cols, rows]))), shape=(n, n))
        a = O.laplacian_adj(adj, 0.5)
        x = rng.standard_normal((n, d)).astype(np.float32)
        if mode == "halo" and chunks == 4:      # plan built collectively from this rank's rows only (device tensors)
            from sgl_b200.dist import build_plan_collective, partition_rows
            bounds = partition_rows(a.indptr, world)
            b0, b1 = int(bounds[rank]), int(bounds[rank + 1])
            j0, j1 = int(a.indptr[b0]), int(a.indptr[b1])
            plan = build_plan_collective(torch.from_numpy(a.indptr[b0:b1 + 1] - a.indptr[b0]).cuda(),
                                         torch.from_numpy(a.indices[j0:j1].astype(np.int64)).cuda(),
                                         torch.from_numpy(a.data[j0:j1].astype(np.float32)).cuda(), bounds, n_chunks=chunks)
        else:
            plan = build_plan(a.indptr, a.indices, a.data, n, world, rank, mode, n_chunks=chunks)
        op = DistOperator(plan, mode="exact")
        lo, hi = plan.bounds[rank], plan.bounds[rank + 1]
        hops = op.propagate(torch.from_numpy(x[lo:hi]).cuda(), K)
        again = op.propagate(torch.from_numpy(x[lo:hi]).cuda(), K)        # slabs and flags are reused across calls
        assert all(torch.equal(a, b) for a, b in zip(hops, again))
        assert op.transport == (transport if mode == "halo" else "nccl")
        np.save(os.path.join(tmp, f"{mode}_{rank}.npy"), np.stack([h.cpu().numpy() for h in hops]))
        if rank == 0:
            np.save(os.path.join(tmp, "ref.npy"), np.stack(O.propagate(a, x, K, "fma")))
            np.save(os.path.join(tmp, "bounds.npy"), plan.bounds)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,chunks,transport", [("halo", 1, "peer"), ("halo", 4, "peer"), ("halo", 1, "nccl"),
                                                   ("halo", 4, "nccl"), ("allgather", 1, "nccl")])
def test_two_rank_nccl_row_partition(tmp_path, mode, chunks, transport):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, mode, str(tmp_path), chunks, transport), nprocs=2, join=True)
    ref, bounds = np.load(tmp_path / "ref.npy"), np.load(tmp_path / "bounds.npy")
    for r in range(2):
        got = np.load(tmp_path / f"{mode}_{r}.npy")
        assert np.array_equal(got, ref[:, bounds[r]:bounds[r + 1]])       # exact mode: bit-identical to one GPU


def _feature_worker(rank, world, port, tmp):
    import scipy.sparse as sp
    import torch.distributed as dist
    from sgl_b200.dist import FeatureSplitOperator
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rng = np.random.default_rng(6)
        n, d, K = 12000, 100, 3
        rows = rng.integers(0, n, 90000)
        cols = (rng.zipf(1.25, rows.size) - 1) % n
        adj = sp.csr_matrix((np.ones(2 * rows.size, dtype=np.float32),
                             (np.concatenate([rows, cols]), np.concatenate([cols, rows]))), shape=(n, n))
        a = O.laplacian_adj(adj, 0.5)
        x = rng.standard_normal((n, d)).astype(np.float32)
        import scipy.sparse
        fs = FeatureSplitOperator(adj_norm=scipy.sparse.csr_matrix((a.data, a.indices, a.indptr), shape=a.shape), world=world,
                                  rank=rank, mode="exact")
        cb = fs.column_bounds(d, world)
        blk = torch.from_numpy(np.ascontiguousarray(x[:, cb[rank]:cb[rank + 1]])).cuda()
        hops = fs.propagate(blk, K)                       # lane-group kernel on the column block, no exchange
        shard = fs.rows_from_columns(hops[-1], d)         # the one all-to-all: this rank's rows, all columns
        np.save(os.path.join(tmp, f"fs_rows_{rank}.npy"), shard.cpu().numpy())
        if rank == 0:
            np.save(os.path.join(tmp, "fs_ref.npy"), O.propagate(a, x, K, "fma")[-1])
        fs.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_feature_split_is_bit_identical_to_one_gpu(tmp_path):
    """Default multi-GPU partition: A^ replicated, feature columns split, rows_from_columns over NCCL; EXACT mode equals the
    oracle's single-GPU chain bit for bit on every rank's row shard."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from sgl_b200.dist import FeatureSplitOperator
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_feature_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ref = np.load(tmp_path / "fs_ref.npy")
    rb = FeatureSplitOperator.row_bounds(ref.shape[0], 2)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"fs_rows_{r}.npy"), ref[rb[r]:rb[r + 1]])

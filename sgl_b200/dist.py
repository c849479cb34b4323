"""dist.py -- 1-D row partition of the normalised adjacency over the GPUs of one box.

The reference has no parallelism on this path (its only collective is DDP of the dense head,
sgl/tasks/node_classification_dist.py:61-70).  Here rank p owns the contiguous row range [b_p, b_{p+1}) of A^ and
of every feature slab; hop k needs X_{k-1}[j] for every column j its rows reference.  Two exchange plans:

  "halo"      (default) only the rows a rank actually references travel, de-duplicated, received straight behind the
              local shard of an "extended" feature slab.  Transport "peer" (default on CUDA): our own kernels store
              the rows into the peers' slabs over NVLink (CUDA IPC memory, sgl_b200/csrc/peer.cu), pipelined against
              the hop itself in `n_chunks` tile ranges; transport "nccl": pack + all_to_all_single;
  "allgather" every rank receives every shard (torch.distributed.all_gather_into_tensor) -- what north_star names,
              kept as the simple fallback and as the comparison the halo plan is measured against.

Column ids are renumbered once at construction: local columns -> [0, n_local), remote columns -> n_local + position
in the receive layout (halo: chunk-major, then peer, then row) or rank*max_rows + offset (allgather, shards padded to
equal length).  The per-hop arithmetic is the single-GPU kernel on a rectangular operator (sglb200_spmm /
sglb200_spmm_tiles).  The storage order of the non-zeros of a row is not changed by the renumbering, so EXACT-mode
results are bit-identical to the single-GPU result and independent of plan and transport; FAST mode differs only by
where long rows are cut (tolerance 1e-5).

Everything that does not touch CUDA (partitioning, plans, renumbering, chunk bounds) is plain numpy and is exercised on
CPU with the gloo backend (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------------------------
# partitioning and exchange plans (numpy, no CUDA)
# ---------------------------------------------------------------------------------------------------------------
def partition_rows(indptr: np.ndarray, world: int, row_cost: float = 4.0) -> np.ndarray:
    """Contiguous row ranges with equal work: cost(row) = nnz(row) + row_cost.  Returns bounds[world+1]."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = indptr.shape[0] - 1
    cost = indptr + (np.arange(n + 1, dtype=np.float64) * row_cost).astype(np.int64)
    targets = cost[-1] * np.arange(1, world, dtype=np.float64) / world
    inner = np.searchsorted(cost, targets, side="left").astype(np.int64)
    bounds = np.concatenate([[0], inner, [n]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


TILE_ITEMS = 256  # tile size the partitioned operators are created with (host and device chunk bounds must agree)


def chunk_row_bounds(loc_ptr: np.ndarray, n_chunks: int, tile_items: int = TILE_ITEMS) -> np.ndarray:
    """Rows finished by each of n_chunks consecutive tile ranges of the merge-path schedule (mirrors
    sglb200_graph_chunks / build_tiles_kernel: boundary tile t sits at item t*tile_items of the merged
    row-end + non-zero stream; the rows finished before it are those with indptr[r+1] + r <= t*tile_items - 1)."""
    loc_ptr = np.asarray(loc_ptr, dtype=np.int64)
    n = loc_ptr.shape[0] - 1
    total = n + int(loc_ptr[-1])
    n_tiles = (total + tile_items - 1) // tile_items if total > 0 else 0
    key = loc_ptr[1:] + np.arange(n, dtype=np.int64)
    bounds = np.zeros(n_chunks + 1, dtype=np.int64)
    for c in range(1, n_chunks):
        k = min((n_tiles * c // n_chunks) * tile_items, total)
        bounds[c] = np.searchsorted(key, k - 1, side="right")
    bounds[n_chunks] = n
    return bounds


@dataclass
class RankPlan:
    """Everything rank `rank` needs: its rectangular CSR with renumbered columns and its exchange lists."""
    rank: int
    world: int
    bounds: np.ndarray            # [world+1] global row bounds
    indptr: np.ndarray            # [n_local+1] int64
    indices: np.ndarray           # [nnz_local] int32, renumbered
    data: np.ndarray              # [nnz_local] float32
    n_ext: int                    # columns of the rectangular operator = rows of the extended feature buffer
    mode: str                     # "halo" | "allgather"
    n_chunks: int                 # halo: the hop and its exchange are pipelined over this many row chunks
    send_rows: List[List[np.ndarray]]   # halo: [chunk][peer] LOCAL row ids this rank sends (sorted)
    recv_counts: List[List[int]]        # halo: [chunk][peer] rows received
    chunk_rows: np.ndarray        # halo: [n_chunks+1] local rows finished by each chunk of this rank's hop
    max_rows: int                 # allgather: padded shard length

    @property
    def n_local(self) -> int:
        return int(self.bounds[self.rank + 1] - self.bounds[self.rank])


def _owner_of(cols: np.ndarray, bounds: np.ndarray) -> np.ndarray:
    return np.searchsorted(bounds, cols, side="right") - 1


def needed_remote_rows(indptr, indices, bounds, rank) -> List[np.ndarray]:
    """Per peer q: sorted global ids of the rows of q referenced by the rows of `rank`."""
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    cols = np.asarray(indices[indptr[lo]:indptr[hi]], dtype=np.int64)
    uniq = np.unique(cols)
    owner = _owner_of(uniq, bounds)
    world = len(bounds) - 1
    return [uniq[owner == q] if q != rank else np.zeros(0, dtype=np.int64) for q in range(world)]


def build_plan(indptr, indices, data, n_cols: int, world: int, rank: int, mode: str = "halo",
               bounds: Optional[np.ndarray] = None, need_from_all: Optional[List[List[np.ndarray]]] = None,
               n_chunks: int = 1) -> RankPlan:
    """Plan of one rank from the FULL CSR (every rank holds or can regenerate it; SGL's preprocess is single-process,
    models/base_model.py:23).  Send lists are derived from the full matrix so that no handshake is needed.
    Halo layout of the extended buffer: [local rows | chunk 0: peer 0 rows, peer 1 rows, ... | chunk 1: ... ] where
    chunk c of peer q are the rows q's hop finishes in its c-th tile range -- each (chunk) block is one all-to-all."""
    indptr = np.asarray(indptr, dtype=np.int64)
    if bounds is None:
        bounds = partition_rows(indptr, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    loc_ptr = indptr[lo:hi + 1] - indptr[lo]
    cols = np.asarray(indices[indptr[lo]:indptr[hi]], dtype=np.int64)
    vals = np.asarray(data[indptr[lo]:indptr[hi]], dtype=np.float32)
    n_local = hi - lo
    if mode == "allgather":
        max_rows = int(np.max(np.diff(bounds)))
        owner = _owner_of(cols, bounds)
        new_cols = owner * max_rows + (cols - bounds[owner])
        return RankPlan(rank, world, bounds, loc_ptr, new_cols.astype(np.int32), vals, world * max_rows, mode, 1, [], [],
                        np.array([0, n_local]), max_rows)
    if mode != "halo":
        raise ValueError("mode must be 'halo' or 'allgather'")
    C = max(1, int(n_chunks))
    # rows finished per chunk, for every rank (a function of that rank's row pointer only)
    crb = [chunk_row_bounds(indptr[bounds[q]:bounds[q + 1] + 1] - indptr[bounds[q]], C) for q in range(world)]
    need_all = need_from_all if need_from_all is not None else \
        [needed_remote_rows(indptr, indices, bounds, p) for p in range(world)]
    need = need_all[rank]
    # receive layout: chunk-major, then peer, then row
    recv_counts = [[0] * world for _ in range(C)]
    seg_start = np.zeros((C, world), dtype=np.int64)      # offset (in rows, from n_local) of block (c, q)
    split = []                                            # per peer: positions of the chunk borders in need[q]
    for q in range(world):
        local_ids = need[q] - bounds[q]
        split.append(np.searchsorted(local_ids, crb[q]))
    pos = 0
    for c in range(C):
        for q in range(world):
            cnt = int(split[q][c + 1] - split[q][c]) if q != rank else 0
            recv_counts[c][q] = cnt
            seg_start[c, q] = pos
            pos += cnt
    n_halo = pos
    owner = _owner_of(cols, bounds)
    new_cols = np.empty_like(cols)
    is_local = owner == rank
    new_cols[is_local] = cols[is_local] - lo
    for q in range(world):
        m = owner == q
        if q == rank or not m.any():
            continue
        idx = np.searchsorted(need[q], cols[m])                        # position in q's sorted need list
        chunk = np.searchsorted(split[q], idx, side="right") - 1       # which chunk of q finishes that row
        new_cols[m] = n_local + seg_start[chunk, q] + (idx - split[q][chunk])
    # what this rank SENDS: for every peer p, the rows of `rank` that p needs, split by this rank's chunks
    send_rows = [[np.zeros(0, dtype=np.int64) for _ in range(world)] for _ in range(C)]
    for p in range(world):
        if p == rank:
            continue
        theirs = (need_all[p][rank] - lo).astype(np.int64)
        cut = np.searchsorted(theirs, crb[rank])
        for c in range(C):
            send_rows[c][p] = theirs[cut[c]:cut[c + 1]]
    return RankPlan(rank, world, bounds, loc_ptr, new_cols.astype(np.int32), vals, n_local + n_halo, mode, C,
                    send_rows, recv_counts, crb[rank], 0)


def build_plan_collective(loc_ptr, cols, vals, bounds, n_chunks: int = 1, group=None) -> RankPlan:
    """Halo plan built COLLECTIVELY from each rank's own rows only: no rank needs the full matrix.

    loc_ptr [n_local+1], cols [nnz_local] (GLOBAL column ids), vals [nnz_local]: torch tensors on the device the
    process group communicates with (CUDA for NCCL, CPU for gloo); bounds: the global row bounds (numpy, length
    world+1).  The receive side (what this rank references) is a device-side unique / searchsorted pass; the send
    side (what every peer references here) arrives through one all-to-all of the reference lists.  Produces exactly the
    plan build_plan() derives from the full matrix (tests/test_dist_cpu.py)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = cols.device
    C = max(1, int(n_chunks))
    bounds = np.asarray(bounds, dtype=np.int64)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    n_local = hi - lo
    b_t = torch.as_tensor(bounds, device=dev)
    cols = cols.to(torch.int64)
    # rows finished by each tile range of every rank's hop (a function of that rank's row pointer alone)
    mine = chunk_row_bounds(loc_ptr.detach().cpu().numpy(), C)
    crb_all = [None] * world
    dist.all_gather_object(crb_all, [int(v) for v in mine], group=group)
    crb = [torch.as_tensor(np.asarray(c, dtype=np.int64), device=dev) for c in crb_all]
    # receive side: the distinct remote rows this rank references, sorted by global id (= by owner, then local id)
    owner = torch.searchsorted(b_t, cols, right=True) - 1
    remote = owner != rank
    uniq = torch.unique(cols[remote])
    u_owner = torch.searchsorted(b_t, uniq, right=True) - 1
    per_peer = torch.bincount(u_owner, minlength=world)                     # rows referenced per owner
    peer_first = torch.cumsum(per_peer, 0) - per_peer                       # start of each owner's block in uniq
    u_local = uniq - b_t[u_owner]                                           # id inside the owner's shard
    # chunk of the owner's hop that finishes each referenced row
    u_chunk = torch.zeros_like(uniq)
    for q in range(world):
        m = u_owner == q
        if q != rank and bool(m.any()):
            u_chunk[m] = torch.searchsorted(crb[q], u_local[m], right=True) - 1
    # layout: chunk-major, then owner, then row
    counts = torch.zeros((C, world), dtype=torch.int64, device=dev)
    if uniq.numel():
        counts.view(-1).index_add_(0, u_chunk * world + u_owner, torch.ones_like(uniq))
    seg_start = (torch.cumsum(counts.view(-1), 0) - counts.view(-1)).view(C, world)
    # position of every referenced row inside its (chunk, owner) block: rows of one owner are sorted, chunks are
    # consecutive ranges of them, so the rank inside the block is the index in uniq minus the block's first index
    idx_in_uniq = torch.arange(uniq.numel(), device=dev)
    first_of_block = torch.zeros((C, world), dtype=torch.int64, device=dev)
    blk = u_chunk * world + u_owner
    if uniq.numel():
        first_of_block.view(-1).scatter_reduce_(0, blk, idx_in_uniq, reduce="amin", include_self=False)
    pos_u = n_local + seg_start.view(-1)[blk] + (idx_in_uniq - first_of_block.view(-1)[blk])
    new_cols = torch.empty_like(cols)
    new_cols[~remote] = cols[~remote] - lo
    if bool(remote.any()):
        new_cols[remote] = pos_u[torch.searchsorted(uniq, cols[remote])]
    recv_counts = counts.cpu().numpy().tolist()
    n_halo = int(counts.sum())
    # send side: every owner learns which of its rows each peer references (one all-to-all of the lists)
    send_counts_t = per_peer.clone()
    send_counts_t[rank] = 0
    got_counts = torch.empty_like(send_counts_t)
    dist.all_to_all_single(got_counts, send_counts_t, group=group)
    out_list = u_local                                                      # already grouped by owner, sorted
    in_list = torch.empty(int(got_counts.sum()), dtype=torch.int64, device=dev)
    dist.all_to_all_single(in_list, out_list, output_split_sizes=got_counts.cpu().tolist(),
                           input_split_sizes=send_counts_t.cpu().tolist(), group=group)
    send_rows = [[np.zeros(0, dtype=np.int64) for _ in range(world)] for _ in range(C)]
    starts = (torch.cumsum(got_counts, 0) - got_counts).cpu().tolist()
    my_crb = crb[rank]
    for p in range(world):
        if p == rank or int(got_counts[p]) == 0:
            continue
        theirs = in_list[starts[p]:starts[p] + int(got_counts[p])]
        cut = torch.searchsorted(theirs, my_crb).cpu().tolist()
        host = theirs.cpu().numpy()
        for c in range(C):
            send_rows[c][p] = host[cut[c]:cut[c + 1]]
    return RankPlan(rank, world, bounds, loc_ptr.to(torch.int64), new_cols.to(torch.int32), vals.to(torch.float32),
                    n_local + n_halo, "halo", C, send_rows, recv_counts, mine, 0)


# ---------------------------------------------------------------------------------------------------------------
# NVLink peer memory (CUDA IPC) for the NCCL-free halo exchange
# ---------------------------------------------------------------------------------------------------------------
class _RawCuda:
    """Exposes a raw device allocation to torch through the CUDA array interface (no ownership)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerSlabs:
    """Two extended feature slabs + one flag word per rank, allocated with sglb200_ipc_alloc and mapped by every other
    rank of the box.  Collective constructor (all ranks of `group`)."""

    FLAG_BYTES = 4096

    def __init__(self, n_ext: int, d: int, device, group=None):
        import ctypes
        from . import _lib
        self._lib = _lib.load()
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_ext, self.d = int(n_ext), int(d)
        slab_bytes = ((self.n_ext * self.d * 4 + 255) // 256) * 256
        self.slab_bytes = slab_bytes
        total = self.FLAG_BYTES + 2 * slab_bytes
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(device):
            _lib.check(self._lib.sglb200_ipc_alloc(total, ctypes.byref(ptr), handle), "ipc_alloc")
        self.base = int(ptr.value)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (bytes(handle), slab_bytes), group=group)
        self.peer_base, self.peer_slab_bytes = [], []
        for q, (h, sb) in enumerate(gathered):
            if q == self.rank:
                self.peer_base.append(self.base)
            else:
                p = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                with torch.cuda.device(device):
                    _lib.check(self._lib.sglb200_ipc_open(buf, ctypes.byref(p)), "ipc_open")
                self.peer_base.append(int(p.value))
            self.peer_slab_bytes.append(int(sb))
        self.slabs = [torch.as_tensor(_RawCuda(self.base + self.FLAG_BYTES + i * slab_bytes, (self.n_ext, self.d), "<f4"),
                                      device=device) for i in range(2)]
        self.flags = torch.as_tensor(_RawCuda(self.base, (self.world,), "<i8"), device=device)
        # device array of the addresses of MY flag word inside every rank's flag block
        self.flag_ptrs = torch.tensor([b + 8 * self.rank for b in self.peer_base], dtype=torch.int64, device=device)
        dist.barrier(group=group)

    def peer_slab_ptr(self, q: int, which: int, row: int) -> int:
        return self.peer_base[q] + self.FLAG_BYTES + which * self.peer_slab_bytes[q] + row * self.d * 4

    def close(self):
        for q, b in enumerate(self.peer_base):
            if q != self.rank and b:
                self._lib.sglb200_ipc_close(ctypes_void(b))
        if self.base:
            self._lib.sglb200_ipc_free(ctypes_void(self.base))
        self.base = 0
        self.peer_base = []


def ctypes_void(v):
    import ctypes
    return ctypes.c_void_p(int(v))


# ---------------------------------------------------------------------------------------------------------------
# distributed operator
# ---------------------------------------------------------------------------------------------------------------
class DistOperator:
    """Row-partitioned A^ on this rank.  `local_hop(x_ext, out)` computes out = A_local @ x_ext; the default is the
    CUDA kernel through CsrOperator.  CPU/gloo tests inject a checker hop to exercise the exchange logic.

    With a chunked halo plan on CUDA the hop is pipelined against its own exchange: the tile ranges of the hop run
    back to back on the compute stream, and as soon as range c has produced its rows their pack + all-to-all runs on
    a second stream while range c+1 computes."""

    def __init__(self, plan: RankPlan, device: Optional[torch.device] = None, group=None,
                 local_hop: Optional[Callable] = None, mode: str = "fast"):
        self.plan = plan
        self.group = group
        self.mode = mode
        self._tiles = None
        if local_hop is None:
            from .runtime import CsrOperator, require_cuda
            require_cuda()
            self.device = device or torch.device("cuda", torch.cuda.current_device())
            self._op = CsrOperator(plan.indptr, plan.indices, plan.data, (plan.n_local, plan.n_ext),
                                   tile_items=TILE_ITEMS)
            self._hop = lambda x_ext, out: self._op.spmm(x_ext, out=out, mode=self.mode)
            if plan.mode == "halo" and plan.n_chunks > 1 and plan.world > 1:
                tiles, rows = self._op.chunks(plan.n_chunks, self.mode)
                if list(rows) != [int(v) for v in plan.chunk_rows]:
                    raise RuntimeError(f"chunk row bounds differ between host plan {plan.chunk_rows} and device {rows}")
                self._tiles = tiles
                self._xstream = torch.cuda.Stream(device=self.device, priority=-1)
        else:
            self.device = device or torch.device("cpu")
            self._op = None
            self._hop = local_hop
        C = plan.n_chunks if plan.mode == "halo" else 0
        self._send_idx = [torch.from_numpy(np.concatenate(plan.send_rows[c])).to(self.device) if plan.world > 1 else None
                          for c in range(C)]
        self._send_counts = [[int(r.size) for r in plan.send_rows[c]] for c in range(C)]
        self._recv_start = []
        pos = plan.n_local
        for c in range(C):
            self._recv_start.append(pos)
            pos += int(sum(plan.recv_counts[c]))
        self._bufs = {}
        # transport of the halo rows: "peer" = stores into the peers' slabs over NVLink (CUDA IPC, own kernels),
        # "nccl" = pack + all_to_all_single.  Peer needs CUDA, a halo plan and more than one rank.
        want = os.environ.get("SGLB200_DIST_TRANSPORT", "peer")
        self.transport = "peer" if (want == "peer" and local_hop is None and plan.mode == "halo" and plan.world > 1) \
            else "nccl"
        self._peer = {}
        self._epoch = 0
        if self.transport == "peer":
            # where MY rows land inside every receiver: receiver q's block (chunk c, sender me) starts at row
            # recv_pos[q][c][me] of q's extended slab
            mine = [[self._recv_start[c] + int(sum(plan.recv_counts[c][:q])) for q in range(plan.world)] for c in range(C)]
            table = [None] * plan.world
            dist.all_gather_object(table, mine, group=group)
            self._dst_row = [[table[q][c][plan.rank] for q in range(plan.world)] for c in range(C)]
            self._send_per_peer = [[torch.from_numpy(plan.send_rows[c][q]).to(self.device) for q in range(plan.world)]
                                   for c in range(C)]
            if self._tiles is None:
                self._xstream = torch.cuda.Stream(device=self.device, priority=-1)
            # grid cap of pushes that overlap a hop (0 = none: measured best on 2 x B200, profiles/)
            self._push_blocks = int(os.environ.get("SGLB200_PUSH_BLOCKS", "0"))  # 0 = full grid (measured best)

    # -- exchange over peer memory ---------------------------------------------------------------------------------
    def _peer_slabs(self, d: int) -> PeerSlabs:
        if d not in self._peer:
            self._peer[d] = PeerSlabs(self.plan.n_ext, d, self.device, self.group)
        return self._peer[d]

    def _push_chunk(self, ps: PeerSlabs, which: int, c: int, blocks: int = 0) -> None:
        """Copy the rows of chunk c that each peer references from my slab `which` into that peer's slab `which`."""
        from . import _lib
        lib = _lib.load()
        p = self.plan
        src = ps.slabs[which]
        stream = ctypes_void(torch.cuda.current_stream().cuda_stream)
        # staggered peer order (rank+1, rank+2, ...): at every step each receiver has exactly one sender, instead of all
        # ranks storing into rank 0 first (incast on one NVLink port: measured 195 GB/s per sender at 8 GPUs)
        for step in range(1, p.world):
            q = (p.rank + step) % p.world
            rows = self._send_per_peer[c][q]
            if rows.numel() == 0:
                continue
            _lib.check(lib.sglb200_push_rows(ctypes_void(src.data_ptr()), ps.d, ps.d, ctypes_void(rows.data_ptr()),
                                             int(rows.numel()), ctypes_void(ps.peer_slab_ptr(q, which, self._dst_row[c][q])),
                                             ps.d, blocks, stream), "push_rows")

    def _signal(self, ps: PeerSlabs) -> int:
        """Publish 'everything I enqueued so far on this stream has been written' to every rank; returns the epoch."""
        from . import _lib
        self._epoch += 1
        _lib.check(_lib.load().sglb200_signal_peers(ctypes_void(ps.flag_ptrs.data_ptr()), self.plan.world, self._epoch,
                                                    ctypes_void(torch.cuda.current_stream().cuda_stream)), "signal_peers")
        return self._epoch

    def _wait(self, ps: PeerSlabs, epoch: int) -> None:
        from . import _lib
        _lib.check(_lib.load().sglb200_wait_flags(ctypes_void(ps.flags.data_ptr()), self.plan.world, epoch,
                                                  ctypes_void(torch.cuda.current_stream().cuda_stream)), "wait_flags")

    def _propagate_peer(self, x_local: torch.Tensor, prop_steps: int, keep: str) -> List[torch.Tensor]:
        p = self.plan
        d = int(x_local.shape[1])
        ps = self._peer_slabs(d)
        trace = os.environ.get("SGLB200_DIST_TRACE") == "1"
        marks = []
        cs, xs = torch.cuda.current_stream(), self._xstream
        n_chunks = p.n_chunks
        tiles = self._tiles
        # nobody may still be reading the slabs of a previous call when the first rows arrive
        self._wait(ps, self._signal(ps))
        ps.slabs[0][:p.n_local].copy_(x_local)
        for c in range(n_chunks):
            self._push_chunk(ps, 0, c)
        arrived = self._signal(ps)
        outs = [x_local]
        for k in range(1, prop_steps + 1):
            src, dst = ps.slabs[(k - 1) % 2], ps.slabs[k % 2]
            self._wait(ps, arrived)                        # every peer has delivered the halo rows of src
            y = dst[:p.n_local]
            last = k == prop_steps
            if tiles is None:
                self._op.spmm(src, out=y, mode=self.mode)
                if not last:
                    for c in range(n_chunks):
                        self._push_chunk(ps, k % 2, c)
                    arrived = self._signal(ps)
            else:
                for c in range(n_chunks):
                    if trace:
                        e0 = torch.cuda.Event(enable_timing=True); e0.record(cs)
                    self._op.spmm_tiles(src, y, tiles[c], tiles[c + 1], mode=self.mode)
                    if trace:
                        e1 = torch.cuda.Event(enable_timing=True); e1.record(cs)
                        marks.append(("hop%d.compute%d" % (k, c), e0, e1))
                    if not last:
                        done = cs.record_event()
                        with torch.cuda.stream(xs):
                            xs.wait_event(done)
                            if trace:
                                x0 = torch.cuda.Event(enable_timing=True); x0.record(xs)
                            self._push_chunk(ps, k % 2, c, self._push_blocks if c + 1 < n_chunks else 0)
                            if trace:
                                x1 = torch.cuda.Event(enable_timing=True); x1.record(xs)
                                marks.append(("hop%d.push%d" % (k, c), x0, x1))
                if not last:
                    with torch.cuda.stream(xs):
                        arrived = self._signal(ps)          # ordered behind all pushes of this hop
            if keep == "all" or last:
                outs.append(y.clone())
        cs.wait_stream(xs)
        if trace and marks:
            torch.cuda.synchronize()
            t0 = marks[0][1]
            print("[dist trace rank %d] " % p.rank + "  ".join(
                "%s@%.2f+%.2fms" % (n, t0.elapsed_time(a), a.elapsed_time(b)) for n, a, b in marks[:4 * n_chunks]),
                flush=True)
        return outs

    # -- exchange ----------------------------------------------------------------------------------------------
    def _exchange_chunk(self, ext: torch.Tensor, c: int) -> None:
        """Send the rows of chunk c every peer needs and receive the peers' chunk-c rows into ext's halo block c."""
        p = self.plan
        local = ext[:p.n_local]
        idx = self._send_idx[c]
        if idx.numel():
            if local.is_cuda:
                from .runtime import gather_rows
                send = gather_rows([local], idx)[0]               # pack kernel (sglb200_gather_rows)
            else:
                send = local[idx]
        else:
            send = local.new_zeros((0, ext.shape[1]))
        n_recv = int(sum(p.recv_counts[c]))
        recv = ext[self._recv_start[c]:self._recv_start[c] + n_recv]
        dist.all_to_all_single(recv, send, output_split_sizes=p.recv_counts[c], input_split_sizes=self._send_counts[c],
                               group=self.group)

    def _exchange(self, ext: torch.Tensor) -> None:
        """Fill the remote part of the extended buffer `ext` ([n_ext, d]); the local shard is already in place."""
        p = self.plan
        if p.world == 1:
            return
        if p.mode == "allgather":
            # shards are padded to max_rows so that the fast equal-size collective applies
            mine = ext[p.rank * p.max_rows:(p.rank + 1) * p.max_rows]
            dist.all_gather_into_tensor(ext, mine if mine.is_cuda else mine.clone(), group=self.group)
            return
        for c in range(p.n_chunks):
            self._exchange_chunk(ext, c)

    def _local_view(self, ext: torch.Tensor) -> torch.Tensor:
        p = self.plan
        if p.mode == "allgather":
            return ext[p.rank * p.max_rows:p.rank * p.max_rows + p.n_local]
        return ext[:p.n_local]

    def propagate(self, x_local: torch.Tensor, prop_steps: int, keep: str = "all") -> List[torch.Tensor]:
        """[X_p, (A^X)_p, ..., (A^^K X)_p] for this rank's rows (keep='last': only the last hop is retained)."""
        p = self.plan
        if self.transport == "peer":
            return self._propagate_peer(x_local, prop_steps, keep)
        d = int(x_local.shape[1])
        if d not in self._bufs:  # two extended slabs, reused across calls (padding rows stay zero)
            self._bufs[d] = [torch.zeros((p.n_ext, d), dtype=torch.float32, device=self.device) for _ in range(2)]
        bufs = self._bufs[d]
        self._local_view(bufs[0]).copy_(x_local)
        outs = [x_local]
        if self._tiles is None:
            for k in range(1, prop_steps + 1):
                src, dst = bufs[(k - 1) % 2], bufs[k % 2]
                self._exchange(src)
                y = self._local_view(dst)
                self._hop(src, y)
                if keep == "all" or k == prop_steps:
                    outs.append(y.clone())
            return outs
        # pipelined: the exchange of hop k's rows runs chunk by chunk behind the hop itself
        trace = os.environ.get("SGLB200_DIST_TRACE") == "1"
        marks = []
        cs, xs = torch.cuda.current_stream(), self._xstream
        xs.wait_stream(cs)
        with torch.cuda.stream(xs):
            self._exchange(bufs[0])
        arrived = xs.record_event()
        for k in range(1, prop_steps + 1):
            src, dst = bufs[(k - 1) % 2], bufs[k % 2]
            cs.wait_event(arrived)                      # every halo row of src is in place
            y = dst[:p.n_local]
            last = k == prop_steps
            for c in range(p.n_chunks):
                if trace:
                    e0 = torch.cuda.Event(enable_timing=True); e0.record(cs)
                self._op.spmm_tiles(src, y, self._tiles[c], self._tiles[c + 1], mode=self.mode)
                if trace:
                    e1 = torch.cuda.Event(enable_timing=True); e1.record(cs)
                    marks.append(("hop%d.compute%d" % (k, c), e0, e1))
                if not last:
                    done = cs.record_event()
                    with torch.cuda.stream(xs):
                        xs.wait_event(done)
                        if trace:
                            x0 = torch.cuda.Event(enable_timing=True); x0.record(xs)
                        self._exchange_chunk(dst, c)
                        if trace:
                            x1 = torch.cuda.Event(enable_timing=True); x1.record(xs)
                            marks.append(("hop%d.exchange%d" % (k, c), x0, x1))
            if not last:
                arrived = xs.record_event()
            if keep == "all" or last:
                outs.append(y.clone())
        cs.wait_stream(xs)
        if trace and marks:
            torch.cuda.synchronize()
            t0 = marks[0][1]
            print("[dist trace rank %d] " % p.rank + "  ".join(
                "%s@%.2f+%.2fms" % (n, t0.elapsed_time(a), a.elapsed_time(b)) for n, a, b in marks[:4 * p.n_chunks]),
                flush=True)
        return outs

    def close(self):
        for ps in self._peer.values():
            ps.close()
        self._peer = {}
        if self._op is not None:
            self._op.close()


class FeatureSplitOperator:
    """The zero-communication alternative of SURVEY.md section 8(e): A^ is replicated on every GPU (17 GB even for the
    2-billion-edge config) and the FEATURE dimension is split, rank p propagating columns [c_p, c_{p+1}) of X.  Column
    blocks of A^^k X are independent, so the K hops need no exchange at all; one all-gather at the end assembles the
    rows (or the caller keeps the column blocks).  Kept next to the row partition as the measured comparison -- it
    trades a per-hop halo exchange for gathers of narrow rows (d / world floats each).

    `local_hop(x_block, out)` computes out = A^ @ x_block; default: the CUDA kernel through CsrOperator."""

    def __init__(self, adj_norm=None, world: int = 1, rank: int = 0, group=None, local_hop: Optional[Callable] = None,
                 operator=None, mode: str = "fast"):
        self.world, self.rank, self.group, self.mode = world, rank, group, mode
        if local_hop is None:
            from .runtime import CsrOperator, require_cuda
            require_cuda()
            self._op = operator if operator is not None else CsrOperator.from_scipy(adj_norm)
            self._hop = lambda x, out: self._op.spmm(x, out=out, mode=self.mode)
        else:
            self._op = None
            self._hop = local_hop

    @staticmethod
    def column_bounds(d: int, world: int, align: int = 4) -> np.ndarray:
        """Contiguous column blocks, multiples of `align` floats (16-byte rows) wherever d allows."""
        units = (d + align - 1) // align
        cuts = [min(d, align * ((units * p) // world)) for p in range(world)] + [d]
        return np.asarray(cuts, dtype=np.int64)

    @staticmethod
    def block_slab(n: int, width: int, device, dtype=torch.float32) -> torch.Tensor:
        """[n, width] view of a slab whose row stride is rounded up to 16 floats for narrow blocks: every gathered row
        then starts on a 64-byte boundary (a 48-byte row of a 12-column block touches two 32-byte sectors and one 128-byte
        line instead of up to three sectors across two lines).  Blocks wider than 64 columns keep their own stride."""
        ld = ((width + 15) // 16) * 16 if 0 < width <= 64 else width
        return torch.empty((n, ld), dtype=dtype, device=device)[:, :width]

    def propagate(self, x_block: torch.Tensor, prop_steps: int) -> List[torch.Tensor]:
        """K hops of this rank's column block: [X[:, blk], (A^X)[:, blk], ...] -- no communication."""
        n, w = int(x_block.shape[0]), int(x_block.shape[1])
        padded = x_block.is_cuda and w % 16 != 0 and w <= 64
        if padded:
            first = self.block_slab(n, w, x_block.device, x_block.dtype)
            first.copy_(x_block)
        else:
            first = x_block.contiguous()
        hops = [first]
        for _ in range(prop_steps):
            out = self.block_slab(n, w, x_block.device, x_block.dtype) if padded else torch.empty_like(hops[-1])
            self._hop(hops[-1], out)
            hops.append(out)
        return hops

    def gather_columns(self, block: torch.Tensor, d: int) -> torch.Tensor:
        """Assemble the full [N, d] matrix of one hop from every rank's column block (one all-gather, not per hop)."""
        if self.world == 1:
            return block
        bounds = self.column_bounds(d, self.world)
        width = int(np.max(np.diff(bounds)))
        padded = block.new_zeros((block.shape[0], width))
        padded[:, :block.shape[1]] = block
        parts = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(parts, padded, group=self.group)
        return torch.cat([parts[q][:, :int(bounds[q + 1] - bounds[q])] for q in range(self.world)], dim=1)

    def rows_from_columns(self, block: torch.Tensor, d: int) -> torch.Tensor:
        """The one real exchange of the feature split: turn this rank's column block of a hop ([N, d_p]) into its ROW shard
        with all d columns ([N_p, d], rows [row_bounds[p], row_bounds[p+1])) -- the layout a data-parallel head trains on
        (reference tasks/node_classification_dist.py:73-85 shards the node ids over ranks).  One all_to_all_single moving
        N*d*4/world bytes per rank (an all-gather of the blocks would move world times more)."""
        if self.world == 1:
            return block
        n = int(block.shape[0])
        cb = self.column_bounds(d, self.world)
        rb = self.row_bounds(n, self.world)
        widths = [int(cb[q + 1] - cb[q]) for q in range(self.world)]
        rows = [int(rb[q + 1] - rb[q]) for q in range(self.world)]
        mine_w, mine_r = widths[self.rank], rows[self.rank]
        send = block.contiguous().reshape(-1)
        recv = torch.empty(mine_r * d, dtype=block.dtype, device=block.device)
        dist.all_to_all_single(recv, send, output_split_sizes=[mine_r * w for w in widths],
                               input_split_sizes=[r * mine_w for r in rows], group=self.group)
        out = torch.empty((mine_r, d), dtype=block.dtype, device=block.device)
        off = 0
        for q in range(self.world):
            out[:, int(cb[q]):int(cb[q + 1])] = recv[off:off + mine_r * widths[q]].view(mine_r, widths[q])
            off += mine_r * widths[q]
        return out

    @staticmethod
    def row_bounds(n: int, world: int) -> np.ndarray:
        return np.asarray([(n * p) // world for p in range(world + 1)], dtype=np.int64)

    def close(self):
        if self._op is not None:
            self._op.close()


def exchange_volume_bytes(plan: RankPlan, d: int) -> int:
    """Bytes this rank RECEIVES per hop."""
    if plan.mode == "allgather":
        return (plan.world - 1) * plan.max_rows * d * 4
    return int(sum(sum(c) for c in plan.recv_counts)) * d * 4

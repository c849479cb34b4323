"""dist.py -- 1-D row partition of the normalised adjacency over the GPUs of one box.

The reference has no parallelism on this path (its only collective is DDP of the dense head,
sgl/tasks/node_classification_dist.py:61-70).  Here rank p owns the contiguous row range [b_p, b_{p+1}) of A^ and
of every feature slab; hop k needs X_{k-1}[j] for every column j its rows reference.  Two exchange plans:

  "halo"      (default) only the rows a rank actually references travel: a packed all-to-all-v of de-duplicated
              remote rows (torch.distributed.all_to_all_single over NCCL), received straight behind the local shard;
  "allgather" every rank receives every shard (torch.distributed.all_gather_into_tensor) -- what north_star names,
              kept as the simple fallback and as the comparison the halo plan is measured against.

Column ids are renumbered once at construction: local columns -> [0, n_local), remote columns -> n_local + position
in the receive buffer (halo) or rank*max_rows + offset (allgather, shards padded to equal length).  The per-hop
arithmetic is the single-GPU kernel on a rectangular operator (sglb200_spmm); row order inside a row is preserved
for the local block and grouped by owner rank for remote columns, so results equal the single-GPU FAST-mode result
up to the re-association tolerance (1e-5), and are independent of the exchange plan.

Everything that does not touch CUDA (partitioning, plans, renumbering) is plain numpy and is exercised on CPU with
the gloo backend (tests/test_dist_cpu.py).
"""
from __future__ import annotations

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
    send_rows: List[np.ndarray]   # halo: per peer, LOCAL row ids this rank must send to that peer
    recv_counts: List[int]        # halo: rows received from each peer (in peer order)
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
               bounds: Optional[np.ndarray] = None, need_from_all: Optional[List[List[np.ndarray]]] = None) -> RankPlan:
    """Plan of one rank from the FULL CSR (every rank holds or can regenerate it; SGL's preprocess is single-process,
    models/base_model.py:23).  `need_from_all[p][q]` (rows of q needed by p) may be passed when already known;
    otherwise send lists are derived from the full matrix so that no handshake is needed."""
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
        return RankPlan(rank, world, bounds, loc_ptr, new_cols.astype(np.int32), vals, world * max_rows, mode, [], [],
                        max_rows)
    if mode != "halo":
        raise ValueError("mode must be 'halo' or 'allgather'")
    need = needed_remote_rows(indptr, indices, bounds, rank) if need_from_all is None else need_from_all[rank]
    recv_counts = [int(a.size) for a in need]
    # column renumbering: local -> [0, n_local); remote -> n_local + offset(peer) + position in that peer's list
    offsets = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
    owner = _owner_of(cols, bounds)
    new_cols = np.empty_like(cols)
    is_local = owner == rank
    new_cols[is_local] = cols[is_local] - lo
    for q in range(world):
        m = owner == q
        if q == rank or not m.any():
            continue
        new_cols[m] = n_local + offsets[q] + np.searchsorted(need[q], cols[m])
    # what this rank must SEND to every peer p: the rows of `rank` that p needs (derived from the full matrix)
    send_rows = []
    for p in range(world):
        if p == rank:
            send_rows.append(np.zeros(0, dtype=np.int64))
            continue
        theirs = needed_remote_rows(indptr, indices, bounds, p)[rank] if need_from_all is None else need_from_all[p][rank]
        send_rows.append((theirs - lo).astype(np.int64))
    return RankPlan(rank, world, bounds, loc_ptr, new_cols.astype(np.int32), vals, n_local + int(offsets[-1]), mode,
                    send_rows, recv_counts, 0)


# ---------------------------------------------------------------------------------------------------------------
# distributed operator
# ---------------------------------------------------------------------------------------------------------------
class DistOperator:
    """Row-partitioned A^ on this rank.  `local_hop(x_ext, out)` computes out = A_local @ x_ext; the default is the
    CUDA kernel through CsrOperator.  CPU/gloo tests inject a checker hop to exercise the exchange logic."""

    def __init__(self, plan: RankPlan, device: Optional[torch.device] = None, group=None,
                 local_hop: Optional[Callable] = None, mode: str = "fast"):
        self.plan = plan
        self.group = group
        self.mode = mode
        if local_hop is None:
            from .runtime import CsrOperator, require_cuda
            require_cuda()
            self.device = device or torch.device("cuda", torch.cuda.current_device())
            self._op = CsrOperator(plan.indptr, plan.indices, plan.data, (plan.n_local, plan.n_ext))
            self._hop = lambda x_ext, out: self._op.spmm(x_ext, out=out, mode=self.mode)
        else:
            self.device = device or torch.device("cpu")
            self._op = None
            self._hop = local_hop
        self._send_idx = [torch.from_numpy(r).to(self.device) for r in plan.send_rows]
        self._send_all = torch.cat(self._send_idx) if self._send_idx else None
        self._send_counts = [int(r.size) for r in plan.send_rows]
        self._bufs = {}

    # -- exchange ----------------------------------------------------------------------------------------------
    def _exchange(self, ext: torch.Tensor) -> None:
        """Fill the remote part of the extended buffer `ext` ([n_ext, d]); the local shard is already in place."""
        p = self.plan
        n_local, d = p.n_local, ext.shape[1]
        if p.world == 1:
            return
        if p.mode == "allgather":
            # shards are padded to max_rows so that the fast equal-size collective applies
            mine = ext[p.rank * p.max_rows:(p.rank + 1) * p.max_rows]
            dist.all_gather_into_tensor(ext, mine if mine.is_cuda else mine.clone(), group=self.group)
            return
        local = ext[:n_local]
        if self._send_all.numel():
            if local.is_cuda:
                from .runtime import gather_rows
                send = gather_rows([local], self._send_all)[0]      # pack kernel (sglb200_gather_rows)
            else:
                send = local[self._send_all]
        else:
            send = local.new_zeros((0, d))
        recv = ext[n_local:]
        dist.all_to_all_single(recv, send, output_split_sizes=p.recv_counts, input_split_sizes=self._send_counts,
                               group=self.group)

    def _local_view(self, ext: torch.Tensor) -> torch.Tensor:
        p = self.plan
        if p.mode == "allgather":
            return ext[p.rank * p.max_rows:p.rank * p.max_rows + p.n_local]
        return ext[:p.n_local]

    def propagate(self, x_local: torch.Tensor, prop_steps: int, keep: str = "all") -> List[torch.Tensor]:
        """[X_p, (A^X)_p, ..., (A^^K X)_p] for this rank's rows (keep='last': only the last hop is retained)."""
        p = self.plan
        d = int(x_local.shape[1])
        if d not in self._bufs:  # two extended slabs, reused across calls (padding rows stay zero)
            self._bufs[d] = [torch.zeros((p.n_ext, d), dtype=torch.float32, device=self.device) for _ in range(2)]
        bufs = self._bufs[d]
        self._local_view(bufs[0]).copy_(x_local)
        outs = [x_local]
        for k in range(1, prop_steps + 1):
            src, dst = bufs[(k - 1) % 2], bufs[k % 2]
            self._exchange(src)
            y = self._local_view(dst)
            self._hop(src, y)
            if keep == "all" or k == prop_steps:
                outs.append(y.clone())
        return outs

    def close(self):
        if self._op is not None:
            self._op.close()


def exchange_volume_bytes(plan: RankPlan, d: int) -> int:
    """Bytes this rank RECEIVES per hop."""
    if plan.mode == "allgather":
        return (plan.world - 1) * plan.max_rows * d * 4
    return int(sum(plan.recv_counts)) * d * 4

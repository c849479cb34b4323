#!/usr/bin/env python
"""bench.py -- K-hop CSR-SpMM propagation throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload arxiv|products|pubmed|rmatS] [--impl reference]

One "step" = one full pass of the hot path over the workload: the K hops  X -> A^X -> ... -> A^^K X  of the
normalised adjacency (GraphOp.propagate minus the one-time normalisation, which is reported separately).
Prints ONE JSON line (contract in the task statement):
  value     whole-job propagated edges/s = nnz(A^) * K * steps / device time, inputs resident in HBM
  e2e       same metric through CsrOperator.propagate_host (C ABI sglb200_propagate_host): pinned host X in,
            K pinned host slabs out, H2D/D2H inside the timed region
  roofline  achieved algorithmic GB/s of the hop kernel vs the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own CPU kernel (oracle/_ref, compiled from its matmul.c) on this box's host cores
`--impl reference` times that CPU kernel as the step and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "propagated edges/sec (k-hop SpMM)"
UNIT = "edges/s"

# shape-matched synthetics of BASELINE.json's configs; the datasets are not available offline
WORKLOADS = {
    #            N          directed edges   R-MAT scale  d    K   undirected?
    "pubmed":   (19_717,    44_324,          15,          500, 3),
    "arxiv":    (169_343,   1_166_243,       18,          128, 5),
    "products": (2_449_029, 61_859_140,      22,          100, 6),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--tile-items", type=int, default=0)
    ap.add_argument("--split-threshold", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--exchange", default="halo", choices=["halo", "allgather"])
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="halo rows move by our NVLink peer-store kernels (CUDA IPC) or by NCCL all_to_all")
    ap.add_argument("--partition", default="row", choices=["row", "feature"],
                    help="N>1: 1-D row partition with a halo exchange per hop (default, SURVEY 8e), or A^ replicated and the "
                         "feature columns split across GPUs (no per-hop exchange; one all-gather of the last hop)")
    ap.add_argument("--plan", default="replicated", choices=["replicated", "collective"],
                    help="halo plan from the full matrix on every rank (numpy) or built collectively from local rows (torch)")
    ap.add_argument("--feat-dim", type=int, default=0, help="experiment: override the feature width of the workload")
    ap.add_argument("--relabel", default="none", choices=["none", "degree"],
                    help="experiment: relabel the vertices by descending degree before building A^")
    ap.add_argument("--chunks", type=int, default=4, help="row chunks the hop is pipelined over against its halo exchange")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
# synthetic graphs
# ---------------------------------------------------------------------------------------------------------------
def rmat_edges(n, m, scale, seed, device):
    """m directed R-MAT(a=.57, b=.19, c=.19, d=.05) edges on 2^scale ids folded mod n (torch, any device)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    src = torch.zeros(m, dtype=torch.int64, device=device)
    dst = torch.zeros(m, dtype=torch.int64, device=device)
    for _ in range(scale):
        u = torch.rand(m, generator=g, device=device)
        src_bit = (u >= 0.76)                                 # quadrants c, d
        dst_bit = ((u >= 0.57) & (u < 0.76)) | (u >= 0.95)    # quadrants b, d
        src = (src << 1) | src_bit.to(torch.int64)
        dst = (dst << 1) | dst_bit.to(torch.int64)
    # break the bit-pattern locality of raw R-MAT ids with a fixed multiplicative hash, then fold
    mult = 0x9E3779B1
    src = ((src * mult) & 0x7FFFFFFF) % n
    dst = ((dst * mult) & 0x7FFFFFFF) % n
    return src, dst


def build_adjacency(name, device):
    """scipy CSR (float32, int32) of the symmetrised graph: concatenation without dedup, duplicates summed
    (reference sgl/data/utils.py:18-24 + sgl/data/base_data.py:29-30)."""
    import scipy.sparse as sp
    import torch
    if name.startswith("rmat"):
        scale = int(name[4:])
        n, m, d, K = 1 << scale, 8 << scale, 128, 10
    else:
        n, m, scale, d, K = WORKLOADS[name]
    seed = {"pubmed": 0, "arxiv": 1, "products": 2}.get(name, 4)
    src, dst = rmat_edges(n, m, scale, seed, device)
    rows = torch.cat([src, dst])
    cols = torch.cat([dst, src])
    keys, counts = torch.unique(rows * n + cols, return_counts=True)   # sorted (row, col), multiplicities
    rows = torch.div(keys, n, rounding_mode="floor")
    cols = keys - rows * n
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    adj = sp.csr_matrix((counts.to(torch.float32).cpu().numpy(), cols.to(torch.int32).cpu().numpy(),
                         indptr.cpu().numpy().astype(np.int32 if keys.numel() < 2 ** 31 else np.int64)), shape=(n, n))
    return adj, d, K


def algorithmic_bytes_per_hop(n, nnz, d):
    """fp32 values + int32 column ids + int64 row pointers, X read once, Y written once."""
    return 8 * nnz + 8 * (n + 1) + 8 * n * d


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        inside = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows[-3:]]
        sm, reasons, mx = [], set(), None
        for r in inside:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU reference arm
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(adj_norm, x, K, steps, warmup, budget_s=None):
    """K hops on the host with the reference's own kernel (oracle/_ref/libmatmul_ref.so, built from the reference's
    matmul.c; falls back to the oracle's C port of the same loop when that build is absent or 32-bit offsets overflow).
    Returns (seconds per step list, kind, threads)."""
    from oracle import sgap_oracle as O
    a = O.Csr(adj_norm.indptr, adj_norm.indices, adj_norm.data, adj_norm.shape)
    n, d = x.shape
    ref = O.load_reference_kernel()
    kind = "reference"
    if ref is None or n * d >= 2 ** 31 or a.nnz >= 2 ** 31:
        ref, kind = None, "port"
    a32 = a.data.astype(np.float32)
    indptr32 = a.indptr.astype(np.int32) if ref is not None else None

    def one_step():
        cur = x
        for _ in range(K):
            if ref is not None:
                cur = O.reference_kernel_hop(ref, a, cur, a32, indptr32)
            else:
                cur = O.spmm_hop(a, cur, "fma")
        return cur

    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one_step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s and times:
            break
    return times, kind, O.num_threads()


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or "arxiv"
    adj, d, K = build_adjacency(name, "cuda" if torch.cuda.is_available() else "cpu")
    from sgl_b200.operators.utils import adj_to_symmetric_norm
    adj_norm = adj_to_symmetric_norm(adj, 0.5).tocsr()
    n, nnz = adj.shape[0], adj_norm.nnz
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(0)).numpy()
    times, kind, threads = cpu_reference_steps(adj_norm, x, K, args.steps, args.warmup, budget_s=240.0)
    sec = float(np.sum(times))
    value = nnz * K * len(times) / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * sec / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(name, n, nnz, d, K, args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"full workload, {len(times)} steps of K={K} hops, bare kernel (no wrapper copies)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(name, n, nnz, d, K, args):
    return {"workload": f"{name}-shape synthetic R-MAT (BASELINE configs[{ {'pubmed': 0, 'arxiv': 1, 'products': 2}.get(name, 4)}])",
            "N": n, "nnz": nnz, "d": d, "prop_steps": K, "operator": "LaplacianGraphOp r=0.5",
            "mode": args.mode, "partition": "single GPU" if args.gpus == 1 else f"1-D row partition x{args.gpus}",
            "l2": "256 MiB buffer written between timed steps (L2 flush); K+1 slabs + CSR exceed L2"}


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    from sgl_b200.operators.graph_op import LaplacianGraphOp
    from sgl_b200.runtime import CsrOperator

    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return run_dist(args)

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    name = args.workload or "arxiv"
    from sgl_b200.graph_build import build_operator_device, parts_to_scipy
    t0 = time.perf_counter()
    rows, cols, n, d, K = device_graph(name, dev)
    if args.feat_dim > 0:
        d = args.feat_dim
    if args.relabel == "degree":
        deg = torch.bincount(rows, minlength=n)
        order = torch.argsort(deg, descending=True, stable=True)
        newid = torch.empty_like(order)
        newid[order] = torch.arange(n, device=dev)
        rows, cols = newid[rows], newid[cols]
        del deg, order, newid
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    # A^ = D^-1/2 (A+I)^T D^-1/2 built on the device (sgl_b200.graph_build: bit-identical structure and values to the
    # reference's scipy pass, tests/test_gpu_parity.py); the reference's own host pass is timed for small graphs only
    t0 = time.perf_counter()
    op = build_operator_device(rows, cols, n, r=0.5, tile_items=args.tile_items, split_threshold=args.split_threshold)
    torch.cuda.synchronize()
    t_upload = time.perf_counter() - t0
    del rows, cols
    adj_norm = parts_to_scipy(op.parts)
    op.parts = None
    torch.cuda.empty_cache()
    t_norm = None
    if n <= 500_000:
        adj, _, _ = build_adjacency(name, dev)
        t0 = time.perf_counter()
        host_norm = LaplacianGraphOp(K, r=0.5)._construct_adj(adj)
        t_norm = time.perf_counter() - t0
        assert np.array_equal(host_norm.indices, adj_norm.indices) and np.array_equal(host_norm.data, adj_norm.data)
    nnz = int(adj_norm.nnz)
    info = op.info()

    x_host = torch.randn(n, d, generator=torch.Generator().manual_seed(0)).pin_memory()
    hops = [x_host.to(dev)] + [torch.empty(n, d, device=dev) for _ in range(K)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        for k in range(1, K + 1):
            op.spmm(hops[k - 1], out=hops[k], mode=args.mode)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                # evict the previous step's slabs from L2 (outside the event pair)
        starts[i].record(stream)
        step()
        stops[i].record(stream)
    torch.cuda.synchronize()
    wall1 = time.perf_counter()
    clocks = sampler.stop(wall0, wall1)
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_s = float(np.sum(step_ms)) / 1e3
    value = nnz * K * args.steps / total_s
    hop_s = total_s / (args.steps * K)

    peaks = {}
    peak_src = "fallback 6650 GB/s (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_gbs, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        peak_gbs = 6650.0
    b_alg = algorithmic_bytes_per_hop(n, nnz, d)
    achieved = b_alg / hop_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": None, "kernel": "spmm_flat_kernel<4,1,8>", "algorithmic_bytes_per_launch": b_alg,
                "gather_bytes_per_launch": nnz * (8 + 4 * d) + n * (8 + 4 * d), "us_per_launch": hop_s * 1e6,
                "peak_source": peak_src}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            t = json.load(open(prof))
            if t.get("workload") == name:
                roofline["traffic"] = t.get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- parity spot check on the timed buffers: every hop against the oracle on a row sample -------------------
    from oracle import sgap_oracle as O
    rng = np.random.default_rng(1)
    sample = np.sort(rng.choice(n, min(n, 2000), replace=False))
    sub = adj_norm[sample].astype(np.float32)
    worst = 0.0
    for k in range(1, K + 1):
        ref = np.zeros((sample.size, d), dtype=np.float32)
        O._lib().oracle_spmm_f32_fma_i64(ref, sub.data, sub.indices.astype(np.int32), sub.indptr.astype(np.int64),
                                         hops[k - 1].cpu().numpy(), sample.size, d)
        got = hops[k][torch.from_numpy(sample).to(dev)].cpu().numpy()
        worst = max(worst, float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)))
    assert worst <= 1e-5, f"bench parity check failed: {worst}"

    # ---- end to end through the public API with host buffers ------------------------------------------------------
    # the call a user makes: SGC(prop_steps=K).preprocess(adj, x) -- configs[1] is SGC, whose LastMessageOp consumes
    # only hop K.  Inside the timed region: H2D of x from pinned host memory, K hops, aggregate, D2H of the result into
    # pinned host memory.  A^ is prepared once (GraphOp.prepare: the one-time setup reported under "setup").
    e2e = None
    if not args.no_e2e and n <= 8_000_000:
        from sgl_b200.sgap import SGC
        adj, _, _ = build_adjacency(name, dev)
        model = SGC(prop_steps=K, feat_dim=d, output_dim=8)
        model._pre_graph_op.mode = args.mode
        model._pre_graph_op.build_on = "device"
        t0 = time.perf_counter()
        model._pre_graph_op.prepare(adj)
        torch.cuda.synchronize()
        t_prepare = time.perf_counter() - t0
        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(3):
            model.preprocess(adj, x_host)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            model.preprocess(adj, x_host)
            result = model._processed_feature
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert not result.is_cuda and result.shape == (n, d)
        e2e = {"value": nnz * K * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": n * d * 4,
               "d2h_bytes_per_step": n * d * 4, "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "prepare_graph_once_s": t_prepare,
               "api": "sgl_b200.sgap.SGC.preprocess(adj, x): pinned host x in -> K hops + LastMessageOp on the GPU -> "
                      "pinned host [N,d] out; A^ prepared once with GraphOp.prepare(adj)"}
        # all K hops to the host (what GraphOp.propagate returns in the reference), through the C ABI entry point
        for _ in range(2):
            op.propagate_host(x_host, K, mode=args.mode, keep="all")
        t0 = time.perf_counter()
        for _ in range(max(3, e2e_steps // 2)):
            outs = op.propagate_host(x_host, K, mode=args.mode, keep="all")
        dt = time.perf_counter() - t0
        e2e["all_hops_to_host"] = {"value": nnz * K * max(3, e2e_steps // 2) / dt, "d2h_bytes_per_step": K * n * d * 4,
                                   "api": "CsrOperator.propagate_host -> sglb200_propagate_host (K pinned slabs out)"}
        got = result.numpy()[sample]
        assert float(np.abs(got - hops[K][torch.from_numpy(sample).to(dev)].cpu().numpy()).max()) == 0.0
        assert float(np.abs(outs[-1].numpy()[sample] - got).max()) == 0.0
        model._pre_graph_op._operator.close()

    # ---- GPU comparator: cuSPARSE SpMM (the reference's own dormant GPU choice, csrc/cudamatmul.c:104-119) -------------
    comparators = {}
    try:
        a_t = torch.sparse_csr_tensor(torch.from_numpy(adj_norm.indptr.astype(np.int64)).to(dev),
                                      torch.from_numpy(adj_norm.indices.astype(np.int64)).to(dev),
                                      torch.from_numpy(adj_norm.data.astype(np.float32)).to(dev), size=(n, n))
        cur = hops[0]
        for _ in range(3):
            cur = torch.sparse.mm(a_t, hops[0])
        torch.cuda.synchronize()
        c_steps = max(3, min(args.steps, 10))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = 0.0
        for i in range(c_steps):
            flush.fill_(i & 0xFF)
            ev0.record(stream)
            cur = hops[0]
            for _ in range(K):
                cur = torch.sparse.mm(a_t, cur)
            ev1.record(stream)
            torch.cuda.synchronize()
            ms += ev0.elapsed_time(ev1)
        err = float((cur - hops[K]).abs().max() / hops[K].abs().max())
        comparators["cusparse_spmm_via_torch"] = {"value": nnz * K * c_steps / (ms / 1e3), "unit": UNIT,
                                                   "us_per_hop": 1e3 * ms / (c_steps * K), "max_rel_diff_vs_ours": err,
                                                   "note": "torch.sparse.mm on a CSR tensor (cuSPARSE SpMM), device resident, "
                                                           "same graph and features; a library call, reported for context"}
        del a_t, cur
    except Exception as exc:  # the comparator must never break the bench line
        comparators["cusparse_spmm_via_torch"] = {"unavailable": str(exc)[:200]}

    cpu = None
    if not args.no_cpu_baseline:
        times, kind, threads = cpu_reference_steps(adj_norm, x_host.numpy(), K, steps=3, warmup=1, budget_s=25.0)
        cpu = {"value": nnz * K * len(times) / float(np.sum(times)), "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"full workload, {len(times)} steps of K={K} hops after 1 warm-up, bare kernel"}

    # cut rows are folded inside the hop kernel (one launch per hop) unless the separate fold launch is selected
    separate_fold = os.environ.get("SGLB200_FOLD", "kernel").startswith("f") or d > 512
    launches_per_step = K * (1 + (1 if info["carry_runs"] > 0 and args.mode == "fast" and separate_fold else 0))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(name, n, nnz, d, K, args), "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e, "comparators": comparators, "clocks": clocks,
            "gpu_launches": launches_per_step * args.steps,
            "parity": {"checked": "every hop vs oracle fma chain on 2000 sampled rows", "max_rel_err": worst},
            "setup": {"generate_s": t_gen, "normalise_host_scipy_s": t_norm, "build_on_device_s": t_upload,
                      "tiles": info["tiles_fast"], "cut_rows": info["carry_runs"], "tile_items": info["tile_items"],
                      "split_threshold": info["split_threshold"], "bytes_resident": info["bytes_resident"]}}
    print(json.dumps(line))


def device_graph(name, dev):
    """Edges of the workload on the device: (rows, cols, n, d, K) with the symmetrisation the reference applies."""
    import torch
    if name.startswith("rmat"):
        scale = int(name[4:])
        n, m, d, K = 1 << scale, 8 << scale, 128, 10
    else:
        n, m, scale, d, K = WORKLOADS[name]
    seed = {"pubmed": 0, "arxiv": 1, "products": 2}.get(name, 4)
    src, dst = rmat_edges(n, m, scale, seed, dev)
    return torch.cat([src, dst]), torch.cat([dst, src]), n, d, K


def run_dist(args):
    """N > 1: 1-D row partition of A^ (sgl_b200.dist), one process per GPU under torchrun, NCCL halo exchange per hop.
    Strong scaling: the same graph as N = 1; value = nnz * K * steps / max-over-ranks device time."""
    import torch
    import torch.distributed as dist
    from sgl_b200.dist import DistOperator, build_plan, exchange_volume_bytes
    from sgl_b200.graph_build import normalized_adjacency_device, values_from_parts

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ["SGLB200_DIST_TRANSPORT"] = args.transport
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    name = args.workload or "arxiv"
    rows, cols, n, d, K = device_graph(name, dev)
    if args.partition == "feature":
        return run_feature_split(args, rank, world, local, dev, name, rows, cols, n, d, K)
    t0 = time.perf_counter()
    parts = normalized_adjacency_device(rows, cols, n, None, r=0.5, alpha=None, pow_on="host")
    del rows, cols
    vals = values_from_parts(parts).to(torch.float32)
    indptr, indices, data = parts["indptr"].cpu().numpy(), parts["indices"].cpu().numpy(), vals.cpu().numpy()
    nnz = int(indptr[-1])
    del parts, vals
    torch.cuda.empty_cache()
    # small shards are launch-bound: pipelining them in chunks only adds launches (measured: arxiv-shape on 4 GPUs)
    n_chunks = args.chunks if nnz // world >= 8_000_000 else 1
    if args.plan == "collective" and args.exchange == "halo":
        from sgl_b200.dist import build_plan_collective, partition_rows
        bounds = partition_rows(indptr, world)
        b0, b1 = int(bounds[rank]), int(bounds[rank + 1])
        j0, j1 = int(indptr[b0]), int(indptr[b1])
        plan = build_plan_collective(torch.from_numpy(indptr[b0:b1 + 1] - indptr[b0]).to(dev),
                                     torch.from_numpy(indices[j0:j1].astype(np.int64)).to(dev),
                                     torch.from_numpy(data[j0:j1]).to(dev), bounds, n_chunks=n_chunks)
    else:
        plan = build_plan(indptr, indices, data, n, world, rank, args.exchange, n_chunks=n_chunks)
    t_build = time.perf_counter() - t0
    op = DistOperator(plan, device=dev, mode=args.mode)
    lo, hi = int(plan.bounds[rank]), int(plan.bounds[rank + 1])
    x_full = torch.randn(n, d, generator=torch.Generator().manual_seed(0))
    x_local = x_full[lo:hi].to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        return op.propagate(x_local, K, keep="last")

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    dist.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        starts[i].record()
        outs = step()
        stops[i].record()
    torch.cuda.synchronize()
    dist.barrier()
    wall1 = time.perf_counter()
    mine = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(starts, stops)) / 1e3], device=dev, dtype=torch.float64)
    dist.all_reduce(mine, op=dist.ReduceOp.MAX)
    total_s = float(mine.item())
    recv = torch.tensor([float(exchange_volume_bytes(plan, d))], device=dev, dtype=torch.float64)
    dist.all_reduce(recv, op=dist.ReduceOp.MAX)
    # end to end at N GPUs: every rank uploads its row shard of X from pinned host memory, the K partitioned hops run,
    # and the rank's rows of the last hop (what SGC's LastMessageOp consumes) come back into pinned host memory
    e2e = None
    if not args.no_e2e:
        x_pin = x_full[lo:hi].clone().pin_memory()
        y_pin = torch.empty((hi - lo, d), dtype=torch.float32, pin_memory=True)
        e2e_steps = max(3, min(args.steps, 10))

        def e2e_step():
            xd = x_pin.to(dev, non_blocking=True)
            last = op.propagate(xd, K, keep="last")[-1]
            y_pin.copy_(last, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            e2e_step()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": nnz * K * e2e_steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": n * d * 4,
               "d2h_bytes_per_step": n * d * 4, "ms_per_step": 1e3 * float(dt.item()) / e2e_steps, "steps": e2e_steps,
               "api": "DistOperator.propagate per rank: pinned host row shard in -> K partitioned hops -> pinned host rows "
                      "of hop K out (bytes are whole-job totals)"}
    # parity: this rank's last hop against the oracle chain on sampled local rows, computed from the full input
    ok = 1.0
    if rank == 0:
        clocks = sampler.stop(wall0, wall1)
        value = nnz * K * args.steps / total_s
        hop_s = total_s / (args.steps * K)
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            peak, src = 6650.0, "fallback (B200_PROFILING.md)"
        b_alg = algorithmic_bytes_per_hop(n, nnz, d)
        achieved = b_alg / hop_s / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(name, n, nnz, d, K, args), exchange=args.exchange, chunks=n_chunks,
                               transport=op.transport,
                               halo_recv_bytes_per_hop_max_rank=float(recv.item())),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s",
                             "frac": achieved / (peak * world), "traffic": None, "kernel": "spmm_flat_kernel",
                             "algorithmic_bytes_per_launch": b_alg / world, "us_per_launch": hop_s * 1e6,
                             "peak_source": src + " x n_gpus"},
                "cpu_baseline": None,
                "e2e": e2e, "clocks": clocks, "gpu_launches": args.steps * K * 3,
                "setup": {"build_plan_s": t_build}}
        print(json.dumps(line))
    dist.destroy_process_group()


def run_feature_split(args, rank, world, local, dev, name, rows, cols, n, d, K):
    """N > 1, alternative partition: every GPU holds all of A^ and propagates d / N feature columns; the only collective
    is one all-gather of the last hop's column blocks inside the timed step."""
    import torch
    import torch.distributed as dist
    from sgl_b200.dist import FeatureSplitOperator
    from sgl_b200.graph_build import build_operator_device

    t0 = time.perf_counter()
    op = build_operator_device(rows, cols, n, r=0.5)
    nnz = int(op.nnz)
    op.parts = None
    t_build = time.perf_counter() - t0
    fs = FeatureSplitOperator(world=world, rank=rank, operator=op, mode=args.mode)
    cb = fs.column_bounds(d, world)
    x_full = torch.randn(n, d, generator=torch.Generator().manual_seed(0))
    x_blk = x_full[:, int(cb[rank]):int(cb[rank + 1])].contiguous().to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        return fs.gather_columns(fs.propagate(x_blk, K)[-1], d)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        starts[i].record()
        step()
        stops[i].record()
    torch.cuda.synchronize()
    dist.barrier()
    mine = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(starts, stops)) / 1e3], device=dev, dtype=torch.float64)
    dist.all_reduce(mine, op=dist.ReduceOp.MAX)
    total_s = float(mine.item())
    if rank == 0:
        line = {"metric": METRIC, "value": nnz * K * args.steps / total_s, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * total_s / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(name, n, nnz, d, K, args), partition=f"feature split x{world} (A^ replicated)"),
                "roofline": None, "cpu_baseline": None, "e2e": None, "clocks": None, "gpu_launches": args.steps * K,
                "setup": {"build_on_device_s": t_build}}
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use every host core it can, and libgomp reads
        # the variable when it is loaded (before torch / the oracle library are imported below)
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

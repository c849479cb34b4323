"""oracle/gen_golden.py -- generates tests/golden/*.npz from the UNMODIFIED reference (test infrastructure).

Run in the build container (where /root/reference exists):

    python oracle/gen_golden.py

It imports sgl.operators straight from /root/reference (nothing is copied), runs LaplacianGraphOp / PprGraphOp
.propagate (Linux branch -> shipped libmatmul.so, and the scipy float64 branch) and every MessageOp on small
seeded graphs, and stores inputs + outputs.  The fixtures pin oracle/sgap_oracle.py (tests/test_oracle.py) and,
on the GPU box where /root/reference does not exist, the CUDA path (tests/test_gpu_parity.py).
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

REF = os.environ.get("SGL_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from sgl.operators.graph_op import LaplacianGraphOp, PprGraphOp  # noqa: E402
from sgl.operators.message_op import (ConcatMessageOp, LastMessageOp, LearnableWeightedMessageOp,  # noqa: E402
                                      MaxMessageOp, MeanMessageOp, MinMessageOp, OverSmoothDistanceWeightedOp,
                                      SimpleWeightedMessageOp, SumMessageOp, IterateLearnableWeightedMessageOp)
from sgl.operators.utils import csr_sparse_dense_matmul  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def csr_from_edges(n, rows, cols, vals=None):
    """Same construction as the reference's Edge (sgl/data/base_data.py:29-30): duplicates are summed."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.ones(len(rows), dtype=np.float32) if vals is None else np.asarray(vals, dtype=np.float32)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))


def undirected(rows, cols):
    """to_undirected without dedup (sgl/data/utils.py:18-24): reciprocal pairs end up with weight 2."""
    return np.concatenate([rows, cols]), np.concatenate([cols, rows])


def graphs():
    rng = np.random.default_rng(20260925)
    g = {}
    # path 0-1-2-3-4-5
    r, c = undirected(np.arange(5), np.arange(1, 6))
    g["path6"] = csr_from_edges(6, r, c)
    # star with centre 0 and an isolated node 7
    r, c = undirected(np.zeros(6, dtype=np.int64), np.arange(1, 7))
    g["star_isolated"] = csr_from_edges(8, r, c)
    # pre-existing self loops + reciprocal directed pairs (weight 2 after symmetrisation) + duplicates
    rows = np.array([0, 1, 1, 2, 2, 3, 3, 3, 4, 0])
    cols = np.array([0, 0, 2, 1, 2, 4, 4, 0, 3, 1])
    r, c = undirected(rows, cols)
    g["selfloop_dup"] = csr_from_edges(5, r, c)
    # directed (propagates along the transpose), not symmetrised
    g["directed"] = csr_from_edges(6, [0, 0, 1, 2, 4, 4, 5], [1, 2, 3, 3, 0, 5, 5])
    # random non-integer weights, directed, with an empty row and an empty column
    n = 40
    m = 160
    rows = rng.integers(0, n - 1, size=m)
    cols = rng.integers(1, n, size=m)
    g["weighted40"] = csr_from_edges(n, rows, cols, rng.uniform(0.1, 3.0, size=m))
    # skewed undirected graph, a few hundred nodes, hubs
    n = 200
    m = 1000
    p = 1.0 / np.arange(1, n + 1) ** 0.9
    p /= p.sum()
    rows = rng.choice(n, size=m, p=p)
    cols = rng.choice(n, size=m, p=p)
    r, c = undirected(rows, cols)
    g["skewed200"] = csr_from_edges(n, r, c)
    return g


def feats(n, d, seed):
    gen = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=gen, dtype=torch.float32).numpy()


def pack_csr(prefix, m, out):
    m = m.tocsr()
    out[prefix + "_indptr"] = m.indptr.astype(np.int64)
    out[prefix + "_indices"] = m.indices.astype(np.int32)
    out[prefix + "_data"] = np.asarray(m.data)
    out[prefix + "_shape"] = np.asarray(m.shape, dtype=np.int64)


def gen_graph_cases():
    dims = {"path6": 3, "star_isolated": 8, "selfloop_dup": 5, "directed": 4, "weighted40": 100, "skewed200": 128}
    for name, adj in graphs().items():
        out = {}
        pack_csr("adj", adj, out)
        n = adj.shape[0]
        x = feats(n, dims[name], seed=sum(map(ord, name)) + n)
        out["x"] = x
        K = 3
        cfgs = [("lap_r0.5", LaplacianGraphOp(K, r=0.5)), ("lap_r0.3", LaplacianGraphOp(K, r=0.3)),
                ("lap_r0", LaplacianGraphOp(K, r=0.0)), ("ppr_r0.5_a0.15", PprGraphOp(K, r=0.5, alpha=0.15))]
        for tag, op in cfgs:
            hops = op.propagate(adj, x)                      # Linux branch: shipped libmatmul.so, fp32 fma per hop
            pack_csr(tag + "_norm", op._adj, out)            # float64 normalised CSR (structure must match bit-exact)
            out[tag + "_hops_fma"] = np.stack([h.numpy() for h in hops])
            if n * dims[name] > 4096 and tag != "lap_r0.5":
                continue  # keep the big fixture small: the other flavours only for the default operator
            # scipy float64 branch (base_op.py:34): fp64 across hops, one cast at the end
            f64 = [x]
            for _ in range(K):
                f64.append(op._adj.dot(f64[-1]))
            out[tag + "_hops_f64"] = np.stack([np.asarray(f, dtype=np.float32) for f in f64])
            # scipy float32 csr.dot (north_star's "scipy.sparse CPU path"), rounded per hop
            a32 = op._adj.astype(np.float32)
            f32 = [x]
            for _ in range(K):
                f32.append(a32.dot(f32[-1]))
            out[tag + "_hops_scipy32"] = np.stack(f32)
        # the bare wrapper on the first config
        op = LaplacianGraphOp(1, r=0.5)
        op.propagate(adj, x)
        out["wrapper_hop"] = csr_sparse_dense_matmul(op._adj, x)
        np.savez_compressed(os.path.join(OUT, f"graph_{name}.npz"), **out)
        print("wrote", name, {k: v.shape for k, v in out.items() if k.endswith("hops_fma")})


def gen_message_cases():
    torch.manual_seed(7)
    adj = graphs()["skewed200"]
    n, d, K = adj.shape[0], 16, 4
    x = feats(n, d, seed=11)
    hops = LaplacianGraphOp(K, r=0.5).propagate(adj, x)
    out = {"hops": np.stack([h.numpy() for h in hops])}
    out["last"] = LastMessageOp().aggregate(hops).numpy()
    for (s, e) in [(0, K + 1), (1, 4)]:
        tag = f"_{s}_{e}"
        out["sum" + tag] = SumMessageOp(s, e).aggregate(hops).numpy()
        out["mean" + tag] = MeanMessageOp(s, e).aggregate(hops).numpy()
        out["max" + tag] = MaxMessageOp(s, e).aggregate(hops).numpy()
        out["min" + tag] = MinMessageOp(s, e).aggregate(hops).numpy()
        out["concat" + tag] = ConcatMessageOp(s, e).aggregate(hops).numpy()
        out["alpha0.85" + tag] = SimpleWeightedMessageOp(s, e, "alpha", 0.85).aggregate(hops).numpy()
        out["alpha0.1" + tag] = SimpleWeightedMessageOp(s, e, "alpha", 0.1).aggregate(hops).numpy()
    hw = [0.5, -0.25, 1.5, 0.125, 2.0]
    out["hand_weights"] = np.asarray(hw, dtype=np.float32)
    out["hand_0_5"] = SimpleWeightedMessageOp(0, K + 1, "hand_crafted", hw).aggregate(hops).numpy()
    # NAFS over-smoothing distance weights: the reference loops over nodes in python (over_smooth_distance_op.py:27-31)
    out["osd"] = OverSmoothDistanceWeightedOp().aggregate(hops).numpy()
    # learnable family, forward values with the randomly initialised parameters stored alongside
    B = 37
    idx = torch.arange(0, n, 8)[:B]
    batch = [h[idx] for h in hops]
    out["batch_idx"] = idx.numpy()
    for kind, args in [("simple", (K,)), ("simple_allow_neg", (K,)), ("gate", (d,)), ("ori_ref", (d,)),
                       ("jk", (K, d))]:
        for (s, e) in [(0, K + 1), (1, 4)]:
            op = LearnableWeightedMessageOp(s, e, kind, *args)
            lw = op._LearnableWeightedMessageOp__learnable_weight
            tag = f"lw_{kind}_{s}_{e}"
            if isinstance(lw, torch.nn.Linear):
                with torch.no_grad():
                    lw.weight.mul_(3.0)  # spread the scores so the softmax is not near-uniform
                out[tag + "_w"] = lw.weight.detach().numpy().copy()
                out[tag + "_b"] = lw.bias.detach().numpy().copy()
            else:
                out[tag + "_w"] = lw.detach().numpy().copy()
            res = op.aggregate(batch)
            out[tag + "_out"] = res.detach().numpy()
            # gradients w.r.t. the parameters and the inputs for the autograd parity of the fused op
            feats_req = [b.clone().requires_grad_(True) for b in batch]
            res = op.aggregate(feats_req)
            gout = torch.linspace(-1, 1, res.numel()).reshape(res.shape)
            res.backward(gout)
            out[tag + "_gin"] = np.stack([f.grad.numpy() if f.grad is not None else np.zeros_like(f.detach().numpy())
                                          for f in feats_req])
            if isinstance(lw, torch.nn.Linear):
                out[tag + "_gw"] = lw.weight.grad.numpy().copy()
                out[tag + "_gb"] = lw.bias.grad.numpy().copy()
            else:
                out[tag + "_gw"] = lw.grad.numpy().copy()
    op = IterateLearnableWeightedMessageOp(0, K + 1, "recursive", d)
    lw = op._IterateLearnableWeightedMessageOp__learnable_weight
    out["iter_w"] = lw.weight.detach().numpy().copy()
    out["iter_b"] = lw.bias.detach().numpy().copy()
    out["iter_out"] = op.aggregate(batch).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "message_ops.npz"), **out)
    print("wrote message_ops", len(out), "arrays")


def gen_trick_cases():
    """label propagation (sgl/tricks/utils.py:40-58), an adjacent consumer of the same hop."""
    from sgl.tricks.utils import adj_to_symmetric_norm as tricks_norm, label_propagation
    adj = graphs()["skewed200"]
    n = adj.shape[0]
    gen = torch.Generator().manual_seed(5)
    labels = torch.randint(0, 6, (n,), generator=gen)
    mask = torch.rand(n, generator=gen) < 0.4
    norm = tricks_norm(adj, 0.5)
    out = {"labels": labels.numpy(), "mask": mask.numpy()}
    out["lp_masked"] = label_propagation(labels, norm, 4, 0.75, mask=mask).numpy()
    out["lp_full"] = label_propagation(labels, norm, 3, 0.5).numpy()
    np.savez_compressed(os.path.join(OUT, "tricks.npz"), **out)
    print("wrote tricks", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--tricks-only" not in sys.argv:
        gen_graph_cases()
        gen_message_cases()
    gen_trick_cases()

"""io.py -- graph inputs in the on-disk layout of the reference's user-defined dataset.

`Custom_Homo` (reference sgl/dataset/custom_dataset.py:12-87) reads, under ``root/name/raw/``:
    x.npy              [N, d] features                                   (:39-40)
    adj_matrix.npz     COO adjacency with arrays row, col, data          (:52-54)
    label.npy          [N] labels or [N, C] one-hot                      (:60-64, optional)
    indices.npz        train_idx / val_idx / test_idx                    (:76-85, optional)
and builds `Graph(row, col, edge_weight, ...)`, whose adjacency is csr_matrix((w, (row, col))) with duplicates summed
(sgl/data/base_data.py:29-30).  write_custom_homo() emits exactly these files, so the synthetic / R-MAT inputs of the
benchmark (and any graph held as scipy or COO arrays) can be fed to an unmodified SGL through Custom_Homo;
read_custom_homo() loads the same directory straight into what the hot path takes: (scipy CSR float32, features, labels,
splits) -- without SGL's pickled Graph and without torch_geometric.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import scipy.sparse as sp


def raw_dir(root: str, name: str) -> str:
    return os.path.join(root, name, "raw")


def write_custom_homo(root: str, name: str, adj, x: Optional[np.ndarray] = None, labels: Optional[np.ndarray] = None,
                      train_idx=None, val_idx=None, test_idx=None) -> str:
    """adj: scipy sparse matrix (any format) or a (row, col, data) triple.  Returns the raw directory."""
    d = raw_dir(root, name)
    os.makedirs(d, exist_ok=True)
    if isinstance(adj, tuple):
        row, col, data = (np.asarray(a) for a in adj)
    else:
        coo = adj.tocoo()
        row, col, data = coo.row, coo.col, coo.data
    np.savez(os.path.join(d, "adj_matrix.npz"), row=np.asarray(row, dtype=np.int64), col=np.asarray(col, dtype=np.int64),
             data=np.asarray(data, dtype=np.float32))
    if x is not None:
        np.save(os.path.join(d, "x.npy"), np.ascontiguousarray(x, dtype=np.float32))
    if labels is not None:
        np.save(os.path.join(d, "label.npy"), np.asarray(labels))
    splits = {k: np.asarray(v, dtype=np.int64) for k, v in
              (("train_idx", train_idx), ("val_idx", val_idx), ("test_idx", test_idx)) if v is not None}
    if splits:
        np.savez(os.path.join(d, "indices.npz"), **splits)
    return d


def read_custom_homo(root: str, name: str, num_node: int = 0):
    """-> dict(adj=scipy CSR float32 with duplicates summed, x, y, train_idx, val_idx, test_idx); the same checks and
    errors as Custom_Homo._process."""
    d = raw_dir(root, name)
    x = np.load(os.path.join(d, "x.npy")) if os.path.exists(os.path.join(d, "x.npy")) else None
    if x is not None:
        if num_node:
            assert num_node == x.shape[0], 'every node should have a feature vector'
        else:
            num_node = x.shape[0]
    elif not num_node:
        raise ValueError('please provide either feature matrix or number of node')
    path = os.path.join(d, "adj_matrix.npz")
    if not os.path.exists(path):
        raise ValueError('the adjacency matrix in coo-format is necessary')
    f = np.load(path)
    adj = sp.csr_matrix((f["data"].astype(np.float32), (f["row"], f["col"])), shape=(num_node, num_node))
    adj.sum_duplicates()
    y = None
    if os.path.exists(os.path.join(d, "label.npy")):
        y = np.load(os.path.join(d, "label.npy"))
        if y.ndim == 2:
            y = np.argmax(y, 1)
    out = {"adj": adj, "x": x, "y": y, "train_idx": None, "val_idx": None, "test_idx": None}
    if os.path.exists(os.path.join(d, "indices.npz")):
        s = np.load(os.path.join(d, "indices.npz"))
        for k in ("train_idx", "val_idx", "test_idx"):
            if k in s:
                out[k] = s[k]
    return out

"""cache.py -- on-disk cache of propagated features.

The reference recomputes  [X, A^X, ..., A^^K X]  on every run: only the raw Graph is pickled (sgl/dataset/ogbn.py:53-59),
propagated features never reach the disk.  Here the K+1 hop matrices of one (adjacency, operator, features) triple are
stored as plain .npy shards (one per hop, memory-mappable, readable from any numpy) under a key derived from the
content: sha1 of the adjacency's CSR arrays, the operator kind / r / alpha / prop_steps, and the feature matrix.

    op = LaplacianGraphOp(3); op.cache_dir = "/data/sgap_cache"      # opt-in
    hops = op.propagate(adj, x)       # first call computes on the GPU and writes; later calls (or runs) read

A hit returns CPU tensors without touching the GPU.  Invalidation is by content: any change of A, X or the operator
parameters changes the key.
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import List, Optional

import numpy as np
import torch

FORMAT_VERSION = 1


def _sha1_arrays(*arrays) -> str:
    h = hashlib.sha1()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(memoryview(a).cast("B"))
    return h.hexdigest()


def graph_fingerprint(adj) -> str:
    csr = adj.tocsr()
    return _sha1_arrays(np.asarray(csr.shape), csr.indptr, csr.indices, csr.data)


def feature_fingerprint(x) -> str:
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return _sha1_arrays(x)


class HopCache:
    def __init__(self, directory: str):
        self.dir = directory
        os.makedirs(directory, exist_ok=True)

    @staticmethod
    def key(adj, x, kind: str, prop_steps: int, **params) -> str:
        spec = json.dumps({"v": FORMAT_VERSION, "kind": kind, "K": int(prop_steps),
                           "params": {k: (None if v is None else float(v)) for k, v in sorted(params.items())},
                           "graph": graph_fingerprint(adj), "x": feature_fingerprint(x)}, sort_keys=True)
        return hashlib.sha1(spec.encode()).hexdigest()

    def _path(self, key: str) -> str:
        return os.path.join(self.dir, key)

    def load(self, key: str, mmap: bool = False) -> Optional[List[torch.Tensor]]:
        path = self._path(key)
        meta_file = os.path.join(path, "meta.json")
        if not os.path.exists(meta_file):
            return None
        meta = json.load(open(meta_file))
        if meta.get("v") != FORMAT_VERSION:
            return None
        hops = []
        for k in range(meta["n_hops"]):
            arr = np.load(os.path.join(path, f"hop_{k}.npy"), mmap_mode="r" if mmap else None)
            if list(arr.shape) != meta["shape"] or arr.dtype != np.float32:
                return None
            hops.append(torch.from_numpy(np.ascontiguousarray(arr)) if not mmap else torch.from_numpy(np.array(arr)))
        return hops

    def save(self, key: str, hops) -> None:
        path = self._path(key)
        tmp = path + ".tmp%d" % os.getpid()
        os.makedirs(tmp, exist_ok=True)
        for k, h in enumerate(hops):
            np.save(os.path.join(tmp, f"hop_{k}.npy"), h.detach().cpu().numpy().astype(np.float32, copy=False))
        with open(os.path.join(tmp, "meta.json"), "w") as f:
            json.dump({"v": FORMAT_VERSION, "n_hops": len(hops), "shape": list(hops[0].shape)}, f)
        import shutil
        if os.path.exists(path):        # another process won the race: keep theirs
            shutil.rmtree(tmp, ignore_errors=True)
        else:
            try:
                os.replace(tmp, path)
            except OSError:             # lost the race between the check and the rename
                shutil.rmtree(tmp, ignore_errors=True)

"""pytest configuration: markers, path setup and shared fixtures for the SGAP propagate/aggregate parity suite."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a host without CUDA, so a plain `pytest` works everywhere."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_graph_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "graph_*.npz")))


@pytest.fixture(scope="session")
def message_golden():
    return dict(np.load(os.path.join(GOLDEN, "message_ops.npz")))


class SimpleCsr:
    """Duck-typed stand-in for scipy.sparse.csr_matrix when a test wants to avoid scipy objects."""
    format = "csr"

    def __init__(self, indptr, indices, data, shape):
        self.indptr, self.indices, self.data, self.shape = indptr, indices, data, tuple(int(s) for s in shape)


def load_graph(path):
    z = dict(np.load(path))
    adj = SimpleCsr(z["adj_indptr"], z["adj_indices"], z["adj_data"], z["adj_shape"])
    return z, adj


GRAPH_TAGS = [("lap_r0.5", "lap", 0.5, None), ("lap_r0.3", "lap", 0.3, None), ("lap_r0", "lap", 0.0, None),
              ("ppr_r0.5_a0.15", "ppr", 0.5, 0.15)]

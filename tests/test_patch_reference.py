"""CPU test: sgl_b200.patch re-routes an importable reference `sgl.operators` (only where the reference tree exists:
the build container; skipped on the GPU box, which has no /root/reference)."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch

REF = os.environ.get("SGL_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "sgl", "operators")), reason="reference tree not present")
def test_install_and_uninstall_round_trip():
    sys.path.insert(0, REF)
    try:
        import sgl.operators.base_op as ref_base
        import sgl.operators.utils as ref_utils
        from sgl.operators.graph_op import LaplacianGraphOp as RefLaplacian
        from sgl.operators.message_op import SumMessageOp as RefSum
        import sgl_b200.patch as patch
        from sgl_b200 import SglB200Error

        orig_propagate, orig_matmul, orig_combine = ref_base.GraphOp.propagate, ref_utils.csr_sparse_dense_matmul, RefSum._combine
        adj = sp.csr_matrix(np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]], dtype=np.float32))
        x = np.arange(6, dtype=np.float32).reshape(3, 2)
        want = [h.numpy() for h in RefLaplacian(2).propagate(adj, x)]       # the reference's own CPU path
        patch.install()
        try:
            assert ref_base.GraphOp.propagate is not orig_propagate
            assert ref_utils.csr_sparse_dense_matmul is not orig_matmul and RefSum._combine is not orig_combine
            op = RefLaplacian(2)                                             # unmodified reference class
            if torch.cuda.is_available():
                op.mode = "exact"
                got = [h.numpy() for h in op.propagate(adj, x)]
                assert all(np.array_equal(a, b) for a, b in zip(got, want))
            else:
                with pytest.raises(SglB200Error):                            # routed to the GPU path: no silent CPU result
                    op.propagate(adj, x)
                with pytest.raises(TypeError):                               # the reference's own argument errors survive
                    op.propagate(adj, [[1.0, 2.0]] * 3)
        finally:
            patch.uninstall()
        assert ref_base.GraphOp.propagate is orig_propagate and ref_utils.csr_sparse_dense_matmul is orig_matmul
        assert RefSum._combine is orig_combine
        again = [h.numpy() for h in RefLaplacian(2).propagate(adj, x)]
        assert all(np.array_equal(a, b) for a, b in zip(again, want))
    finally:
        sys.path.remove(REF)

"""sgl_b200 -- B200-native (sm_100a) SGAP propagate/aggregate path for PKU-DAIR/SGL.

Layout (only what the hot path needs):
  csrc/        hand-written CUDA kernels + the C ABI (include/sglb200.h) -> libsglb200.so
  _lib.py      ctypes binding of the C ABI
  runtime.py   CsrOperator: device-resident CSR handle, K-hop drivers, aggregation launchers
  operators/   host-side mirror of the reference's sgl.operators (same names, arguments, errors)
  sgap.py      BaseSGAPModel glue mirror (preprocess / forward contract) + SGC / GAMLP / NAFS style wiring
  patch.py     install(): routes an importable reference `sgl.operators` through this implementation
  graph_build.py  construction of the normalised adjacency on the device (sort / segment passes + the float64 value kernel)
  dist.py      1-D row partition of the operator over the GPUs of one box: halo plans, NVLink peer-memory exchange
               (CUDA IPC + our push / flag kernels) pipelined against the hop, NCCL fallback
  tricks.py    label propagation and the NAFS task-level feature construction on the same hop handle
  cache.py     content-keyed on-disk cache of propagated features
"""
from ._lib import SglB200Error  # noqa: F401

__version__ = "0.1.0"

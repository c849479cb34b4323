"""Builds sgl_b200/libsglb200.so for sm_100a with nvcc (in-tree; the .so travels to the GPU box, it is git-ignored).

    python sgl_b200/csrc/build.py [--force] [--verbose]

Called by __graft_entry__.build().  No torch headers are involved: the library is a plain C-ABI CUDA library.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
SOURCES = ["graph.cu", "build_adj.cu", "spmm.cu", "spmm_group.cu", "spmm_tma.cu", "propagate.cu", "aggregate.cu", "learnable.cu", "iterate.cu", "legacy.cu", "peer.cu"]
HEADERS = ["common.cuh", "spmm_common.cuh", "trace.cuh", os.path.join(ROOT, "include", "sglb200.h")]
OUT = os.path.join(PKG, "libsglb200.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "-I", os.path.join(ROOT, "include"), "-I", HERE,
         "--expt-relaxed-constexpr", "-DSGLB200_BUILD"]


def _digest(paths):
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    objs = []
    rebuilt = False
    for src in SOURCES:
        sp = os.path.join(HERE, src)
        op = os.path.join(OBJ, src.replace(".cu", ".o"))
        stamp = op + ".sha"
        dig = _digest([sp] + hdrs)
        if force or not os.path.exists(op) or not os.path.exists(stamp) or open(stamp).read() != dig:
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", op]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            with open(stamp, "w") as f:
                f.write(dig)
            rebuilt = True
        objs.append(op)
    if rebuilt or not os.path.exists(OUT):
        # host compiler: the distro g++ (the image's CC/CXX wrappers lack some specs)
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

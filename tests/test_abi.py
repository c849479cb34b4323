"""CPU tests: the C-ABI library loads and exports exactly what include/sglb200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from sgl_b200 import _lib


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "sglb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"SGLB200_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in sglb200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "python binding table and header disagree"


def test_version_and_error_channel():
    lib = _lib.load()
    assert lib.sglb200_version() == 100
    assert isinstance(_lib.last_error(), str)


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.sglb200_device_count() < 0
    assert "no CPU fallback" in _lib.last_error()
    h = ctypes.c_void_p()
    import numpy as np
    indptr = np.zeros(2, dtype=np.int64)
    st = lib.sglb200_graph_create(ctypes.byref(h), 1, 1, 0, indptr.ctypes.data, 1, None, None, 0, 0, 0, None)
    assert st == _lib.ERR_NO_DEVICE and not h.value


def test_argument_validation_without_gpu():
    lib = _lib.load()
    st = lib.sglb200_graph_create(None, 1, 1, 0, None, 1, None, None, 0, 0, 0, None)
    assert st == _lib.ERR_INVALID
    assert "NULL" in _lib.last_error()


def test_header_is_valid_c_and_cpp():
    """The boundary is a C ABI: the header must compile as plain C (and as C++) without any CUDA or torch include."""
    import subprocess
    import tempfile
    inc = os.path.join(ROOT, "include")
    for lang, comp, std in (("c", "gcc", "-std=c99"), ("c++", "g++", "-std=c++11")):
        with tempfile.NamedTemporaryFile("w", suffix=".c" if lang == "c" else ".cpp", delete=False) as f:
            f.write('#include "sglb200.h"\nint main(void) { return SGLB200_VERSION == sglb200_version() ? 0 : 1; }\n')
            src = f.name
        try:
            exe = "/usr/bin/" + comp if os.path.exists("/usr/bin/" + comp) else comp
            r = subprocess.run([exe, std, "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-I", inc, src],
                               capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
        finally:
            os.unlink(src)

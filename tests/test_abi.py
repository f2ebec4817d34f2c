"""CPU-side checks of the drop-in boundary: the library builds, loads, and exports every symbol
include/mitre_b200.h declares; the Python binding table matches the header."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    hdr = open(os.path.join(ROOT, "include", "mitre_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(mitre_[a-z0-9_A-Z]+)\s*\(", hdr))


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib)
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libmitre_b200.so does not export %s" % n


def test_binding_table_matches_header(built_lib):
    from mitre_b200 import _lib
    assert set(_lib.SIGNATURES) == header_symbols()
    L = _lib.lib()
    assert b"sm_100a" in L.mitre_version()
    assert L.mitre_error_string(0) == b"ok"
    assert L.mitre_error_string(-4) == b"matrix is not positive definite"


def test_workspace_size_queries(built_lib):
    from mitre_b200 import _lib
    L = _lib.lib()
    assert L.mitre_score_workspace_bytes(35, 4) >= 8 * (8 + 35 * 6)
    assert L.mitre_score_workspace_bytes(35, 25) == 0          # p > MITRE_MAX_P
    assert L.mitre_scan_workspace_bytes(10 ** 6) > 0
    assert L.mitre_lexsort_workspace_bytes(10 ** 6, 1) > 8 * 10 ** 6
    assert L.mitre_draw_workspace_bytes(10 ** 6) > 0


def test_argument_validation_without_gpu(built_lib):
    """Bad arguments are rejected before any CUDA call, so this runs on a CPU-only box."""
    from mitre_b200 import _lib
    L = _lib.lib()
    null = ctypes.c_void_p(0)
    assert L.mitre_score_all(null, null, 10, 1, null, null, null, null, 35, 2, 100.0, null, 0, null,
                             null) == -1
    assert L.mitre_pack_rows_f64(null, null, 5, 0, null) == -1
    assert L.mitre_lexsort_rows(null, null, 5, 2, 35, null, null, null, 0) == -1


def test_no_cpu_fallback_without_device(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mitre_b200 import device
    with pytest.raises(RuntimeError):
        device.pack_rows(np.zeros((3, 5), dtype=bool))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mitre_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "scoring_oracle" not in src and "liboracle" not in src, f

"""The C-ABI library loads and exports every symbol include/c2c_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "c2c_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(c2c_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("c2c_build_ccc", "c2c_group_aggregate", "c2c_mttkrp", "c2c_gram_hadamard", "c2c_mu_update",
                 "c2c_hals_update", "c2c_recon_sums", "c2c_cp_normalize", "c2c_ncp_run", "c2c_ncp_workspace_bytes",
                 "c2c_sumsq", "c2c_last_error", "c2c_version"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from cell2cell_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from cell2cell_b200._build import build
        build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), name
    assert set(_lib.SIGNATURES) == set(declared_symbols())
    assert _lib.lib().c2c_version() >= 100


def test_argument_errors_are_reported_without_a_gpu():
    from cell2cell_b200 import _lib
    L = _lib.lib()
    shape = (ctypes.c_int64 * 2)(3, 4)
    assert L.c2c_ncp_workspace_bytes(2, shape, 3, 0) == 0            # ndim < 3
    assert b"ndim" in L.c2c_last_error()
    shape4 = (ctypes.c_int64 * 4)(4, 30, 6, 6)
    assert L.c2c_ncp_workspace_bytes(4, shape4, 4, 0) > 0
    assert L.c2c_ncp_workspace_bytes(4, shape4, 4, 1) >= L.c2c_ncp_workspace_bytes(4, shape4, 4, 0)
    assert L.c2c_ncp_workspace_bytes(4, shape4, 1000, 0) == 0        # rank out of range
    with pytest.raises(_lib.C2CError):
        _lib.check(-1, "probe")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from cell2cell_b200._lib import C2CError
    from cell2cell_b200.tensor import PreBuiltTensor
    with pytest.raises(C2CError):
        PreBuiltTensor(np.ones((3, 4, 2, 2)), order_names=[list("abc"), list("abcd"), list("ab"), list("ab")])

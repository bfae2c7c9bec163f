"""ctypes binding of ``libc2c_b200.so`` (the C ABI declared in ``include/c2c_b200.h``).

There is no CPU implementation behind this module: if the shared library is missing or no CUDA device is present the
product raises -- it never falls back (contrast the reference's silent CPU fallback at cell2cell/tensor/tensor.py:205-214).
"""
import ctypes
import os
import threading

__all__ = ["lib", "check", "C2CError", "LIB_PATH", "SCORE", "AGG", "ALGO", "CVG", "MAX_RANK", "MAX_NDIM"]

# C2C_B200_LIB selects another build of the same library (kernel tuning variants); the default is the in-tree product build
LIB_PATH = os.environ.get("C2C_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libc2c_b200.so")

SCORE = {"expression_product": 0, "expression_mean": 1, "expression_gmean": 2}
AGG = {"min": 0, "mean": 1, "gmean": 2, "sum": 3}
ALGO = {"non_negative_cp": 0, "non_negative_cp_hals": 1}
CVG = {"abs_rec_error": 0, "rec_error": 1}
MAX_RANK = 128
MAX_NDIM = 8


class C2CError(RuntimeError):
    """An entry point of libc2c_b200.so returned non-zero (< 0 argument error, > 0 cudaError_t)."""


_c = ctypes
_p = _c.c_void_p
_i = _c.c_int
_i64 = _c.c_int64
_sz = _c.c_size_t

# name -> (restype, argtypes); one entry per declaration in include/c2c_b200.h
SIGNATURES = {
    "c2c_version": (_i, []),
    "c2c_last_error": (_c.c_char_p, []),
    "c2c_launch_count": (_c.c_longlong, []),
    "c2c_release_pending": (_i, []),
    "c2c_time_pass_kernel": (_i, [_p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _i, _i, _i, _c.POINTER(_c.c_float), _c.POINTER(_i), _p, _sz, _p]),
    "c2c_set_mma_min_rank": (_i, [_i]),
    "c2c_plane_contract": (_i, [_p, _p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _i, _p, _p, _sz, _p]),
    "c2c_fiber_contract": (_i, [_p, _p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _i, _p, _p, _sz, _p]),
    "c2c_build_ccc": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "c2c_group_aggregate": (_i, [_p, _p, _p, _p, _i, _i, _i, _i64, _i, _p, _p, _p]),
    "c2c_ncp_workspace_bytes": (_sz, [_i, _c.POINTER(_i64), _i, _i]),
    "c2c_sumsq": (_i, [_p, _i64, _p, _p, _sz, _p]),
    "c2c_mttkrp": (_i, [_p, _p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _i, _i, _p, _p, _sz, _p]),
    "c2c_mttkrp_from_contraction": (_i, [_p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _i, _i, _p, _p]),
    "c2c_gram_hadamard": (_i, [_c.POINTER(_p), _i, _c.POINTER(_i64), _i, _i, _i, _p, _p, _p, _sz, _p]),
    "c2c_gram_hadamard_update": (_i, [_c.POINTER(_p), _i, _c.POINTER(_i64), _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "c2c_mu_update": (_i, [_p, _p, _p, _i64, _i, _i, _c.c_float, _p, _sz, _p]),
    "c2c_mu_update_coupled": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _i, _c.c_float, _p, _p]),
    "c2c_coupled_ncp_run": (_i, [_p, _p, _i, _p, _p, _p, _p, _i, _p, _p, _i, _i, _p, _i, _i, _c.c_double, _i, _c.c_double,
                                 _c.c_double, _p, _p, _p, _i, _p, _sz, _p, _sz, _p]),
    "c2c_hals_update": (_i, [_p, _p, _p, _i64, _i, _i, _i, _c.c_float, _p]),
    "c2c_recon_sums": (_i, [_p, _p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _i, _p, _p, _sz, _p]),
    "c2c_missing_rec_sumsq": (_i, [_p, _p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _i, _p, _p, _sz, _p]),
    "c2c_cp_normalize": (_i, [_c.POINTER(_p), _i, _c.POINTER(_i64), _i, _i, _p, _p]),
    "c2c_admm_update": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _c.c_double, _i, _c.c_double, _p, _p]),
    "c2c_corrindex_pairs": (_i, [_p, _i, _i64, _i, _p, _i, _i, _c.c_double, _p, _p, _p]),
    "c2c_ncp_run": (_i, [_p, _p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _i, _i, _i, _c.c_double, _i, _p, _p, _p, _i,
                         _p, _sz, _p]),
    "c2c_ncp_run_batched": (_i, [_p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _c.POINTER(_i), _i, _i, _c.c_double, _i, _p, _p, _p, _i,
                                 _p, _sz, _p]),
    "c2c_ncp_multi_workspace_bytes": (_sz, [_i, _c.POINTER(_i64), _c.POINTER(_i), _i]),
    "c2c_ncp_run_multi": (_i, [_p, _i, _c.POINTER(_i64), _c.POINTER(_p), _c.POINTER(_i), _i, _i, _c.c_double, _i, _p, _p, _p, _p,
                               _p, _sz, _p]),
    "c2c_shard_comm_bytes": (_sz, [_i, _c.POINTER(_i64), _i, _i]),
    "c2c_enable_peer_access": (_i, [_i]),
    "c2c_comm_alloc": (_i, [_sz, _c.POINTER(_p), _c.POINTER(_c.c_ubyte)]),
    "c2c_comm_open": (_i, [_c.POINTER(_c.c_ubyte), _c.POINTER(_p)]),
    "c2c_comm_close": (_i, [_p]),
    "c2c_comm_free": (_i, [_p]),
    "c2c_comm_probe": (_i, [_i, _i, _c.POINTER(_p), _sz, _c.c_uint, _c.c_float, _p, _p]),
    "c2c_ncp_run_sharded": (_i, [_p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _i, _i, _c.c_double, _i, _p, _p, _p, _i, _p, _sz,
                                 _i, _i, _c.POINTER(_p), _sz, _c.c_uint, _p]),
    "c2c_mask_plan_header_bytes": (_sz, [_i, _c.POINTER(_i64)]),
    "c2c_mask_plan_scan": (_i, [_p, _p, _i, _c.POINTER(_i64), _p, _sz, _c.POINTER(_i64), _p]),
    "c2c_mask_plan_fill": (_i, [_p, _i, _c.POINTER(_i64), _p, _sz, _c.POINTER(_i64), _p, _sz, _p]),
    "c2c_ncp_run_planned": (_i, [_p, _p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _i, _i, _c.c_double, _i, _p, _p, _p, _i,
                                 _p, _sz, _p, _sz, _p, _sz, _c.POINTER(_i64), _p]),
    "c2c_masked_contract_planned": (_i, [_p, _i, _c.POINTER(_i64), _c.POINTER(_p), _i, _i, _i, _p, _p, _sz, _p, _sz, _p, _sz,
                                         _c.POINTER(_i64), _p]),
}

_lock = threading.Lock()
_lib = None


def lib():
    """The loaded library (loaded once).  Raises if it has not been built: run ``python -m cell2cell_b200._build``."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise C2CError("%s is missing: build it with `python -m cell2cell_b200._build` (nvcc, sm_100a). "
                                   "cell2cell_b200 has no CPU fallback." % LIB_PATH)
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().c2c_last_error().decode("utf-8", "replace")
        err = C2CError("%s failed (%d): %s" % (what, rc, msg))
        err.code = int(rc)
        raise err

"""Import the *reference's own* tensor-build code (read-only, /root/reference) for golden-vector generation.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module.  It is used by
`tests/golden/make_golden.py` (the committed generator of the golden fixtures) and, when /root/reference is
present (this container, never the GPU box), by the CPU tests that re-check the numpy restatement in
`oracle/build_oracle.py` against the live reference.

How: `cell2cell/__init__.py` of the reference imports matplotlib/seaborn/tensorly/kneed, all absent from this image
(SURVEY.md section 8c).  We therefore register *empty* package stubs whose ``__path__`` points at the reference
directories (so the package ``__init__`` files never run) plus dummy ``tensorly`` / ``kneed`` /
``cell2cell.plotting.tensor_plot`` modules, and then import ``cell2cell.tensor.tensor`` by name.  The code that runs
for the build path -- tensor.py:796-883, 973-1252, core/communication_scores.py:159-244,
preprocessing/rnaseq.py:144-196, preprocessing/ppi.py:185-244, 370-523, preprocessing/find_elements.py:32-76 --
is the reference's, byte for byte.
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("C2C_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cell2cell", "tensor"))


def _stub_package(name, path):
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


_loaded = {}


def _add_coupled_shim(tl):
    """What ``cell2cell/tensor/coupled_factorization.py`` needs from tensorly (its imports at :4-12 and the ``tl.*`` calls
    of :222-482), on numpy.  The array helpers are numpy's; ``initialize_cp`` / ``unfolding_dot_khatri_rao`` /
    ``error_calc`` / ``cp_normalize`` / ``cp_to_tensor`` are the restatements of ``oracle/ncp_oracle.py`` (tensorly itself is
    not installable here), so a run of the reference's function under this shim pins the reference's coupling logic, not
    tensorly's primitives."""
    from . import ncp_oracle as o
    tl.eps = lambda dtype: np.finfo(dtype).eps
    tl.ndim = np.ndim
    tl.shape = np.shape
    tl.ones = lambda shape, **k: np.ones(shape, dtype=k.get("dtype", np.float64))
    tl.dot = np.dot
    tl.transpose = np.transpose
    tl.reshape = np.reshape
    tl.clip = lambda x, a_min=None, a_max=None: np.clip(x, a_min, a_max)
    tl.copy = np.copy
    tl.abs = np.abs
    tl.cp_to_tensor = lambda cp, mask=None: o.cp_to_tensor(cp[0], cp[1], mask=mask)

    def initialize_cp(tensor, rank, init="svd", svd="truncated_svd", non_negative=False, random_state=None,
                      normalize_factors=False, mask=None, svd_mask_repeats=5):
        assert non_negative and not normalize_factors
        return o.CPResult(np.ones(rank), o.initialize_cp(np.asarray(tensor, dtype=np.float64), rank, init=init,
                                                         random_state=random_state))

    def error_calc(tensor, norm_tensor, weights, factors, sparsity, mask, mttkrp=None):
        # tensorly.decomposition._cp.error_calc: "If mttkrp=None or masking is being performed, then the full tensor must be
        # constructed": the DIRECT norm of (tensor - reconstruction); the Gram form is used only when an MTTKRP is passed in.
        # The reference calls it with mttkrp=None (coupled_factorization.py:420-431) on the tensor it has been re-imputing.
        assert not sparsity and mask is None
        if mttkrp is None:
            low_rank = o.cp_to_tensor(weights, factors)
            return np.sqrt(np.sum((np.asarray(tensor, dtype=np.float64) - low_rank) ** 2)), tensor, norm_tensor
        iprod = np.sum(np.sum(mttkrp * factors[-1], axis=0) * weights)
        return np.sqrt(np.abs(norm_tensor ** 2 + o.cp_norm(weights, factors) ** 2 - 2 * iprod)), tensor, norm_tensor

    cp_mod = types.ModuleType("tensorly.decomposition._cp")
    cp_mod.initialize_cp = initialize_cp
    cp_mod.error_calc = error_calc
    ct = types.ModuleType("tensorly.cp_tensor")
    ct.CPTensor = lambda cp: o.CPResult(cp[0], cp[1])
    ct.unfolding_dot_khatri_rao = lambda tensor, cp, mode: o.mttkrp(tensor, cp[0], cp[1], mode)
    ct.cp_normalize = lambda cp: o.cp_normalize(cp[0], cp[1])
    ct.validate_cp_rank = lambda shape, rank="same", rounding="round": int(rank)
    tl.cp_tensor = ct
    sys.modules["tensorly.decomposition._cp"] = cp_mod
    sys.modules["tensorly.cp_tensor"] = ct


def load_reference_coupled():
    """The reference's own ``cell2cell/tensor/coupled_factorization.py`` (file imported unmodified) on the tensorly shim."""
    load_reference()
    if "coupled_factorization" not in _loaded:
        _loaded["coupled_factorization"] = importlib.import_module("cell2cell.tensor.coupled_factorization")
    return _loaded["coupled_factorization"]


def load_reference():
    """Returns a namespace with the reference's build-path modules (tensor, communication_scores, rnaseq, ppi)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    pkg = os.path.join(REFERENCE_ROOT, "cell2cell")
    saved = {k: sys.modules.get(k) for k in ("tensorly", "tensorly.decomposition", "kneed")}
    _stub_package("cell2cell", pkg)
    for sub in ("core", "preprocessing", "tensor", "plotting", "io"):
        _stub_package("cell2cell." + sub, os.path.join(pkg, sub))

    tl = types.ModuleType("tensorly")
    tl.tensor = lambda x, **k: np.asarray(x)
    tl.to_numpy = lambda x: np.asarray(x)
    tl.get_backend = lambda: "numpy"
    tl.context = lambda x: {"dtype": np.asarray(x).dtype}
    tl.sum = np.sum
    tl.prod = np.prod
    tl.mean = np.mean
    tl.norm = lambda x, order=2: np.sqrt(np.sum(np.asarray(x, dtype=float) ** 2))
    _add_coupled_shim(tl)
    dec = types.ModuleType("tensorly.decomposition")
    for n in ("non_negative_parafac", "non_negative_parafac_hals", "parafac", "constrained_parafac"):
        setattr(dec, n, None)
    tl.decomposition = dec
    sys.modules["tensorly"] = tl
    sys.modules["tensorly.decomposition"] = dec
    kneed = types.ModuleType("kneed")
    kneed.KneeLocator = None
    sys.modules["kneed"] = kneed
    tp = types.ModuleType("cell2cell.plotting.tensor_plot")
    tp.plot_elbow = None
    tp.plot_multiple_run_elbow = None
    sys.modules["cell2cell.plotting.tensor_plot"] = tp

    _loaded["tensor"] = importlib.import_module("cell2cell.tensor.tensor")
    _loaded["factorization"] = importlib.import_module("cell2cell.tensor.factorization")
    _loaded["metrics"] = importlib.import_module("cell2cell.tensor.metrics")
    _loaded["communication_scores"] = importlib.import_module("cell2cell.core.communication_scores")
    _loaded["rnaseq"] = importlib.import_module("cell2cell.preprocessing.rnaseq")
    _loaded["ppi"] = importlib.import_module("cell2cell.preprocessing.ppi")
    _loaded["find_elements"] = importlib.import_module("cell2cell.preprocessing.find_elements")
    _loaded["signal"] = importlib.import_module("cell2cell.preprocessing.signal")
    return types.SimpleNamespace(**_loaded)

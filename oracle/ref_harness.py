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

"""Golden fixtures of the factorization-side API, produced by the REFERENCE'S OWN drivers.

What runs here is the reference's code, files imported unmodified through ``oracle/ref_harness.load_reference()``:

* ``cell2cell/tensor/factorization.py``: ``_compute_tensor_factorization`` (:14-143), ``_run_elbow_analysis`` (:186-268),
  ``_multiple_runs_elbow_analysis`` (:271-373, both metrics -- the similarity metric goes through the reference's own
  ``cell2cell/tensor/metrics.py`` CorrIndex), ``_compute_norm_error`` / ``normalized_error`` (:376-458);
* ``cell2cell/tensor/tensor.py``: ``PreBuiltTensor.__init__`` (:886-968) and ``BaseTensor.compute_tensor_factorization``
  (:216-362: best-of-runs, ``cp_normalize``, weight ordering, ``explained_variance_ratio_``, the masked
  ``explained_variance`` quirk of :659-688, the factor DataFrames).

The only thing injected is tensorly's decomposition entry points (``non_negative_parafac`` / ``_hals`` / ``parafac``), which
are the numpy restatements of ``oracle/ncp_oracle.py`` -- tensorly itself is not installable here, so this pins the
reference's DRIVERS and post-processing, not tensorly's primitives (those stay "unpinned", see oracle/ncp_oracle.py).

The second half holds full-size oracle runs that are too slow to recompute inside a test (the fp64 numpy oracle on BASELINE
configs 3, 4 and 5-shaped tensors); the inputs are regenerated from seeds by ``tests/cases.ncp_golden_inputs``.

Run in the build container only:   python tests/golden/make_golden_ncp.py [--only api|full]   ->  tests/golden/ncp_api.npz,
tests/golden/ncp_full.npz
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ncp_oracle as o                                    # noqa: E402
from oracle.ref_harness import load_reference                          # noqa: E402
from tests.cases import NCP_API_CASES, NCP_FULL_CASES, ncp_api_inputs, ncp_golden_inputs  # noqa: E402


def inject_decompositions(ref):
    """tensorly.decomposition.* as the reference's factorization.py imported them (factorization.py:8), on the oracle."""
    def nn_parafac(tensor, rank, init="svd", svd="truncated_svd", random_state=None, mask=None, n_iter_max=100, tol=10e-7,
                   verbose=False, return_errors=False, **kw):
        assert not kw, kw
        cp, errs = o.non_negative_parafac(np.asarray(tensor, dtype=np.float64), rank, n_iter_max=n_iter_max, init=init, tol=tol,
                                          random_state=random_state, mask=None if mask is None else np.asarray(mask, float))
        return (cp, errs) if return_errors else cp

    def nn_hals(tensor, rank, init="svd", svd="truncated_svd", random_state=None, n_iter_max=100, tol=10e-8, verbose=False,
                return_errors=False, **kw):
        assert not kw, kw
        cp, errs = o.non_negative_parafac_hals(np.asarray(tensor, dtype=np.float64), rank, n_iter_max=n_iter_max, init=init,
                                               tol=tol, random_state=random_state)
        return (cp, errs) if return_errors else cp

    def parafac(tensor, rank, init="svd", svd="truncated_svd", random_state=None, mask=None, n_iter_max=100, tol=1e-8,
                verbose=False, return_errors=False, **kw):
        assert not kw, kw
        cp, errs = o.parafac(np.asarray(tensor, dtype=np.float64), rank, n_iter_max=n_iter_max, init=init, tol=tol,
                             random_state=random_state, mask=None if mask is None else np.asarray(mask, float))
        return (cp, errs) if return_errors else cp

    ref.factorization.non_negative_parafac = nn_parafac
    ref.factorization.non_negative_parafac_hals = nn_hals
    ref.factorization.parafac = parafac


def make_api(store):
    ref = load_reference()
    inject_decompositions(ref)
    F, T = ref.factorization, ref.tensor
    for name, case in NCP_API_CASES.items():
        x, mask, names = ncp_api_inputs(name)
        x64 = x.astype(np.float64)
        m64 = None if mask is None else mask.astype(np.float64)
        kind = case["kind"]
        kw = dict(case["kwargs"])
        if kind == "factorize":
            t = T.PreBuiltTensor(x64, order_names=names, mask=m64)
            t.compute_tensor_factorization(**kw)
            for i, (label, df) in enumerate(t.factors.items()):
                store["%s/factor%d" % (name, i)] = df.values.astype(np.float64)
            store[name + "/labels"] = np.array(list(t.factors.keys()))
            store[name + "/columns"] = np.array(list(t.factors[label].columns))
            evr = t.explained_variance_ratio_
            store[name + "/explained_variance_ratio"] = np.asarray([] if evr is None else evr, dtype=np.float64)
            store[name + "/explained_variance"] = np.float64(t.explained_variance_)
            store[name + "/raw_weights"] = np.asarray(t.tl_object[0], dtype=np.float64)
            for i, f in enumerate(t.tl_object[1]):
                store["%s/raw_factor%d" % (name, i)] = np.asarray(f, dtype=np.float64)
            store[name + "/norm_error"] = np.float64(F._compute_norm_error(t.tensor, t.tl_object, t.mask))
            store[name + "/fractions"] = np.asarray([t.excluded_value_fraction(), t.sparsity_fraction(), t.missing_fraction()],
                                                    dtype=np.float64)
        elif kind == "elbow_single":
            loss = F._run_elbow_analysis(x64, mask=m64, disable_pbar=True, **kw)
            store[name + "/loss"] = np.asarray([[r, float(l)] for r, l in loss], dtype=np.float64)
        elif kind == "elbow_multi":
            store[name + "/all_loss"] = np.asarray(F._multiple_runs_elbow_analysis(x64, mask=m64, **kw), dtype=np.float64)
        else:
            raise ValueError(kind)
        print(name, "done")


def make_full(store):
    for name, case in NCP_FULL_CASES.items():
        x, mask = ncp_golden_inputs(name)
        t0 = time.time()
        x64 = x.astype(np.float64)
        init = o.random_init(x.shape, case["rank"], case["seed"])
        cp, errs = o.non_negative_parafac(x64, case["rank"], n_iter_max=case["n_iter"], init=(np.ones(case["rank"]), init),
                                          tol=-1.0, mask=None if mask is None else mask.astype(np.float64))
        for i, f in enumerate(cp.factors):
            store["%s/factor%d" % (name, i)] = np.asarray(f, dtype=np.float64)
        store[name + "/errors"] = np.asarray(errs, dtype=np.float64)
        if mask is not None:
            store[name + "/masked_norm_error"] = np.float64(o.masked_norm_error(x64, cp, mask.astype(np.float64)))
        print(name, x.shape, "rank", case["rank"], "%d iterations" % len(errs), "last error %.6f" % errs[-1],
              "%.0f s" % (time.time() - t0), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", choices=["api", "full"], default=None)
    args = ap.parse_args()
    if args.only in (None, "api"):
        store = {}
        make_api(store)
        np.savez_compressed(os.path.join(HERE, "ncp_api.npz"), **store)
    if args.only in (None, "full"):
        store = {}
        make_full(store)
        np.savez_compressed(os.path.join(HERE, "ncp_full.npz"), **store)


if __name__ == "__main__":
    main()

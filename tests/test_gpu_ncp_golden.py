"""GPU parity against the factorization-side goldens (run with ``-m gpu``).

``tests/golden/ncp_api.npz``: outputs of the REFERENCE'S OWN drivers -- ``BaseTensor.compute_tensor_factorization``
(tensor.py:216-362), ``_run_elbow_analysis`` / ``_multiple_runs_elbow_analysis`` (factorization.py:186-373, CorrIndex of
metrics.py included), ``_compute_norm_error`` (:376-432) -- run on the oracle's decomposition functions
(``tests/golden/make_golden_ncp.py``).  ``ncp_full.npz``: fp64 oracle runs at full BASELINE sizes (config 3 masked, config 4
at rank 20, the K = 30 family of config 5), too slow to recompute inside a test.  Tolerance: 1e-4 relative on factors and
kept errors (north_star)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from tests.cases import NCP_API_CASES, NCP_FULL_CASES, ncp_api_inputs, ncp_golden_inputs  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-4


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("name", [n for n, c in NCP_API_CASES.items() if c["kind"] == "factorize"])
def test_compute_tensor_factorization_matches_the_reference_driver(name, capsys):
    from cell2cell_b200.tensor import PreBuiltTensor
    from cell2cell_b200.tensor.factorization import _compute_norm_error
    api = np.load(os.path.join(GOLD, "ncp_api.npz"))
    case = NCP_API_CASES[name]
    x, mask, names = ncp_api_inputs(name)
    t = PreBuiltTensor(x, order_names=names, mask=mask)
    kw = dict(case["kwargs"])
    stop_rule = kw["tol"] > 0
    t.compute_tensor_factorization(**kw)
    assert list(t.factors.keys()) == list(api[name + "/labels"])
    tol = 2e-3 if stop_rule else RTOL                  # a real stop rule may end an iteration apart in fp32
    for i, (label, df) in enumerate(t.factors.items()):
        want = api["%s/factor%d" % (name, i)]
        assert list(df.columns) == list(api[name + "/columns"]) and list(df.index) == names[i]
        assert rel_err(df.values, want) < tol, (label, rel_err(df.values, want))
    evr = api[name + "/explained_variance_ratio"]
    if evr.size:
        np.testing.assert_allclose(t.explained_variance_ratio_, evr, rtol=tol)
    else:
        assert t.explained_variance_ratio_ is None
    assert abs(t.explained_variance_ - float(api[name + "/explained_variance"])) < max(tol, 1e-4)
    for i, f in enumerate(t.tl_object.factors):
        assert rel_err(f, api["%s/raw_factor%d" % (name, i)]) < tol
    e = float(_compute_norm_error(t.tensor, t.tl_object, t.mask))
    assert abs(e - float(api[name + "/norm_error"])) < tol * float(api[name + "/norm_error"])
    np.testing.assert_allclose([t.excluded_value_fraction(), t.sparsity_fraction(), t.missing_fraction()],
                               api[name + "/fractions"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name", [n for n, c in NCP_API_CASES.items() if c["kind"] != "factorize"])
def test_elbow_drivers_match_the_reference_drivers(name):
    from cell2cell_b200.tensor.factorization import _multiple_runs_elbow_analysis, _run_elbow_analysis
    api = np.load(os.path.join(GOLD, "ncp_api.npz"))
    case = NCP_API_CASES[name]
    x, mask, _ = ncp_api_inputs(name)
    kw = dict(case["kwargs"])
    if case["kind"] == "elbow_single":
        loss = _run_elbow_analysis(x, mask=mask, disable_pbar=True, **kw)
        want = api[name + "/loss"]
        assert [r for r, _ in loss] == [int(r) for r in want[:, 0]]
        np.testing.assert_allclose([float(l) for _, l in loss], want[:, 1], rtol=RTOL)
    else:
        got = _multiple_runs_elbow_analysis(x, mask=mask, **kw)
        want = api[name + "/all_loss"]
        assert got.shape == want.shape
        if kw["metric"] == "similarity":
            # 1 - CorrIndex of the runs' factor sets (metrics.py:12-179); CorrIndex amplifies the 1e-4 factor differences
            np.testing.assert_allclose(got, want, rtol=5e-3, atol=5e-4)
        else:
            np.testing.assert_allclose(got, want, rtol=RTOL)


@pytest.mark.parametrize("name", list(NCP_FULL_CASES))
def test_full_size_runs_match_the_oracle_goldens(name):
    """Whole runs at full BASELINE sizes from identical initial factors, fixed iteration count."""
    from cell2cell_b200.tensor.factorization import _compute_norm_error, non_negative_parafac
    full = np.load(os.path.join(GOLD, "ncp_full.npz"))
    case = NCP_FULL_CASES[name]
    x, mask = ncp_golden_inputs(name)
    X = torch.as_tensor(x).cuda()
    M = None if mask is None else torch.as_tensor(mask).cuda()
    del x
    got, errs = non_negative_parafac(X, case["rank"], n_iter_max=case["n_iter"], init="random", random_state=case["seed"], tol=-1.0,
                                     mask=M, return_errors=True)
    want_errs = full[name + "/errors"]
    assert len(errs) == len(want_errs) == case["n_iter"]
    for i, f in enumerate(got.factors):
        want = full["%s/factor%d" % (name, i)]
        assert rel_err(f, want) < RTOL, (i, rel_err(f, want))
    assert abs(errs[-1] - want_errs[-1]) <= RTOL * want_errs[-1]
    np.testing.assert_allclose(np.asarray(errs), want_errs, rtol=RTOL, atol=3e-6)
    if mask is not None:
        e = float(_compute_norm_error(X, got, M))
        assert abs(e - float(full[name + "/masked_norm_error"])) <= RTOL * float(full[name + "/masked_norm_error"])

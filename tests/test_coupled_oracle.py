"""Coupled factorization on CPU: (1) the numpy restatement ``oracle/coupled_oracle.py`` against the golden outputs of the
reference's own ``coupled_non_negative_parafac`` (``tests/golden/coupled.npz``, made by ``make_golden_coupled.py``) and,
when /root/reference is present, against the live reference; (2) the product's host-side orchestration
(``cell2cell_b200.tensor.coupled_factorization._coupled_mu``: update order, contraction caches, combined stop rule) with
numpy stand-ins for the kernels; (3) mode-mapping handling and validation errors (coupled_factorization.py:15-120)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from cell2cell_b200.tensor import coupled_factorization as cf      # noqa: E402
from oracle import coupled_oracle as co                            # noqa: E402
from oracle import ncp_oracle as o                                 # noqa: E402
from oracle.ref_harness import reference_available                 # noqa: E402
from tests.cases import COUPLED_CASES, coupled_inputs              # noqa: E402
from tests.numpy_ops import NumpyOps                               # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "coupled.npz"))


def _oracle_run(name):
    case = COUPLED_CASES[name]
    x1, x2, m1, m2 = coupled_inputs(name)
    return co.coupled_non_negative_parafac(x1, x2, case["rank"], case["mode_mapping"], mask1=m1, mask2=m2,
                                           n_iter_max=case["n_iter_max"], init="random", tol=case["tol"], random_state=7,
                                           **case.get("kwargs", {}))


@pytest.mark.parametrize("name", sorted(COUPLED_CASES))
def test_oracle_matches_reference_golden(name):
    cp1, cp2, (e1, e2) = _oracle_run(name)
    assert len(e1) == len(GOLD[name + "/errors1"]) == len(e2)
    np.testing.assert_allclose(e1, GOLD[name + "/errors1"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(e2, GOLD[name + "/errors2"], rtol=1e-12, atol=0)
    for k, cp in ((1, cp1), (2, cp2)):
        for i, f in enumerate(cp[1]):
            np.testing.assert_allclose(f, GOLD["%s/factors%d_%d" % (name, k, i)], rtol=1e-12, atol=1e-300)


@pytest.mark.skipif(not reference_available(), reason="/root/reference is only present in the build container")
def test_oracle_matches_live_reference():
    from oracle.ref_harness import load_reference_coupled
    ref = load_reference_coupled()
    rs = np.random.RandomState(3)
    x1, x2 = rs.rand(3, 6, 4, 5), rs.rand(6, 5, 8)
    mm = {"shared": [(1, 0), (3, 1)]}
    a1, a2, (ea1, ea2) = ref.coupled_non_negative_parafac(x1, x2, 2, mm, n_iter_max=12, init="svd", tol=1e-12,
                                                          return_errors=True)
    b1, b2, (eb1, eb2) = co.coupled_non_negative_parafac(x1, x2, 2, mm, n_iter_max=12, init="svd", tol=1e-12)
    np.testing.assert_allclose(ea1, eb1, rtol=1e-13)
    np.testing.assert_allclose(ea2, eb2, rtol=1e-13)
    for f, g in zip(a1[1] + a2[1], b1[1] + b2[1]):
        np.testing.assert_allclose(f, g, rtol=1e-13)


def test_shared_factors_stay_identical_and_errors_decrease():
    cp1, cp2, (e1, e2) = _oracle_run("different_orders")
    for a, b in COUPLED_CASES["different_orders"]["mode_mapping"]["shared"]:
        assert np.array_equal(cp1[1][a], cp2[1][b])
    comb = np.array(e1) + np.array(e2)
    assert comb[-1] < comb[0]
    assert all((f >= 0).all() for f in cp1[1] + cp2[1])


def _driver_run(name, ops=None):
    case = COUPLED_CASES[name]
    x1, x2, m1, m2 = coupled_inputs(name)
    kw = case.get("kwargs", {})
    mm = co.process_mode_mapping(x1.ndim, x2.ndim, case["mode_mapping"])
    shared = [tuple(p) for p in mm["shared"]]
    w1, w2 = cf._balance_weights(x1.shape, x2.shape, shared, kw.get("balance_errors", True))
    rank = case["rank"]
    ops = ops or NumpyOps(rank)
    t = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))   # noqa: E731
    s1, s2, e1, e2 = cf._coupled_mu(t(x1), t(x2), t(m1), t(m2), o.random_init(x1.shape, rank, 7), o.random_init(x2.shape, rank, 7),
                                    rank, shared, case["n_iter_max"], case["tol"], kw.get("cvg_criterion", "abs_rec_error"),
                                    w1, w2, ops)
    return s1, s2, e1, e2, ops


@pytest.mark.parametrize("name", sorted(COUPLED_CASES))
def test_driver_orchestration_matches_golden(name):
    """fp64 numpy kernels behind the product's driver: same factors, error histories and stop iteration as the reference.
    (The inner product of the error goes through the last shared update instead of a fresh last-mode MTTKRP -- equal in
    exact arithmetic, 1e-9 here.)"""
    s1, s2, e1, e2, ops = _driver_run(name)
    assert len(e1) == len(GOLD[name + "/errors1"])
    np.testing.assert_allclose(e1, GOLD[name + "/errors1"], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(e2, GOLD[name + "/errors2"], rtol=1e-7, atol=1e-10)
    for k, side in ((1, s1), (2, s2)):
        for i, a in enumerate(side.A):
            np.testing.assert_allclose(a.numpy(), GOLD["%s/factors%d_%d" % (name, k, i)], rtol=1e-9, atol=1e-12)


def test_unmasked_tensors_are_read_twice_per_iteration():
    """4-way tensors sharing contexts / senders / receivers: update order LR1, LR2, receivers, senders, contexts.  One plane
    pass serves the LR mode and (from the previous iteration) nothing else changes it; one fiber pass serves both cell
    modes; the context update needs a fresh plane pass -- which is still valid for the LR mode of the next iteration."""
    s1, s2, e1, e2, ops = _driver_run("same_order_int_mapping")
    n = COUPLED_CASES["same_order_int_mapping"]["n_iter_max"]
    assert s1.passes == s2.passes == 2 * n + 1
    assert ops.calls[:6] == ["plane", "plane", "fiber", "fiber", "plane", "plane"]
    assert ops.calls[6:10] == ["fiber", "fiber", "plane", "plane"]


def test_masked_tensor_is_reimputed_before_every_mode():
    s1, s2, e1, e2, ops = _driver_run("masked_first")
    n = COUPLED_CASES["masked_first"]["n_iter_max"]
    assert s1.passes == 4 * n and s2.passes == 2 * n + 1


def test_mode_mapping_forms_and_errors():
    a, b = np.zeros((2, 3, 4, 4)), np.zeros((2, 5, 4, 4))
    assert cf._process_mode_mapping(a, b, 1) == {"shared": [(0, 0), (2, 2), (3, 3)]}
    assert cf._process_mode_mapping(a, b, [1, 1]) == {"shared": [(0, 0), (2, 2), (3, 3)]}
    assert cf._process_mode_mapping(a, b, {"shared": [(0, 0)]}) == {"shared": [(0, 0)]}
    with pytest.raises(ValueError):
        cf._process_mode_mapping(a, np.zeros((2, 4, 4)), 1)
    with pytest.raises(ValueError):
        cf._process_mode_mapping(a, b, "1")
    cf._validate_tensors(a, b, {"shared": [(0, 0), (2, 2)]})
    with pytest.raises(ValueError, match="same size"):
        cf._validate_tensors(a, b, {"shared": [(1, 1)]})
    with pytest.raises(ValueError, match="At least one"):
        cf._validate_tensors(a, b, {"shared": []})
    with pytest.raises(ValueError, match="invalid tensor2"):
        cf._validate_tensors(a, b, {"shared": [(0, 7)]})

    class Named:
        def __init__(self, t, names):
            self.tensor, self.order_names = t, names
    n1 = Named(a, [["c1", "c2"], list("abc"), list("wxyz"), list("wxyz")])
    n2 = Named(b, [["c1", "c3"], list("abcde"), list("wxyz"), list("wxyz")])
    with pytest.raises(ValueError, match="Element names"):
        cf._validate_tensors(n1, n2, {"shared": [(0, 0)]})
    cf._validate_tensors(n1, n2, {"shared": [(2, 2), (3, 3)]})


def test_balance_weights_and_combined_error():
    sh = [(0, 0), (2, 2), (3, 3)]
    w1, w2 = cf._balance_weights((5, 12, 6, 6), (5, 30, 6, 6), sh)
    assert (w1, w2) == co.balance_weights((5, 12, 6, 6), (5, 30, 6, 6), sh) == (42 / 12, 42 / 30)
    assert cf._balance_weights((5, 12, 6, 6), (5, 30, 6, 6), sh, False) == (1.0, 1.0)
    assert cf._balance_weights((5, 10, 6, 6), (10, 6, 6), [(1, 0), (2, 1), (3, 2)]) == (6 / 5, 6.0)
    e = cf._combined_error(0.2, 0.4, (5, 12, 6, 6), (5, 30, 6, 6), {"shared": sh})
    assert e == pytest.approx(co.combined_error(0.2, 0.4, (5, 12, 6, 6), (5, 30, 6, 6), sh))
    assert cf._combined_error(0.2, 0.4, (5, 12, 6, 6), (5, 30, 6, 6), {"shared": sh}, False) == pytest.approx(0.3)

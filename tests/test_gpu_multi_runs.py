"""One-CTA-per-run factorization of small tensors (cell2cell_b200/csrc/c2c_multi.cu, c2c_ncp_run_multi) through the C ABI:
every run against the whole-GPU path (c2c_ncp_run) and the oracle from identical initial factors, the stop rule, and the
elbow analysis / best-of-runs drivers that use it (cell2cell/tensor/factorization.py:335-352, tensor.py:294-320)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ncp_oracle as o                                   # noqa: E402
from tests.test_gpu_parity import lowrank, rel_err                   # noqa: E402

pytestmark = pytest.mark.gpu

MULTI_SHAPES = [((4, 300, 6, 6), [4, 1, 7, 4]), ((12, 1054, 6, 6), [10, 25, 3]), ((5, 40, 7, 9), [3, 5]), ((30, 9, 11), [2, 6]),
                ((3, 4, 20, 8, 8), [5, 12]), ((6, 200, 17, 17), [10, 32]), ((200, 7, 16, 12), [8])]


@pytest.mark.parametrize("shape,ranks", MULTI_SHAPES)
def test_multi_runs_match_single_runs_and_oracle(shape, ranks):
    from cell2cell_b200 import engine
    from cell2cell_b200.tensor.factorization import non_negative_parafac
    x = lowrank(shape, 4, seed=7)
    X = torch.as_tensor(x).cuda()
    n_iter = 30
    inits = [o.random_init(shape, r, 100 + i) for i, r in enumerate(ranks)]
    res = engine.ncp_run_multi(X, inits, n_iter_max=n_iter, tol=-1.0)
    for i, (r, (fs, errs, it, conv, sums)) in enumerate(zip(ranks, res)):
        assert it == n_iter and not conv and len(errs) == n_iter
        ref, ref_err = non_negative_parafac(x, r, n_iter_max=n_iter, init=(np.ones(r), inits[i]), tol=-1.0, return_errors=True)
        for a, b in zip(fs, ref.factors):
            assert a.shape == b.shape and (a >= 0).all()
            assert rel_err(a, b) < 2e-5, rel_err(a, b)
        np.testing.assert_allclose(errs[:-1], np.asarray(ref_err)[:-1], rtol=1e-4, atol=3e-6)
        # the directly accumulated sums of the final factors
        want = engine.recon_sums(X, None, [engine.pack_factor(f, r, X.device) for f in fs], r)
        np.testing.assert_allclose(sums[1:], want[1:], rtol=1e-5)
        assert abs(np.sqrt(sums[1] / sums[3]) - ref_err[-1]) < 1e-4 * ref_err[-1]
    # first run also against the fp64 oracle
    oref, oerr = o.non_negative_parafac(x.astype(np.float64), ranks[0], n_iter_max=n_iter, init=(np.ones(ranks[0]), inits[0]),
                                        tol=-1.0)
    for a, b in zip(res[0][0], oref.factors):
        assert rel_err(a, b) < 1e-4
    np.testing.assert_allclose(res[0][1], oerr, rtol=1e-4, atol=3e-6)


def test_multi_runs_stop_on_their_own():
    from cell2cell_b200 import engine
    shape = (6, 80, 6, 6)
    x = lowrank(shape, 3, seed=2, noise=0.01)
    X = torch.as_tensor(x).cuda()
    ranks = [2, 3, 6]
    inits = [o.random_init(shape, r, 5 + i) for i, r in enumerate(ranks)]
    # tol well above the fp32 noise of the Gram-form error (~1e-6): the stop iteration is then the oracle's, give or take one
    res = engine.ncp_run_multi(X, inits, n_iter_max=300, tol=1e-4)
    for i, r in enumerate(ranks):
        _, oerr = o.non_negative_parafac(x.astype(np.float64), r, n_iter_max=300, init=(np.ones(r), inits[i]), tol=1e-4)
        fs, errs, it, conv, sums = res[i]
        assert conv and it == len(errs) < 300
        assert abs(it - len(oerr)) <= 2, (it, len(oerr))
        # the Gram-form error cancels in fp32 once it is small (||X||^2 + ||rec||^2 - 2<X, rec> from fp32 factors): the
        # SQUARED error carries an absolute uncertainty of a few fp32 eps, so that is what is compared
        n = min(it, len(oerr)) - 1
        np.testing.assert_allclose(errs[:n] ** 2, np.asarray(oerr[:n]) ** 2, rtol=2e-4, atol=5e-7)


def test_elbow_and_best_of_runs_use_the_multi_kernel_and_agree_with_run_by_run():
    from cell2cell_b200 import engine
    from cell2cell_b200.tensor import PreBuiltTensor
    from cell2cell_b200.tensor.factorization import _multi_eligible, _multiple_runs_elbow_analysis
    shape = (12, 200, 6, 6)
    x = lowrank(shape, 4, seed=9)
    X = torch.as_tensor(x).cuda()
    assert _multi_eligible(X, [r for r in range(1, 9) for _ in range(3)], 20)
    n0 = engine.launch_count()
    got = _multiple_runs_elbow_analysis(x, upper_rank=8, runs=3, init="random", random_state=3, n_iter_max=20, tol=-1.0)
    n_multi = engine.launch_count() - n0
    want = _multiple_runs_elbow_analysis(x, upper_rank=8, runs=3, init="random", random_state=3, n_iter_max=20, tol=-1.0,
                                         batch_runs=False)
    assert got.shape == want.shape == (3, 8)
    np.testing.assert_allclose(got, want, rtol=1e-4)
    assert n_multi < 20                                          # one launch (+ the norm of the tensor), not 24 x 20 x 6
    names = [list(range(s)) for s in shape]
    a = PreBuiltTensor(x, order_names=names)
    a.compute_tensor_factorization(rank=5, init="random", random_state=1, runs=4, n_iter_max=25, tol=-1.0)
    b = PreBuiltTensor(x, order_names=names)
    b.compute_tensor_factorization(rank=5, init="random", random_state=1, runs=4, n_iter_max=25, tol=-1.0, batch_runs=False)
    for k in a.factors:
        assert rel_err(a.factors[k].values, b.factors[k].values) < 1e-4
    assert abs(a.explained_variance_ - b.explained_variance_) < 1e-5

"""Self-consistency of the NCP oracle (oracle/ncp_oracle.py).  The arithmetic it restates lives in tensorly (not
installable here; the reference holds no golden vectors for it -- parity unpinned, see the oracle's header), so these are
the properties the algorithm must have, plus an independent einsum restatement of the same update."""
import numpy as np
import pytest

from oracle import ncp_oracle as o


def lowrank(shape, rank, seed=0, noise=0.0):
    rs = np.random.RandomState(seed)
    fs = [rs.gamma(1.0, 1.0, size=(s, rank)) for s in shape]
    x = o.cp_to_tensor(None, fs)
    if noise:
        x = x + noise * rs.random_sample(x.shape)
    return x, fs


def test_mttkrp_equals_einsum():
    x, fs = lowrank((4, 7, 3, 5), 3, noise=0.1)
    ref = [np.einsum("abcd,br,cr,dr->ar", x, fs[1], fs[2], fs[3]), np.einsum("abcd,ar,cr,dr->br", x, fs[0], fs[2], fs[3]),
           np.einsum("abcd,ar,br,dr->cr", x, fs[0], fs[1], fs[3]), np.einsum("abcd,ar,br,cr->dr", x, fs[0], fs[1], fs[2])]
    for m in range(4):
        np.testing.assert_allclose(o.mttkrp(x, None, fs, m), ref[m], rtol=1e-12)


def test_random_init_is_numpy_mt19937_stream():
    fs = o.random_init((3, 4, 2), 2, random_state=5)
    rs = np.random.RandomState(5)
    for f, s in zip(fs, (3, 4, 2)):
        np.testing.assert_array_equal(f, rs.random_sample((s, 2)))


def test_mu_errors_monotone_and_match_direct_error():
    x, _ = lowrank((5, 9, 4, 4), 3, noise=0.3)
    cp, errs = o.non_negative_parafac(x, 3, n_iter_max=40, init="random", tol=-1.0, random_state=1)
    assert len(errs) == 40
    assert all(errs[i + 1] <= errs[i] + 1e-12 for i in range(len(errs) - 1))
    assert all((f >= 0).all() for f in cp.factors)
    direct = np.linalg.norm(x - cp.to_tensor()) / np.linalg.norm(x)
    assert abs(direct - errs[-1]) < 1e-10


def test_fixed_point_stays():
    x, fs = lowrank((4, 6, 3, 3), 2)
    cp, errs = o.non_negative_parafac(x, 2, n_iter_max=5, init=(np.ones(2), fs), tol=-1.0)
    for a, b in zip(cp.factors, fs):
        np.testing.assert_allclose(a, b, rtol=1e-9)
    assert errs[-1] < 1e-7


def test_tol_stops_early_and_masked_runs():
    x, _ = lowrank((5, 8, 4, 4), 2, noise=0.05)
    cp, errs = o.non_negative_parafac(x, 2, n_iter_max=500, init="random", tol=1e-6, random_state=0)
    assert 2 <= len(errs) < 500 and abs(errs[-2] - errs[-1]) < 1e-6
    rs = np.random.RandomState(3)
    mask = (rs.rand(*x.shape) > 0.1).astype(float)
    cp, errs = o.non_negative_parafac(x * mask, 2, n_iter_max=80, init="random", tol=-1.0, random_state=0, mask=mask)
    assert o.masked_norm_error(x * mask, cp, mask) < 0.2


def test_hals_decreases_and_nonneg():
    x, _ = lowrank((5, 9, 4, 4), 3, noise=0.3)
    cp, errs = o.non_negative_parafac_hals(x, 3, n_iter_max=20, init="random", tol=-1.0, random_state=2)
    assert errs[-1] < errs[0]
    assert all((f >= 0).all() for f in cp.factors)
    direct = np.linalg.norm(x - cp.to_tensor()) / np.linalg.norm(x)
    assert abs(direct - errs[-1]) < 1e-8


def test_cp_normalize_roundtrip():
    x, fs = lowrank((4, 6, 3, 3), 3)
    w, nf = o.cp_normalize(None, fs)
    for f in nf:
        np.testing.assert_allclose(np.linalg.norm(f, axis=0), 1.0, rtol=1e-12)
    np.testing.assert_allclose(o.cp_to_tensor(w, nf), x, rtol=1e-10)


def test_svd_init_shapes_nonneg():
    x, _ = lowrank((3, 8, 4, 4), 2, noise=0.1)
    fs = o.initialize_cp(x, 5, init="svd", random_state=0)       # rank > I_0 -> random padding columns
    assert [f.shape for f in fs] == [(3, 5), (8, 5), (4, 5), (4, 5)]
    assert all((f >= 0).all() for f in fs)


def test_parafac_als_recovers_low_rank_and_gram_error_matches_direct():
    """ALS restatement (tensorly parafac): exact low-rank data is fitted to ~0, the error sequence is non-increasing
    (unmasked: every mode update is an exact least-squares step) and the Gram-form error equals ||X - rec|| / ||X||."""
    x, _ = lowrank((5, 8, 6, 7), 3, seed=1)
    cp, errs = o.parafac(x, 3, init="random", random_state=2, n_iter_max=200, tol=1e-14)
    assert errs[-1] < 1e-6
    xn, _ = lowrank((5, 8, 6, 7), 3, seed=1, noise=0.05)
    cp, errs = o.parafac(xn, 3, init="random", random_state=2, n_iter_max=30, tol=-1.0)
    assert len(errs) == 30 and all(b <= a + 1e-12 for a, b in zip(errs, errs[1:]))
    direct = np.linalg.norm(xn - cp.to_tensor()) / np.linalg.norm(xn)
    assert abs(errs[-1] - direct) < 1e-10
    cp_svd, errs_svd = o.parafac(xn, 3, init="svd", n_iter_max=30, tol=-1.0)
    assert errs_svd[-1] < 1.05 * errs[-1] + 1e-3
    # masked: missing entries are imputed once per iteration; the fit on the observed entries is close to the full fit
    m = (np.random.RandomState(3).rand(*xn.shape) > 0.2).astype(float)
    cpm, errm, n_it = o.parafac(xn * m, 3, init="random", random_state=2, n_iter_max=60, tol=1e-9, mask=m, return_iters=True)
    assert n_it == len(errm) <= 60
    assert o.masked_norm_error(xn * m, cpm, m) < 2.0 * direct + 1e-3


def test_masked_mttkrp_is_dense_pass_plus_missing_entry_correction():
    """The identity behind the masked device path (cell2cell_b200/csrc/c2c_corr.cu): with X zero at its missing entries,
    MTTKRP(X*m + rec*(1-m)) = MTTKRP(X) + the contraction of the re-imputed entries alone, where the dense parts T / U do
    not depend on the factors of their own side (one pass serves all leading modes, one both trailing modes)."""
    rs = np.random.RandomState(4)
    shape, rank = (4, 7, 5, 6), 3
    x, _ = lowrank(shape, 3, seed=2, noise=0.05)
    m = (rs.rand(*shape) > 0.25).astype(float)
    m[1, :, 2, :] = 0
    m[1, :, :, 2] = 0
    m[0, 3] = 0
    xz = x * m
    fs = [rs.random_sample((s, rank)) for s in shape]
    rec = o.cp_to_tensor(None, fs)
    xhat = xz * m + rec * (1 - m)
    P, E = shape[0] * shape[1], shape[2] * shape[3]
    W = (fs[0][:, None, :] * fs[1][None, :, :]).reshape(P, rank)            # Khatri-Rao rows of the leading factors
    B = (fs[2][:, None, :] * fs[3][None, :, :]).reshape(E, rank)            # ... of the trailing factors
    miss = (1 - m).reshape(P, E)
    recm = (W @ B.T) * miss                                                 # the re-imputed entries alone
    T = xz.reshape(P, E) @ B + recm @ B                                     # plane contraction + correction
    U = xz.reshape(P, E).T @ W + recm.T @ W                                 # fiber contraction + correction
    T4, U4 = T.reshape(shape[0], shape[1], rank), U.reshape(shape[2], shape[3], rank)
    want = [o.mttkrp(xhat, None, fs, mode) for mode in range(4)]
    np.testing.assert_allclose(np.einsum("abr,br->ar", T4, fs[1]), want[0], rtol=1e-12)
    np.testing.assert_allclose(np.einsum("abr,ar->br", T4, fs[0]), want[1], rtol=1e-12)
    np.testing.assert_allclose(np.einsum("svr,vr->sr", U4, fs[3]), want[2], rtol=1e-12)
    np.testing.assert_allclose(np.einsum("svr,sr->vr", U4, fs[2]), want[3], rtol=1e-12)


def test_constrained_parafac_oracle_self_consistency():
    """AO-ADMM restatement: non-negative factors under `non_negative`, a decreasing error tail, and with no constraint the
    first ADMM step is the regularised least-squares step (rho-damped), converging to the ALS fixed point on an exact tensor."""
    rs = np.random.RandomState(0)
    fs = [rs.gamma(1.0, 1.0, size=(s, 3)) for s in (5, 12, 4, 4)]
    x = o.cp_to_tensor(None, fs)
    cp, errs = o.constrained_parafac(x, 3, n_iter_max=60, init="random", random_state=1, non_negative=True, tol_outer=0)
    assert all((f >= 0).all() for f in cp.factors)
    assert errs[-1] < 0.05 and errs[-1] <= errs[5]
    direct = np.linalg.norm(x - o.cp_to_tensor(None, cp.factors)) / np.linalg.norm(x)
    assert abs(direct - errs[-1]) < 1e-8
    # soft thresholding shrinks: a large l1 penalty on mode 0 zeroes entries of that factor only
    cp1, _ = o.constrained_parafac(x, 3, n_iter_max=10, init="random", random_state=1, l1_reg={0: 0.5}, tol_outer=0)
    assert (cp1.factors[0] == 0).any() and not (cp1.factors[1] == 0).any()
    with pytest.raises(ValueError):
        o.constrained_parafac(x, 3, n_iter_max=1, init="random", non_negative=True, l1_reg=0.1)

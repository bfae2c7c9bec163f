"""GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path, called through the C ABI (ctypes -> libc2c_b200.so),
against the oracles on identical seeded inputs and against the reference-made golden fixtures.

Tolerances (BASELINE.json north_star): tensor build -- mask / names / ordering bit-exact, scores within 1 ulp fp32 of the
fp32-rounded reference value; factorization -- factors and final reconstruction error within 1e-4 relative after a fixed
iteration count from identical initial factors.
"""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ncp_oracle as o                                   # noqa: E402
from oracle.build_oracle import interaction_tensor_oracle            # noqa: E402
from tests.cases import OPTION_GRID, hard_case, toy_ppi, toy_rnaseq  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FACTOR_RTOL = 1e-4
ERROR_RTOL = 1e-4


def ulp_distance(a32, b32):
    """Units in the last place between two fp32 arrays of equal sign (scores are >= 0 or both -/+)."""
    a = np.ascontiguousarray(a32, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b32, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-300))


def lowrank(shape, rank, seed=0, noise=0.05):
    rs = np.random.RandomState(seed)
    fs = [rs.gamma(1.0, 1.0, size=(s, rank)) for s in shape]
    x = o.cp_to_tensor(None, fs)
    x = x / x.mean() + noise * rs.random_sample(x.shape)
    return np.ascontiguousarray(x, dtype=np.float32)


# ----------------------------------------------------------------------------------------------------------------------
# K1 / K1b: tensor build
# ----------------------------------------------------------------------------------------------------------------------
def _check_build(t, ref_tensor, ref_mask, ref_names, ref_genes, max_ulp=1):
    from cell2cell_b200.tensor import InteractionTensor  # noqa: F401
    X = t.tensor.cpu().numpy()
    assert X.dtype == np.float32 and X.shape == ref_tensor.shape
    want = np.asarray(ref_tensor, dtype=np.float64)
    want32 = np.clip(want, -np.finfo(np.float32).max, np.finfo(np.float32).max).astype(np.float32)
    d = ulp_distance(X, want32)
    assert d.max() <= max_ulp, "max ulp distance %d" % d.max()
    if ref_mask is None:
        assert t.mask is None
    else:
        np.testing.assert_array_equal(t.mask.cpu().numpy(), np.asarray(ref_mask).astype(np.uint8))
    assert [list(x) for x in t.order_names] == [list(x) for x in ref_names]
    assert list(t.genes) == list(ref_genes)
    return float((d == 0).mean())


def test_build_toy_golden_bit_exact():
    """Integer-valued toy data (cell2cell/datasets/toy_data.py): every score is exactly representable -> 0 ulp allowed
    for product / mean, 1 ulp for gmean."""
    from cell2cell_b200.tensor import InteractionTensor
    store = np.load(os.path.join(GOLD, "build_toy.npz"))
    for score in ("expression_product", "expression_mean", "expression_gmean"):
        for cplx in (True, False):
            key = "%s|%d" % (score, cplx)
            t = InteractionTensor([toy_rnaseq()], toy_ppi(cplx), communication_score=score,
                                  complex_sep="&" if cplx else None, verbose=False)
            names = json.loads(str(store[key + "/names"]))
            _check_build(t, store[key + "/tensor"], None, names, json.loads(str(store[key + "/genes"])),
                         max_ulp=1 if score == "expression_gmean" else 0)
            np.testing.assert_array_equal(t.loc_zeros.cpu().numpy(), store[key + "/loc_zeros"])
    a = toy_rnaseq()
    b = toy_rnaseq().drop(columns=["C2"]).drop(index=["Protein-D"]) * 2
    for how in ("inner", "outer", "outer_genes", "outer_cells"):
        key = "two|" + how
        t = InteractionTensor([a, b], toy_ppi(True), communication_score="expression_gmean", complex_sep="&", how=how,
                              verbose=False)
        mask = store[key + "/mask"] if bool(store[key + "/has_mask"]) else None
        _check_build(t, store[key + "/tensor"], mask, json.loads(str(store[key + "/names"])),
                     json.loads(str(store[key + "/genes"])))
        np.testing.assert_array_equal(t.loc_zeros.cpu().numpy(), store[key + "/loc_zeros"])


@pytest.mark.parametrize("seed", [2, 3])
def test_build_hard_golden_all_options(seed):
    """Every how x score x complex aggregation x grouping option on the quirk-laden seeded case, fp32-representable
    inputs, against the reference's own outputs (tests/golden/build_hard32.npz)."""
    from cell2cell_b200.tensor import InteractionTensor
    store = np.load(os.path.join(GOLD, "build_hard32.npz"))
    mats, ppi = hard_case(seed, fp32_values=True)
    exact = []
    for i, opt in enumerate(OPTION_GRID):
        key = "%d|%d|1" % (seed, i)
        t = InteractionTensor([m.copy() for m in mats], ppi.copy(), complex_sep="&", verbose=False, **opt)
        mask = store[key + "/mask"] if bool(store[key + "/has_mask"]) else None
        # grouped scores aggregate fp32-rounded members (<= 1 ulp each) -> 2 ulp on the aggregate
        exact.append(_check_build(t, store[key + "/tensor"], mask, json.loads(str(store[key + "/names"])),
                                  json.loads(str(store[key + "/genes"])), max_ulp=1 if opt["group_ppi_by"] is None else 2))
        np.testing.assert_array_equal(t.loc_zeros.cpu().numpy(), store[key + "/loc_zeros"])
    assert np.mean(exact) > 0.9


def test_build_matches_oracle_on_synthetic_config_shape():
    """BASELINE config 3 shape family (how='outer', complexes, gmean score), L scaled down so the pandas oracle is fast."""
    from cell2cell_b200.synthetic import synthetic_case
    from cell2cell_b200.tensor import InteractionTensor
    mats, ppi, kw = synthetic_case("asd", scale=0.03)
    mats64 = [m.astype(np.float64) for m in mats]
    ref = interaction_tensor_oracle(mats64, ppi, **{k: v for k, v in kw.items() if k != "verbose"})
    t = InteractionTensor(mats, ppi, **kw)
    _check_build(t, ref["tensor"], ref["mask"], ref["order_names"], ref["genes"])
    assert abs(t.excluded_value_fraction() - (1 - ref["mask"].mean())) < 1e-12
    assert abs(t.sparsity_fraction() - ref["loc_zeros"].mean()) < 1e-12
    assert abs(t.missing_fraction() - ref["loc_nans"].mean()) < 1e-12


def test_build_edge_cases():
    import pandas as pd
    from cell2cell_b200.tensor import InteractionTensor, build_context_ccc_tensor
    a = toy_rnaseq().astype(float)
    # no LR pair survives the filter -> empty LR mode
    ppi = pd.DataFrame({"A": ["X1"], "B": ["X2"]})
    t = InteractionTensor([a], ppi, verbose=False)
    assert t.tensor.shape == (1, 0, 5, 5)
    # NaN and inf already in the input: mask 0 / FLT_MAX (numpy.nan_to_num semantics in fp32)
    b = a.copy()
    b.iloc[0, 0] = np.nan
    b.iloc[1, 1] = np.inf
    # direct build_context_ccc_tensor upper-cases the PPI names only (reference quirk Q3): the lookup of the upper-cased
    # partner in the original-case index raises KeyError, as the reference's .loc does
    with pytest.raises(KeyError):
        build_context_ccc_tensor([a, b], toy_ppi(False), how="outer", communication_score="expression_product")
    ref = interaction_tensor_oracle([a, b], toy_ppi(False), how="outer", communication_score="expression_product",
                                    upper_letter_comparison=True)
    up = [m.rename(index=str.upper) for m in (a, b)]
    X, genes, cells, names, mask = build_context_ccc_tensor(up, toy_ppi(False), how="outer",
                                                            communication_score="expression_product")
    np.testing.assert_array_equal(mask.cpu().numpy(), ref["mask"].astype(np.uint8))
    got = X.cpu().numpy()
    want = ref["tensor"]
    big = want > 1e300
    assert (got[big] == np.finfo(np.float32).max).all()
    np.testing.assert_array_equal(got[~big], want[~big].astype(np.float32))


# ----------------------------------------------------------------------------------------------------------------------
# K2..K8 one at a time
# ----------------------------------------------------------------------------------------------------------------------
SHAPES = [(4, 30, 6, 6), (3, 5, 7, 9), (5, 11, 17, 17), (2, 3, 4, 8, 4), (10, 6, 8), (3, 40, 40, 40)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("rank", [1, 4, 10, 25, 40])
@pytest.mark.parametrize("masked", [False, True])
def test_mttkrp_every_mode(shape, rank, masked):
    from cell2cell_b200 import engine
    rs = np.random.RandomState(len(shape) * 100 + rank)
    x = lowrank(shape, 3, seed=rank)
    fs = [rs.random_sample((s, rank)) for s in shape]
    mask = (rs.rand(*shape) > 0.2).astype(np.uint8) if masked else None
    X = torch.as_tensor(x).cuda()
    Mk = torch.as_tensor(mask).cuda() if masked else None
    dfs = [engine.pack_factor(f.astype(np.float32), rank, X.device) for f in fs]
    f32 = [f.astype(np.float32).astype(np.float64) for f in fs]
    xh = x.astype(np.float64)
    if masked:
        xh = xh * mask + o.cp_to_tensor(None, f32) * (1 - mask)
    for mode in range(len(shape)):
        got = engine.unpack_factor(engine.mttkrp(X, Mk, dfs, rank, mode), rank).cpu().numpy()
        want = o.mttkrp(xh, None, f32, mode)
        assert rel_err(got, want) < 2e-6, (mode, rel_err(got, want))


def test_small_steps_match_numpy():
    from cell2cell_b200 import engine
    rs = np.random.RandomState(0)
    shape, rank = (7, 33, 5, 6), 6
    fs = [rs.random_sample((s, rank)) for s in shape]
    dfs = [engine.pack_factor(f.astype(np.float32), rank, "cuda") for f in fs]
    f32 = [f.astype(np.float32).astype(np.float64) for f in fs]
    for skip in (-1, 0, 2):
        G, n2 = engine.gram_hadamard(dfs, shape, rank, skip)
        want = np.ones((rank, rank))
        for k, f in enumerate(f32):
            if k != skip:
                want = want * (f.T @ f)
        assert rel_err(G[:rank, :rank].cpu().numpy(), want) < 1e-6
        assert abs(n2 - o.cp_norm(None, f32) ** 2) / n2 < 1e-12
    # MU and HALS single updates
    M = rs.random_sample((33, rank)) * 5
    G = (f32[0].T @ f32[0]) * (f32[2].T @ f32[2]) * (f32[3].T @ f32[3])
    dM = engine.pack_factor(M.astype(np.float32), rank, "cuda")
    dG = engine.pack_factor(G.astype(np.float32), rank, "cuda")
    eps = np.finfo(np.float32).eps
    A = dfs[1].clone()
    engine.mu_update(A, dM, dG, rank)
    M32, G32 = M.astype(np.float32).astype(np.float64), G.astype(np.float32).astype(np.float64)
    want = f32[1] * np.clip(M32, eps, None) / np.clip(f32[1] @ G32, eps, None)
    assert rel_err(A[:, :rank].cpu().numpy(), want) < 1e-6
    A = dfs[1].clone()
    engine.hals_update(A, dM, dG, rank)
    want = o.hals_nnls(M32.T, G32, f32[1].T, n_iter_max=100).T
    assert rel_err(A[:, :rank].cpu().numpy(), want) < 1e-5
    # cp_normalize
    w = engine.cp_normalize(dfs, rank)
    wr, nf = o.cp_normalize(None, f32)
    assert rel_err(w.cpu().numpy(), wr) < 1e-6
    for a, b in zip(dfs, nf):
        assert rel_err(a[:, :rank].cpu().numpy(), b) < 1e-6


@pytest.mark.parametrize("shape,rank", [((4, 30, 6, 6), 4), ((5, 11, 17, 17), 10), ((2, 3, 4, 8, 4), 3), ((6, 9, 8, 12), 40),
                                        # planes split over 2 / 2 / 3 CTAs of the streaming kernel (positions > threads per CTA)
                                        ((3, 50, 40, 40), 20), ((2, 37, 48, 48), 10), ((2, 21, 44, 60), 32)])
@pytest.mark.parametrize("masked", [False, True])
def test_recon_sums_and_sumsq(shape, rank, masked):
    from cell2cell_b200 import engine
    rs = np.random.RandomState(3)
    x = lowrank(shape, 3, seed=5)
    fs = [rs.random_sample((s, rank)) for s in shape]
    mask = (rs.rand(*shape) > 0.3).astype(np.uint8) if masked else None
    X = torch.as_tensor(x).cuda()
    Mk = torch.as_tensor(mask).cuda() if masked else None
    dfs = [engine.pack_factor(f.astype(np.float32), rank, X.device) for f in fs]
    f32 = [f.astype(np.float32).astype(np.float64) for f in fs]
    rec = o.cp_to_tensor(None, f32)
    m = np.ones(shape) if mask is None else mask.astype(np.float64)
    d = (x.astype(np.float64) - rec) * m
    want = np.array([d.sum(), (d * d).sum(), (x * m).sum(), (x.astype(np.float64) ** 2 * m).sum(), m.sum()])
    got = engine.recon_sums(X, Mk, dfs, rank)
    np.testing.assert_allclose(got, want, rtol=2e-5)
    assert abs(engine.sumsq(X) - float((x.astype(np.float64) ** 2).sum())) / want[3] < 1e-12 or masked


# ----------------------------------------------------------------------------------------------------------------------
# whole factorization against the oracle, identical initial factors, fixed iteration count
# ----------------------------------------------------------------------------------------------------------------------
def _compare_run(x, rank, n_iter, mask=None, hals=False, seed=888, init="random"):
    from cell2cell_b200.tensor.factorization import non_negative_parafac, non_negative_parafac_hals, _compute_norm_error
    x64 = x.astype(np.float64)
    if hals:
        ref, ref_err = o.non_negative_parafac_hals(x64, rank, n_iter_max=n_iter, init=init, tol=-1.0, random_state=seed)
        got, got_err = non_negative_parafac_hals(x, rank, n_iter_max=n_iter, init=init, tol=-1.0, random_state=seed,
                                                 return_errors=True)
    else:
        m64 = None if mask is None else mask.astype(np.float64)
        ref, ref_err = o.non_negative_parafac(x64, rank, n_iter_max=n_iter, init=init, tol=-1.0, random_state=seed, mask=m64)
        got, got_err = non_negative_parafac(x, rank, n_iter_max=n_iter, init=init, tol=-1.0, random_state=seed,
                                            mask=mask, return_errors=True)
    assert len(got_err) == len(ref_err) == n_iter
    for a, b in zip(got.factors, ref.factors):
        assert a.shape == b.shape and (a >= 0).all()
        assert rel_err(a, b) < FACTOR_RTOL, rel_err(a, b)
    assert abs(got_err[-1] - ref_err[-1]) <= ERROR_RTOL * abs(ref_err[-1])
    # per-iteration errors come from the Gram form sqrt(|nx2 + rn2 - 2 ip|), which cancels in fp32 when the error is small
    np.testing.assert_allclose(np.asarray(got_err), np.asarray(ref_err), rtol=ERROR_RTOL, atol=3e-6)
    if mask is not None:
        e_ref = o.masked_norm_error(x64, ref, mask.astype(np.float64))
        e_got = _compute_norm_error(x, got, mask)
        assert abs(e_got - e_ref) <= ERROR_RTOL * e_ref
    return got, got_err


def test_ncp_config1_toy_rank4():
    """BASELINE config 1: 4 x 300 x 6 x 6, rank 4, random init, full size, 100 iterations."""
    _compare_run(lowrank((4, 300, 6, 6), 5, seed=1), 4, 100)


def test_ncp_config2_balf_rank10():
    """BASELINE config 2: 12 x 1054 x 6 x 6, rank 10, full size, 100 iterations."""
    _compare_run(lowrank((12, 1054, 6, 6), 5, seed=2), 10, 100)


def test_ncp_config3_asd_masked_rank10():
    """BASELINE config 3 family: 13 x L x 17 x 17 masked (L scaled to 150 for the numpy oracle), rank 10, 60 iterations."""
    shape = (13, 150, 17, 17)
    x = lowrank(shape, 5, seed=3)
    rs = np.random.RandomState(4)
    mask = np.ones(shape, dtype=np.uint8)
    mask[2::3, :, 5, :] = 0          # a cell type missing from every 3rd context (sender rows and receiver columns)
    mask[2::3, :, :, 5] = 0
    mask[rs.rand(13, 150) < 0.05] = 0  # gene pairs missing from a context: whole planes
    mask[rs.rand(*shape) < 0.02] = 0   # and arbitrary user-masked entries
    _compare_run(x * mask, 10, 60, mask=mask)


def test_ncp_config4_shape_family_rank20():
    """BASELINE config 4 family (K = 40 planes, rank 20), C and L scaled down for the oracle."""
    _compare_run(lowrank((6, 50, 40, 40), 6, seed=5), 20, 40)


@pytest.mark.parametrize("shape,rank", [((3, 5, 7, 9), 3), ((10, 6, 8), 2), ((2, 3, 4, 8, 4), 5), ((4, 12, 6, 6), 40),
                                        ((5, 7, 3, 3), 1)])
@pytest.mark.parametrize("masked", [False, True])
def test_ncp_odd_shapes(shape, rank, masked):
    x = lowrank(shape, 3, seed=7)
    mask = None
    if masked:
        mask = (np.random.RandomState(8).rand(*shape) > 0.15).astype(np.uint8)
        x = x * mask
    _compare_run(x, rank, 30, mask=mask)


@pytest.mark.parametrize("shape,rank", [((4, 60, 6, 6), 4), ((5, 20, 17, 17), 10), ((10, 6, 8), 3)])
def test_hals_parity(shape, rank):
    _compare_run(lowrank(shape, 4, seed=9), rank, 25, hals=True)


def test_tol_stop_matches_oracle_iteration_count():
    x = lowrank((6, 40, 6, 6), 3, seed=11)
    from cell2cell_b200.tensor.factorization import non_negative_parafac
    ref, ref_err = o.non_negative_parafac(x.astype(np.float64), 3, n_iter_max=500, init="random", tol=1e-5, random_state=1)
    got, got_err = non_negative_parafac(x, 3, n_iter_max=500, init="random", tol=1e-5, random_state=1, return_errors=True)
    # the stop rule compares successive fp32-derived errors with tol: the count may move by a few iterations
    assert abs(len(got_err) - len(ref_err)) <= 0.15 * len(ref_err)
    # ... and the error is still falling by ~tol per iteration when the rule fires
    assert abs(got_err[-1] - ref_err[-1]) < 1e-4 * ref_err[-1] + 1e-5 * (abs(len(got_err) - len(ref_err)) + 2)


def test_svd_init_statistically_equivalent():
    """svd init depends on LAPACK sign / ordering conventions: compare the reached error, not the factors."""
    x = lowrank((6, 40, 8, 8), 3, seed=13, noise=0.02)
    from cell2cell_b200.tensor.factorization import non_negative_parafac
    ref, ref_err = o.non_negative_parafac(x.astype(np.float64), 3, n_iter_max=60, init="svd", tol=-1.0)
    got, got_err = non_negative_parafac(x, 3, n_iter_max=60, init="svd", tol=-1.0, return_errors=True)
    assert abs(got_err[-1] - ref_err[-1]) < 0.05 * ref_err[-1] + 1e-3


# ----------------------------------------------------------------------------------------------------------------------
# the public API end to end
# ----------------------------------------------------------------------------------------------------------------------
def test_api_interaction_tensor_to_factor_dataframes():
    from cell2cell_b200.synthetic import synthetic_case
    from cell2cell_b200.tensor import InteractionTensor
    mats, ppi, kw = synthetic_case("toy", scale=0.2)
    t = InteractionTensor(mats, ppi, **kw)
    t.compute_tensor_factorization(rank=4, init="random", random_state=888, n_iter_max=50, tol=-1.0)
    ref_in = interaction_tensor_oracle([m.astype(np.float64) for m in mats], ppi, **{k: v for k, v in kw.items() if k != "verbose"})
    x64 = ref_in["tensor"].astype(np.float32).astype(np.float64)
    cp, errs = o.non_negative_parafac(x64, 4, n_iter_max=50, init="random", tol=-1.0, random_state=888)
    w, nf = o.cp_normalize(cp.weights, cp.factors)
    order = np.argsort(w)[::-1]
    assert list(t.factors.keys()) == ["Contexts", "Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"]
    for (label, df), f, names in zip(t.factors.items(), nf, ref_in["order_names"]):
        assert list(df.index) == list(names) and list(df.columns) == ["Factor %d" % i for i in range(1, 5)]
        assert rel_err(df.values, f[:, order]) < FACTOR_RTOL
    np.testing.assert_allclose(t.explained_variance_ratio_, w[order] / w.sum(), rtol=1e-4)
    rec = cp.to_tensor()
    d = x64 - rec
    ev = 1 - np.linalg.norm(d - d.mean()) / np.linalg.norm(x64 - x64.mean())
    assert abs(t.explained_variance_ - ev) < 1e-4
    assert t.rank == 4
    # best-of-runs picks the lowest error and is reproducible
    t.compute_tensor_factorization(rank=3, init="random", random_state=0, runs=3, n_iter_max=20, tol=-1.0)
    errs = [o.non_negative_parafac(x64, 3, n_iter_max=20, init="random", tol=-1.0, random_state=s)[1][-1] for s in range(3)]
    best = int(np.argmin(errs))
    cp = o.non_negative_parafac(x64, 3, n_iter_max=20, init="random", tol=-1.0, random_state=best)[0]
    for a, b in zip(t.tl_object.factors, cp.factors):
        assert rel_err(a, b) < FACTOR_RTOL


def test_api_elbow_matches_oracle_losses():
    from cell2cell_b200.tensor import PreBuiltTensor
    shape = (5, 40, 6, 6)
    x = lowrank(shape, 4, seed=21, noise=0.1)
    names = [["n%d_%d" % (k, i) for i in range(s)] for k, s in enumerate(shape)]
    t = PreBuiltTensor(x, order_names=names)
    fig, loss = t.elbow_rank_selection(upper_rank=6, runs=2, init="random", random_state=5, n_iter_max=30, tol=-1.0,
                                       output_fig=False)
    assert fig is None and t.elbow_metric_raw.shape == (2, 6)
    want = np.array([[o.non_negative_parafac(x.astype(np.float64), r, n_iter_max=30, init="random", tol=-1.0,
                                             random_state=5 + run)[1][-1] for r in range(1, 7)] for run in range(2)])
    np.testing.assert_allclose(t.elbow_metric_raw, want, rtol=ERROR_RTOL)
    assert [l[0] for l in loss] == list(range(1, 7)) and isinstance(t.rank, int)
    # runs == 1 path, masked: loss is the masked normalized error (factorization.py:263-267)
    mask = (np.random.RandomState(1).rand(*shape) > 0.1).astype(np.uint8)
    t2 = PreBuiltTensor(x * mask, order_names=names, mask=mask)
    fig, loss = t2.elbow_rank_selection(upper_rank=3, runs=1, init="random", random_state=2, n_iter_max=20, tol=-1.0,
                                        output_fig=False, automatic_elbow=False, manual_elbow=2)
    for r, l in loss:
        cp = o.non_negative_parafac((x * mask).astype(np.float64), r, n_iter_max=20, init="random", tol=-1.0,
                                    random_state=2, mask=mask.astype(np.float64))[0]
        assert abs(l - o.masked_norm_error((x * mask).astype(np.float64), cp, mask.astype(np.float64))) < ERROR_RTOL * l
    assert t2.rank == 2 and t2.explained_variance_ is None
    # similarity metric
    fig, loss = t.elbow_rank_selection(upper_rank=3, runs=3, metric="similarity", init="random", random_state=5,
                                       n_iter_max=30, tol=-1.0, output_fig=False, automatic_elbow=False, manual_elbow=2)
    assert t.elbow_metric_raw.shape == (3, 3)


def test_prebuilt_tensor_nan_handling_and_pickle(tmp_path):
    import pickle
    from cell2cell_b200.tensor import PreBuiltTensor
    x = lowrank((3, 6, 4, 4), 2, seed=1).astype(np.float64)
    x[0, 0, 0, 0] = np.nan
    x[1, 1, 1, 1] = 0.0
    names = [["n%d_%d" % (k, i) for i in range(s)] for k, s in enumerate(x.shape)]
    t = PreBuiltTensor(x, order_names=names)
    assert t.loc_nans.sum().item() == 1 and t.loc_zeros.sum().item() == 1 and t.tensor[0, 0, 0, 0].item() == 0.0
    assert t.order_labels == ["Dimension-1", "Dimension-2", "Dimension-3", "Dimension-4"]
    t.compute_tensor_factorization(rank=2, init="random", random_state=0, n_iter_max=10)
    p = tmp_path / "t.pkl"
    t.write_file(str(p))
    t2 = pickle.load(open(p, "rb"))
    assert t2.tensor.is_cuda and torch.equal(t2.tensor, t.tensor) and list(t2.factors) == list(t.factors)


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE full sizes: size-independent properties (the numpy oracle would take minutes per iteration here)
# ----------------------------------------------------------------------------------------------------------------------
def _device_lowrank(shape, rank, seed, noise):
    g = torch.Generator(device="cuda").manual_seed(seed)
    fs = [torch.rand((s, rank), generator=g, device="cuda") + 0.1 for s in shape]
    x = torch.einsum("ar,br,cr,dr->abcd", *fs)
    if noise:
        x += noise * x.mean() * torch.rand(x.shape, generator=g, device="cuda")
    return x.contiguous(), fs


@pytest.mark.parametrize("shape,rank", [((60, 2000, 40, 40), 10), ((12, 1054, 6, 6), 10), ((13, 1900, 17, 17), 10)])
def test_full_size_properties(shape, rank):
    from cell2cell_b200 import engine
    X, fs = _device_lowrank(shape, rank, 1, 0.0)
    # (1) exact low-rank tensor + true factors: a fixed point of the multiplicative update, error ~ 0
    dfs = [engine.pack_factor(f, rank, X.device) for f in fs]
    errs, n_it, conv = engine.ncp_run(X, None, dfs, rank, n_iter_max=3, tol=-1.0)
    assert n_it == 3 and errs[-1] < 5e-3
    for a, b in zip(dfs, fs):
        assert float((a[:, :rank] - b).norm() / b.norm()) < 1e-4
    # (2) MTTKRP is linear in X and matches einsum on a mode
    M1 = engine.mttkrp(X, None, dfs, rank, 1)
    M2 = engine.mttkrp(X * 2.0, None, dfs, rank, 1)
    assert float((M2 - 2 * M1).norm() / M1.norm()) < 1e-6
    want = torch.einsum("abcd,ar,cr,dr->br", X.double(), dfs[0][:, :rank].double(), dfs[2][:, :rank].double(),
                        dfs[3][:, :rank].double())
    assert float((M1[:, :rank].double() - want).norm() / want.norm()) < 2e-6
    # (3) noisy tensor, random start: error sequence monotone non-increasing, Gram-form error == direct masked-sum error
    X, _ = _device_lowrank(shape, 5, 2, 0.2)
    g = torch.Generator(device="cuda").manual_seed(3)
    dfs = [engine.pack_factor(torch.rand((s, rank), generator=g, device="cuda"), rank, X.device) for s in shape]
    errs, n_it, conv = engine.ncp_run(X, None, dfs, rank, n_iter_max=12, tol=-1.0)
    assert all(errs[i + 1] <= errs[i] * (1 + 1e-6) for i in range(len(errs) - 1))
    s = engine.recon_sums(X, None, dfs, rank)
    direct = np.sqrt(s[1]) / np.sqrt(s[3])
    assert abs(direct - errs[-1]) < 1e-4 * direct
    assert all(bool((f >= 0).all()) for f in dfs)
    # (4) masked path with an all-ones mask reproduces the unmasked run
    ones = torch.ones(shape, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    dfs2 = [engine.pack_factor(torch.rand((s, rank), generator=g, device="cuda"), rank, X.device) for s in shape]
    errs2, _, _ = engine.ncp_run(X, ones, dfs2, rank, n_iter_max=12, tol=-1.0)
    np.testing.assert_allclose(errs2, errs, rtol=1e-5)


@pytest.mark.parametrize("K", [8, 40])
@pytest.mark.parametrize("score", ["expression_product", "expression_mean", "expression_gmean"])
def test_build_fast_kernel_equals_general_kernel_on_special_values(score, K):
    """K % 4 == 0 with fp32 partner values takes build_ccc_fast (branch-free quads); the general kernel
    (C2C_B200_BUILD_GENERAL=1) is the one the golden grid pinned in round 1.  Same bits -- scores and mask -- on values that
    exercise every rare case: zeros, NaN (absent genes / cell types), infinities, subnormal / tiny / huge products, negative
    values, two-subunit 'min' complexes, tiles that straddle contexts (L not a multiple of the 16-plane tile); and against
    a float64 numpy statement of compute_ccc_matrix (communication_scores.py:195-203)."""
    from cell2cell_b200 import engine
    C, G, L = 3, 64, 100
    rs = np.random.RandomState(5)
    expr = rs.gamma(1.0, 2.0, size=(C, G, K)).astype(np.float32)
    special = np.array([0.0, 0.0, 0.0, np.nan, np.inf, -np.inf, 1e-30, 3e-38, 1e-44, 1e-20, 1e20, 1e19, -1.5, -0.0, 1.0, 2.5e-16],
                       dtype=np.float32)
    pick = rs.rand(C, G, K) < 0.35
    expr[pick] = special[rs.randint(0, len(special), size=int(pick.sum()))]
    lig, rec, rec2 = rs.randint(0, G, size=L), rs.randint(0, G, size=L), rs.randint(0, G, size=L)
    two = rs.rand(L) < 0.4                                           # receptor l is a two-subunit complex ('min')
    cell_col = np.tile(np.arange(K, dtype=np.int32), (C, 1))
    cell_col[1, 3] = -1                                              # a cell type absent from context 1
    sub_rows, a_first, b_first, b_count = [], np.zeros(L, np.int32), np.zeros(L, np.int32), np.ones(L, np.int32)
    for l in range(L):
        a_first[l] = len(sub_rows); sub_rows.append(lig[l])
        b_first[l] = len(sub_rows); sub_rows.append(rec[l])
        if two[l]:
            sub_rows.append(rec2[l]); b_count[l] = 2
    a_count = np.ones((C, L), dtype=np.int32)
    b_cnt = np.tile(b_count, (C, 1))
    a_count[2, 5] = 0; b_cnt[2, 5] = 0                               # a gene absent from context 2
    a_count[0, 7] = -1; b_cnt[0, 7] = -1                             # complexes with a missing subunit: zeros
    args = (expr.reshape(-1), np.arange(C, dtype=np.int64) * G * K, np.full(C, K, np.int32), cell_col,
            np.tile(a_first, C), a_count.reshape(-1), np.tile(b_first, C), b_cnt.reshape(-1), np.asarray(sub_rows, np.int32),
            C, L, K, score, "min", True, torch.device("cuda", 0))
    Xf, Mf = engine.build_ccc(*args)
    os.environ["C2C_B200_BUILD_GENERAL"] = "1"
    try:
        Xg, Mg = engine.build_ccc(*args)
    finally:
        del os.environ["C2C_B200_BUILD_GENERAL"]
    xf, xg = Xf.cpu().numpy(), Xg.cpu().numpy()
    np.testing.assert_array_equal(Mf.cpu().numpy(), Mg.cpu().numpy())
    np.testing.assert_array_equal(xf.view(np.uint32), xg.view(np.uint32))
    # float64 statement of the reference's score on the same operands
    e64 = expr.astype(np.float64)
    with np.errstate(all="ignore"):
        a = np.stack([e64[c][lig] for c in range(C)])
        b = np.stack([np.where(two[:, None], np.fmin(e64[c][rec], e64[c][rec2]), e64[c][rec]) for c in range(C)])
    for side in (a, b):
        side[1, :, 3] = np.nan
        side[2, 5] = np.nan
        side[0, 7] = 0.0
    with np.errstate(all="ignore"):
        va, vb = a[:, :, :, None], b[:, :, None, :]
        ref = va * vb if score == "expression_product" else ((va + vb) / 2.0 if score == "expression_mean" else np.sqrt(va * vb))
    want_mask = (~np.isnan(ref)).astype(np.uint8)
    np.testing.assert_array_equal(Mf.cpu().numpy(), want_mask)
    fmax = float(np.finfo(np.float32).max)
    with np.errstate(all="ignore"):
        want = np.clip(np.nan_to_num(ref, nan=0.0, posinf=fmax, neginf=-fmax), -fmax, fmax).astype(np.float32)
    ok = np.isfinite(want) & (want_mask == 1)
    d = ulp_distance(np.abs(xf[ok]), np.abs(want[ok]))
    assert d.max() <= 1, "max ulp distance %d" % d.max()


def test_full_size_build_roundtrip():
    """Config 4 size build (60 x 2000 x 40 x 40): sampled planes against a numpy recomputation, mask all ones under 'outer'
    when nothing is missing, checksum stable across two builds."""
    from cell2cell_b200.synthetic import synthetic_contexts, synthetic_ppi
    from cell2cell_b200.tensor import InteractionTensor
    L, K, C = 2000, 40, 60
    ppi = synthetic_ppi(L, 2 * L)
    mats = synthetic_contexts(C, 2 * L, K)
    t = InteractionTensor(mats, ppi, how="outer", communication_score="expression_gmean", complex_sep="&", verbose=False)
    assert t.tensor.shape == (C, L, K, K) and bool((t.mask == 1).all())
    rs = np.random.RandomState(0)
    for _ in range(20):
        c, l = rs.randint(C), rs.randint(L)
        a, b = ppi["A"][l], ppi["B"][l]
        v = mats[c].loc[a].values.astype(np.float64)
        w = np.min([mats[c].loc[p].values for p in b.split("&")], axis=0).astype(np.float64)
        want = np.sqrt(v[:, None] * w[None, :]).astype(np.float32)
        assert ulp_distance(t.tensor[c, l].cpu().numpy(), want).max() <= 1
    t2 = InteractionTensor(mats, ppi, how="inner", communication_score="expression_gmean", complex_sep="&", verbose=False)
    assert t2.mask is None and torch.equal(t.tensor, t2.tensor)


# ----------------------------------------------------------------------------------------------------------------------
# tensor-pipe streaming passes (TMA + tcgen05.mma kind::tf32 with the three-term split, c2c_mma.cuh): taken when the rank
# reaches c2c_set_mma_min_rank, the in-plane size is a multiple of 32 and there are >= 128 planes per SM
# ----------------------------------------------------------------------------------------------------------------------
MMA_SHAPES = [(40, 500, 16, 8), (25, 801, 8, 32), (19000, 16, 16), (10, 1900, 40, 40), (20000, 12, 12), (22, 870, 30, 30)]


@pytest.fixture
def mma_from_rank_1():
    from cell2cell_b200 import engine
    prev = engine.set_mma_min_rank(1)
    yield
    engine.set_mma_min_rank(prev)


@pytest.mark.parametrize("shape", MMA_SHAPES)
@pytest.mark.parametrize("rank", [1, 5, 16, 17, 20, 24, 25, 32])       # 17-24: plane_mma at N = 24
def test_mma_passes_match_einsum(shape, rank, mma_from_rank_1):
    """T and U of the tensor-pipe passes against float64 einsum (every mode's MTTKRP is derived from them)."""
    from cell2cell_b200 import engine
    g = torch.Generator(device="cuda").manual_seed(rank)
    X = torch.rand(shape, generator=g, device="cuda") * (torch.rand(shape, generator=g, device="cuda") > 0.3)
    fs = [engine.pack_factor(torch.rand((s, rank), generator=g, device="cuda"), rank, X.device) for s in shape]
    n0 = engine.launch_count()
    T = engine.plane_contract(X, None, fs, rank)
    U = engine.fiber_contract(X, None, fs, rank)
    assert engine.launch_count() - n0 == 3            # plane_mma, fiber_mma, fiber_reduce: no FFMA tail kernels
    letters = "abcdefgh"[: len(shape)]
    lead, tr = letters[:-2], letters[-2:]
    f64 = [f[:, :rank].double() for f in fs]
    Tw = torch.einsum(letters + "," + tr[0] + "r," + tr[1] + "r->" + lead + "r", X.double(), f64[-2], f64[-1]).reshape(-1, rank)
    Uw = torch.einsum(letters + "," + ",".join(l + "r" for l in lead) + "->" + tr + "r", X.double(), *f64[:-2]).reshape(-1, rank)
    eT = float((T[:, :rank].double() - Tw).norm() / Tw.norm())
    eU = float((U[:, :rank].double() - Uw).norm() / Uw.norm())
    assert eT < 2e-6 and eU < 2e-6, (eT, eU)
    # padding columns of T / U stay zero (the factor updates read whole float4 rows)
    ldr = engine.ldr_of(rank)
    if ldr > rank:
        assert float(T[:, rank:].abs().max()) == 0.0 and float(U[:, rank:].abs().max()) == 0.0


@pytest.mark.parametrize("shape,rank,n_iter", [((40, 500, 16, 8), 16, 30), ((40, 500, 16, 8), 20, 30), ((25, 801, 8, 32), 5, 30),
                                               ((19000, 16, 16), 32, 15), ((20000, 12, 12), 10, 20)])
def test_mma_ncp_matches_oracle(shape, rank, n_iter, mma_from_rank_1):
    """Whole factorization through the tensor-pipe passes against the float64 oracle: factors and error within 1e-4."""
    _compare_run(lowrank(shape, 6, seed=11), rank, n_iter)


def test_mma_full_size_rank20():
    """BASELINE config 4 at its named rank (60 x 2000 x 40 x 40, rank 20): default dispatch takes the tensor-pipe passes."""
    test_full_size_properties((60, 2000, 40, 40), 20)


# ----------------------------------------------------------------------------------------------------------------------
# batched runs (c2c_ncp_run_batched): the runs of one rank side by side, one read of the tensor per pass for all of them
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,rank,nruns", [((6, 50, 8, 8), 3, 5), ((12, 1054, 6, 6), 4, 8), ((40, 500, 16, 8), 6, 5),
                                              ((40, 500, 16, 8), 16, 2), ((5, 7, 3, 3), 1, 10)])
@pytest.mark.parametrize("tol", [-1.0, 2e-4])
def test_batched_runs_match_single_runs(shape, rank, nruns, tol):
    """Every run of a batch equals the same run factorized alone: factors, error history, stop iteration."""
    from cell2cell_b200.tensor.factorization import _factorize_batch, as_device_tensor, non_negative_parafac
    x = lowrank(shape, 4, seed=21)
    X = as_device_tensor(x)
    seeds = [300 + 7 * b for b in range(nruns)]
    n_iter = 40
    batch = _factorize_batch(X, [(rank, sd) for sd in seeds], n_iter, tol, check_every=5)
    assert len(batch) == nruns
    n_its = set()
    for (cp_b, err_b), seed in zip(batch, seeds):
        cp_s, err_s = non_negative_parafac(X, rank, n_iter_max=n_iter, init="random", tol=tol, random_state=seed, return_errors=True)
        assert len(err_b) == len(err_s), (len(err_b), len(err_s))
        n_its.add(len(err_b))
        # the per-iteration values are the fp32 Gram form (cancels at the 1e-6 level, see _compare_run); the kept last one is direct
        np.testing.assert_allclose(err_b, err_s, rtol=2e-5, atol=5e-6)
        assert abs(err_b[-1] - err_s[-1]) <= 2e-5 * err_s[-1]
        for a, b in zip(cp_b.factors, cp_s.factors):
            assert a.shape == b.shape and rel_err(a, b) < FACTOR_RTOL, rel_err(a, b)
    if tol < 0:
        assert n_its == {n_iter}


def test_mixed_rank_batch_matches_single_runs():
    """Runs of different ranks sharing a pass (17 + 9 + 6 = 32 columns, tensor-pipe path): each equals the run alone."""
    from cell2cell_b200.tensor.factorization import _factorize_batch, _pack_runs, as_device_tensor, non_negative_parafac
    X = as_device_tensor(lowrank((40, 500, 16, 8), 5, seed=23))
    runs = [(17, 5), (9, 6), (6, 7)]
    batch = _factorize_batch(X, runs, 25, -1.0)
    for (cp_b, err_b), (rank, seed) in zip(batch, runs):
        cp_s, err_s = non_negative_parafac(X, rank, n_iter_max=25, init="random", tol=-1.0, random_state=seed, return_errors=True)
        assert cp_b.rank == rank and len(err_b) == len(err_s) == 25
        np.testing.assert_allclose(err_b, err_s, rtol=2e-5, atol=5e-6)
        for a, b in zip(cp_b.factors, cp_s.factors):
            assert rel_err(a, b) < FACTOR_RTOL, rel_err(a, b)
    bins = _pack_runs([(r, run) for r in range(1, 26) for run in range(10)])
    assert sorted(u for b in bins for u in b) == sorted((r, run) for r in range(1, 26) for run in range(10))
    assert all(sum(r for r, _ in b) <= 28 for b in bins) and len(bins) <= 125


def test_batched_elbow_equals_unbatched_elbow():
    from cell2cell_b200.tensor.factorization import _multiple_runs_elbow_analysis
    x = lowrank((12, 300, 6, 6), 4, seed=5)
    a = _multiple_runs_elbow_analysis(x, upper_rank=7, runs=4, init="random", random_state=11, n_iter_max=25, tol=-1.0)
    b = _multiple_runs_elbow_analysis(x, upper_rank=7, runs=4, init="random", random_state=11, n_iter_max=25, tol=-1.0, batch_runs=False)
    assert a.shape == b.shape == (4, 7)
    np.testing.assert_allclose(a, b, rtol=2e-5)


def test_best_of_runs_batched_equals_run_by_run(capsys):
    """``compute_tensor_factorization(runs=N)`` (the 'robust' pipeline mode, tensor.py:294-323): the runs packed side by side
    pick the same best model as the runs one after the other."""
    from cell2cell_b200.tensor import PreBuiltTensor
    x = lowrank((8, 120, 6, 6), 4, seed=6).astype(np.float64)
    names = [["n%d_%d" % (k, i) for i in range(s)] for k, s in enumerate(x.shape)]
    a, b = PreBuiltTensor(x, order_names=names), PreBuiltTensor(x, order_names=names)
    a.compute_tensor_factorization(rank=5, init="random", random_state=3, runs=7, n_iter_max=30, tol=-1.0)
    out_a = capsys.readouterr().out
    b.compute_tensor_factorization(rank=5, init="random", random_state=3, runs=7, n_iter_max=30, tol=-1.0, batch_runs=False)
    out_b = capsys.readouterr().out
    assert "Best model has a normalized error of" in out_a and out_a == out_b
    for k in a.factors:
        np.testing.assert_allclose(a.factors[k].values, b.factors[k].values, rtol=2e-4, atol=1e-6)
    assert a.explained_variance_ == pytest.approx(b.explained_variance_, rel=1e-5)
    # the best run is the oracle's best run
    errs = [o.non_negative_parafac(x, 5, n_iter_max=30, init="random", tol=-1.0, random_state=3 + r)[1][-1] for r in range(7)]
    ref = o.non_negative_parafac(x, 5, n_iter_max=30, init="random", tol=-1.0, random_state=3 + int(np.argmin(errs)))[0]
    for f, g in zip(a.tl_object.factors, ref.factors):
        assert rel_err(f, g) < FACTOR_RTOL


def test_mma_hals_matches_oracle(mma_from_rank_1):
    """HALS (unfused update kernels) over the tensor-pipe passes."""
    _compare_run(lowrank((40, 500, 16, 8), 6, seed=13), 16, 12, hals=True)


def test_mma_short_runs_without_graph(mma_from_rank_1):
    """n_iter_max < 3: the iteration is enqueued directly (no CUDA graph), dependent launches on the caller's stream."""
    _compare_run(lowrank((40, 500, 16, 8), 6, seed=14), 20, 2)


def test_api_shortcuts_keep_the_reference_values():
    """The API-path shortcuts (finite check through sum(x^2), cached ||X||^2, recon sums shared between the factorization's last
    pass and explained_variance, lazy loc_nans / loc_zeros) return what the long way returns."""
    from cell2cell_b200 import engine
    from cell2cell_b200.tensor import PreBuiltTensor
    x = lowrank((6, 40, 9, 9), 3, seed=31)
    x[x < 0.35] = 0.0
    names = [list(range(s)) for s in x.shape]
    t = PreBuiltTensor(x, order_names=names)
    assert t.mask is None and abs(t.tensor._c2c_norm_x2 - float((x.astype(np.float64) ** 2).sum())) < 1e-6 * float((x ** 2).sum())
    assert int(t.loc_nans.sum()) == 0 and np.array_equal(t.loc_zeros.cpu().numpy(), (x == 0).astype(np.uint8))
    assert t.missing_fraction() == 0.0 and abs(t.sparsity_fraction() - float((x == 0).mean())) < 1e-12
    t.compute_tensor_factorization(rank=3, init="random", random_state=4, n_iter_max=30, tol=-1.0)
    cached = t.explained_variance_
    assert getattr(t.tl_object, "recon_sums_", None) is not None
    t.tl_object.recon_sums_ = None
    assert abs(t.explained_variance() - cached) < 1e-9
    rec = t.tl_object.to_tensor().double()
    X = t.tensor.double()
    want = 1.0 - float(((X - rec) - (X - rec).mean()).norm() / (X - X.mean()).norm())
    assert abs(cached - want) < 1e-5
    # a tensor with NaN / inf takes the full clean-up path (tensor.py:933-946)
    y = x.copy()
    y[0, 1, 2, 3] = np.nan
    y[1, 0, 0, 0] = np.inf
    u = PreBuiltTensor(y, order_names=names)
    assert int(u.loc_nans.sum()) == 1 and bool(torch.isfinite(u.tensor).all()) and not hasattr(u.tensor, "_c2c_norm_x2")
    assert float(u.tensor[0, 1, 2, 3]) == 0.0 and float(u.tensor[1, 0, 0, 0]) == float(np.finfo(np.float32).max)
    assert abs(engine.sumsq(t.tensor) - t.tensor._c2c_norm_x2) == 0.0


@pytest.mark.parametrize("shape,rank", [((4, 30, 6, 6), 3), ((5, 11, 17, 17), 5), ((10, 6, 8), 2), ((40, 500, 16, 8), 16)])
@pytest.mark.parametrize("masked", [False, True])
def test_parafac_als_matches_oracle(shape, rank, masked):
    """tf_type='parafac' (factorization.py:116-126): ALS over the two streaming passes + fp64 R x R solves against the
    numpy restatement of tensorly's parafac, identical random initial factors, fixed iteration count."""
    from cell2cell_b200.tensor.factorization import _compute_tensor_factorization
    n_iter = 12
    x = lowrank(shape, min(rank, 4), seed=5, noise=0.05)
    mask = (np.random.RandomState(9).rand(*shape) > 0.15).astype(np.uint8) if masked else None
    xin = x * mask if masked else x
    ref, ref_err = o.parafac(xin.astype(np.float64), rank, n_iter_max=n_iter, init="random", tol=-1.0, random_state=888,
                             mask=None if mask is None else mask.astype(np.float64))
    got, got_err = _compute_tensor_factorization(xin, rank, tf_type="parafac", init="svd" if masked else "random",
                                                 random_state=888, mask=mask, n_iter_max=n_iter, tol=-1.0,
                                                 return_errors=True)
    assert len(got_err) == len(ref_err) == n_iter
    # ALS factors are only defined up to the conditioning of the normal equations: compare the reconstructions
    rec_ref, rec_got = ref.to_tensor(), o.cp_to_tensor(None, [f.astype(np.float64) for f in got.factors])
    assert rel_err(rec_got, rec_ref) < 1e-3, rel_err(rec_got, rec_ref)
    # the per-iteration errors are tensorly's Gram form sqrt(|nx2 + rn2 - 2 ip|) / sqrt(nx2): in fp32 it cancels once the
    # error is small (1e-7 relative on the squares is 5e-5 on an error of 1e-3)
    np.testing.assert_allclose(np.asarray(got_err), np.asarray(ref_err), rtol=1e-3, atol=1e-4 if masked else 1e-5)
    if masked:
        from cell2cell_b200.tensor.factorization import _compute_norm_error
        e_ref = o.masked_norm_error(xin.astype(np.float64), ref, mask.astype(np.float64))
        assert abs(_compute_norm_error(xin, got, mask) - e_ref) <= 1e-3 * e_ref + 1e-6
    else:
        direct = np.linalg.norm(x - rec_got) / np.linalg.norm(x)
        assert abs(got_err[-1] - direct) <= 1e-4 * direct + 1e-7


def test_parafac_through_the_api_and_rejected_options():
    from cell2cell_b200.tensor import PreBuiltTensor
    from cell2cell_b200.tensor.factorization import _compute_tensor_factorization, parafac
    shape = (6, 40, 8, 8)
    x = lowrank(shape, 3, seed=2, noise=0.02)
    t = PreBuiltTensor(x, order_names=[["e%d_%d" % (m, i) for i in range(s)] for m, s in enumerate(shape)])
    t.compute_tensor_factorization(rank=3, tf_type="parafac", init="svd", n_iter_max=50, tol=1e-7)
    assert t.rank == 3 and list(t.factors.keys())[0].startswith("Dimension") or len(t.factors) == 4
    assert t.explained_variance_ > 0.99
    with pytest.raises(NotImplementedError):
        parafac(x, 3, orthogonalise=True)
    with pytest.raises(NotImplementedError):
        _compute_tensor_factorization(x, 3, tf_type="constrained_parafac", unimodality=True)
    with pytest.raises(ValueError):
        _compute_tensor_factorization(x, 3, tf_type="tucker")


@pytest.mark.parametrize("shape,rank", [((13, 60, 17, 17), 10), ((9, 50, 16, 12), 4), ((4, 5, 7, 8, 12), 6), ((40, 500, 16, 8), 20)])
def test_masked_run_dense_plus_correction_equals_fused_imputation(shape, rank, monkeypatch):
    """The masked factorization as unmasked passes + missing-entry corrections (c2c_corr.cu) against the kernels that
    impute inside the pass (C2C_B200_NO_MASK_CORR=1) and against the oracle; values at the missing entries of the input do
    not matter (a tensor that is not zero there takes the imputing kernels)."""
    from cell2cell_b200.tensor.factorization import non_negative_parafac
    rs = np.random.RandomState(11)
    x = lowrank(shape, 4, seed=3, noise=0.05)
    mask = (rs.rand(*shape) > 0.2).astype(np.uint8)
    mask[1::3, ..., 2, :] = 0
    mask[1::3, ..., :, 2] = 0
    mask[0, 1] = 0
    xz = x * mask
    got, got_err = _compare_run(xz, rank, 25, mask=mask)
    monkeypatch.setenv("C2C_B200_NO_MASK_CORR", "1")
    old, old_err = non_negative_parafac(xz, rank, n_iter_max=25, init="random", tol=-1.0, random_state=888, mask=mask,
                                        return_errors=True)
    monkeypatch.delenv("C2C_B200_NO_MASK_CORR")
    garbage = xz + (1 - mask) * rs.rand(*shape).astype(x.dtype)
    dirty, dirty_err = non_negative_parafac(garbage, rank, n_iter_max=25, init="random", tol=-1.0, random_state=888, mask=mask,
                                            return_errors=True)
    for a, b, c in zip(got.factors, old.factors, dirty.factors):
        assert rel_err(a, b) < 2e-5, rel_err(a, b)
        assert rel_err(c, b) < 2e-5, rel_err(c, b)
    np.testing.assert_allclose(np.asarray(got_err), np.asarray(old_err), rtol=1e-4, atol=3e-6)


# ----------------------------------------------------------------------------------------------------------------------
# tf_type='constrained_parafac' (tensorly's AO-ADMM; factorization.py:128-136): passes + the single-CTA ADMM kernel
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,rank", [((4, 30, 6, 6), 3), ((5, 11, 17, 17), 5), ((10, 6, 8), 2), ((40, 500, 16, 8), 16)])
@pytest.mark.parametrize("cons", [dict(non_negative=True), dict(l1_reg=0.05), dict(l2_square_reg={0: 0.1}), dict(),
                                  dict(non_negative={0: True, 1: True}, l1_reg={2: 0.02})])
def test_constrained_parafac_matches_oracle(shape, rank, cons):
    from cell2cell_b200.tensor.factorization import constrained_parafac
    x = lowrank(shape, max(2, rank - 1), seed=31, noise=0.05)
    if "l1_reg" in cons and len(shape) == 3:
        cons = {k: ({m: v for m, v in val.items() if m < 3} if isinstance(val, dict) else val) for k, val in cons.items()}
    n_iter = 12
    ref, ref_err = o.constrained_parafac(x.astype(np.float64), rank, n_iter_max=n_iter, init="random", random_state=5,
                                         tol_outer=0, **cons)
    got, got_err = constrained_parafac(x, rank, n_iter_max=n_iter, init="random", random_state=5, tol_outer=0,
                                       return_errors=True, **cons)
    assert len(got_err) == len(ref_err) == n_iter
    np.testing.assert_allclose(np.asarray(got_err), np.asarray(ref_err), rtol=1e-3, atol=3e-6)
    rec = o.cp_to_tensor(None, [np.asarray(f, dtype=np.float64) for f in got.factors])
    rec_ref = o.cp_to_tensor(None, ref.factors)
    assert rel_err(rec, rec_ref) < 1e-3
    if cons.get("non_negative") is True:
        assert all((np.asarray(f) >= 0).all() for f in got.factors)


def test_constrained_parafac_admm_kernel_alone():
    """One call of c2c_admm_update against the numpy restatement of tensorly's admm (rank not a multiple of 4, every prox)."""
    from cell2cell_b200 import engine
    rs = np.random.RandomState(3)
    I, R = 700, 7
    dev = torch.device("cuda")
    for name, param in ((None, 0.0), ("non_negative", 0.0), ("l1_reg", 0.3), ("l2_square_reg", 0.25)):
        B = rs.rand(50, R)
        G = (B.T @ B).astype(np.float32)
        M = rs.randn(I, R).astype(np.float32) * 3
        x0, d0 = rs.randn(I, R), 0.1 * rs.randn(I, R)
        want_x, want_split, want_dual = o.admm(M.astype(np.float64), G.astype(np.float64), x0.copy(), d0.copy(), 10, name, param, 1e-6)
        A = engine.pack_factor(x0.astype(np.float32), R, dev)
        X64, dual = torch.as_tensor(x0.copy()).to(dev), torch.as_tensor(d0.copy()).to(dev)
        aux = torch.zeros_like(X64)
        Gp = torch.zeros((engine.ldr_of(R), engine.ldr_of(R)), dtype=torch.float32, device=dev)
        Gp[:R, :R] = torch.as_tensor(G).to(dev)
        out = engine.admm_update(A, X64, dual, aux, engine.pack_factor(M, R, dev), Gp, R, constraint=name, param=param,
                                 n_inner=10, tol=1e-6)
        np.testing.assert_allclose(X64.cpu().numpy(), want_x, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(aux.cpu().numpy(), want_split.T, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(dual.cpu().numpy(), want_dual, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(A[:, :R].cpu().numpy(), want_x.astype(np.float32), rtol=1e-6, atol=1e-7)
        want_c = np.linalg.norm(want_x - want_split.T) / np.linalg.norm(want_x)
        assert abs(float(out[0]) - want_c) <= 1e-8 * max(want_c, 1e-30) + 1e-14 and 1 <= int(out[1]) <= 10


def test_constrained_parafac_through_the_api():
    from cell2cell_b200.tensor import PreBuiltTensor
    shape = (6, 40, 8, 8)
    x = lowrank(shape, 3, seed=2, noise=0.02)
    t = PreBuiltTensor(x, order_names=[["e%d_%d" % (m, i) for i in range(s)] for m, s in enumerate(shape)])
    t.compute_tensor_factorization(rank=3, tf_type="constrained_parafac", init="random", random_state=0, n_iter_max=60,
                                   non_negative=True)
    assert t.rank == 3 and t.explained_variance_ > 0.95     # (the oracle: 0.985 after its 11 outer iterations)
    assert all((df.values >= 0).all() for df in t.factors.values())


# ----------------------------------------------------------------------------------------------------------------------
# unpolled runs are only enqueued; the bench's kernel-only timing aid
# ----------------------------------------------------------------------------------------------------------------------
def test_unpolled_runs_are_stream_ordered_and_equal_polled_runs():
    """check_every <= 0 (what fixed-iteration runs pass): the call returns once the iterations are enqueued and the CUDA-graph
    objects are parked; a second run enqueued right behind it on the same stream, sharing the workspace, must see the first one's
    work finished (stream order) -- both give exactly what polled runs give."""
    from cell2cell_b200 import engine
    from cell2cell_b200._lib import lib
    shape, rank, n_it = (6, 300, 40, 40), 10, 25
    X = torch.as_tensor(lowrank(shape, 4, seed=9)).cuda()
    ws = engine.Workspace(shape, rank, False, X.device)
    mk = lambda sd: [engine.pack_factor(np.random.RandomState(sd + k).random_sample((s, rank)).astype(np.float32), rank, X.device)  # noqa: E731
                     for k, s in enumerate(shape)]
    a1, a2, b1, b2 = mk(1), mk(50), mk(1), mk(50)
    e1 = engine.ncp_run(X, None, a1, rank, n_iter_max=n_it, tol=-1.0, check_every=0, ws=ws, sync=False)
    e2 = engine.ncp_run(X, None, a2, rank, n_iter_max=n_it, tol=-1.0, check_every=0, ws=ws, sync=False)
    torch.cuda.synchronize()
    assert lib().c2c_release_pending() == 0
    p1 = engine.ncp_run(X, None, b1, rank, n_iter_max=n_it, tol=1e-30, check_every=5, ws=ws)      # polled every 5 iterations
    p2 = engine.ncp_run(X, None, b2, rank, n_iter_max=n_it, tol=1e-30, check_every=5, ws=ws)
    for got, want in ((a1, b1), (a2, b2)):
        for g, w in zip(got, want):
            assert rel_err(g.cpu().numpy(), w.cpu().numpy()) < 1e-5
    np.testing.assert_allclose(e1[0].cpu().numpy()[:n_it - 1], p1[0][:n_it - 1], rtol=1e-5)
    np.testing.assert_allclose(e2[0].cpu().numpy()[:n_it - 1], p2[0][:n_it - 1], rtol=1e-5)
    assert int(e1[1].cpu()[0]) == n_it and p1[1] == n_it


def test_time_pass_kernel_times_the_streaming_kernels_alone():
    from cell2cell_b200 import engine
    from cell2cell_b200._lib import C2CError
    shape, rank = (60, 400, 40, 40), 10
    X = torch.as_tensor(lowrank(shape, 3, seed=2)).cuda()
    f = [engine.pack_factor(np.random.RandomState(k).random_sample((s, rank)).astype(np.float32), rank, X.device)
         for k, s in enumerate(shape)]
    for which, names in ((0, ("plane_stream", "plane_mma")), (1, ("fiber_stream", "fiber_mma"))):
        ms, name = engine.time_pass_kernel(X, f, rank, which, reps=5)
        assert name in names and 0.0 < ms < 5.0, (name, ms)
    small = torch.as_tensor(lowrank((3, 7, 5, 6), 2, seed=1)).cuda()               # register-staged kernels: refused
    fs = [engine.pack_factor(np.random.RandomState(k).random_sample((s, 3)).astype(np.float32), 3, small.device)
          for k, s in enumerate(small.shape)]
    with pytest.raises(C2CError):
        engine.time_pass_kernel(small, fs, 3, 0, reps=2)

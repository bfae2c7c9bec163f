"""Mask plan of the masked factorization (cell2cell_b200/csrc/c2c_plan.cu) on the GPU, through the C ABI: the separable model
and the residual counts against their numpy statement (tests/mask_model.py), the corrected contractions against the einsum
of the imputed tensor (tensorly's `tensor*mask + cp_to_tensor(cp, mask=1-mask)`, cell2cell/tensor/coupled_factorization.py:
342-345), and whole masked runs against the oracle and against the kernels that impute inside the pass."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ncp_oracle as o                                   # noqa: E402
from tests.mask_model import fit_separable, khatri_rao_rows, list_slots   # noqa: E402
from tests.test_gpu_parity import _compare_run, lowrank, rel_err     # noqa: E402

pytestmark = pytest.mark.gpu


def make_mask(shape, kind, seed=0):
    """kind: 'outer' (separable: planes / cell rows + columns of a context missing), 'random' (unstructured),
    'mixed' (separable + scattered entries), 'sparse' (a handful of scattered entries)."""
    rs = np.random.RandomState(seed)
    C0, I2, I3 = shape[0], shape[-2], shape[-1]
    P = int(np.prod(shape[:-2]))
    m = np.ones((P, I2, I3), dtype=np.uint8)
    ppc = P // C0
    if kind in ("outer", "mixed"):
        for c in range(C0):
            if c % 3 == 1:
                k = rs.randint(0, min(I2, I3))
                m[c * ppc:(c + 1) * ppc, k, :] = 0
                m[c * ppc:(c + 1) * ppc, :, k] = 0
            if c % 4 == 2 and I2 > 2:
                m[c * ppc:(c + 1) * ppc, I2 - 1, :] = 0
        m[rs.rand(P) < 0.1] = 0
    if kind == "random":
        m[rs.rand(P, I2, I3) < 0.2] = 0
    if kind == "mixed":
        m[rs.rand(P, I2, I3) < 0.03] = 0
    if kind == "sparse":
        idx = rs.randint(0, m.size, size=7)
        m.reshape(-1)[idx] = 0
    return m.reshape(shape)


PLAN_SHAPES = [((5, 23, 8, 12), 6), ((13, 40, 17, 17), 10), ((7, 9, 16), 4), ((3, 4, 6, 8, 8), 5), ((6, 300, 40, 40), 20),
               ((4, 700, 6, 6), 32), ((9, 64, 16, 16), 13)]


@pytest.mark.parametrize("shape,rank", PLAN_SHAPES)
@pytest.mark.parametrize("kind", ["outer", "random", "mixed", "sparse"])
def test_plan_model_and_corrected_contractions(shape, rank, kind):
    from cell2cell_b200 import engine
    mask = make_mask(shape, kind, seed=len(shape) + rank)
    x = lowrank(shape, 3, seed=1) * mask
    rs = np.random.RandomState(5)
    fs = [rs.random_sample((s, rank)) for s in shape]
    X = torch.as_tensor(x).cuda()
    M = torch.as_tensor(mask).cuda()
    plan = engine.MaskPlan(X, M)
    g, hs, hv, resid = fit_separable(mask)
    assert plan.usable
    assert plan.n_missing == int((mask == 0).sum())
    assert plan.n_residual == int(resid.sum())
    assert plan.model_missing == bool((~g).any() or (~hs).any() or (~hv).any())
    by_plane, by_pos = list_slots(resid)                           # class padding of both residual lists
    assert plan.n_list_slots == by_plane and int(plan.info[3]) == by_pos
    if kind == "outer":
        assert plan.n_residual == 0
    A = [engine.pack_factor(f.astype(np.float32), rank, X.device) for f in fs]
    f32 = [f.astype(np.float32).astype(np.float64) for f in fs]
    W, B = khatri_rao_rows(f32[:-2]), khatri_rao_rows(f32[-2:])
    m2 = mask.reshape(W.shape[0], B.shape[0]).astype(np.float64)
    ximp = x.reshape(m2.shape).astype(np.float64) * m2 + (W @ B.T) * (1 - m2)
    T = engine.masked_contract_planned(X, M, A, rank, 0, plan=plan).cpu().numpy()[:, :rank]
    U = engine.masked_contract_planned(X, M, A, rank, 1, plan=plan).cpu().numpy()[:, :rank]
    assert rel_err(T, ximp @ B) < 2e-6, rel_err(T, ximp @ B)
    assert rel_err(U, ximp.T @ W) < 2e-6, rel_err(U, ximp.T @ W)
    # a tensor that is not zero under its mask cannot use the plan
    dirty = torch.as_tensor(x + (1 - mask) * 0.5).float().cuda()
    assert not engine.MaskPlan(dirty, M).usable


@pytest.mark.parametrize("shape,rank", [((13, 60, 17, 17), 10), ((9, 50, 16, 12), 4), ((4, 5, 7, 8, 12), 6), ((40, 500, 16, 8), 20),
                                        ((10, 6, 8), 2)])
@pytest.mark.parametrize("kind", ["outer", "mixed", "random"])
def test_masked_run_with_plan_matches_oracle_and_imputing_kernels(shape, rank, kind, monkeypatch):
    from cell2cell_b200.tensor.factorization import non_negative_parafac
    mask = make_mask(shape, kind, seed=3)
    xz = lowrank(shape, 4, seed=3, noise=0.05) * mask
    got, got_err = _compare_run(xz, rank, 25, mask=mask)                       # the plan path, against the oracle
    monkeypatch.setenv("C2C_B200_NO_MASK_PLAN", "1")                           # mask-scanning corrections (c2c_corr.cu)
    scan, scan_err = non_negative_parafac(xz, rank, n_iter_max=25, init="random", tol=-1.0, random_state=888, mask=mask,
                                          return_errors=True)
    monkeypatch.setenv("C2C_B200_NO_MASK_CORR", "1")                           # kernels that impute inside the pass
    old, old_err = non_negative_parafac(xz, rank, n_iter_max=25, init="random", tol=-1.0, random_state=888, mask=mask,
                                        return_errors=True)
    for a, b, c in zip(got.factors, scan.factors, old.factors):
        assert rel_err(a, c) < 2e-5, rel_err(a, c)
        assert rel_err(b, c) < 2e-5, rel_err(b, c)
    np.testing.assert_allclose(np.asarray(got_err), np.asarray(old_err), rtol=1e-4, atol=3e-6)


def test_all_ones_mask_plan_runs_the_unmasked_iteration():
    from cell2cell_b200.tensor.factorization import non_negative_parafac
    shape, rank = (6, 40, 8, 8), 5
    x = lowrank(shape, 3, seed=2)
    a, ea = non_negative_parafac(x, rank, n_iter_max=20, init="random", tol=-1.0, random_state=1, mask=np.ones(shape, np.uint8),
                                 return_errors=True)
    b, eb = non_negative_parafac(x, rank, n_iter_max=20, init="random", tol=-1.0, random_state=1, return_errors=True)
    for p, q in zip(a.factors, b.factors):               # same kernels; the fp64 Gram atomics may differ in the last bit
        assert rel_err(p, q) < 1e-6
    np.testing.assert_allclose(np.asarray(ea[:-1]), np.asarray(eb[:-1]), rtol=1e-6)


def test_masked_stop_rule_with_plan():
    """tol > 0: the planned masked run stops where the oracle stops (within one iteration: fp32 vs fp64 error differences)."""
    from cell2cell_b200.tensor.factorization import non_negative_parafac
    shape, rank = (8, 30, 9, 9), 3
    mask = make_mask(shape, "mixed", seed=8)
    xz = lowrank(shape, 3, seed=4, noise=0.01) * mask
    ref, ref_err = o.non_negative_parafac(xz.astype(np.float64), rank, n_iter_max=200, init="random", tol=1e-5, random_state=3,
                                          mask=mask.astype(np.float64))
    got, got_err = non_negative_parafac(xz, rank, n_iter_max=200, init="random", tol=1e-5, random_state=3, mask=mask,
                                        return_errors=True)
    assert abs(len(got_err) - len(ref_err)) <= 2


@pytest.mark.parametrize("max_parts", ["1", "3"])
def test_position_lists_walked_several_chunks_per_cta(max_parts, monkeypatch):
    """Beyond ~300 k planes a CTA of the position-side correction walks several 1024-plane chunks and accumulates into its
    partial (same lane, fixed order).  C2C_B200_PLAN_MAX_PARTS forces that path on a 7-chunk tensor."""
    from cell2cell_b200 import engine
    shape, rank = (7, 1000, 8, 8), 10
    mask = make_mask(shape, "random", seed=11)
    x = lowrank(shape, 3, seed=1) * mask
    rs = np.random.RandomState(5)
    fs = [rs.random_sample((s, rank)) for s in shape]
    X, M = torch.as_tensor(x).cuda(), torch.as_tensor(mask).cuda()
    A = [engine.pack_factor(f.astype(np.float32), rank, X.device) for f in fs]
    plan = engine.MaskPlan(X, M)
    ref = engine.masked_contract_planned(X, M, A, rank, 1, plan=plan).cpu().numpy()
    monkeypatch.setenv("C2C_B200_PLAN_MAX_PARTS", max_parts)
    got = engine.masked_contract_planned(X, M, A, rank, 1, plan=plan).cpu().numpy()
    assert rel_err(got, ref) < 1e-6, rel_err(got, ref)
    f32 = [f.astype(np.float32).astype(np.float64) for f in fs]
    W, B = khatri_rao_rows(f32[:-2]), khatri_rao_rows(f32[-2:])
    m2 = mask.reshape(W.shape[0], B.shape[0]).astype(np.float64)
    ximp = x.reshape(m2.shape).astype(np.float64) * m2 + (W @ B.T) * (1 - m2)
    assert rel_err(got[:, :rank], ximp.T @ W) < 2e-6

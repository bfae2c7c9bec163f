"""The algebra behind the mask plan (cell2cell_b200/csrc/c2c_plan.cu), in numpy: for any mask, the contraction of the
reconstruction over the missing entries (tensorly's imputation, cell2cell/tensor/coupled_factorization.py:342-345) equals the
restricted-Gram form over the separable model's missing set plus the list sum over the residual."""
import numpy as np
import pytest

from tests.mask_model import direct_corrections, fit_separable, gram_corrections, list_corrections


@pytest.mark.parametrize("shape", [(5, 7, 4, 3), (4, 6, 5), (3, 2, 4, 3, 5)])
@pytest.mark.parametrize("extra", [0.0, 0.1])
def test_gram_form_plus_residual_list_equals_direct_sum(shape, extra):
    rs = np.random.RandomState(len(shape))
    C0, I2, I3 = shape[0], shape[-2], shape[-1]
    P = int(np.prod(shape[:-2]))
    ppc = P // C0
    g = rs.rand(P) > 0.3
    hs, hv = rs.rand(C0, I2) > 0.3, rs.rand(C0, I3) > 0.2
    m = g[:, None, None] & np.repeat(hs, ppc, 0)[:, :, None] & np.repeat(hv, ppc, 0)[:, None, :]
    m = (m & (rs.rand(P, I2, I3) >= extra)).reshape(shape)
    fs = [rs.rand(s, 3) for s in shape]
    g2, hs2, hv2, resid = fit_separable(m)
    if extra == 0.0:
        assert resid.sum() == 0
    T, U = direct_corrections(m, fs)
    Tg, Ug = gram_corrections(g2, hs2, hv2, fs)
    Tl, Ul = list_corrections(resid, fs)
    np.testing.assert_allclose(Tg + Tl, T, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(Ug + Ul, U, rtol=1e-12, atol=1e-12)


def test_unstructured_mask_is_all_residual():
    rs = np.random.RandomState(0)
    m = rs.rand(6, 8, 5, 5) > 0.2
    g, hs, hv, resid = fit_separable(m)
    assert g.all() and hs.all() and hv.all() and resid.sum() == (~m).sum()

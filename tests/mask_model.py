"""numpy statement of the mask plan of the masked factorization (cell2cell_b200/csrc/c2c_plan.cu) -- test infrastructure.

A `how='outer'` tensor misses whole (context, LR pair) planes (a partner gene absent from the context) and whole rows /
columns of every plane of a context (a cell type absent from it): its mask is *separable*,

    observed[p, s, v] = g[p] * hs[c(p), s] * hv[c(p), v]          c(p) = index of the first mode of plane p.

`fit_separable` finds the tightest such model of ANY mask (g / hs / hv = "some entry observed"); the model never calls an
observed entry missing, so   missing = model-missing  +  residual   (disjoint), with the residual = entries the model calls
observed but the mask does not (empty for cell2cell's own masks; everything for an unstructured mask).

tensorly's imputation adds, to the contraction of the zero-filled tensor, the contraction of the reconstruction over the
missing entries (cell2cell/tensor/coupled_factorization.py:342-345).  Over the model-missing set that sum is a difference of
Gram products (`gram_corrections`); over the residual it is a sum over a list (`list_corrections`).
"""
import numpy as np


def fit_separable(mask):
    """(g [P], hs [C0, I2], hv [C0, I3], resid bool [P, I2, I3]) of a 0/1 mask of shape [I_0, ..., I2, I3]."""
    m = np.asarray(mask) != 0
    shape = m.shape
    C0, I2, I3 = shape[0], shape[-2], shape[-1]
    P = int(np.prod(shape[:-2]))
    m3 = m.reshape(P, I2, I3)
    g = m3.any(axis=(1, 2))
    mc = m.reshape(C0, -1, I2, I3)
    hs = mc.any(axis=(1, 3))
    hv = mc.any(axis=(1, 2))
    ppc = P // C0
    model = g[:, None, None] & np.repeat(hs, ppc, axis=0)[:, :, None] & np.repeat(hv, ppc, axis=0)[:, None, :]
    assert not (m3 & ~model).any()
    return g, hs, hv, model & ~m3


def khatri_rao_rows(factors):
    kr = factors[0]
    for f in factors[1:]:
        kr = (kr[:, None, :] * f[None, :, :]).reshape(-1, kr.shape[1])
    return kr


def direct_corrections(mask, factors):
    """(Tcorr [P, R], Ucorr [I2*I3, R]): contraction of (1 - mask) * reconstruction, the definition."""
    lead, (A2, A3) = factors[:-2], factors[-2:]
    W = khatri_rao_rows(lead)
    B = khatri_rao_rows([A2, A3])
    miss = (np.asarray(mask) == 0).reshape(W.shape[0], B.shape[0])
    rec = (W @ B.T) * miss
    return rec @ B, rec.T @ W


def gram_corrections(g, hs, hv, factors):
    """The same sums over the model-missing set only, in Gram form."""
    lead, (A2, A3) = factors[:-2], factors[-2:]
    W = khatri_rao_rows(lead)
    B = khatri_rao_rows([A2, A3])
    P, R = W.shape
    C0, I2 = hs.shape
    I3 = hv.shape[1]
    ppc = P // C0
    F = (A2.T @ A2) * (A3.T @ A3)
    Q = np.stack([((A2 * hs[c][:, None]).T @ A2) * ((A3 * hv[c][:, None]).T @ A3) for c in range(C0)])
    D = np.where(g[:, None, None], F[None] - np.repeat(Q, ppc, axis=0), F[None])         # [P, R, R]
    Tcorr = np.einsum("pq,pqr->pr", W, D)
    Wg = W * g[:, None]
    Y = np.stack([Wg[c * ppc:(c + 1) * ppc].T @ Wg[c * ppc:(c + 1) * ppc] for c in range(C0)])
    FL = W.T @ W
    msv = (hs[:, :, None] & hv[:, None, :]).reshape(C0, I2 * I3).astype(np.float64)        # [C0, E]
    Z = FL[None] - np.einsum("ce,cqr->eqr", msv, Y)
    Ucorr = np.einsum("eq,eqr->er", B, Z)
    return Tcorr, Ucorr


def list_corrections(resid, factors):
    lead, (A2, A3) = factors[:-2], factors[-2:]
    W = khatri_rao_rows(lead)
    B = khatri_rao_rows([A2, A3])
    r2 = resid.reshape(W.shape[0], B.shape[0])
    rec = (W @ B.T) * r2
    return rec @ B, rec.T @ W


def list_slots(resid, chunk=1024):
    """(slots of the by-plane lists, slots of the by-position lists) of the plan's class-ordered residual lists: a list row is
    padded to 8 x its longest class -- class = in-plane position mod 8 for a plane's row, plane number within the chunk mod 8 for
    the row of a (chunk of `chunk` planes, position) pair (c2c_plan.cu: plan_count_rows / plan_count_pos)."""
    P = resid.shape[0]
    r2 = resid.reshape(P, -1)
    E = r2.shape[1]
    cls_e = np.arange(E) % 8
    by_plane = sum(8 * int(max(np.bincount(cls_e[row], minlength=8))) for row in r2 if row.any())
    by_pos = 0
    for p0 in range(0, P, chunk):
        blk = r2[p0:p0 + chunk]
        cls_p = np.arange(blk.shape[0]) % 8
        counts = np.stack([blk[cls_p == c].sum(axis=0) for c in range(8)])               # [8, E]
        by_pos += int(8 * counts.max(axis=0).sum())
    return by_plane, by_pos

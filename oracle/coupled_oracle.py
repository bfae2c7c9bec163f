"""CPU oracle for Coupled Tensor Component Analysis: a numpy (float64) restatement of the reference's
``coupled_non_negative_parafac`` (cell2cell/tensor/coupled_factorization.py:123-482) and of the combined-error bookkeeping
around it (``coupled_factorization.py:598-641``, ``coupled_tensor.py:218-244``).

TEST INFRASTRUCTURE ONLY -- imported by ``tests/`` (and ``tests/golden/make_golden_coupled.py``), never by the product
package ``cell2cell_b200``.

PARITY: the *coupling logic* (mode mapping, update order, averaging of the numerators / denominators of shared modes,
balancing weights, combined stop rule) is PINNED: ``tests/golden/coupled.npz`` holds outputs of the reference's own
``coupled_non_negative_parafac`` -- its file imported unmodified through ``oracle/ref_harness.load_reference_coupled()`` --
and ``tests/test_coupled_oracle.py`` checks this restatement against them to 1e-12.  The tensorly primitives that function
calls (``initialize_cp``, ``unfolding_dot_khatri_rao``, ``error_calc``, ``cp_normalize``, ``cp_to_tensor``) are absent
from this image, so under the harness they are the restatements of ``oracle/ncp_oracle.py`` -- that part stays UNPINNED,
exactly as for the single-tensor factorization (see the header of ncp_oracle.py).
"""
import numpy as np

from . import ncp_oracle as o

__all__ = ["process_mode_mapping", "split_modes", "balance_weights", "combined_error", "error_calc",
           "coupled_non_negative_parafac"]


def process_mode_mapping(ndim1, ndim2, mode_mapping):
    """coupled_factorization.py:15-65: dict passes through; int / list = non-shared modes of two same-order tensors."""
    if isinstance(mode_mapping, dict):
        return mode_mapping
    if ndim1 != ndim2:
        raise ValueError("When tensors have different dimensions, mode_mapping must be a dict.")
    if isinstance(mode_mapping, int):
        non_shared = [mode_mapping]
    elif isinstance(mode_mapping, (list, tuple)):
        non_shared = list(mode_mapping)
    else:
        raise ValueError("mode_mapping must be dict, int, list, or tuple")
    if len(set(non_shared)) != len(non_shared):
        non_shared = list(set(non_shared))
    return {"shared": [(i, i) for i in range(ndim1) if i not in non_shared]}


def split_modes(ndim1, ndim2, shared_pairs):
    """coupled_factorization.py:230-235: the tensor-specific modes, in increasing order."""
    s1 = set(p[0] for p in shared_pairs)
    s2 = set(p[1] for p in shared_pairs)
    return [i for i in range(ndim1) if i not in s1], [i for i in range(ndim2) if i not in s2]


def balance_weights(shape1, shape2, shared_pairs, balance_errors=True):
    """coupled_factorization.py:258-269."""
    if not balance_errors:
        return 1.0, 1.0
    only1, only2 = split_modes(len(shape1), len(shape2), shared_pairs)
    n1 = np.prod([shape1[i] for i in only1]) if only1 else 1
    n2 = np.prod([shape2[i] for i in only2]) if only2 else 1
    total = n1 + n2
    if total > 0:
        return (total / n1 if n1 > 0 else 1.0), (total / n2 if n2 > 0 else 1.0)
    return 1.0, 1.0


def combined_error(error1, error2, shape1, shape2, shared_pairs, balance_errors=True):
    """coupled_factorization.py:614-639 / coupled_tensor.py:229-244 (what the elbow and best-of-runs compare)."""
    if not balance_errors:
        return (error1 + error2) / 2
    w1, w2 = balance_weights(shape1, shape2, shared_pairs, True)
    return (w1 * error1 + w2 * error2) / (w1 + w2)


def error_calc(tensor, norm_tensor, weights, factors):
    """tensorly.decomposition._cp.error_calc(sparsity=None, mask=None, mttkrp=None) as called at
    coupled_factorization.py:420-431: with no MTTKRP passed in, tensorly builds the full reconstruction and returns the
    DIRECT norm ||tensor - cp||.  ``tensor`` is the array the loop has been re-imputing (observed entries of the input, the
    reconstruction of the factors before the tensor's last update at the missing ones), ``norm_tensor`` the norm of the
    tensor as passed in -- so for a masked tensor the value differs from the Gram form on ``norm_tensor`` by the squared
    norm of the imputed entries."""
    return np.sqrt(np.sum((tensor - o.cp_to_tensor(weights, factors)) ** 2))


def coupled_non_negative_parafac(tensor1, tensor2, rank, mode_mapping, mask1=None, mask2=None, n_iter_max=100,
                                 init="svd", tol=10e-7, random_state=None, cvg_criterion="abs_rec_error",
                                 balance_errors=True, return_iters=False):
    """coupled_factorization.py:123-482 with ``normalize_factors=False`` (weights stay ones).  Returns
    ``(cp1, cp2, (errors1, errors2))``."""
    x1 = np.array(tensor1, dtype=np.float64)
    x2 = np.array(tensor2, dtype=np.float64)
    eps = np.finfo(x1.dtype).eps                                           # :222 tl.eps(tensor1.dtype)
    mode_mapping = process_mode_mapping(x1.ndim, x2.ndim, mode_mapping)
    shared = list(mode_mapping.get("shared", []))
    if len(shared) == 0:
        raise ValueError("At least one mode must be shared between tensors")
    for a, b in shared:
        if x1.shape[a] != x2.shape[b]:
            raise ValueError("Shared modes must have same size")
    only1, only2 = split_modes(x1.ndim, x2.ndim, shared)
    w1, w2 = balance_weights(x1.shape, x2.shape, shared, balance_errors)
    f1 = o.initialize_cp(x1, rank, init=init, random_state=random_state)   # :272-280, same random_state for both
    f2 = o.initialize_cp(x2, rank, init=init, random_state=random_state)
    wt1 = np.ones(rank)
    wt2 = np.ones(rank)
    for a, b in shared:                                                    # :294-295
        f2[b] = f1[a].copy()
    m1 = None if mask1 is None else np.asarray(mask1, dtype=np.float64)
    m2 = None if mask2 is None else np.asarray(mask2, dtype=np.float64)
    n1 = np.sqrt(np.sum(x1 ** 2))                                          # :298-299, of the tensors as passed
    n2 = np.sqrt(np.sum(x2 ** 2))
    e1, e2 = [], []
    order = [("t1", m, None) for m in only1] + [("t2", None, m) for m in only2] + \
            [("sh", a, b) for a, b in reversed(shared)]                    # :309-320

    def accum(fs, skip):
        g = np.ones((rank, rank))
        for i, f in enumerate(fs):
            if i != skip:
                g = g * (f.T @ f)
        return g

    it_done = 0
    for it in range(n_iter_max):
        for kind, a, b in order:
            if kind == "t1":                                               # :334-348
                g = accum(f1, a)
                if m1 is not None:
                    x1 = x1 * m1 + o.cp_to_tensor(wt1, f1, mask=1 - m1)
                num = np.clip(o.mttkrp(x1, wt1, f1, a), eps, None)
                den = np.clip(f1[a] @ g, eps, None)
                f1[a] = f1[a] * num / den
            elif kind == "t2":                                             # :350-365
                g = accum(f2, b)
                if m2 is not None:
                    x2 = x2 * m2 + o.cp_to_tensor(wt2, f2, mask=1 - m2)
                num = np.clip(o.mttkrp(x2, wt2, f2, b), eps, None)
                den = np.clip(f2[b] @ g, eps, None)
                f2[b] = f2[b] * num / den
            else:                                                          # :367-405
                g1 = accum(f1, a)
                if m1 is not None:
                    x1 = x1 * m1 + o.cp_to_tensor(wt1, f1, mask=1 - m1)
                g2 = accum(f2, b)
                if m2 is not None:
                    x2 = x2 * m2 + o.cp_to_tensor(wt2, f2, mask=1 - m2)
                k1 = o.mttkrp(x1, wt1, f1, a)
                k2 = o.mttkrp(x2, wt2, f2, b)
                num = np.clip((k1 + k2) / 2.0, eps, None)
                den = np.clip((f1[a] @ g1 + f2[b] @ g2) / 2.0, eps, None)
                new = f1[a] * num / den
                f1[a] = new
                f2[b] = new.copy()
        it_done = it + 1
        if tol:                                                            # :421-458
            r1 = error_calc(x1, n1, wt1, f1) / n1
            r2 = error_calc(x2, n2, wt2, f2) / n2
            e1.append(r1)
            e2.append(r2)
            comb = (w1 * r1 + w2 * r2) / (w1 + w2)
            if it >= 1:
                prev = (w1 * e1[-2] + w2 * e2[-2]) / (w1 + w2)
                dec = prev - comb
                if cvg_criterion == "abs_rec_error":
                    stop = abs(dec) < tol
                elif cvg_criterion == "rec_error":
                    stop = dec < tol
                else:
                    raise ValueError("Unknown convergence criterion")
                if stop:
                    break
    out = (o.CPResult(wt1, f1), o.CPResult(wt2, f2), (e1, e2))
    return out + (it_done,) if return_iters else out

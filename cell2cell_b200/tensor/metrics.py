"""CorrIndex similarity between decompositions (Sobhani et al. 2022), for ``elbow_rank_selection(metric='similarity')``.

Same functions as cell2cell/tensor/metrics.py:12-179.  The reference scores the ``n (n - 1) / 2`` pairs of a rank's runs one by
one on the host (normalise, stack, ``|X1^T X2|``, row / column maxima); here the runs' factor matrices are stacked into one
``[n, S, R]`` device array and ONE launch of ``corrindex_pairs_kernel`` (csrc/c2c_metrics.cu, fp64, a CTA per pair) scores
them all.  Argument checks and error messages are the reference's.
"""
import numpy as np
import pandas as pd

from .. import engine

__all__ = ["correlation_index", "pairwise_correlation_index"]

_METHODS = ("stacked", "max_score", "min_score", "avg_score")


def _as_list(factors):
    fs = list(factors.values()) if hasattr(factors, "values") else list(factors)
    return [np.asarray(getattr(a, "values", a), dtype=np.float64) for a in fs]


def _stack(decompositions, method):
    """[n, S, R] float64 + the mode sizes, after the reference's shape checks (metrics.py:55-75)."""
    sets = [_as_list(f) for f in decompositions]
    for fs in sets:
        if len({a.shape[1] for a in fs}) != 1:
            raise ValueError("Factors should be a list of loading matrices of the same rank")
    if method not in _METHODS:
        raise ValueError("The `method` must be either option among {}".format(list(_METHODS)))
    first = sets[0]
    for fs in sets[1:]:
        if method == "stacked":
            ok = sum(a.shape[0] for a in fs) == sum(a.shape[0] for a in first) and fs[0].shape[1] == first[0].shape[1]
        else:
            ok = len(fs) == len(first) and all(a.shape == b.shape for a, b in zip(fs, first))
        if not ok:
            raise ValueError("Factor matrices should be of the same shapes")
    sizes = [a.shape[0] for a in first]
    if method == "stacked":
        sizes = [sum(sizes)]
    return np.stack([np.concatenate(fs, 0) for fs in sets], 0), sizes


def correlation_index(factors_1, factors_2, tol=5e-16, method="stacked"):
    """CorrIndex in [0, 1] (0 = same decomposition up to column scaling and permutation) (metrics.py:12-97)."""
    stacked, sizes = _stack([factors_1, factors_2], method)
    score = float(engine.corrindex_pairs(stacked, sizes, method=method, tol=tol)[0, 1])
    return 0 if score == 0.0 else score


def pairwise_correlation_index(factors, tol=5e-16, method="stacked"):
    """Symmetric DataFrame of CorrIndex scores between all pairs of decompositions (metrics.py:131-179)."""
    n = len(factors)
    if n < 2:
        return pd.DataFrame(np.zeros((n, n)), index=range(n), columns=range(n))
    stacked, sizes = _stack(list(factors), method)
    scores = engine.corrindex_pairs(stacked, sizes, method=method, tol=tol)
    return pd.DataFrame(scores, index=range(n), columns=range(n))

"""CorrIndex similarity between decompositions (Sobhani et al. 2022), for ``elbow_rank_selection(metric='similarity')``.

Same functions as cell2cell/tensor/metrics.py:12-179; the inputs are the small factor matrices (sum(I_n) x R), so this is
host numpy.
"""
from itertools import combinations

import numpy as np
import pandas as pd

__all__ = ["correlation_index", "pairwise_correlation_index"]

_METHODS = ("stacked", "max_score", "min_score", "avg_score")


def _as_list(factors):
    return list(factors.values()) if hasattr(factors, "values") else list(factors)


def _score(x1, x2, tol):
    c = np.abs(np.conj(np.asarray(x1).T) @ np.asarray(x2))
    n = c.shape[0] + c.shape[1]
    s = (np.sum(np.abs(c.max(axis=1) - 1)) + np.sum(np.abs(c.max(axis=0) - 1))) / n
    return 0 if s < tol else s


def correlation_index(factors_1, factors_2, tol=5e-16, method="stacked"):
    """CorrIndex in [0, 1] (0 = same decomposition up to column scaling and permutation) (metrics.py:12-97)."""
    f1 = [np.asarray(a, dtype=float) for a in _as_list(factors_1)]
    f2 = [np.asarray(a, dtype=float) for a in _as_list(factors_2)]
    for fs in (f1, f2):
        if len({a.shape[1] for a in fs}) != 1:
            raise ValueError("Factors should be a list of loading matrices of the same rank")
    if method not in _METHODS:
        raise ValueError("The `method` must be either option among {}".format(list(_METHODS)))
    if method == "stacked":
        f1, f2 = [np.concatenate(f1, 0)], [np.concatenate(f2, 0)]
    for a, b in zip(f1, f2):
        if a.shape != b.shape:
            raise ValueError("Factor matrices should be of the same shapes")
    scores = []
    for a, b in zip(f1, f2):
        na, nb = np.linalg.norm(a, axis=0), np.linalg.norm(b, axis=0)
        if np.any(na == 0) or np.any(nb == 0):
            raise ValueError("Column norms must be non-zero")
        scores.append(_score(a / na, b / nb, tol))
    if method == "stacked":
        return scores[0]
    if method == "max_score":
        return np.max(scores)
    if method == "min_score":
        return np.min(scores)
    return np.mean(scores)


def pairwise_correlation_index(factors, tol=5e-16, method="stacked"):
    """Symmetric DataFrame of CorrIndex scores between all pairs of decompositions (metrics.py:131-179)."""
    n = len(factors)
    scores = pd.DataFrame(np.zeros((n, n)), index=range(n), columns=range(n))
    for p1, p2 in combinations(range(n), 2):
        s = correlation_index(factors[p1], factors[p2], tol=tol, method=method)
        scores.at[p1, p2] = s
        scores.at[p2, p1] = s
    return scores

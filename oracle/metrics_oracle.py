"""CPU oracle of CorrIndex (cell2cell/tensor/metrics.py:12-179): numpy float64 restatement.

TEST INFRASTRUCTURE ONLY -- imported by ``tests/``, never by the product package.  PARITY PINNED: ``tests/test_elbow_metrics.py``
compares it with the reference's own ``metrics.py`` (imported unmodified through ``oracle/ref_harness.py``) for the four
methods; the GPU tests compare ``corrindex_pairs_kernel`` with it."""
from itertools import combinations

import numpy as np

_METHODS = ("stacked", "max_score", "min_score", "avg_score")


def _as_list(factors):
    return [np.asarray(getattr(a, "values", a), dtype=np.float64)
            for a in (factors.values() if hasattr(factors, "values") else factors)]


def _score(x1, x2, tol):
    """metrics.py:100-128: (sum |rowmax(C) - 1| + sum |colmax(C) - 1|) / (rows + cols) of C = |x1^H x2|."""
    c = np.abs(np.conj(x1.T) @ x2)
    s = (np.sum(np.abs(c.max(axis=1) - 1)) + np.sum(np.abs(c.max(axis=0) - 1))) / (c.shape[0] + c.shape[1])
    return 0 if s < tol else s


def correlation_index(factors_1, factors_2, tol=5e-16, method="stacked"):
    """metrics.py:12-97."""
    f1, f2 = _as_list(factors_1), _as_list(factors_2)
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
    return {"stacked": lambda s: s[0], "max_score": np.max, "min_score": np.min, "avg_score": np.mean}[method](scores)


def pairwise_correlation_index(factors, tol=5e-16, method="stacked"):
    """metrics.py:131-179, as a plain [n, n] array."""
    n = len(factors)
    out = np.zeros((n, n))
    for i, j in combinations(range(n), 2):
        out[i, j] = out[j, i] = correlation_index(factors[i], factors[j], tol=tol, method=method)
    return out

"""numpy (float64) emulation of what ``c2c_build_ccc`` / ``c2c_group_aggregate`` do with the index arrays produced by
``cell2cell_b200.tensor.build_index`` -- lets the CPU-only suite check the product's host-side name logic against the
reference-made golden fixtures without a GPU.  Test infrastructure only."""
import warnings

import numpy as np


def emulate_build(mats, idx, C, L, K, score, agg):
    X = np.full((C, L, K, K), np.nan)
    for c in range(C):
        m = np.asarray(mats[c], dtype=np.float64)
        cols = idx["cell_col"][c]
        def side(first, count):
            out = np.full((L, K), np.nan)
            for l in range(L):
                n = count[c * L + l]
                if n == 0:
                    continue
                if n < 0:
                    v = np.zeros(m.shape[1])
                else:
                    rows = idx["sub_rows"][first[c * L + l]: first[c * L + l] + n]
                    sub = m[rows, :]
                    with np.errstate(all="ignore"), warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        if n == 1:
                            v = sub[0]
                        elif agg == "min":
                            v = np.nanmin(sub, axis=0)
                        elif agg == "mean":
                            v = np.nanmean(sub, axis=0)
                        else:
                            v = np.exp(np.nanmean(np.log(sub), axis=0))
                for k in range(K):
                    if cols[k] >= 0:
                        out[l, k] = v[cols[k]]
            return out
        a = side(idx["a_first"], idx["a_count"])
        b = side(idx["b_first"], idx["b_count"])
        with np.errstate(all="ignore"):
            if score == "expression_product":
                X[c] = a[:, :, None] * b[:, None, :]
            elif score == "expression_mean":
                X[c] = (a[:, :, None] + b[:, None, :]) / 2.0
            else:
                X[c] = np.sqrt(a[:, :, None] * b[:, None, :])
    return X


def emulate_group(X, g_off, g_rows, method):
    C = X.shape[0]
    out = np.zeros((C, len(g_off) - 1) + X.shape[2:])
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for g in range(len(g_off) - 1):
            sub = X[:, g_rows[g_off[g]:g_off[g + 1]]]
            if method == "gmean":
                out[:, g] = np.exp(np.mean(np.log(sub), axis=1))
            elif method == "sum":
                out[:, g] = np.nansum(sub, axis=1)
            else:
                out[:, g] = np.nanmean(sub, axis=1)
    return out

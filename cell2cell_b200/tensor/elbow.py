"""Elbow detection and curve smoothing for ``elbow_rank_selection`` (host side; a few dozen points).

``kneedle_elbow`` restates kneed's ``KneeLocator(x, y, S, curve, direction, interp_method='interp1d').elbow`` as called at
cell2cell/tensor/factorization.py:181 (kneed is an un-vendored, un-pinned dependency -- setup.py:102 -- that is not
installable in this image; the algorithm is Satopaa et al. 2011 "Finding a Kneedle in a Haystack" as published in kneed
0.8.x; PARITY UNPINNED).  ``smooth_curve`` restates cell2cell/preprocessing/signal.py:6-34.
"""
import numpy as np
from scipy.signal import argrelextrema, savgol_filter

__all__ = ["kneedle_elbow", "smooth_curve"]


def _normalize(a):
    return (a - a.min()) / (a.max() - a.min())


def kneedle_elbow(x, y, S=1.0, curve="convex", direction="decreasing", interp_method="interp1d"):
    """x value of the first knee/elbow found walking left to right (kneed offline mode), or None."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    if interp_method != "interp1d":
        raise NotImplementedError("only interp_method='interp1d' (the reference's setting) is supported")
    if curve not in ("convex", "concave") or direction not in ("increasing", "decreasing"):
        raise ValueError("Please check that the curve and direction arguments are valid.")
    n = len(x)
    if n < 2:
        return None
    ds_y = y.copy()                       # interp1d evaluated at its own knots
    x_n, y_n = _normalize(x), _normalize(ds_y)
    # transform so that the knee is a maximum of the difference curve (kneed.transform_y)
    if direction == "decreasing":
        if curve == "concave":
            y_n = np.flip(y_n)
        else:
            y_n = y_n.max() - y_n
    elif curve == "convex":
        y_n = np.flip(y_n.max() - y_n)
    y_d = y_n - x_n
    x_d = x_n.copy()
    maxima = argrelextrema(y_d, np.greater_equal)[0]
    minima = argrelextrema(y_d, np.less_equal)[0]
    if maxima.size == 0:
        return None
    tmx = y_d[maxima] - S * np.abs(np.diff(x_n).mean())
    threshold = 0.0
    threshold_index = 0
    maxima_idx = 0
    minima_idx = 0
    knee = None
    for i, _ in enumerate(x_d):
        if i < maxima[0]:
            continue
        j = i + 1
        if i == n - 1:
            break
        if (maxima == i).any():
            threshold = tmx[maxima_idx]
            threshold_index = i
            maxima_idx += 1
        if (minima == i).any():
            threshold = 0.0
            minima_idx += 1
        if y_d[j] < threshold:
            if curve == "convex":
                knee = x[threshold_index] if direction == "decreasing" else x[-(threshold_index + 1)]
            else:
                knee = x[-(threshold_index + 1)] if direction == "decreasing" else x[threshold_index]
            break
    if knee is None:
        return None
    return knee.item() if hasattr(knee, "item") else knee


def smooth_curve(values, window_length=None, polyorder=3, **kwargs):
    """Savitzky-Golay smoothing with the reference's default window (signal.py:27-33)."""
    size = len(values)
    if window_length is None:
        window_length = int(size / min([2, size]))
    if window_length % 2 == 0:
        window_length += 1
    assert polyorder < window_length, "polyorder must be less than window_length."
    return savgol_filter(values, window_length, polyorder, **kwargs)

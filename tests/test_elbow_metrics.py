"""Host-side pieces of elbow_rank_selection: Kneedle restatement and smooth_curve, the CorrIndex oracle (each against the
reference's own functions when /root/reference is present), and the seeded synthetic generators.  The product's CorrIndex is
a device kernel: its parity test is tests/test_gpu_metrics.py."""
import numpy as np
import pytest

from cell2cell_b200.tensor.elbow import kneedle_elbow, smooth_curve
from oracle.metrics_oracle import correlation_index, pairwise_correlation_index
from oracle.ref_harness import load_reference, reference_available


def test_kneedle_on_textbook_curves():
    x = np.arange(1, 26)
    y = 1.0 / x                                   # convex decreasing: sharp elbow early
    assert kneedle_elbow(x, y) in (3, 4, 5, 6)
    y2 = np.concatenate([np.linspace(1.0, 0.2, 5), np.linspace(0.19, 0.15, 20)])   # piecewise linear, knee at x = 5
    assert kneedle_elbow(x, y2) == 5
    assert kneedle_elbow(x, np.linspace(1, 0, 25)) is None or isinstance(kneedle_elbow(x, np.linspace(1, 0, 25)), (int, float))


def test_kneedle_known_answer_from_the_paper_dataset():
    # kneed's README / test-suite dataset (Satopaa et al. figure 2): concave increasing, knee at x = 0.22
    x = np.array([0.0, 0.111, 0.222, 0.333, 0.444, 0.556, 0.667, 0.778, 0.889, 1.0])
    y = np.array([-5.0, 0.263, 1.897, 2.692, 3.163, 3.475, 3.696, 3.861, 3.989, 4.091])
    assert kneedle_elbow(x, y, S=1.0, curve="concave", direction="increasing") == pytest.approx(0.222)


def test_smooth_curve_default_window():
    v = np.random.RandomState(0).rand(25)
    out = smooth_curve(v)
    assert out.shape == v.shape
    if reference_available():
        ref = load_reference()
        np.testing.assert_array_equal(out, ref.signal.smooth_curve(v))


def test_corrindex_invariances_and_reference():
    rs = np.random.RandomState(1)
    fs = [rs.rand(s, 4) for s in (5, 9, 3, 3)]
    perm = [2, 0, 3, 1]
    gs = [f[:, perm] * np.array([2.0, 0.5, 3.0, 1.0]) for f in fs]
    assert correlation_index(fs, gs) == 0
    hs = [rs.rand(s, 4) for s in (5, 9, 3, 3)]
    s = correlation_index(fs, hs)
    assert 0 < s <= 1
    df = pairwise_correlation_index([fs, gs, hs])
    assert df.shape == (3, 3) and df[0, 2] == df[2, 0] == s
    if reference_available():
        ref = load_reference()
        d = lambda l: dict(zip(range(len(l)), l))
        for m in ("stacked", "max_score", "min_score", "avg_score"):
            assert correlation_index(fs, hs, method=m) == pytest.approx(ref.metrics.correlation_index(d(fs), d(hs), method=m), rel=1e-14)
        want = ref.metrics.pairwise_correlation_index([d(fs), d(gs), d(hs)]).values
        np.testing.assert_allclose(df, want, rtol=1e-14, atol=0)


def test_synthetic_generators_are_seeded():
    from cell2cell_b200.synthetic import synthetic_case, synthetic_lowrank_tensor
    m1, p1, kw = synthetic_case("toy", scale=0.05)
    m2, p2, _ = synthetic_case("toy", scale=0.05)
    assert p1.equals(p2) and all(a.equals(b) for a, b in zip(m1, m2))
    x = synthetic_lowrank_tensor((3, 5, 2, 2))
    assert x.dtype == np.float32 and (x >= 0).all()

"""CorrIndex on the device (csrc/c2c_metrics.cu through ``c2c_corrindex_pairs``) against the pinned numpy oracle
(oracle/metrics_oracle.py == the reference's cell2cell/tensor/metrics.py, checked on CPU in tests/test_elbow_metrics.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import metrics_oracle as mo  # noqa: E402

pytestmark = pytest.mark.gpu


def _sets(rs, sizes, R, n):
    return [[rs.rand(s, R) for s in sizes] for _ in range(n)]


@pytest.mark.parametrize("sizes,R,n", [((5, 9, 3, 3), 4, 3), ((60, 2000, 40, 40), 10, 5), ((12, 1054, 6, 6), 25, 10),
                                       ((7, 5, 11), 1, 4), ((13, 190, 17, 17), 33, 3), ((6, 40, 9, 9), 128, 2)])
@pytest.mark.parametrize("method", ["stacked", "max_score", "min_score", "avg_score"])
def test_pairwise_corrindex_matches_the_oracle(sizes, R, n, method):
    from cell2cell_b200.tensor.metrics import correlation_index, pairwise_correlation_index
    rs = np.random.RandomState(R + n)
    sets = _sets(rs, sizes, R, n)
    # a permuted + rescaled copy of the first decomposition: CorrIndex 0 against it
    perm = rs.permutation(R)
    scale = 0.5 + rs.rand(R)                 # ONE scaling per column for every mode: the stacked matrix stays proportional
    sets[-1] = [f[:, perm] * scale for f in sets[0]]
    got = pairwise_correlation_index(sets, method=method)
    want = mo.pairwise_correlation_index(sets, method=method)
    assert got.shape == (n, n) and list(got.index) == list(range(n))
    np.testing.assert_allclose(got.values, want, rtol=1e-11, atol=1e-13)
    # identical up to permutation and scaling: 0, or rounding at the level of the reference's own `tol` (5e-16) switch
    assert got.values[0, n - 1] < 1e-14 and np.all(np.diag(got.values) == 0)
    assert correlation_index(sets[0], sets[1], method=method) == pytest.approx(want[0, 1], rel=1e-11)
    assert correlation_index(sets[0], sets[-1], method=method) < 1e-14


def test_corrindex_argument_errors_are_the_reference_ones():
    from cell2cell_b200.tensor.metrics import correlation_index
    rs = np.random.RandomState(0)
    fs = [rs.rand(5, 3), rs.rand(4, 3)]
    with pytest.raises(ValueError, match="same rank"):
        correlation_index([rs.rand(5, 3), rs.rand(4, 2)], fs)
    with pytest.raises(ValueError, match="must be either option"):
        correlation_index(fs, fs, method="median")
    with pytest.raises(ValueError, match="same shapes"):
        correlation_index(fs, [rs.rand(6, 3), rs.rand(4, 3)])
    zero = [f.copy() for f in fs]
    zero[0][:, 1] = 0
    zero[1][:, 1] = 0
    with pytest.raises(ValueError, match="non-zero"):
        correlation_index(fs, zero)
    # dict of DataFrames, as BaseTensor.factors holds them (metrics.py:53-54)
    import pandas as pd
    d = {"a": pd.DataFrame(fs[0]), "b": pd.DataFrame(fs[1])}
    assert correlation_index(d, d) == 0


def test_similarity_elbow_uses_the_device_corrindex():
    from cell2cell_b200 import engine
    from cell2cell_b200.tensor import PreBuiltTensor
    from tests.cases import _lowrank32
    shape = (5, 40, 6, 6)
    x = _lowrank32(shape, 4, 21, noise=0.1)
    t = PreBuiltTensor(x, order_names=[["n%d_%d" % (k, i) for i in range(s)] for k, s in enumerate(shape)])
    before = engine.launch_count()
    t.elbow_rank_selection(upper_rank=3, runs=3, metric="similarity", init="random", random_state=5, n_iter_max=30, tol=-1.0,
                           output_fig=False, automatic_elbow=False, manual_elbow=2)
    assert t.elbow_metric_raw.shape == (3, 3) and engine.launch_count() > before
    assert np.all(t.elbow_metric_raw >= 0) and np.all(t.elbow_metric_raw <= 1)

"""``run_tensor_cell2cell_pipeline`` (cell2cell/analysis/tensor_pipelines.py:10-215) and the pickle helpers
(cell2cell/io/save_data.py:8-28, read_data.py:467-583): argument handling on CPU, the full flow on the GPU."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from cell2cell_b200 import analysis, io                       # noqa: E402


class _Fake:
    """Records what the pipeline asks of a tensor object."""

    def __init__(self, dim=4):
        self.tensor = np.zeros((2,) * dim)
        self.calls = []
        self.rank = None

    def copy(self):
        c = _Fake(self.tensor.ndim)
        c.calls = self.calls + ["copy"]
        return c

    def to_device(self, device):
        self.calls.append(("to_device", device))

    def elbow_rank_selection(self, **kw):
        self.calls.append(("elbow", kw))
        self.rank = 3

    def compute_tensor_factorization(self, **kw):
        self.calls.append(("factorize", kw))

    def export_factor_loadings(self, filename):
        self.calls.append(("export", filename))


def test_pipeline_regular_and_robust_parameters(tmp_path):
    t = analysis.run_tensor_cell2cell_pipeline(_Fake(), None, tf_optimization="regular", random_state=4, output_fig=False,
                                               upper_rank=7, device="cuda:0")
    assert t.calls[0] == ("to_device", "cuda:0")
    kind, kw = t.calls[1]
    assert kind == "elbow" and kw["runs"] == 10 and kw["tol"] == 1e-7 and kw["n_iter_max"] == 100 and kw["upper_rank"] == 7
    assert kw["init"] == "random" and kw["metric"] == "error" and kw["automatic_elbow"] and kw["random_state"] == 4
    kind, kw = t.calls[2]
    assert kind == "factorize" and kw["rank"] == 3 and kw["runs"] == 1 and kw["normalize_loadings"] and kw["tol"] == 1e-7
    assert len(t.calls) == 3
    t = analysis.run_tensor_cell2cell_pipeline(_Fake(), None, copy_tensor=True, rank=5, tf_optimization="robust",
                                               output_fig=False, output_folder=str(tmp_path))
    assert t.calls[0] == "copy"
    kind, kw = t.calls[1]
    assert kind == "factorize" and kw["rank"] == 5 and kw["runs"] == 100 and kw["tol"] == 1e-8 and kw["n_iter_max"] == 500
    assert t.calls[2] == ("export", str(tmp_path) + "/Loadings.xlsx")


def test_pipeline_argument_errors():
    with pytest.raises(ValueError, match="robust"):
        analysis.run_tensor_cell2cell_pipeline(_Fake(), None, tf_optimization="fast", output_fig=False)
    with pytest.raises(ValueError, match="higher to 5"):
        analysis.run_tensor_cell2cell_pipeline(_Fake(dim=6), None, output_fig=False)
    with pytest.raises(AssertionError):
        analysis.run_tensor_cell2cell_pipeline(_Fake(), None, cmaps=["tab10"], output_fig=False)
    with pytest.raises(ValueError, match="backend"):
        analysis.run_tensor_cell2cell_pipeline(_Fake(), None, backend="numpy", output_fig=False)


def test_pickle_helpers_round_trip(tmp_path):
    p = str(tmp_path / "v.pkl")
    io.export_variable_with_pickle({"a": np.arange(5), "b": "x"}, p)
    back = io.load_variable_with_pickle(p)
    assert back["b"] == "x" and np.array_equal(back["a"], np.arange(5))


@pytest.mark.gpu
def test_pipeline_on_device_equals_direct_calls(tmp_path, capsys):
    from cell2cell_b200.tensor import PreBuiltTensor
    rs = np.random.RandomState(0)
    fs = [rs.gamma(1.0, 1.0, size=(s, 3)) for s in (5, 40, 6, 6)]
    x = np.einsum("ar,br,cr,dr->abcd", *fs) + 0.05 * rs.random_sample((5, 40, 6, 6))
    names = [["n%d_%d" % (k, i) for i in range(s)] for k, s in enumerate(x.shape)]
    t = PreBuiltTensor(x, order_names=names)
    out = analysis.run_tensor_cell2cell_pipeline(t, None, copy_tensor=True, rank=None, tf_optimization="regular",
                                                 random_state=1, upper_rank=6, output_fig=False, device="cuda:0")
    assert out is not t and t.factors is None and out.factors is not None
    ref = PreBuiltTensor(x, order_names=names)
    ref.elbow_rank_selection(upper_rank=6, runs=10, init="random", svd="numpy_svd", automatic_elbow=True, metric="error",
                             output_fig=False, random_state=1, tol=1e-7, n_iter_max=100)
    assert out.rank == ref.rank
    np.testing.assert_allclose([l[1] for l in out.elbow_metric_mean], [l[1] for l in ref.elbow_metric_mean], rtol=1e-5)
    ref.compute_tensor_factorization(rank=ref.rank, init="random", svd="numpy_svd", random_state=1, runs=1,
                                     normalize_loadings=True, tol=1e-7, n_iter_max=100)
    for k in ref.factors:
        np.testing.assert_allclose(out.factors[k].values, ref.factors[k].values, rtol=1e-4, atol=1e-7)
    text = capsys.readouterr().out
    assert "Running Elbow Analysis" in text and "Running Tensor Factorization" in text
    # pickle -> load_tensor puts everything back on the device
    p = str(tmp_path / "tensor.pkl")
    out.write_file(p)
    back = io.load_tensor(p, device="cuda:0")
    assert back.tensor.is_cuda and torch.equal(back.tensor, out.tensor) and list(back.factors) == list(out.factors)
    with pytest.raises(ValueError):
        io.load_tensor(p, backend="numpy")


def test_generate_tensor_metadata_equals_the_reference_function():
    """``generate_tensor_metadata`` against the reference's own function (file imported unmodified through the harness) on a
    stand-in tensor object: mapped orders, short / long unmapped orders, order labels given or not, 3- / 4- / 5-way."""
    import types
    import numpy as np
    import pandas as pd
    import pytest
    from oracle.ref_harness import load_reference, reference_available
    from cell2cell_b200.tensor.tensor import generate_tensor_metadata
    if not reference_available():
        pytest.skip("reference not present (GPU box)")
    ref = load_reference().tensor.generate_tensor_metadata
    for shape, labels in (((3, 80, 4, 4), None), ((3, 80, 4, 4), ["Ctx", "LR", "S", "R"]), ((80, 4, 4), None), ((2, 3, 90, 5, 5), None)):
        names = [["e%d_%d" % (k, i) for i in range(n)] for k, n in enumerate(shape)]
        obj = types.SimpleNamespace(tensor=np.zeros(shape), order_names=names, order_labels=labels)
        dicts = [None] * len(shape)
        dicts[0] = {n: "g%d" % (i % 2) for i, n in enumerate(names[0])}
        for fill in (True, False):
            want = ref(obj, list(dicts), fill_with_order_elements=fill)
            got = generate_tensor_metadata(obj, list(dicts), fill_with_order_elements=fill)
            assert len(got) == len(want)
            for a, b in zip(got, want):
                assert (a is None) == (b is None)
                if a is not None:
                    pd.testing.assert_frame_equal(a.reset_index(drop=True), b.reset_index(drop=True), check_dtype=False)
    with pytest.raises(ValueError):
        generate_tensor_metadata(types.SimpleNamespace(tensor=np.zeros((4, 4)), order_names=[["a"] * 4] * 2, order_labels=None),
                                 [None, None])

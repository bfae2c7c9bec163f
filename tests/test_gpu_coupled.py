"""GPU parity of Coupled Tensor Component Analysis (run with ``-m gpu`` on a B200): the device path
(``cell2cell_b200.tensor.coupled_factorization`` -> libc2c_b200.so) against the golden outputs of the reference's own
``coupled_non_negative_parafac`` (tests/golden/coupled.npz) and against ``oracle/coupled_oracle.py`` on seeded inputs.

Tolerance (BASELINE.json north_star): factors and final reconstruction errors within 1e-4 relative after a fixed iteration
count from identical initial factors; the stop iteration of a real stop rule within 2 iterations.  The per-iteration error
history is tensorly's Gram form sqrt(|nx2 + rn2 - 2 ip|), which cancels in fp32 once the error is small: it is compared at
3e-6 absolute for the fixed-count cases (errors >= 0.03) and 1e-5 for the stop-rule cases (errors down to 0.01); the kept
(last) error of an unmasked tensor is the directly accumulated one and is held to 1e-4 relative."""
import os
import pickle

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import coupled_oracle as co                       # noqa: E402
from oracle import ncp_oracle as o                            # noqa: E402
from tests.cases import COUPLED_CASES, coupled_inputs         # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "coupled.npz"))
FACTOR_RTOL = 1e-4
ERROR_RTOL = 1e-4


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-300))


def test_mu_update_coupled_kernel_matches_numpy():
    from cell2cell_b200 import engine
    rs = np.random.RandomState(0)
    for I, R in ((7, 3), (100, 10), (2000, 25)):
        a = rs.rand(I, R).astype(np.float32)
        m1, m2 = rs.rand(I, R).astype(np.float32), rs.rand(I, R).astype(np.float32)
        g1, g2 = rs.rand(R, R).astype(np.float32), rs.rand(R, R).astype(np.float32)
        m1[0, 0] = m2[0, 0] = 0.0                                   # numerator clipped to eps
        dev = torch.device("cuda")
        A1, A2 = engine.pack_factor(a, R, dev), engine.pack_factor(a, R, dev)
        G1, G2 = engine.pack_factor(g1, R, dev)[:R].contiguous(), engine.pack_factor(g2, R, dev)[:R].contiguous()
        ldr = engine.ldr_of(R)
        G1p = torch.zeros((ldr, ldr), device=dev); G1p[:R] = G1
        G2p = torch.zeros((ldr, ldr), device=dev); G2p[:R] = G2
        ip = torch.full((2,), 123.0, dtype=torch.float64, device=dev)
        engine.mu_update_coupled(A1, A2, engine.pack_factor(m1, R, dev), engine.pack_factor(m2, R, dev), G1p, G2p, R, iprod=ip)
        eps = np.finfo(np.float32).eps
        a64 = a.astype(np.float64)
        want = a64 * np.clip((m1.astype(np.float64) + m2) / 2, eps, None) / np.clip((a64 @ g1 + a64 @ g2) / 2, eps, None)
        got = A1[:, :R].cpu().numpy()
        assert torch.equal(A1, A2)
        np.testing.assert_allclose(got, want, rtol=5e-6)
        assert (A1[:, R:] == 0).all()
        np.testing.assert_allclose(ip.cpu().numpy(), [(m1 * want).sum(), (m2 * want).sum()], rtol=1e-5)


@pytest.fixture(params=["library", "host"])
def driver(request, monkeypatch):
    """``library``: the run inside libc2c_b200.so (c2c_coupled_ncp_run, the product path); ``host``: the same sequence driven
    from Python over the single-step entry points (ranks > 32 take it)."""
    if request.param == "host":
        monkeypatch.setenv("C2C_B200_COUPLED_HOST_DRIVER", "1")
    else:
        monkeypatch.delenv("C2C_B200_COUPLED_HOST_DRIVER", raising=False)
    return request.param


@pytest.mark.parametrize("name", sorted(COUPLED_CASES))
def test_coupled_matches_reference_golden(name, driver):
    from cell2cell_b200.tensor.coupled_factorization import coupled_non_negative_parafac
    case = COUPLED_CASES[name]
    x1, x2, m1, m2 = coupled_inputs(name)
    cp1, cp2, (e1, e2) = coupled_non_negative_parafac(x1, x2, case["rank"], case["mode_mapping"], mask1=m1, mask2=m2,
                                                      n_iter_max=case["n_iter_max"], init="random", tol=case["tol"],
                                                      random_state=7, return_errors=True, **case.get("kwargs", {}))
    g1, g2 = GOLD[name + "/errors1"], GOLD[name + "/errors2"]
    assert len(e1) == len(e2)
    if case["tol"] > 0:
        assert abs(len(e1) - len(g1)) <= 2, (len(e1), len(g1))
        n = min(len(e1), len(g1))
        np.testing.assert_allclose(e1[:n - 1], g1[:n - 1], rtol=ERROR_RTOL, atol=1e-5)
        np.testing.assert_allclose(e2[:n - 1], g2[:n - 1], rtol=ERROR_RTOL, atol=1e-5)
        if len(e1) != len(g1):
            return                                             # stopped an iteration apart: factors differ by that step
    else:
        assert len(e1) == len(g1) == case["n_iter_max"]
        np.testing.assert_allclose(e1, g1, rtol=ERROR_RTOL, atol=3e-6)
        np.testing.assert_allclose(e2, g2, rtol=ERROR_RTOL, atol=3e-6)
    assert abs(e1[-1] - g1[-1]) <= ERROR_RTOL * g1[-1] + (3e-6 if m1 is not None else 0)
    assert abs(e2[-1] - g2[-1]) <= ERROR_RTOL * g2[-1] + (3e-6 if m2 is not None else 0)
    for k, cp in ((1, cp1), (2, cp2)):
        for i, f in enumerate(cp.factors):
            ref = GOLD["%s/factors%d_%d" % (name, k, i)]
            assert f.shape == ref.shape and (f >= 0).all()
            assert rel_err(f, ref) < FACTOR_RTOL, (name, k, i, rel_err(f, ref))


def _lowrank_pair(shape1, shape2, shared, seed, true_rank=5):
    rs = np.random.RandomState(seed)
    f1 = [rs.gamma(1.0, 1.0, size=(s, true_rank)) for s in shape1]
    f2 = [rs.gamma(1.0, 1.0, size=(s, true_rank)) for s in shape2]
    for a, b in shared:
        f2[b] = f1[a]
    out = []
    for fs in (f1, f2):
        x = o.cp_to_tensor(None, fs)
        out.append(np.ascontiguousarray(x / x.mean() + 0.05 * rs.random_sample(x.shape), dtype=np.float32))
    return out


@pytest.mark.parametrize("shape1,shape2,shared,rank,n_iter", [
    ((6, 50, 40, 40), (6, 70, 40, 40), [(0, 0), (2, 2), (3, 3)], 10, 20),      # BASELINE config 4 family, streaming FFMA passes
    ((150, 128, 16, 8), (150, 128, 20), [(0, 0), (1, 1)], 16, 15),              # tensor-pipe passes on tensor 1 (P >= 128 * 148)
    ((12, 300, 6, 6), (300, 6, 6), [(1, 0), (2, 1), (3, 2)], 12, 25),           # BALF family + a context-free tensor
])
def test_coupled_matches_oracle_on_streaming_shapes(shape1, shape2, shared, rank, n_iter, driver):
    from cell2cell_b200.tensor.coupled_factorization import coupled_non_negative_parafac
    x1, x2 = _lowrank_pair(shape1, shape2, shared, seed=21)
    mm = {"shared": shared}
    r1, r2, (re1, re2) = co.coupled_non_negative_parafac(x1, x2, rank, mm, n_iter_max=n_iter, init="random", tol=-1.0,
                                                         random_state=3)
    c1, c2, (e1, e2) = coupled_non_negative_parafac(x1, x2, rank, mm, n_iter_max=n_iter, init="random", tol=-1.0,
                                                    random_state=3, return_errors=True)
    np.testing.assert_allclose(e1, re1, rtol=ERROR_RTOL, atol=3e-6)
    np.testing.assert_allclose(e2, re2, rtol=ERROR_RTOL, atol=3e-6)
    assert abs(e1[-1] - re1[-1]) <= ERROR_RTOL * re1[-1] and abs(e2[-1] - re2[-1]) <= ERROR_RTOL * re2[-1]
    for got, ref in ((c1, r1), (c2, r2)):
        for f, g in zip(got.factors, ref[1]):
            assert rel_err(f, g) < FACTOR_RTOL, rel_err(f, g)
    for a, b in shared:
        assert np.array_equal(c1.factors[a], c2.factors[b])
    if driver == "host":
        assert c1.passes_ <= 2 * 2 + 1                           # eager first iteration + the captured one: two reads each


def test_coupled_rank_above_32_takes_the_host_driver():
    from cell2cell_b200.tensor.coupled_factorization import coupled_non_negative_parafac
    shared = [(0, 0), (2, 2), (3, 3)]
    x1, x2 = _lowrank_pair((4, 12, 6, 6), (4, 9, 6, 6), shared, seed=2)
    r1, r2, (re1, re2) = co.coupled_non_negative_parafac(x1, x2, 40, 1, n_iter_max=10, init="random", tol=-1.0, random_state=1)
    c1, c2, (e1, e2) = coupled_non_negative_parafac(x1, x2, 40, 1, n_iter_max=10, init="random", tol=-1.0, random_state=1,
                                                    return_errors=True)
    assert c1.passes_ is not None
    for got, ref in ((c1, r1), (c2, r2)):
        for f, g in zip(got.factors, ref[1]):
            assert rel_err(f, g) < FACTOR_RTOL, rel_err(f, g)
    assert abs(e1[-1] - re1[-1]) <= ERROR_RTOL * re1[-1] and abs(e2[-1] - re2[-1]) <= ERROR_RTOL * re2[-1]


def test_coupled_library_launch_count():
    """Two 4-way tensors sharing three modes: 22 small kernels per replayed iteration (8 MTTKRP tails, 5 updates, 8 Gram
    finalizes, 1 error kernel) + the 4 streaming passes (1-3 launches each, depending on the shape's tail / reduce kernels)."""
    from cell2cell_b200 import engine
    shared = [(0, 0), (2, 2), (3, 3)]
    x1, x2 = _lowrank_pair((6, 40, 8, 8), (6, 30, 8, 8), shared, seed=2)
    dev = torch.device("cuda")
    X1, X2 = torch.as_tensor(x1).to(dev), torch.as_tensor(x2).to(dev)
    counts = []
    for n in (10, 20):
        A1 = [engine.pack_factor(np.random.RandomState(0).rand(s, 4), 4, dev) for s in x1.shape]
        A2 = [engine.pack_factor(np.random.RandomState(0).rand(s, 4), 4, dev) for s in x2.shape]
        l0 = engine.launch_count()
        errs, n_iter, conv = engine.coupled_ncp_run(X1, None, A1, X2, None, A2, 4, shared, n_iter_max=n, tol=-1.0)
        counts.append(engine.launch_count() - l0)
        assert n_iter == n and not conv and errs.shape == (n, 2) and (np.diff(errs.sum(axis=1)) < 1e-6).all()
    per_it = (counts[1] - counts[0]) / 10.0
    assert per_it == int(per_it) and 26 <= per_it <= 34, per_it


def _named(x, prefix, shared_names=None, mask=None):
    from cell2cell_b200.tensor import PreBuiltTensor
    names = [["%s%d_%d" % (prefix, k, i) for i in range(s)] for k, s in enumerate(x.shape)]
    for k, n in (shared_names or {}).items():
        names[k] = n
    return PreBuiltTensor(x if mask is None else x * mask, order_names=names, mask=mask,
                          order_labels=["Contexts", "Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"][-x.ndim:]
                          if x.ndim < 4 else None)


def test_coupled_interaction_tensor_api(tmp_path, capsys):
    from cell2cell_b200.tensor import CoupledInteractionTensor
    shared = [(0, 0), (2, 2), (3, 3)]
    x1, x2 = _lowrank_pair((5, 14, 6, 6), (5, 9, 6, 6), shared, seed=4, true_rank=3)
    ctx = ["c%d" % i for i in range(5)]
    cells = ["k%d" % i for i in range(6)]
    t1 = _named(x1.astype(np.float64), "a", {0: ctx, 2: cells, 3: cells})
    t2 = _named(x2.astype(np.float64), "b", {0: ctx, 2: cells, 3: cells})
    t1.order_labels = t2.order_labels = ["Contexts", "Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"]
    ct = CoupledInteractionTensor(t1, t2, mode_mapping=1, tensor2_name="Metab")
    assert ct.mode_mapping == {"shared": shared} and ct.shape == ((5, 14, 6, 6), (5, 9, 6, 6))
    ct.compute_tensor_factorization(rank=3, init="random", random_state=2, runs=2, n_iter_max=40, tol=-1.0)
    # best of the two runs by the balanced combined error, exactly as the oracle picks it
    best = None
    for run in range(2):
        r1, r2, (re1, re2) = co.coupled_non_negative_parafac(x1, x2, 3, 1, n_iter_max=40, init="random", tol=-1.0,
                                                             random_state=2 + run)
        err = co.combined_error(re1[-1], re2[-1], x1.shape, x2.shape, shared)
        if best is None or err < best[0]:
            best = (err, r1, r2)
    for f, g in zip(ct.tl_object1.factors + ct.tl_object2.factors, best[1][1] + best[2][1]):
        assert rel_err(f, g) < FACTOR_RTOL
    assert "Best coupled model has a combined normalized error of: %.3f" % best[0] in capsys.readouterr().out
    assert list(ct.factors) == ["Contexts", "Sender Cells", "Receiver Cells", "Ligand-Receptor Pairs",
                                "Ligand-Receptor Pairs_Metab"]
    assert ct.factors["Ligand-Receptor Pairs_Metab"].shape == (9, 3) and ct.factors["Ligand-Receptor Pairs"].shape == (14, 3)
    assert ct.factors["Contexts"].equals(ct.factors1["Contexts"])
    assert list(ct.factors["Contexts"].index) == ctx and list(ct.factors["Contexts"].columns) == ["Factor 1", "Factor 2", "Factor 3"]
    # normalised loadings ordered by the mean of the two tensors' weights
    n1, n2 = o.cp_normalize(*best[1]), o.cp_normalize(*best[2])
    avg = (n1[0] + n2[0]) / 2
    order = avg.argsort()[::-1]
    np.testing.assert_allclose(ct.explained_variance_ratio_, avg[order] / avg.sum(), rtol=1e-4)
    np.testing.assert_allclose(ct.factors2["Ligand-Receptor Pairs"].values, n2[1][1][:, order], rtol=2e-4, atol=1e-6)

    def ev(x, cp):
        d = x.astype(np.float64) - o.cp_to_tensor(cp[0], cp[1])
        return 1.0 - np.linalg.norm(d - d.mean()) / np.linalg.norm(x - x.astype(np.float64).mean())
    want = (ev(x1, best[1]) * x1.size + ev(x2, best[2]) * x2.size) / (x1.size + x2.size)
    assert ct.explained_variance_ == pytest.approx(want, rel=1e-4)
    assert ct.get_top_factor_elements("Contexts", "Factor 1", 2, tensor="tensor2").shape == (2,)
    with pytest.raises(ValueError):
        ct.get_top_factor_elements("nope", "Factor 1")
    assert ct.reorder_metadata(list("abcd"), list("wxyz")) == ["a", "c", "d", "b", "x"]
    assert ct.excluded_value_fraction() == {"tensor1": 0.0, "tensor2": 0.0} and ct.missing_fraction("combined") == 0.0
    p = tmp_path / "coupled.pkl"
    ct.write_file(str(p))
    back = pickle.load(open(p, "rb"))
    assert back.tensor1.is_cuda and torch.equal(back.tensor2, ct.tensor2) and list(back.factors) == list(ct.factors)
    # mismatching element names of a shared mode are refused
    t3 = _named(x2.astype(np.float64), "z")
    with pytest.raises(ValueError, match="Element names"):
        CoupledInteractionTensor(t1, t3, mode_mapping=1)


def test_coupled_masked_objects_and_elbow():
    from cell2cell_b200.tensor import CoupledInteractionTensor
    shared = [(1, 0), (2, 1), (3, 2)]
    x1, x2 = _lowrank_pair((4, 8, 5, 5), (8, 5, 5), shared, seed=9, true_rank=2)
    rs = np.random.RandomState(1)
    m1 = (rs.rand(*x1.shape) > 0.1).astype(np.uint8)
    lr = ["lr%d" % i for i in range(8)]
    cells = ["k%d" % i for i in range(5)]
    t1 = _named(x1.astype(np.float64), "a", {1: lr, 2: cells, 3: cells}, mask=m1)
    t2 = _named(x2.astype(np.float64), "b", {0: lr, 1: cells, 2: cells})
    ct = CoupledInteractionTensor(t1, t2, {"shared": shared})
    assert ct.excluded_value_fraction("tensor1") == pytest.approx(1.0 - m1.mean())
    assert ct.missing_fraction()["tensor2"] == 0.0
    fig, loss = ct.elbow_rank_selection(upper_rank=3, runs=1, init="random", random_state=5, n_iter_max=20, tol=-1.0,
                                        automatic_elbow=False, manual_elbow=2, output_fig=False)
    assert ct.rank == 2 and [l[0] for l in loss] == [1, 2, 3]
    for r, got in loss:
        c1, c2, (e1, e2) = co.coupled_non_negative_parafac(x1 * m1, x2, r, {"shared": shared}, mask1=m1, n_iter_max=20,
                                                           init="random", tol=-1.0, random_state=5)
        want = co.combined_error(o.masked_norm_error(x1 * m1, c1, m1.astype(np.float64)), e2[-1], x1.shape, x2.shape, shared)
        assert got == pytest.approx(want, rel=2e-4)
    fig, loss2 = ct.elbow_rank_selection(upper_rank=2, runs=2, init="random", random_state=5, n_iter_max=20, tol=-1.0,
                                         automatic_elbow=False, manual_elbow=1, output_fig=False)
    assert ct.elbow_metric_raw.shape == (2, 2)
    assert ct.elbow_metric_raw[0, 0] == pytest.approx(loss[0][1], rel=1e-6)          # run 0 of rank 1 = the single-run value
    ct.compute_tensor_factorization(rank=2, init="svd", random_state=0, n_iter_max=10)   # a mask forces init='random' (:503)
    assert 0.0 < ct.explained_variance_ <= 1.0
    with pytest.raises(AssertionError):
        ct.elbow_rank_selection(upper_rank=2, runs=2, metric="similarity")

"""The fused context-sharded factorization (c2c_ncp_run_sharded, csrc/c2c_xgpu.cuh) on ONE GPU: with a world of one rank
the exchange kernels run with no peer, so their arithmetic (partial rows rounded to fp32, summed in rank order, the Grams
taken from the communication buffer) is checked against the oracle and against the single-GPU run.  The multi-GPU exchange
itself is covered by tests/test_gpu_multi.py (>= 2 GPUs)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ncp_oracle as o  # noqa: E402

pytestmark = pytest.mark.gpu


def lowrank(shape, rank, seed=0, noise=0.05):
    rs = np.random.RandomState(seed)
    fs = [rs.gamma(1.0, 1.0, size=(s, rank)) for s in shape]
    letters = "abcdefgh"[:len(shape)]
    x = np.einsum(",".join(l + "r" for l in letters) + "->" + letters, *fs)
    x = x / x.mean() + noise * rs.random_sample(shape)
    return np.ascontiguousarray(x, dtype=np.float32)


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b))


@pytest.mark.parametrize("shape,rank", [((9, 120, 16, 16), 6), ((7, 300, 12, 8), 16), ((5, 64, 40, 40), 20),
                                         ((4, 6, 50, 9, 7), 5), ((3, 700, 6, 6), 3)])
def test_world_of_one_matches_oracle_and_single_gpu_run(shape, rank):
    from cell2cell_b200 import engine
    from cell2cell_b200.sharded import sharded_non_negative_parafac
    x = lowrank(shape, 4, seed=1)
    X = torch.as_tensor(x).cuda()
    init = o.random_init(shape, rank, 11)
    n0 = engine.launch_count()
    cp, errs = sharded_non_negative_parafac(X, shape, rank, init, n_iter_max=25, tol=-1.0, fused=True)
    assert engine.launch_count() - n0 > 25 * 5
    ocp, oerr = o.non_negative_parafac(x.astype(np.float64), rank, n_iter_max=25, init=(np.ones(rank), init), tol=-1.0)
    for a, b in zip(cp.factors, ocp.factors):
        assert rel_err(a, b) < 1e-4
    np.testing.assert_allclose(np.asarray(errs) ** 2, np.asarray(oerr) ** 2, rtol=2e-4, atol=5e-7)
    # the single-GPU product run
    fs = [engine.pack_factor(np.asarray(f, dtype=np.float32), rank, X.device) for f in init]
    engine.ncp_run(X, None, fs, rank, n_iter_max=25, tol=-1.0, check_every=0)
    for a, f in zip(cp.factors, fs):
        assert rel_err(a, engine.unpack_factor(f, rank).cpu().numpy().astype(np.float64)) < 2e-5


def test_world_of_one_stops_where_the_oracle_stops():
    from cell2cell_b200.sharded import sharded_non_negative_parafac
    shape, rank = (6, 80, 10, 10), 4
    x = lowrank(shape, 3, seed=2, noise=0.01)
    init = o.random_init(shape, rank, 5)
    cp, errs = sharded_non_negative_parafac(torch.as_tensor(x).cuda(), shape, rank, init, n_iter_max=300, tol=1e-4, fused=True)
    _, oerr = o.non_negative_parafac(x.astype(np.float64), rank, n_iter_max=300, init=(np.ones(rank), init), tol=1e-4)
    assert len(errs) < 300 and abs(len(errs) - len(oerr)) <= 2, (len(errs), len(oerr))


def test_fused_path_refuses_what_it_does_not_cover():
    from cell2cell_b200.sharded import sharded_non_negative_parafac
    x = lowrank((5, 30, 8), 2)
    with pytest.raises(ValueError):
        sharded_non_negative_parafac(torch.as_tensor(x).cuda(), x.shape, 3, o.random_init(x.shape, 3, 1), n_iter_max=5, fused=True)

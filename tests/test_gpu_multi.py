"""Multi-GPU paths on real devices (NCCL, one process per GPU); skipped with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tensor(shape, seed=3):
    rs = np.random.RandomState(seed)
    fs = [rs.gamma(1.0, 1.0, size=(s, 4)) for s in shape]
    x = np.einsum("ar,br,cr,dr->abcd", *fs)
    x = x / x.mean() + 0.05 * rs.random_sample(shape)
    return np.ascontiguousarray(x, dtype=np.float32)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from cell2cell_b200.sharded import shard_bounds, sharded_non_negative_parafac
    from cell2cell_b200.tensor import PreBuiltTensor
    from cell2cell_b200.tensor.cp import random_factors
    shape, R = (9, 120, 16, 16), 6
    x = _tensor(shape)
    lo, hi = shard_bounds(shape[0], world, rank)
    init = random_factors(shape, R, 11)
    Xl = torch.as_tensor(x[lo:hi].copy()).cuda()
    cp, errs = sharded_non_negative_parafac(Xl, shape, R, init, n_iter_max=25, tol=-1.0)             # fused: peer-memory exchanges
    cp_n, errs_n = sharded_non_negative_parafac(Xl, shape, R, init, n_iter_max=25, tol=-1.0, fused=False)   # NCCL all-reduces
    # a second fused run on the same communication buffers (growing epoch), tensor-pipe rank, real stop rule
    init20 = random_factors(shape, 20, 12)
    cp20, errs20 = sharded_non_negative_parafac(Xl, shape, 20, init20, n_iter_max=200, tol=1e-4)
    # the elbow sweep spread over the ranks
    t = PreBuiltTensor(x, order_names=[list(range(s)) for s in shape])
    t.elbow_rank_selection(upper_rank=4, runs=3, init="random", random_state=5, n_iter_max=15, tol=-1.0, output_fig=False,
                           automatic_elbow=False, manual_elbow=2)
    # every rank builds ITS context slab of the interaction tensor and the slabs feed one sharded factorization
    from cell2cell_b200.sharded import build_sharded_interaction_tensor, compute_sharded_tensor_factorization
    from cell2cell_b200.synthetic import synthetic_case
    mats, ppi, kw = synthetic_case("balf", scale=0.05)
    st = build_sharded_interaction_tensor(mats, ppi, **kw)
    compute_sharded_tensor_factorization(st, 4, random_state=888, n_iter_max=30, tol=-1.0)
    slab = (st.context_slab, tuple(st.tensor.shape), [df.values for df in st.factors.values()])
    q.put((rank, [np.asarray(f) for f in cp.factors], errs, np.asarray(t.elbow_metric_raw),
           [np.asarray(f) for f in cp_n.factors], errs_n, [np.asarray(f) for f in cp20.factors], errs20, slab))
    dist.barrier()
    dist.destroy_process_group()


def _run_world(world):
    from cell2cell_b200.tensor import PreBuiltTensor
    from cell2cell_b200.tensor.cp import random_factors
    from oracle import ncp_oracle as o
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    shape, R = (9, 120, 16, 16), 6
    x = _tensor(shape)
    # the oracle (float64 restatement of the reference's tensorly path) from the same initial factors
    init = random_factors(shape, R, 11)
    ref, ref_err = o.non_negative_parafac(x.astype(np.float64), R, n_iter_max=25, init=(np.ones(R), init), tol=-1.0)
    init20 = random_factors(shape, 20, 12)
    ref20, ref20_err = o.non_negative_parafac(x.astype(np.float64), 20, n_iter_max=200, init=(np.ones(20), init20), tol=1e-4)
    t = PreBuiltTensor(x, order_names=[list(range(s)) for s in shape])
    t.elbow_rank_selection(upper_rank=4, runs=3, init="random", random_state=5, n_iter_max=15, tol=-1.0, output_fig=False,
                           automatic_elbow=False, manual_elbow=2)
    from cell2cell_b200.sharded import shard_bounds
    from cell2cell_b200.synthetic import synthetic_case
    from cell2cell_b200.tensor import InteractionTensor
    mats, ppi, kw = synthetic_case("balf", scale=0.05)
    whole = InteractionTensor(mats, ppi, **kw)
    whole.compute_tensor_factorization(rank=4, init="random", random_state=888, n_iter_max=30, tol=-1.0)
    res.sort(key=lambda r: r[0])
    for rank, factors, errs, raw, factors_n, errs_n, factors20, errs20, slab in res:
        assert slab[0] == shard_bounds(len(mats), world, rank) and slab[1][0] == slab[0][1] - slab[0][0]
        for a, b in zip(slab[2], [df.values for df in whole.factors.values()]):
            assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-4
        for got, got_err in ((factors, errs), (factors_n, errs_n)):
            for a, b in zip(got, ref.factors):
                assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-4
            assert abs(got_err[-1] - ref_err[-1]) < 1e-4 * ref_err[-1] + 3e-6
        np.testing.assert_allclose(raw, np.asarray(t.elbow_metric_raw), rtol=1e-6)
        # the fused run's replicated results are bit-identical on every rank, stop iteration included
        assert len(errs20) == len(res[0][7]) and abs(len(errs20) - len(ref20_err)) <= 3, (len(errs20), len(ref20_err))
        assert errs == res[0][2] and errs20 == res[0][7]
        for a, b in zip(factors, res[0][1]):
            assert np.array_equal(a, b)
        rec = np.einsum("ar,br,cr,dr->abcd", *[np.asarray(f, dtype=np.float64) for f in factors20])
        rec_ref = np.einsum("ar,br,cr,dr->abcd", *ref20.factors)
        assert np.linalg.norm(rec - rec_ref) / np.linalg.norm(rec_ref) < 2e-3


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_sharded_and_elbow_on_two_gpus():
    _run_world(2)


@pytest.mark.skipif(_ngpu() < 4, reason="needs >= 4 GPUs")
def test_sharded_and_elbow_on_every_gpu():
    _run_world(min(_ngpu(), 8))

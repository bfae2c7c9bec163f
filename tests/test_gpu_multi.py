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
    cp, errs = sharded_non_negative_parafac(torch.as_tensor(x[lo:hi].copy()).cuda(), shape, R, init, n_iter_max=25, tol=-1.0)
    # the elbow sweep spread over the ranks
    t = PreBuiltTensor(x, order_names=[list(range(s)) for s in shape])
    t.elbow_rank_selection(upper_rank=4, runs=3, init="random", random_state=5, n_iter_max=15, tol=-1.0, output_fig=False,
                           automatic_elbow=False, manual_elbow=2)
    q.put((rank, [np.asarray(f) for f in cp.factors], errs, np.asarray(t.elbow_metric_raw)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_sharded_and_elbow_on_two_gpus():
    from cell2cell_b200.tensor import PreBuiltTensor
    from cell2cell_b200.tensor.factorization import non_negative_parafac
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    shape, R = (9, 120, 16, 16), 6
    x = _tensor(shape)
    ref, ref_err = non_negative_parafac(x, R, n_iter_max=25, init="random", random_state=11, tol=-1.0, return_errors=True)
    t = PreBuiltTensor(x, order_names=[list(range(s)) for s in shape])
    t.elbow_rank_selection(upper_rank=4, runs=3, init="random", random_state=5, n_iter_max=15, tol=-1.0, output_fig=False,
                           automatic_elbow=False, manual_elbow=2)
    for rank, factors, errs, raw in res:
        for a, b in zip(factors, ref.factors):
            assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-4
        assert abs(errs[-1] - ref_err[-1]) < 1e-4 * ref_err[-1] + 3e-6
        np.testing.assert_allclose(raw, np.asarray(t.elbow_metric_raw), rtol=1e-6)

"""Sharded factorization driver (cell2cell_b200/sharded.py): the multi-process logic on the gloo backend, world size 2,
with numpy stand-ins for the kernels; the result must equal the unsharded oracle run.  The GPU version of the same check
(real kernels, NCCL) is tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from cell2cell_b200.sharded import shard_bounds, sharded_non_negative_parafac
from oracle import ncp_oracle as o
from tests.numpy_ops import NumpyOps


def test_shard_bounds_cover():
    for n in (1, 5, 60, 61):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(shape, rank, seed):
    rs = np.random.RandomState(seed)
    fs = [rs.gamma(1.0, 1.0, size=(s, 3)) for s in shape]
    x = o.cp_to_tensor(None, fs) + 0.1 * rs.random_sample(shape)
    init = o.random_init(shape, rank, 7)
    return np.ascontiguousarray(x), init


def _worker(rank, size, port, shape, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    x, init = _case(shape, 3, 0)
    lo, hi = shard_bounds(shape[0], size, rank)
    cp, errs = sharded_non_negative_parafac(torch.as_tensor(x[lo:hi].copy()), shape, 3, init, n_iter_max=15, tol=-1.0,
                                            ops=NumpyOps(3))
    q.put((rank, [np.asarray(f) for f in cp.factors], errs))
    dist.barrier()
    dist.destroy_process_group()


def _run(shape):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shape, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, init = _case(shape, 3, 0)
    ref, ref_err = o.non_negative_parafac(x, 3, n_iter_max=15, init=(np.ones(3), init), tol=-1.0)
    for rank, factors, errs in res:
        np.testing.assert_allclose(errs, ref_err, rtol=1e-9)
        for a, b in zip(factors, ref.factors):
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)


def test_sharded_two_ranks_gloo_4way():
    _run((5, 7, 3, 4))


def test_sharded_two_ranks_gloo_3way_and_5way():
    _run((6, 4, 5))
    _run((3, 2, 4, 3, 2))


def test_sharded_single_process_matches_oracle():
    x, init = _case((4, 6, 3, 3), 3, 0)
    cp, errs = sharded_non_negative_parafac(torch.as_tensor(x), x.shape, 3, init, n_iter_max=10, tol=-1.0, ops=NumpyOps(3))
    ref, ref_err = o.non_negative_parafac(x, 3, n_iter_max=10, init=(np.ones(3), init), tol=-1.0)
    np.testing.assert_allclose(errs, ref_err, rtol=1e-10)

"""Multi-process partitioning of the elbow units (cell2cell_b200/parallel.py) on the gloo backend, world size 2."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from cell2cell_b200.parallel import local_only, lpt_schedule, run_units, world


def test_lpt_schedule_balances_and_is_deterministic():
    costs = [r for r in range(1, 26) for _ in range(10)]
    own = lpt_schedule(costs, 8)
    loads = [sum(c for c, o in zip(costs, own) if o == w) for w in range(8)]
    assert max(loads) - min(loads) <= 25
    assert (own == lpt_schedule(costs, 8)).all()
    assert set(lpt_schedule([1, 1, 1], 1)) == {0}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    units = [(r, run) for r in range(1, 7) for run in range(3)]
    calls = []

    def fn(u):
        calls.append(u)
        return u[0] * 10.0 + u[1]

    vals = run_units(units, fn, cost=lambda u: u[0], gather="float")
    objs = run_units(units, lambda u: {"unit": u}, cost=lambda u: u[0], gather="object")
    # work only ONE rank does must stay out of the collectives (bench.py's rank-0 reports): a world of one inside the guard
    solo = None
    if rank == 0:
        with local_only():
            assert world() == (0, 1)
            solo = run_units([1, 2, 3], lambda u: u * 2.0, gather="object")
    assert world() == (rank, size)
    out.put((rank, vals, len(calls), [o["unit"] for o in objs], solo))
    dist.barrier()
    dist.destroy_process_group()


def test_run_units_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    units = [(r, run) for r in range(1, 7) for run in range(3)]
    expect = [u[0] * 10.0 + u[1] for u in units]
    n_calls = 0
    for rank, vals, calls, objs, solo in res:
        assert solo == ([2.0, 4.0, 6.0] if rank == 0 else None)
        assert np.allclose(vals, expect)
        assert [tuple(o) for o in objs] == units
        n_calls += calls
    assert n_calls == len(units)          # every unit evaluated exactly once across the ranks


def test_run_units_single_process():
    assert run_units([1, 2, 3], lambda u: u * 2.0) == [2.0, 4.0, 6.0]

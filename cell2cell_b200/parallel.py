"""Multi-GPU partitioning of the hot path on one 8 x B200 box: one process per GPU (``torch.distributed``).

The reference has no parallel layer at all (its elbow sweep is a doubly nested serial loop,
cell2cell/tensor/factorization.py:335-352; ``cell2cell/utils/parallel_computing.py`` is unused).  The natural shards are
the independent ``(rank, run)`` factorizations of the elbow sweep / of ``compute_tensor_factorization(runs > 1)``: every
rank holds a replica of the tensor (<= 1.8 GB at the largest config), takes the units a static longest-processing-time
schedule gives it, and the per-unit results (one float, or one small factor set) are gathered on the host side.  There is
no collective on the data path.

``run_units`` is the only entry point; with ``torch.distributed`` not initialised (or world size 1) it is a plain loop.
"""
import numpy as np
import torch

__all__ = ["world", "lpt_schedule", "run_units", "local_only"]

_local_only = 0


class local_only:
    """``with local_only():`` -- inside, this process is a world of one for ``run_units`` / ``world`` even when
    ``torch.distributed`` is initialised: work that only SOME ranks do (a report on rank 0, a per-rank side computation) must
    not enter a collective the other ranks never join."""

    def __enter__(self):
        global _local_only
        _local_only += 1
        return self

    def __exit__(self, *exc):
        global _local_only
        _local_only -= 1
        return False


def world():
    """(rank, world_size) of the initialised default process group, else (0, 1)."""
    import torch.distributed as dist
    if _local_only:
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def lpt_schedule(costs, n_workers):
    """Longest-processing-time-first assignment.  Returns ``owner[i]`` for every unit; deterministic (ties by index), so
    every rank computes the same schedule without talking to the others."""
    costs = np.asarray(costs, dtype=np.float64)
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * n_workers
    owner = np.zeros(len(costs), dtype=np.int64)
    for i in order:
        w = min(range(n_workers), key=lambda k: (load[k], k))
        owner[i] = w
        load[w] += float(costs[i])
    return owner


def _run_local(units, fn, threads):
    """This rank's units, in order.  With ``threads > 1`` they are issued from that many host threads, each on a CUDA stream
    of its own: the device work of the units still runs one kernel after the other (every streaming kernel fills the
    machine), but the host part of one unit (factor init, graph capture, result read-back) overlaps the device part of
    another.  Results do not depend on the interleaving: units share no state."""
    if threads <= 1 or len(units) <= 1 or not torch.cuda.is_available():
        return [fn(u) for u in units]
    from concurrent.futures import ThreadPoolExecutor
    dev = torch.cuda.current_device()
    main = torch.cuda.current_stream(dev)
    start = torch.cuda.Event()
    start.record(main)

    def work(u):
        torch.cuda.set_device(dev)
        st = _thread_stream(dev)
        st.wait_event(start)                        # everything enqueued before the sweep (the tensor upload) is visible
        with torch.cuda.stream(st):
            out = fn(u)
        st.synchronize()
        return out

    with ThreadPoolExecutor(max_workers=threads) as pool:
        return list(pool.map(work, units))


_tls = __import__("threading").local()


def _thread_stream(dev):
    streams = getattr(_tls, "streams", None)
    if streams is None:
        streams = _tls.streams = {}
    if dev not in streams:
        streams[dev] = torch.cuda.Stream(device=dev)
    return streams[dev]


def run_units(units, fn, cost=None, gather="float", threads=1):
    """Evaluate ``fn(unit)`` for every unit, spread over the ranks of the default process group (and, inside a rank, over
    ``threads`` host threads -- see ``_run_local``).

    gather='float': every rank returns the full list of floats (one all-reduce of a ``len(units)`` vector -- results
    only, never tensor data).  gather='object': full list of picklable objects (``all_gather_object``)."""
    rank, size = world()
    n = len(units)
    if size == 1:
        return _run_local(list(units), fn, threads)
    import torch.distributed as dist
    costs = [1.0 if cost is None else float(cost(u)) for u in units]
    owner = lpt_schedule(costs, size)
    mine = [i for i in range(n) if owner[i] == rank]
    local = dict(zip(mine, _run_local([units[i] for i in mine], fn, threads)))
    if gather == "float":
        backend = dist.get_backend()
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        buf = torch.zeros(n, dtype=torch.float64, device=dev)
        for i, v in local.items():
            buf[i] = float(v)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        return buf.cpu().numpy().tolist()
    parts = [None] * size
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged[i] for i in range(n)]

"""One factorization sharded over the GPUs of a box along the first (context) mode  (SURVEY.md section 8e).

Every rank holds a slab ``X[c0:c1]`` of the tensor (contiguous in C order) and the matching rows of the first factor; the
other factors are replicated.  Per NCP iteration (same Gauss-Seidel order and update as the single-GPU path):

    T      = plane contraction of the local slab                         (local pass 1)
    mode 0 : M_0 rows of the local contexts are complete locally -> local update; the full A_0 (C x R, tiny) is re-assembled
             with one all-reduce of a zero-padded buffer so every rank can form the Gram-Hadamard products
    mode k : (other leading modes) M_k = all-reduce of the per-slab partials (I_k x R fp32, <= 240 KB), replicated update
    U      = all-reduce of the per-slab fiber contractions (K*K x R fp32)   (local pass 2)
    trailing modes: replicated updates from the global U; error scalar from the global Grams

i.e. ``N - 1`` small NCCL all-reduces per iteration over NVLink; the tensor itself never moves.  Summation order differs from
the single-GPU run, so results agree to rounding (well inside the 1e-4 parity budget), not bit for bit.

The orchestration is written against a small ``ops`` interface so that the multi-process logic is testable on CPU
(gloo) with numpy stand-ins for the kernels (tests/test_sharded.py); ``DeviceOps`` is the product implementation
(libc2c_b200.so through ``engine``).
"""
import numpy as np
import torch

from . import engine
from .tensor.cp import CPTensor

__all__ = ["shard_bounds", "DeviceOps", "PeerComm", "sharded_non_negative_parafac", "build_sharded_interaction_tensor",
           "compute_sharded_tensor_factorization"]


def shard_bounds(n, world, rank):
    """Contiguous, near-equal split of ``range(n)``: [lo, hi) of ``rank``."""
    return (n * rank) // world, (n * (rank + 1)) // world


class DeviceOps:
    """Kernels of libc2c_b200.so behind the interface the sharded driver needs."""

    def __init__(self, rank_r):
        self.R = int(rank_r)
        self._gws = {}

    def pack(self, f, device):
        return engine.pack_factor(f, self.R, device)

    def unpack(self, a):
        return engine.unpack_factor(a, self.R)

    def sumsq(self, X):
        return engine.sumsq(X)

    def plane(self, X, factors, ws, out=None):
        return engine.plane_contract(X, None, factors, self.R, ws=ws, out=out)

    def fiber(self, X, factors, ws, out=None):
        return engine.fiber_contract(X, None, factors, self.R, ws=ws, out=out)

    def m_from(self, TU, shape, factors, mode):
        return engine.mttkrp_from_contraction(TU, shape, factors, self.R, mode)

    def gram_hadamard(self, factors, shape, skip, changed=None, out=None):
        """``changed``: only that factor changed since the previous call for this shape -- its Gram alone is recomputed."""
        key = (tuple(shape), factors[0].device)
        if key not in self._gws:
            self._gws[key] = engine.Workspace(shape, self.R, False, factors[0].device)
            changed = None
        return engine.gram_hadamard(factors, shape, self.R, skip, ws=self._gws[key], as_tensor=True, changed_mode=changed,
                                    out=out)

    def mu_update(self, A, M, G):
        return engine.mu_update(A, M, G, self.R)

    def workspace(self, shape, device, masked=False):
        return engine.Workspace(shape, self.R, masked, device)

    # ---- used by the coupled factorization (tensor/coupled_factorization.py) ----
    def masked_mttkrp(self, X, mask, factors, mode, ws):
        return engine.mttkrp(X, mask, factors, self.R, mode, ws=ws)

    def mu_update_coupled(self, A1, A2, M1, M2, G1, G2, iprod):
        return engine.mu_update_coupled(A1, A2, M1, M2, G1, G2, self.R, iprod=iprod)

    def recon_sums(self, X, mask, factors, ws):
        return engine.recon_sums(X, mask, factors, self.R, ws=ws)

    def missing_rec_sumsq(self, X, mask, factors, ws, out=None):
        return engine.missing_rec_sumsq(X, mask, factors, self.R, ws=ws, out=out)


class _RawCuda:
    """A device pointer as a ``__cuda_array_interface__`` object (torch wraps it without copying)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerComm:
    """The communication buffers of the fused sharded run (``c2c_ncp_run_sharded``): one zero-filled device buffer per rank
    (``c2c_comm_alloc``), exported as a CUDA IPC handle, exchanged through ``group`` and opened by every peer process with its
    own device current (``c2c_comm_open``: mapped for peer access over NVLink).  The kernels store into the peers' buffers and
    poll flags in their own: no NCCL call on the data path.  ``epoch`` only grows, so the buffers are never cleared between
    runs.  One exchange is run over the mapping when it is made (``c2c_comm_probe``)."""

    _cache = {}

    def __init__(self, nbytes, device, group=None):
        import ctypes
        import torch.distributed as dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.me = dist.get_rank(group) if dist.is_initialized() else 0
        self.nbytes = int(nbytes)
        self.device = device
        lib = engine.lib()
        with torch.cuda.device(device):
            ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
            engine.check(lib.c2c_comm_alloc(self.nbytes, ctypes.byref(ptr), handle), "c2c_comm_alloc")
            self.local = int(ptr.value)
            self.buf = torch.as_tensor(_RawCuda(self.local, self.nbytes), device=device)
            self.ptrs = [self.local]
            self.epoch = 1
            if self.world > 1:
                # A failure on ONE rank must not leave the others waiting in a collective it never joins: every step below ends
                # in an all-reduce of a success flag (which is also the barrier the protocol needs), and all ranks raise together.
                def agree(ok, what):
                    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
                    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                    if int(flag.item()) == 0:
                        raise engine.C2CError("peer-memory setup failed on some rank: " + what)

                objs = [None] * self.world
                dist.all_gather_object(objs, (int(device.index), bytes(handle)), group=group)
                self.ptrs = []
                problem = None
                try:
                    for r, (didx, h) in enumerate(objs):
                        if r == self.me:
                            self.ptrs.append(self.local)
                            continue
                        engine.check(lib.c2c_enable_peer_access(int(didx)), "c2c_enable_peer_access")
                        p = ctypes.c_void_p()
                        engine.check(lib.c2c_comm_open((ctypes.c_ubyte * 64)(*h), ctypes.byref(p)), "c2c_comm_open")
                        self.ptrs.append(int(p.value))
                except Exception as exc:                      # noqa: BLE001 -- reported through the agreement below
                    problem = exc
                # every rank has opened every buffer before anyone writes
                agree(problem is None, "mapping the peers' buffers (CUDA IPC / peer access)%s" % ("" if problem is None else ": %s" % problem))
                out = torch.zeros(1, dtype=torch.float32, device=device)
                arr = (ctypes.c_void_p * self.world)(*self.ptrs)
                rc = lib.c2c_comm_probe(self.world, self.me, arr, self.nbytes, self.epoch, 3.0, ctypes.c_void_p(out.data_ptr()),
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                self.epoch += 2
                got, want = float(out.item()), 3.0 * self.world * (self.world + 1) / 2
                agree(rc == 0 and got == want and not self.failed(), "the probe exchange (%r != %r)" % (got, want))

    @classmethod
    def get(cls, nbytes, device, group=None):
        key = (id(group), str(device))
        c = cls._cache.get(key)
        if c is None or c.nbytes < nbytes:
            c = cls(nbytes, device, group)
            cls._cache[key] = c
        return c

    def failed(self):
        """True if a wait on a peer gave up (the int at offset 0 of the local buffer)."""
        return bool(self.buf[:4].view(torch.int32).item() != 0)


def _fused_sharded(X_local, full_shape, rank, init_factors, n_iter_max, tol, group, cvg_criterion, world, me, lo, hi):
    """The whole loop inside the library (``c2c_ncp_run_sharded``): two streaming passes over the local slab per iteration, the
    cross-slab sums exchanged inside the update kernels through peer memory (csrc/c2c_xgpu.cuh)."""
    import torch.distributed as dist
    dev = X_local.device
    nd = len(full_shape)
    local_shape = (hi - lo,) + tuple(full_shape[1:])
    with torch.cuda.device(dev):
        A = [engine.pack_factor(np.abs(np.asarray(init_factors[0], dtype=np.float64))[lo:hi], rank, dev)]
        A += [engine.pack_factor(np.abs(np.asarray(f, dtype=np.float64)), rank, dev) for f in init_factors[1:]]
        nx2 = torch.tensor([engine.sumsq(X_local)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(nx2, op=dist.ReduceOp.SUM, group=group)
        nbytes = engine.shard_comm_bytes(local_shape, rank, world)
        comm = PeerComm.get(nbytes, dev, group)
        ws = engine.Workspace(local_shape, rank, False, dev)
        if world > 1:
            dist.barrier(group)                     # the ranks' kernels start together (a wait on a peer gives up after ~5 s)
        errs, n_iter, _ = engine.ncp_run_sharded(X_local, A, rank, comm, n_iter_max=n_iter_max, tol=tol,
                                                 cvg_criterion=cvg_criterion, norm_x2=nx2, ws=ws)
        failed = torch.tensor([1 if comm.failed() else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(failed, op=dist.ReduceOp.MAX, group=group)     # every rank raises, or none does
        if int(failed.item()):
            raise engine.C2CError("sharded factorization: a wait on a peer GPU timed out; the results are invalid")
        A0 = torch.zeros((full_shape[0], engine.ldr_of(rank)), dtype=torch.float32, device=dev)
        A0[lo:hi] = A[0]
        if world > 1:
            dist.all_reduce(A0, op=dist.ReduceOp.SUM, group=group)   # assembles the context factor (disjoint rows)
        factors = [engine.unpack_factor(A0, rank).cpu().numpy()] + [engine.unpack_factor(a, rank).cpu().numpy() for a in A[1:]]
    return CPTensor(np.ones(rank, dtype=np.float32), factors, device=dev), [float(e) for e in errs]


def sharded_non_negative_parafac(X_local, full_shape, rank, init_factors, n_iter_max=100, tol=1e-6, group=None, ops=None,
                                 cvg_criterion="abs_rec_error", fused=None):
    """MU NCP of a tensor whose first mode is split over the ranks of ``group``.

    X_local: this rank's slab ``[c1 - c0, I_1, ...]`` (``shard_bounds(full_shape[0], world, rank)``), fp32, on the compute
    device.  init_factors: the FULL initial factors (host arrays, identical on every rank -- e.g. host-seeded random_cp).
    Returns ``(CPTensor with the full factors, errors)`` on every rank.

    ``fused`` (default: whenever it applies -- CUDA, >= 4 modes, rank <= 32, <= 8 ranks, every rank owning a context): the
    loop runs inside the library with the exchanges in the update kernels over peer memory; otherwise (and with ``ops`` given)
    the iteration is driven from here with NCCL / gloo all-reduces between the single-step entry points."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    me = dist.get_rank(group) if dist.is_initialized() else 0
    nd = len(full_shape)
    dev = X_local.device
    can_fuse = (ops is None and dev.type == "cuda" and nd >= 4 and int(rank) <= 32 and world <= 8 and full_shape[0] >= world
                and int(n_iter_max) >= 1)
    if fused is None:
        import os
        fused = can_fuse and os.environ.get("C2C_B200_NO_FUSED_SHARDED", "0") != "1"
    if fused:
        if not can_fuse:
            raise ValueError("the fused sharded run needs a CUDA tensor with >= 4 modes, rank <= 32 and <= 8 ranks")
        lo, hi = shard_bounds(full_shape[0], world, me)
        if X_local.shape[0] != hi - lo:
            raise ValueError("X_local has %d slices of mode 0, expected %d" % (X_local.shape[0], hi - lo))
        return _fused_sharded(X_local, tuple(int(s) for s in full_shape), int(rank), init_factors, int(n_iter_max), float(tol),
                              group, cvg_criterion, world, me, lo, hi)
    ops = ops or DeviceOps(rank)
    lo, hi = shard_bounds(full_shape[0], world, me)
    if X_local.shape[0] != hi - lo:
        raise ValueError("X_local has %d slices of mode 0, expected %d" % (X_local.shape[0], hi - lo))
    local_shape = (hi - lo,) + tuple(int(s) for s in full_shape[1:])
    full_shape = tuple(int(s) for s in full_shape)

    def allreduce(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t

    A_full0 = ops.pack(np.abs(np.asarray(init_factors[0], dtype=np.float64)), dev)         # [C, ldr], replicated
    A = [None] + [ops.pack(np.abs(np.asarray(f, dtype=np.float64)), dev) for f in init_factors[1:]]
    A_loc0 = A_full0[lo:hi].clone().contiguous()
    ws = ops.workspace(local_shape, dev)

    nx2 = torch.tensor([ops.sumsq(X_local)], dtype=torch.float64, device=dev)
    allreduce(nx2)
    norm_x2 = float(nx2.item())
    eps_rule = (lambda d: abs(d) < tol) if cvg_criterion == "abs_rec_error" else (lambda d: d < tol)
    err_dev = torch.zeros(1, dtype=torch.float64, device=dev)            # this iteration's error, written by `iteration`
    errors_dev = torch.zeros(max(int(n_iter_max), 1), dtype=torch.float64, device=dev)

    # Gauss-Seidel changes one factor per update: after each update only that factor's Gram is recomputed and the Hadamard
    # product for the NEXT mode comes out of the same call (2 launches per mode).  G0 (mode 0 of the next iteration) lives in
    # a fixed buffer so that a replayed graph reads what the previous replay wrote.
    G0, _ = ops.gram_hadamard([A_full0] + A[1:], full_shape, 0)

    def iteration():
        loc = [A_loc0] + A[1:]
        full = [A_full0] + A[1:]
        T = ops.plane(X_local, loc, ws)
        # ---- mode 0: rows of the local contexts ----
        M0 = ops.m_from(T, local_shape, loc, 0)
        ops.mu_update(A_loc0, M0, G0)
        A_full0.zero_()
        A_full0[lo:hi] = A_loc0
        allreduce(A_full0)
        G, _ = ops.gram_hadamard(full, full_shape, 1, changed=0)
        # ---- other leading modes: partial M all-reduced, update replicated ----
        for k in range(1, nd - 2):
            Mk = ops.m_from(T, local_shape, loc, k)
            allreduce(Mk)
            ops.mu_update(A[k], Mk, G)
            G, _ = ops.gram_hadamard(full, full_shape, k + 1, changed=k)
        # ---- trailing modes from the all-reduced fiber contraction ----
        U = ops.fiber(X_local, loc, ws)
        allreduce(U)
        M = rn2 = None
        for mode in (nd - 2, nd - 1):
            M = ops.m_from(U, full_shape, full, mode)
            ops.mu_update(A[mode], M, G)
            if mode == nd - 2:
                G, _ = ops.gram_hadamard(full, full_shape, mode + 1, changed=mode)
            else:                                                      # next iteration's mode-0 product + the cp norm
                _, rn2 = ops.gram_hadamard(full, full_shape, 0, changed=mode, out=G0)
        iprod = (ops.unpack(M).double() * ops.unpack(A[nd - 1]).double()).sum()
        err_dev.copy_((torch.sqrt(torch.abs(norm_x2 + rn2 - 2.0 * iprod)) / np.sqrt(norm_x2)).reshape(1))

    # On CUDA the iteration (kernels + the N - 1 NCCL all-reduces) is captured once into a CUDA graph and replayed: no
    # per-iteration host work besides the replay; the error history stays on the device.  With a real stop rule (tol > 0)
    # the history is read back after every iteration so that the run ends on the iteration the reference ends on; with a
    # fixed iteration count (tol <= 0) it is read once at the end.  Off-device (the gloo tests with numpy stand-ins) it
    # runs eagerly.
    graph = None
    check_every = 1 if tol > 0 else int(n_iter_max)
    n_done = 0
    if dev.type == "cuda" and int(n_iter_max) >= 3:
        iteration()                                                    # eager first iteration: lazy initialisations happen here
        errors_dev[0].copy_(err_dev[0])
        n_done = 1
        torch.cuda.synchronize(dev)
        try:
            side = torch.cuda.Stream(device=dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                iteration()
        except Exception:                                              # capture not possible here: stay eager
            graph = None
            torch.cuda.synchronize(dev)
    errors = []
    it = n_done
    stop = False
    if n_done:
        errors = errors_dev[:n_done].cpu().numpy().tolist()
    while it < int(n_iter_max) and not stop:
        if graph is not None:
            graph.replay()
        else:
            iteration()
        errors_dev[it].copy_(err_dev[0])
        it += 1
        if graph is None or it % check_every == 0 or it == int(n_iter_max):
            hist = errors_dev[:it].cpu().numpy().tolist()
            for k in range(max(len(errors), 1), it):
                if eps_rule(hist[k - 1] - hist[k]):
                    hist = hist[:k + 1]
                    stop = True
                    break
            errors = hist
    factors = [ops.unpack(A_full0).cpu().numpy()] + [ops.unpack(a).cpu().numpy() for a in A[1:]]
    return CPTensor(np.ones(rank, dtype=np.float32), factors, device=dev if dev.type == "cuda" else None), errors


def build_sharded_interaction_tensor(rnaseq_matrices, ppi_data, group=None, **kwargs):
    """This rank's context slab of ``InteractionTensor(rnaseq_matrices, ppi_data, **kwargs)`` (tensor
    build row): every rank resolves the same genes / cells / LR pairs from the names of ALL contexts (host side, cheap) and
    uploads and builds only the contexts ``shard_bounds(C, world, rank)`` -- the reference's loop over contexts
    (tensor.py:1126-1129) has no cross-context dependency.  The result feeds ``compute_sharded_tensor_factorization``."""
    import torch.distributed as dist
    from .tensor import InteractionTensor
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    me = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(rnaseq_matrices), world, me)
    return InteractionTensor(rnaseq_matrices, ppi_data, context_slab=(lo, hi), **kwargs)


def compute_sharded_tensor_factorization(interaction_tensor, rank, init="random", random_state=None, n_iter_max=100, tol=10e-7,
                                         normalize_loadings=True, var_ordered_factors=True, group=None):
    """``compute_tensor_factorization`` (tensor.py:216-362, tf_type='non_negative_cp', one run) for a tensor whose context mode
    is spread over the ranks of ``group``: ``interaction_tensor`` holds this rank's slab (``build_sharded_interaction_tensor``).
    Sets ``tl_object``, ``norm_tl_object``, ``factors`` (the FULL factor DataFrames, identical on every rank),
    ``explained_variance_ratio_`` and ``rank`` on it and returns the error history.  Host-seeded random init only (an svd init
    needs the unfoldings of the whole tensor); unmasked tensors (how='inner')."""
    from collections import OrderedDict
    import pandas as pd
    from .tensor.cp import random_factors
    from .tensor.tensor import _cp_normalize
    t = interaction_tensor
    if t.mask is not None:
        raise NotImplementedError("the context-sharded factorization takes unmasked tensors (how='inner')")
    if init != "random":
        raise NotImplementedError("the context-sharded factorization takes init='random' (host-seeded, identical on every rank)")
    full_c = int(getattr(t, "full_contexts", t.tensor.shape[0]))
    full_shape = (full_c,) + tuple(int(s) for s in t.tensor.shape[1:])
    init_factors = random_factors(full_shape, rank, random_state)
    cp, errors = sharded_non_negative_parafac(t.tensor, full_shape, rank, init_factors, n_iter_max=n_iter_max, tol=tol, group=group)
    t.tl_object = cp
    names = [list(getattr(t, "full_context_names", t.order_names[0]))] + [list(n) for n in t.order_names[1:]]
    labels = t.order_labels or ["Contexts", "Ligand-Receptor Pairs", "Sender Cells", "Receiver Cells"]
    if normalize_loadings:
        t.norm_tl_object = _cp_normalize(cp)
        weights, factors = t.norm_tl_object
        weights = np.asarray(weights)
        order = weights.argsort()[::-1] if var_ordered_factors else np.arange(rank)
        factors = [np.asarray(f)[:, order] for f in factors]
        t.explained_variance_ratio_ = weights[order] / sum(weights)
    else:
        factors = [np.asarray(f) for f in cp.factors]
        t.explained_variance_ratio_ = None
    cols = ["Factor {}".format(i) for i in range(1, rank + 1)]
    t.factors = OrderedDict(zip(labels, [pd.DataFrame(f, index=idx, columns=cols) for f, idx in zip(factors, names)]))
    t.rank = rank
    return errors

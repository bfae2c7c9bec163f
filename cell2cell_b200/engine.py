"""Device-side driver: torch owns device memory and streams, ``libc2c_b200.so`` does the work.

Every function here takes CUDA tensors (or uploads numpy arrays) and calls the C ABI of ``include/c2c_b200.h`` through
``_lib``.  Factor matrices travel as ``[I_n, ldr]`` fp32 tensors with ``ldr = rank`` rounded up to a multiple of 4 and zero
padding columns (``pack_factor`` / ``unpack_factor``).
"""
import ctypes

import numpy as np
import torch

from ._lib import AGG, ALGO, CVG, MAX_NDIM, MAX_RANK, SCORE, C2CError, check, lib

__all__ = ["require_cuda", "ldr_of", "pack_factor", "unpack_factor", "Workspace", "sumsq", "mttkrp", "gram_hadamard",
           "plane_contract", "fiber_contract", "mttkrp_from_contraction", "launch_count", "set_mma_min_rank",
           "mu_update", "mu_update_coupled", "hals_update", "recon_sums", "cp_normalize", "ncp_run", "ncp_run_batched", "coupled_ncp_run", "build_ccc", "group_aggregate",
           "missing_rec_sumsq", "MaskPlan", "mask_plan", "masked_contract_planned", "ncp_run_multi", "multi_run_cost",
           "shard_comm_bytes", "ncp_run_sharded"]

_vp = ctypes.c_void_p
BUILD_EVENTS = None       # set to a list to collect (start, end) CUDA events of every c2c_build_ccc launch (bench.py)


def require_cuda(device=None):
    """Resolve ``device`` to a CUDA ``torch.device`` or raise: there is no CPU path."""
    if not torch.cuda.is_available():
        raise C2CError("no CUDA device is visible: cell2cell_b200 runs on B200 (sm_100a) only and has no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device)
    if dev.type != "cuda":
        raise C2CError("device %r is not a CUDA device: cell2cell_b200 has no CPU fallback" % (device,))
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def ldr_of(rank):
    return (int(rank) + 3) // 4 * 4


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return _vp(0) if t is None else _vp(t.data_ptr())


def _shape_arr(shape):
    return (ctypes.c_int64 * len(shape))(*[int(s) for s in shape])


def _ptr_arr(tensors):
    return (_vp * len(tensors))(*[t.data_ptr() for t in tensors])


def pack_factor(f, rank, device):
    """numpy / torch ``[I, rank]`` -> contiguous fp32 ``[I, ldr]`` on ``device`` with zero padding columns."""
    t = torch.as_tensor(np.ascontiguousarray(f) if isinstance(f, np.ndarray) else f)
    out = torch.zeros((t.shape[0], ldr_of(rank)), dtype=torch.float32, device=device)
    out[:, :rank] = t.to(device=device, dtype=torch.float32)
    return out


def unpack_factor(t, rank):
    return t[:, :rank]


def _check_tensor(X, mask):
    if not (isinstance(X, torch.Tensor) and X.is_cuda and X.dtype == torch.float32 and X.is_contiguous()):
        raise C2CError("tensor must be a contiguous fp32 CUDA tensor")
    if X.data_ptr() % 16 != 0:
        raise C2CError("tensor storage must be 16-byte aligned (a view at an odd storage offset is not: use .clone())")
    if not 3 <= X.dim() <= MAX_NDIM:
        raise C2CError("tensor must have between 3 and %d dimensions, got %d" % (MAX_NDIM, X.dim()))
    if mask is not None:
        if not (isinstance(mask, torch.Tensor) and mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous()
                and mask.shape == X.shape and mask.device == X.device):
            raise C2CError("mask must be a contiguous uint8 CUDA tensor of the tensor's shape on the same device")
        if mask.data_ptr() % 4 != 0:
            raise C2CError("mask storage must be 4-byte aligned (use .clone())")


class Workspace:
    """Scratch buffer for the factorization entry points, sized by ``c2c_ncp_workspace_bytes``."""

    def __init__(self, shape, rank, masked, device):
        self.shape = tuple(int(s) for s in shape)
        self.rank = int(rank)
        self.masked = bool(masked)
        nbytes = lib().c2c_ncp_workspace_bytes(len(self.shape), _shape_arr(self.shape), self.rank, int(self.masked))
        if nbytes == 0:
            check(-1, "c2c_ncp_workspace_bytes")
        self.nbytes = int(nbytes)
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)

    @property
    def ptr(self):
        return _vp(self.buf.data_ptr())

    def fits(self, shape, rank, masked):
        return self.shape == tuple(int(s) for s in shape) and self.rank == int(rank) and (self.masked or not masked)


def _ws_for(X, rank, mask, ws):
    if ws is not None and ws.fits(X.shape, rank, mask is not None) and ws.buf.device == X.device:
        return ws
    return Workspace(X.shape, rank, mask is not None, X.device)


def sumsq(X, ws=None):
    """sum(X**2) with fp64 accumulation (tensorly ``tl.norm(tensor, 2) ** 2``)."""
    if not (isinstance(X, torch.Tensor) and X.is_cuda and X.dtype == torch.float32 and X.is_contiguous() and X.data_ptr() % 16 == 0):
        raise C2CError("tensor must be a contiguous, 16-byte aligned fp32 CUDA tensor")
    with torch.cuda.device(X.device):
        out = torch.zeros(1, dtype=torch.float64, device=X.device)
        scratch = torch.empty(1184 * 8, dtype=torch.uint8, device=X.device)
        check(lib().c2c_sumsq(_ptr(X), X.numel(), _ptr(out), _ptr(scratch), scratch.numel(), _stream()), "c2c_sumsq")
        return float(out.item())


def mttkrp(X, mask, factors, rank, mode, ws=None):
    """Fused (masked) MTTKRP of one mode; returns ``M`` as ``[I_mode, ldr]``."""
    _check_tensor(X, mask)
    with torch.cuda.device(X.device):
        ws = _ws_for(X, rank, mask, ws)
        M = torch.empty((X.shape[mode], ldr_of(rank)), dtype=torch.float32, device=X.device)
        check(lib().c2c_mttkrp(_ptr(X), _ptr(mask), X.dim(), _shape_arr(X.shape), _ptr_arr(factors), ldr_of(rank), rank,
                               mode, _ptr(M), ws.ptr, ws.nbytes, _stream()), "c2c_mttkrp")
        return M


def plane_contract(X, mask, factors, rank, ws=None, out=None):
    """T[P, ldr]: contraction of every plane with the two trailing factors (one read of X)."""
    _check_tensor(X, mask)
    with torch.cuda.device(X.device):
        ws = _ws_for(X, rank, mask, ws)
        P = int(np.prod(X.shape[:-2]))
        T = out if out is not None else torch.empty((P, ldr_of(rank)), dtype=torch.float32, device=X.device)
        check(lib().c2c_plane_contract(_ptr(X), _ptr(mask), X.dim(), _shape_arr(X.shape), _ptr_arr(factors), ldr_of(rank),
                                       rank, _ptr(T), ws.ptr, ws.nbytes, _stream()), "c2c_plane_contract")
        return T


def fiber_contract(X, mask, factors, rank, ws=None, out=None):
    """U[I_{N-2} * I_{N-1}, ldr]: contraction of every in-plane position with the leading Khatri-Rao rows (one read of X)."""
    _check_tensor(X, mask)
    with torch.cuda.device(X.device):
        ws = _ws_for(X, rank, mask, ws)
        E = int(X.shape[-2] * X.shape[-1])
        U = out if out is not None else torch.empty((E, ldr_of(rank)), dtype=torch.float32, device=X.device)
        check(lib().c2c_fiber_contract(_ptr(X), _ptr(mask), X.dim(), _shape_arr(X.shape), _ptr_arr(factors), ldr_of(rank),
                                       rank, _ptr(U), ws.ptr, ws.nbytes, _stream()), "c2c_fiber_contract")
        return U


def time_pass_kernel(X, factors, rank, which, reps=20, ws=None):
    """(milliseconds, kernel name) of the streaming kernel of one unmasked pass ALONE (``c2c_time_pass_kernel``): ``which`` 0 =
    plane pass, 1 = fiber pass; average over ``reps`` back-to-back launches, CUDA events on the launching stream."""
    _check_tensor(X, None)
    with torch.cuda.device(X.device):
        ws = _ws_for(X, rank, None, ws)
        ms, kind = ctypes.c_float(0.0), ctypes.c_int(0)
        check(lib().c2c_time_pass_kernel(_ptr(X), X.dim(), _shape_arr(X.shape), _ptr_arr(factors), ldr_of(rank), int(rank), int(which),
                                         int(reps), ctypes.byref(ms), ctypes.byref(kind), ws.ptr, ws.nbytes, _stream()),
              "c2c_time_pass_kernel")
        name = ("plane_" if which == 0 else "fiber_") + ("mma" if kind.value else "stream")
        return float(ms.value), name


def mttkrp_from_contraction(TU, shape, factors, rank, mode):
    """M of ``mode`` from T (leading modes) or U (trailing modes) -- the L2-resident half of the MTTKRP."""
    with torch.cuda.device(TU.device):
        M = torch.empty((int(shape[mode]), ldr_of(rank)), dtype=torch.float32, device=TU.device)
        check(lib().c2c_mttkrp_from_contraction(_ptr(TU), len(shape), _shape_arr(shape), _ptr_arr(factors), ldr_of(rank), rank,
                                                mode, _ptr(M), _stream()), "c2c_mttkrp_from_contraction")
        return M


def set_mma_min_rank(rank):
    """Smallest rank that takes the tensor-pipe (tcgen05) streaming passes; returns the previous value."""
    return int(lib().c2c_set_mma_min_rank(int(rank)))


def launch_count():
    """Kernels enqueued by libc2c_b200.so so far in this process."""
    return int(lib().c2c_launch_count())


def gram_hadamard(factors, shape, rank, skip_mode=-1, ws=None, as_tensor=False, changed_mode=None, out=None):
    """(G, cp_norm^2): Hadamard product of the factor Grams, skipping ``skip_mode`` (-1: none).  ``as_tensor`` returns the
    norm as a 1-element fp64 CUDA tensor without synchronising.  ``changed_mode``: only that factor changed since the last
    call made with this ``ws`` -- its Gram alone is recomputed (``c2c_gram_hadamard_update``).  ``out``: the ``[ldr, ldr]``
    buffer G is written to."""
    dev = factors[0].device
    with torch.cuda.device(dev):
        if ws is None:
            if changed_mode is not None:
                raise C2CError("gram_hadamard(changed_mode=...) needs the workspace of the call that computed the other Grams")
            ws = Workspace(shape, rank, False, dev)
        ldr = ldr_of(rank)
        G = out if out is not None else torch.empty((ldr, ldr), dtype=torch.float32, device=dev)
        n2 = torch.zeros(1, dtype=torch.float64, device=dev)
        if changed_mode is None:
            check(lib().c2c_gram_hadamard(_ptr_arr(factors), len(factors), _shape_arr(shape), ldr, rank, skip_mode, _ptr(G),
                                          _ptr(n2), ws.ptr, ws.nbytes, _stream()), "c2c_gram_hadamard")
        else:
            check(lib().c2c_gram_hadamard_update(_ptr_arr(factors), len(factors), _shape_arr(shape), ldr, rank,
                                                 int(changed_mode), skip_mode, _ptr(G), _ptr(n2), ws.ptr, ws.nbytes,
                                                 _stream()), "c2c_gram_hadamard_update")
        return (G, n2) if as_tensor else (G, float(n2.item()))


def mu_update(A, M, G, rank, eps=float(np.finfo(np.float32).eps)):
    with torch.cuda.device(A.device):
        check(lib().c2c_mu_update(_ptr(A), _ptr(M), _ptr(G), A.shape[0], rank, ldr_of(rank), eps, _vp(0), 0, _stream()),
              "c2c_mu_update")
    return A


def mu_update_coupled(A1, A2, M1, M2, G1, G2, rank, iprod=None, eps=float(np.finfo(np.float32).eps)):
    """Shared-mode update of the coupled factorization: averaged numerators / denominators of the two tensors, the new factor
    written to both copies ``A1`` / ``A2``; ``iprod`` (fp64 ``[2]`` CUDA tensor or None) receives ``sum(M_k * new)``."""
    with torch.cuda.device(A1.device):
        check(lib().c2c_mu_update_coupled(_ptr(A1), _ptr(A2), _ptr(M1), _ptr(M2), _ptr(G1), _ptr(G2), A1.shape[0], rank,
                                          ldr_of(rank), eps, _ptr(iprod), _stream()), "c2c_mu_update_coupled")
    return A1


def hals_update(A, M, G, rank, max_sweeps=100, tol=1e-8):
    with torch.cuda.device(A.device):
        check(lib().c2c_hals_update(_ptr(A), _ptr(M), _ptr(G), A.shape[0], rank, ldr_of(rank), max_sweeps, tol, _stream()),
              "c2c_hals_update")
    return A


def recon_sums(X, mask, factors, rank, ws=None):
    """fp64 ``[sum m(x-rec), sum m(x-rec)^2, sum m x, sum m x^2, sum m]`` in one pass (m = mask, all ones if None)."""
    _check_tensor(X, mask)
    with torch.cuda.device(X.device):
        ws = _ws_for(X, rank, mask, ws)
        out = torch.zeros(5, dtype=torch.float64, device=X.device)
        check(lib().c2c_recon_sums(_ptr(X), _ptr(mask), X.dim(), _shape_arr(X.shape), _ptr_arr(factors), ldr_of(rank), rank,
                                   _ptr(out), ws.ptr, ws.nbytes, _stream()), "c2c_recon_sums")
        return out.cpu().numpy()


def missing_rec_sumsq(X, mask, factors, rank, ws=None, out=None):
    """``||(1 - mask) * reconstruction||^2`` as a 1-element fp64 CUDA tensor (no synchronisation): the squared norm of the
    entries tensorly's imputation writes -- see ``c2c_missing_rec_sumsq``."""
    _check_tensor(X, mask)
    with torch.cuda.device(X.device):
        if ws is None or not ws.fits(X.shape, rank, True):
            ws = Workspace(X.shape, rank, True, X.device)
        out = torch.zeros(1, dtype=torch.float64, device=X.device) if out is None else out
        check(lib().c2c_missing_rec_sumsq(_ptr(X), _ptr(mask), X.dim(), _shape_arr(X.shape), _ptr_arr(factors), ldr_of(rank), rank,
                                          _ptr(out), ws.ptr, ws.nbytes, _stream()), "c2c_missing_rec_sumsq")
        return out


def cp_normalize(factors, rank, weights=None):
    """In-place ``cp_normalize`` of packed factors; returns the weights (fp32 ``[rank]`` CUDA tensor)."""
    dev = factors[0].device
    with torch.cuda.device(dev):
        w = torch.ones(rank, dtype=torch.float32, device=dev) if weights is None else weights.to(dev, torch.float32).clone()
        shape = [f.shape[0] for f in factors]
        check(lib().c2c_cp_normalize(_ptr_arr(factors), len(factors), _shape_arr(shape), ldr_of(rank), rank, _ptr(w),
                                     _stream()), "c2c_cp_normalize")
        return w


ADMM_CONSTRAINTS = {None: 0, "non_negative": 1, "l1_reg": 2, "l2_square_reg": 3}


def admm_update(A, X64, dual, aux, M, G, rank, constraint=None, param=0.0, n_inner=10, tol=1e-6, out=None):
    """One mode's ADMM update of ``constrained_parafac`` in place (``c2c_admm_update``): ``A`` packed fp32 ``[I, ldr]``,
    ``X64`` / ``dual`` / ``aux`` fp64 ``[I, rank]``, ``M`` packed MTTKRP, ``G`` the ``[ldr, ldr]`` Hadamard-Gram.  Returns the
    device ``double[2]`` = (this mode's constraint-error term, inner iterations made)."""
    dev = A.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty(2, dtype=torch.float64, device=dev)
        check(lib().c2c_admm_update(_ptr(A), _ptr(X64), _ptr(dual), _ptr(aux), _ptr(M), _ptr(G), int(A.shape[0]), int(rank),
                                    ldr_of(rank), ADMM_CONSTRAINTS[constraint], float(param), int(n_inner), float(tol), _ptr(out),
                                    _stream()), "c2c_admm_update")
        return out


_CORR_METHODS = {"stacked": 0, "max_score": 1, "min_score": 2, "avg_score": 3}


def corrindex_pairs(stacked, mode_sizes, method="stacked", tol=5e-16, device=None):
    """CorrIndex of every pair of ``n`` decompositions in one launch (``c2c_corrindex_pairs``; metrics.py:131-179).

    stacked: ``[n, S, R]`` float64 (numpy or torch), every decomposition's factor matrices stacked along the rows;
    mode_sizes: the row counts of the stacked matrices (sum = S).  Returns the symmetric ``[n, n]`` float64 numpy matrix."""
    dev = require_cuda(device)
    with torch.cuda.device(dev):
        F = torch.as_tensor(stacked).to(device=dev, dtype=torch.float64).contiguous()
        n, S, R = (int(v) for v in F.shape)
        sizes = [int(v) for v in mode_sizes]
        if sum(sizes) != S:
            raise ValueError("mode sizes do not add up to the stacked row count")
        offs = [0, S] if method == "stacked" else list(np.concatenate([[0], np.cumsum(sizes)]))
        seg = torch.as_tensor(np.asarray(offs, dtype=np.int64)).to(dev)
        out = torch.empty((n, n), dtype=torch.float64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        check(lib().c2c_corrindex_pairs(_ptr(F), n, S, R, _ptr(seg), len(offs) - 1, _CORR_METHODS[method], float(tol), _ptr(out),
                                        _ptr(status), _stream()), "c2c_corrindex_pairs")
        res = out.cpu().numpy()
        if int(status.item()) != 0:
            raise ValueError("Column norms must be non-zero")
        return res


class MaskPlan:
    """The mask plan of ``c2c_mask_plan_scan`` / ``_fill`` (include/c2c_b200.h): the tightest separable model
    ``observed[p, s, v] = g[p] * hs[c(p), s] * hv[c(p), v]`` of the mask plus the residual missing entries as lists, built once
    per (tensor, mask) and used by every masked factorization of that tensor.  ``usable`` is False when the tensor holds
    non-zero values under zero mask bytes (the dense + correction form needs zeros there) or the lists would overflow."""

    def __init__(self, X, mask):
        _check_tensor(X, mask)
        shape = tuple(int(s) for s in X.shape)
        self.shape = shape
        with torch.cuda.device(X.device):
            nb = int(lib().c2c_mask_plan_header_bytes(len(shape), _shape_arr(shape)))
            self.usable = nb > 0
            self.header = self.lists = None
            self.info = (ctypes.c_int64 * 8)()
            if not self.usable:
                return
            self.header = torch.empty(nb, dtype=torch.uint8, device=X.device)
            check(lib().c2c_mask_plan_scan(_ptr(X), _ptr(mask), len(shape), _shape_arr(shape), _ptr(self.header), nb, self.info,
                                           _stream()), "c2c_mask_plan_scan")
            self.usable = self.info[2] == 0 and self.info[5] == 0
            if self.usable and self.info[1] > 0:
                self.lists = torch.empty(int(self.info[6]), dtype=torch.uint8, device=X.device)
                check(lib().c2c_mask_plan_fill(_ptr(mask), len(shape), _shape_arr(shape), _ptr(self.header), nb, self.info,
                                               _ptr(self.lists), self.lists.numel(), _stream()), "c2c_mask_plan_fill")

    n_missing = property(lambda self: int(self.info[0]))
    n_residual = property(lambda self: int(self.info[7]))        # entries the separable model cannot express
    n_list_slots = property(lambda self: int(self.info[1]))      # slots of the by-plane list (class padding included)
    model_missing = property(lambda self: bool(self.info[4]))

    def args(self):
        """(header, header_bytes, lists, lists_bytes, info) as the planned entry points take them."""
        return (_ptr(self.header), self.header.numel(), _ptr(self.lists), 0 if self.lists is None else self.lists.numel(),
                self.info)


def mask_plan(X, mask):
    """The plan of ``mask`` (cached on the mask tensor; rebuilt if the mask was written to since)."""
    cached = getattr(mask, "_c2c_plan", None)
    key = (X.data_ptr(), mask.data_ptr(), mask._version, X._version, tuple(X.shape))
    if cached is not None and cached[0] == key:
        return cached[1]
    plan = MaskPlan(X, mask)
    mask._c2c_plan = (key, plan)
    return plan


def masked_contract_planned(X, mask, factors, rank, which, ws=None, plan=None):
    """T (``which=0``, ``[P, ldr]``) or U (``which=1``, ``[E, ldr]``) of the imputed tensor -- observed entries of X, the
    reconstruction of ``factors`` at the missing ones -- from the unmasked pass plus the mask plan's corrections."""
    _check_tensor(X, mask)
    with torch.cuda.device(X.device):
        plan = plan or mask_plan(X, mask)
        if ws is None or not ws.fits(X.shape, rank, True):
            ws = Workspace(X.shape, rank, True, X.device)
        n = int(np.prod(X.shape[:-2])) if which == 0 else int(X.shape[-2] * X.shape[-1])
        out = torch.empty((n, ldr_of(rank)), dtype=torch.float32, device=X.device)
        check(lib().c2c_masked_contract_planned(_ptr(X), X.dim(), _shape_arr(X.shape), _ptr_arr(factors), ldr_of(rank), rank,
                                                int(which), _ptr(out), ws.ptr, ws.nbytes, *plan.args(), _stream()),
              "c2c_masked_contract_planned")
        return out


def _use_mask_plan():
    import os
    return os.environ.get("C2C_B200_NO_MASK_PLAN", "0") != "1"


def ncp_run(X, mask, factors, rank, tf_type="non_negative_cp", n_iter_max=100, tol=1e-6, cvg_criterion="abs_rec_error",
            check_every=10, norm_x2=None, ws=None, sync=True):
    """Run tensorly-semantics NCP on the device, updating the packed ``factors`` in place.

    Returns ``(errors, n_iter, converged)``; ``errors`` is a float64 numpy array with one entry per iteration done.
    With ``sync=False`` returns the raw device buffers ``(errors_dev, state_dev)`` without waiting."""
    _check_tensor(X, mask)
    if rank < 1 or rank > MAX_RANK:
        raise C2CError("rank must be in [1, %d], got %r" % (MAX_RANK, rank))
    if tf_type not in ALGO:
        raise C2CError("tf_type %r is not one of %s" % (tf_type, sorted(ALGO)))
    if cvg_criterion not in CVG:
        raise C2CError("cvg_criterion %r is not one of %s" % (cvg_criterion, sorted(CVG)))
    if len(factors) != X.dim():
        raise C2CError("need one factor per mode")
    for f, s in zip(factors, X.shape):
        if f.shape != (s, ldr_of(rank)) or f.dtype != torch.float32 or not f.is_contiguous() or f.device != X.device:
            raise C2CError("factors must be packed fp32 [I_n, ldr] tensors on the tensor's device (pack_factor)")
    if tol < 0 and cvg_criterion == "abs_rec_error":
        check_every = 0             # |err[-2] - err[-1]| < tol never holds: nothing to poll, the run is only enqueued
    with torch.cuda.device(X.device):
        ws = _ws_for(X, rank, mask, ws)
        errors = torch.zeros(max(int(n_iter_max), 1), dtype=torch.float64, device=X.device)
        state = torch.zeros(2, dtype=torch.int32, device=X.device)
        nx2 = None
        if norm_x2 is not None:
            nx2 = torch.tensor([float(norm_x2)], dtype=torch.float64, device=X.device)
        done = False
        if mask is not None and tf_type == "non_negative_cp" and rank <= 32 and _use_mask_plan():
            # masked multiplicative updates: the mask plan (built once per tensor) replaces every pass over the mask
            plan = mask_plan(X, mask)
            if plan.usable:
                try:
                    check(lib().c2c_ncp_run_planned(_ptr(X), _ptr(mask), X.dim(), _shape_arr(X.shape), _ptr_arr(factors),
                                                    ldr_of(rank), rank, int(n_iter_max), float(tol), CVG[cvg_criterion],
                                                    _ptr(nx2), _ptr(errors), _ptr(state), int(check_every), ws.ptr, ws.nbytes,
                                                    *plan.args(), _stream()), "c2c_ncp_run_planned")
                    done = True
                except C2CError as err:
                    if getattr(err, "code", 0) != -50:      # -50: the plan path does not take this shape / rank
                        raise
        if not done:
            check(lib().c2c_ncp_run(_ptr(X), _ptr(mask), X.dim(), _shape_arr(X.shape), _ptr_arr(factors), ldr_of(rank), rank,
                                    ALGO[tf_type], int(n_iter_max), float(tol), CVG[cvg_criterion], _ptr(nx2), _ptr(errors),
                                    _ptr(state), int(check_every), ws.ptr, ws.nbytes, _stream()), "c2c_ncp_run")
        if not sync:
            return errors, state
        st = state.cpu().numpy()
        n_iter = int(st[0])
        return errors[:n_iter].cpu().numpy(), n_iter, bool(st[1])


def ncp_run_batched(X, factors, run_ranks, n_iter_max=100, tol=1e-6, cvg_criterion="abs_rec_error", check_every=10,
                    norm_x2=None, ws=None):
    """Independent MU factorizations side by side (``c2c_ncp_run_batched``): run ``b`` has rank ``run_ranks[b]`` and owns the
    next ``run_ranks[b]`` columns of the packed ``[I_n, ldr(sum(run_ranks))]`` factors.  Returns
    ``(errors [nruns, n_iter_max] float64, iterations [nruns] int)``."""
    _check_tensor(X, None)
    run_ranks = [int(r) for r in run_ranks]
    nruns, rank = len(run_ranks), int(sum(run_ranks))
    if nruns < 1 or min(run_ranks) < 1 or rank > 32 or nruns > 32:
        raise C2CError("batched runs need ranks >= 1 adding up to <= 32, got %r" % (run_ranks,))
    if cvg_criterion not in CVG:
        raise C2CError("cvg_criterion %r is not one of %s" % (cvg_criterion, sorted(CVG)))
    for f, s in zip(factors, X.shape):
        if f.shape != (s, ldr_of(rank)) or f.dtype != torch.float32 or not f.is_contiguous() or f.device != X.device:
            raise C2CError("factors must be packed fp32 [I_n, ldr] tensors on the tensor's device (pack_factor)")
    if tol < 0 and cvg_criterion == "abs_rec_error":
        check_every = 0
    with torch.cuda.device(X.device):
        ws = _ws_for(X, rank, None, ws)
        n_it = max(int(n_iter_max), 1)
        errors = torch.zeros((nruns, n_it), dtype=torch.float64, device=X.device)
        state = torch.zeros(2 + 2 * nruns, dtype=torch.int32, device=X.device)
        nx2 = None
        if norm_x2 is not None:
            nx2 = torch.tensor([float(norm_x2)], dtype=torch.float64, device=X.device)
        ranks_arr = (ctypes.c_int * nruns)(*run_ranks)
        check(lib().c2c_ncp_run_batched(_ptr(X), X.dim(), _shape_arr(X.shape), _ptr_arr(factors), ldr_of(rank), ranks_arr,
                                        nruns, int(n_iter_max), float(tol), CVG[cvg_criterion], _ptr(nx2), _ptr(errors),
                                        _ptr(state), int(check_every), ws.ptr, ws.nbytes, _stream()), "c2c_ncp_run_batched")
        st = state.cpu().numpy()
        if nruns == 1:
            iters = np.array([int(st[0])])
        else:
            iters = st[2 + nruns:2 + 2 * nruns].astype(np.int64)
        return errors.cpu().numpy(), iters


def multi_run_cost(shape, ranks, n_iter_max=100, n_sm=148):
    """(seconds as a grid of one-CTA runs, seconds as whole-GPU factorizations of bins of runs): a coarse model used to choose
    between ``ncp_run_multi`` and ``ncp_run`` / ``ncp_run_batched``, calibrated on B200 (gpurun_out r02c: 250 runs x 100
    iterations, ranks 1-25: BALF 0.138 s, ASD 2.1 s as a grid of CTAs) -- a CTA does 2 N_X ldr FMAs per iteration at ~8 % of an
    SM's FMA rate with 2 CTAs per SM; a whole-GPU iteration of a bin of <= 28 factor columns costs two HBM-speed passes but at
    least ~110 us of launch-bound kernels."""
    n_x = float(np.prod([int(s) for s in shape]))
    per_cta = [2.0 * n_x * ldr_of(r) / (128 * 1.9e9 * 0.08) for r in ranks]
    multi = max(n_iter_max * max(per_cta), n_iter_max * sum(per_cta) / n_sm)
    bins = max(1, -(-int(sum(ranks)) // 28)) if len(ranks) > 1 else 1
    whole = bins * n_iter_max * max(110e-6, 1.45 * 2.0 * n_x * 4 / 5.5e12 + 30e-6)
    return multi, whole


def ncp_run_multi(X, inits, n_iter_max=100, tol=1e-6, cvg_criterion="abs_rec_error", norm_x2=None):
    """Independent MU factorizations of a small (L2-resident) tensor as ONE kernel launch, a CTA per run
    (``c2c_ncp_run_multi``).  ``inits``: one list of host ``[I_k, rank_b]`` arrays per run (any ranks <= 32).
    Returns ``[(factors (host fp32 arrays), errors float64 [n_iter], n_iter, converged, sums [5]), ...]`` in the order given.
    Raises ``C2CError`` with ``code == -61`` when the shape is outside this path."""
    _check_tensor(X, None)
    if cvg_criterion not in CVG:
        raise C2CError("cvg_criterion %r is not one of %s" % (cvg_criterion, sorted(CVG)))
    nruns, nd = len(inits), X.dim()
    ranks = [int(np.asarray(f[0]).shape[1]) for f in inits]
    # the kernel takes the runs in grid order: the most expensive first
    order = sorted(range(nruns), key=lambda b: (-ranks[b], b))
    shape = [int(s) for s in X.shape]
    offs, total = [], 0
    for b in order:
        ldr = ldr_of(ranks[b])
        row = []
        for k in range(nd):
            row.append(total)
            total += shape[k] * ldr
        offs.append(row)
    host = np.zeros(total, dtype=np.float32)
    for b, row in zip(order, offs):
        ldr = ldr_of(ranks[b])
        for k in range(nd):
            f = np.abs(np.asarray(inits[b][k], dtype=np.float32))
            host[row[k]:row[k] + shape[k] * ldr].reshape(shape[k], ldr)[:, :ranks[b]] = f
    with torch.cuda.device(X.device):
        flat = torch.from_numpy(host).to(X.device)
        n_it = max(int(n_iter_max), 1)
        ranks_arr = (ctypes.c_int * nruns)(*[ranks[b] for b in order])
        nbytes = int(lib().c2c_ncp_multi_workspace_bytes(nd, _shape_arr(shape), ranks_arr, nruns))
        if nbytes == 0:
            check(-61, "c2c_ncp_multi_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=X.device)
        errors = torch.zeros((nruns, n_it), dtype=torch.float64, device=X.device)
        state = torch.zeros((nruns, 2), dtype=torch.int32, device=X.device)
        sums = torch.zeros((nruns, 5), dtype=torch.float64, device=X.device)
        nx2 = torch.tensor([float(norm_x2) if norm_x2 is not None else sumsq(X)], dtype=torch.float64, device=X.device)
        base = flat.data_ptr()
        ptrs = (_vp * (nruns * nd))(*[base + 4 * o for row in offs for o in row])
        check(lib().c2c_ncp_run_multi(_ptr(X), nd, _shape_arr(shape), ptrs, ranks_arr, nruns, n_it, float(tol), CVG[cvg_criterion],
                                      _ptr(nx2), _ptr(errors), _ptr(state), _ptr(sums), _ptr(ws), nbytes, _stream()),
              "c2c_ncp_run_multi")
        out_host = flat.cpu().numpy()
        errs, st, sm = errors.cpu().numpy(), state.cpu().numpy(), sums.cpu().numpy()
    res = [None] * nruns
    for j, (b, row) in enumerate(zip(order, offs)):
        ldr = ldr_of(ranks[b])
        fs = [out_host[row[k]:row[k] + shape[k] * ldr].reshape(shape[k], ldr)[:, :ranks[b]].copy() for k in range(nd)]
        n_iter = int(st[j, 0])
        res[b] = (fs, errs[j, :n_iter].copy(), n_iter, bool(st[j, 1]), sm[j].copy())
    return res


def shard_comm_bytes(local_shape, rank, world):
    """Bytes of one rank's communication buffer for ``ncp_run_sharded`` (``c2c_shard_comm_bytes``)."""
    n = int(lib().c2c_shard_comm_bytes(len(local_shape), _shape_arr(local_shape), int(rank), int(world)))
    if n == 0:
        check(-71, "c2c_shard_comm_bytes")
    return n


def ncp_run_sharded(X_local, factors, rank, comm, n_iter_max=100, tol=1e-6, cvg_criterion="abs_rec_error", check_every=10,
                    norm_x2=None, ws=None):
    """This rank's part of one MU factorization sharded along mode 0 (``c2c_ncp_run_sharded``): ``X_local`` is the slab,
    ``factors[0]`` the local rows of the first factor, ``factors[1:]`` full replicated factors (packed, updated in place);
    ``comm``: an object with ``world``, ``me``, ``ptrs`` (every rank's buffer as mapped here, ``ptrs[me]`` local), ``nbytes`` and a
    growing ``epoch`` (``sharded.PeerComm``); ``norm_x2``: 1-element fp64 CUDA tensor with the squared norm of the WHOLE tensor.
    Returns ``(errors, n_iter, converged)``."""
    _check_tensor(X_local, None)
    if cvg_criterion not in CVG:
        raise C2CError("cvg_criterion %r is not one of %s" % (cvg_criterion, sorted(CVG)))
    for f, s in zip(factors, X_local.shape):
        if f.shape != (s, ldr_of(rank)) or f.dtype != torch.float32 or not f.is_contiguous() or f.device != X_local.device:
            raise C2CError("factors must be packed fp32 [I_n, ldr] tensors on the tensor's device (pack_factor)")
    if norm_x2 is None or not isinstance(norm_x2, torch.Tensor):
        raise C2CError("ncp_run_sharded needs norm_x2: the squared norm of the whole tensor as a 1-element fp64 CUDA tensor")
    with torch.cuda.device(X_local.device):
        ws = _ws_for(X_local, rank, None, ws)
        n_it = max(int(n_iter_max), 1)
        errors = torch.zeros(n_it, dtype=torch.float64, device=X_local.device)
        state = torch.zeros(2, dtype=torch.int32, device=X_local.device)
        ptrs = (_vp * comm.world)(*[int(p) for p in comm.ptrs])
        base = int(comm.epoch)
        comm.epoch = base + n_it + 2
        check(lib().c2c_ncp_run_sharded(_ptr(X_local), X_local.dim(), _shape_arr(X_local.shape), _ptr_arr(factors), ldr_of(rank),
                                        int(rank), int(n_iter_max), float(tol), CVG[cvg_criterion], _ptr(norm_x2), _ptr(errors),
                                        _ptr(state), int(check_every), ws.ptr, ws.nbytes, int(comm.world), int(comm.me), ptrs,
                                        int(comm.nbytes), base & 0x7fffffff, _stream()), "c2c_ncp_run_sharded")
        st = state.cpu().numpy()
        n_iter = int(st[0])
        return errors[:n_iter].cpu().numpy(), n_iter, bool(st[1])


def coupled_ncp_run(X1, mask1, factors1, X2, mask2, factors2, rank, shared_pairs, n_iter_max=100, tol=1e-6,
                    cvg_criterion="abs_rec_error", weights=(1.0, 1.0), check_every=10, norm_x2=None):
    """Coupled multiplicative-update factorization of two tensors sharing the mode pairs ``shared_pairs``
    (``c2c_coupled_ncp_run``), updating both sets of packed factors in place.  Returns ``(errors [n_iter, 2] float64,
    n_iter, converged)``."""
    _check_tensor(X1, mask1)
    _check_tensor(X2, mask2)
    if X1.device != X2.device:
        raise C2CError("both tensors must live on the same device")
    if rank < 1 or rank > 32:
        raise C2CError("coupled factorization takes ranks 1..32, got %r" % (rank,))
    if cvg_criterion not in CVG:
        raise C2CError("cvg_criterion %r is not one of %s" % (cvg_criterion, sorted(CVG)))
    for X, fs in ((X1, factors1), (X2, factors2)):
        if len(fs) != X.dim():
            raise C2CError("need one factor per mode")
        for f, s in zip(fs, X.shape):
            if f.shape != (s, ldr_of(rank)) or f.dtype != torch.float32 or not f.is_contiguous() or f.device != X.device:
                raise C2CError("factors must be packed fp32 [I_n, ldr] tensors on the tensor's device (pack_factor)")
    pairs = [int(v) for p in shared_pairs for v in p]
    with torch.cuda.device(X1.device):
        ws1 = Workspace(X1.shape, rank, mask1 is not None, X1.device)
        ws2 = Workspace(X2.shape, rank, mask2 is not None, X2.device)
        n_it = max(int(n_iter_max), 1)
        errors = torch.zeros((n_it, 2), dtype=torch.float64, device=X1.device)
        state = torch.zeros(2, dtype=torch.int32, device=X1.device)
        nx2 = None
        if norm_x2 is not None:
            nx2 = torch.tensor([float(norm_x2[0]), float(norm_x2[1])], dtype=torch.float64, device=X1.device)
        arr = (ctypes.c_int * len(pairs))(*pairs)
        check(lib().c2c_coupled_ncp_run(_ptr(X1), _ptr(mask1), X1.dim(), _shape_arr(X1.shape), _ptr_arr(factors1),
                                        _ptr(X2), _ptr(mask2), X2.dim(), _shape_arr(X2.shape), _ptr_arr(factors2),
                                        ldr_of(rank), rank, arr, len(pairs) // 2, int(n_iter_max), float(tol),
                                        CVG[cvg_criterion], float(weights[0]), float(weights[1]), _ptr(nx2), _ptr(errors),
                                        _ptr(state), int(check_every), ws1.ptr, ws1.nbytes, ws2.ptr, ws2.nbytes, _stream()),
              "c2c_coupled_ncp_run")
        st = state.cpu().numpy()
        n_iter = int(st[0])
        return errors[:n_iter].cpu().numpy(), n_iter, bool(st[1])


def build_ccc(expr, expr_off, expr_ld, cell_col, a_first, a_count, b_first, b_count, sub_rows, C, L, K, score, agg,
              with_mask, device):
    """K1: returns ``(X [C, L, K, K] fp32, mask uint8 or None)`` on ``device``.  Index arrays are host numpy arrays."""
    dev = require_cuda(device)
    with torch.cuda.device(dev):
        def up(a, dt):
            return torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).to(dev, non_blocking=True)
        d_expr = expr if isinstance(expr, torch.Tensor) else up(expr, np.float32)
        d = [up(expr_off, np.int64), up(expr_ld, np.int32), up(cell_col, np.int32), up(a_first, np.int32),
             up(a_count, np.int32), up(b_first, np.int32), up(b_count, np.int32),
             up(sub_rows if len(sub_rows) else np.zeros(1, np.int32), np.int32)]
        X = torch.empty((C, L, K, K), dtype=torch.float32, device=dev)
        mask = torch.empty((C, L, K, K), dtype=torch.uint8, device=dev) if with_mask else None
        ev = None
        if BUILD_EVENTS is not None:                 # bench.py: CUDA events around the kernel alone (uploads are before `a`)
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        check(lib().c2c_build_ccc(_ptr(d_expr), *[_ptr(t) for t in d], C, L, K, SCORE[score], AGG[agg], _ptr(X), _ptr(mask),
                                  _stream()), "c2c_build_ccc")
        if ev is not None:
            ev[1].record()
            BUILD_EVENTS.append(ev)
        return X, mask


def group_aggregate(X, mask, grp_off, grp_rows, method):
    """K1b: aggregate LR rows of ``X [C, L, ...]`` into groups; returns ``(X_out, mask_out)``."""
    dev = X.device
    C, L = int(X.shape[0]), int(X.shape[1])
    E = int(np.prod(X.shape[2:]))
    ng = len(grp_off) - 1
    with torch.cuda.device(dev):
        d_off = torch.as_tensor(np.ascontiguousarray(grp_off, dtype=np.int32)).to(dev)
        d_rows = torch.as_tensor(np.ascontiguousarray(grp_rows, dtype=np.int32)).to(dev)
        Xo = torch.empty((C, ng) + tuple(X.shape[2:]), dtype=torch.float32, device=dev)
        mo = torch.empty(Xo.shape, dtype=torch.uint8, device=dev)
        check(lib().c2c_group_aggregate(_ptr(X), _ptr(mask), _ptr(d_off), _ptr(d_rows), C, L, ng, E, AGG[method], _ptr(Xo),
                                        _ptr(mo), _stream()), "c2c_group_aggregate")
        return Xo, mo

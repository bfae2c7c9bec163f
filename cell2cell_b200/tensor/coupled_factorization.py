"""Coupled Tensor Component Analysis on the device: same functions, signatures and results as
``cell2cell/tensor/coupled_factorization.py`` ("next" row 2).

Two tensors that share some of their modes are factorized together by multiplicative updates
(coupled_factorization.py:123-482): per iteration the tensor-specific modes are updated first (each exactly like
``non_negative_parafac``), then the shared modes in reverse order of ``mode_mapping['shared']`` with the numerators
(MTTKRPs) and denominators (``A @ Gram-Hadamard``) of the two tensors averaged before the clip; the stop rule runs on the
size-balanced combination of the two Gram-form reconstruction errors.

How it maps on the B200 kernels (``libc2c_b200.so`` through ``engine``; there is no CPU path):

* an MTTKRP of a leading mode comes from the plane contraction ``T`` of its tensor, one of a trailing mode from the fiber
  contraction ``U`` (the two streaming passes of the single-tensor path).  ``T`` depends only on the two trailing factors
  and ``U`` only on the leading ones, so a contraction is recomputed only when a factor it depends on changed since it was
  formed (``_Side``): whatever the update order the mode mapping induces, an unmasked tensor is read **twice** per
  iteration, not once per mode -- and ``T`` survives from the end of one iteration into the next.
* with a mask the missing entries are re-imputed from the *current* factors before every mode update
  (coupled_factorization.py:342-343), so every MTTKRP is the fused masked single-mode pass (``c2c_mttkrp``).
* shared modes: ``c2c_mu_update_coupled`` -- one kernel forms both denominators, averages, updates and writes the new
  factor into both tensors' copies; on the last update of the iteration it also accumulates the two inner products
  ``<X_k, cp_k>`` of tensorly's ``error_calc`` (the reference spends one more MTTKRP per tensor on them, :473-485).
* the error scalars and the combined stop rule are evaluated on the device (``coupled_error_kernel``); iteration 0 runs
  eagerly (cold caches), then ONE iteration is captured into a CUDA graph and replayed: ~26 launches per iteration for two
  4-way tensors sharing three modes, no host work besides the replay.

Two drivers produce that sequence: ``c2c_coupled_ncp_run`` inside the library (the product path, rank <= 32) and
``_coupled_mu`` below -- the same order / cache logic written against the small ``ops`` interface of
``cell2cell_b200.sharded.DeviceOps`` over the single-step entry points; it serves ranks 33..128 and lets the CPU suite
check the orchestration with numpy stand-ins for the kernels (tests/test_coupled_oracle.py).
``C2C_B200_COUPLED_HOST_DRIVER=1`` forces it on the device as well (the GPU tests run both).
"""
import os

import numpy as np
import torch
from tqdm.auto import tqdm

from .. import engine
from .cp import CPTensor, initialize_factors
from .factorization import _compute_norm_error, as_device_mask, as_device_tensor

__all__ = ["coupled_non_negative_parafac", "_process_mode_mapping", "_validate_tensors", "_compute_coupled_tensor_factorization",
           "_run_coupled_elbow_analysis", "_multiple_runs_coupled_elbow_analysis", "_balance_weights", "_combined_error"]


def _raw(t):
    return t.tensor if hasattr(t, "tensor") else t


def _process_mode_mapping(tensor1, tensor2, mode_mapping):
    """dict ``{'shared': [(t1_mode, t2_mode), ...]}`` passes through; an int / list is the old ``non_shared_modes`` of two
    tensors of the same order (coupled_factorization.py:15-65)."""
    if isinstance(mode_mapping, dict):
        return mode_mapping
    ndim1, ndim2 = len(_raw(tensor1).shape), len(_raw(tensor2).shape)
    if ndim1 != ndim2:
        raise ValueError("When tensors have different dimensions, mode_mapping must be a dict. "
                         "For backward compatibility with non_shared_modes, both tensors must have same dimensions.")
    if isinstance(mode_mapping, (int, np.integer)):
        non_shared_modes = [int(mode_mapping)]
    elif isinstance(mode_mapping, (list, tuple)):
        non_shared_modes = list(mode_mapping)
    else:
        raise ValueError("mode_mapping must be dict, int, list, or tuple")
    if len(set(non_shared_modes)) != len(non_shared_modes):
        original = list(non_shared_modes)
        non_shared_modes = list(set(non_shared_modes))
        print("Warning: Duplicate modes found in non_shared_modes {}. Using deduplicated version: {}".format(
            original, non_shared_modes))
    return {"shared": [(i, i) for i in range(ndim1) if i not in non_shared_modes]}


def _validate_tensors(tensor1, tensor2, mode_mapping):
    """Shared modes must exist, be of equal size and (for tensor objects) carry the same element names
    (coupled_factorization.py:68-120)."""
    t1, t2 = _raw(tensor1), _raw(tensor2)
    ndim1, ndim2 = len(t1.shape), len(t2.shape)
    shared_pairs = mode_mapping.get("shared", [])
    shared_t1 = set(p[0] for p in shared_pairs)
    shared_t2 = set(p[1] for p in shared_pairs)
    if not shared_t1.issubset(set(range(ndim1))):
        raise ValueError("Shared modes contain invalid tensor1 modes: {}".format(sorted(shared_t1 - set(range(ndim1)))))
    if not shared_t2.issubset(set(range(ndim2))):
        raise ValueError("Shared modes contain invalid tensor2 modes: {}".format(sorted(shared_t2 - set(range(ndim2)))))
    if len(shared_pairs) == 0:
        raise ValueError("At least one mode must be shared between tensors")
    for a, b in shared_pairs:
        if t1.shape[a] != t2.shape[b]:
            raise ValueError("Shared modes must have same size: tensor1 mode {} ({}) vs tensor2 mode {} ({})".format(
                a, t1.shape[a], b, t2.shape[b]))
    if hasattr(tensor1, "order_names") and hasattr(tensor2, "order_names"):
        for a, b in shared_pairs:
            if list(tensor1.order_names[a]) != list(tensor2.order_names[b]):
                raise ValueError("Element names must match for shared dimensions: tensor1 mode {} vs tensor2 mode {}".format(a, b))


def _split_modes(ndim1, ndim2, shared_pairs):
    s1 = set(p[0] for p in shared_pairs)
    s2 = set(p[1] for p in shared_pairs)
    return [i for i in range(ndim1) if i not in s1], [i for i in range(ndim2) if i not in s2]


def _balance_weights(shape1, shape2, shared_pairs, balance_errors=True):
    """Weights of the two reconstruction errors from the sizes of the tensor-specific modes (coupled_factorization.py:258-269)."""
    if not balance_errors:
        return 1.0, 1.0
    only1, only2 = _split_modes(len(shape1), len(shape2), shared_pairs)
    n1 = float(np.prod([shape1[i] for i in only1])) if only1 else 1.0
    n2 = float(np.prod([shape2[i] for i in only2])) if only2 else 1.0
    total = n1 + n2
    if total > 0:
        return (total / n1 if n1 > 0 else 1.0), (total / n2 if n2 > 0 else 1.0)
    return 1.0, 1.0


def _combined_error(error1, error2, shape1, shape2, mode_mapping, balance_errors=True):
    """The error best-of-runs and the elbow compare (coupled_factorization.py:614-639, coupled_tensor.py:229-244)."""
    if not balance_errors:
        return (error1 + error2) / 2
    w1, w2 = _balance_weights(shape1, shape2, mode_mapping.get("shared", []), True)
    return (w1 * error1 + w2 * error2) / (w1 + w2)


class _Side:
    """One tensor of the pair: its packed factors, workspaces and the two cached contractions."""

    def __init__(self, X, mask, host_factors, ops):
        self.X, self.mask, self.ops = X, mask, ops
        self.shape = tuple(int(s) for s in X.shape)
        self.nd = len(self.shape)
        self.A = [ops.pack(f, X.device) for f in host_factors]
        self.ws = ops.workspace(self.shape, X.device, masked=mask is not None)
        self.T = self.U = None
        self.T_ok = self.U_ok = False
        self.passes = 0                                    # streaming passes issued (reads of the tensor)

    def mttkrp(self, mode):
        ops = self.ops
        if self.mask is not None:
            self.passes += 1
            return ops.masked_mttkrp(self.X, self.mask, self.A, mode, self.ws)
        if mode < self.nd - 2:
            if not self.T_ok:
                self.T = ops.plane(self.X, self.A, self.ws, out=self.T)     # the buffer persists across (replayed) iterations
                self.T_ok = True
                self.passes += 1
            return ops.m_from(self.T, self.shape, self.A, mode)
        if not self.U_ok:
            self.U = ops.fiber(self.X, self.A, self.ws, out=self.U)
            self.U_ok = True
            self.passes += 1
        return ops.m_from(self.U, self.shape, self.A, mode)

    def updated(self, mode):
        """Factor ``mode`` changed: T depends on the two trailing factors, U on the leading ones."""
        if mode >= self.nd - 2:
            self.T_ok = False
        else:
            self.U_ok = False

    def gram(self, skip):
        return self.ops.gram_hadamard(self.A, self.shape, skip)


def _update_order(ndim1, ndim2, shared_pairs):
    """Tensor-specific modes first, then the shared pairs in reverse (coupled_factorization.py:309-320)."""
    only1, only2 = _split_modes(ndim1, ndim2, shared_pairs)
    return [("tensor1_only", m, None) for m in only1] + [("tensor2_only", None, m) for m in only2] + \
           [("shared", a, b) for a, b in reversed(list(shared_pairs))]


def _coupled_mu(X1, X2, M1, M2, init1, init2, rank, shared_pairs, n_iter_max, tol, cvg_criterion, w1, w2, ops,
                use_graph=True):
    """The iteration loop.  Returns ``(side1, side2, errors1, errors2)``; the factors stay packed on the device."""
    dev = X1.device
    s1, s2 = _Side(X1, M1, init1, ops), _Side(X2, M2, init2, ops)
    for a, b in shared_pairs:                                      # :294-295 shared factors start as tensor 1's
        s2.A[b].copy_(s1.A[a])
    order = _update_order(s1.nd, s2.nd, shared_pairs)
    nx2 = torch.tensor([getattr(X, "_c2c_norm_x2", None) or ops.sumsq(X) for X in (X1, X2)], dtype=torch.float64, device=dev)
    iprod = torch.zeros(2, dtype=torch.float64, device=dev)
    # masked tensors: squared norm of the entries the last imputation of the iteration writes.  tensorly's error_calc is
    # called without an MTTKRP (coupled_factorization.py:420-431): the direct norm of (imputed tensor - reconstruction), whose
    # ||imputed||^2 is ||mask * X||^2 plus this term
    miss2 = torch.zeros(2, dtype=torch.float64, device=dev)
    err_dev = torch.zeros(2, dtype=torch.float64, device=dev)
    n_it = max(int(n_iter_max), 1)
    hist_dev = torch.zeros((n_it, 2), dtype=torch.float64, device=dev)

    def iteration():
        for k, (kind, a, b) in enumerate(order):
            if kind == "tensor1_only":
                ops.mu_update(s1.A[a], s1.mttkrp(a), s1.gram(a)[0])
                s1.updated(a)
            elif kind == "tensor2_only":
                ops.mu_update(s2.A[b], s2.mttkrp(b), s2.gram(b)[0])
                s2.updated(b)
            else:
                last = k == len(order) - 1
                if last:
                    for j, sd in enumerate((s1, s2)):
                        if sd.mask is not None:
                            ops.missing_rec_sumsq(sd.X, sd.mask, sd.A, sd.ws, out=miss2[j:j + 1])
                ops.mu_update_coupled(s1.A[a], s2.A[b], s1.mttkrp(a), s2.mttkrp(b), s1.gram(a)[0], s2.gram(b)[0],
                                      iprod if last else None)
                s1.updated(a)
                s2.updated(b)
        rn2 = torch.cat([s1.gram(-1)[1].reshape(1), s2.gram(-1)[1].reshape(1)]).to(torch.float64)
        err_dev.copy_(torch.sqrt(torch.abs(nx2 + miss2 + rn2 - 2.0 * iprod)) / torch.sqrt(nx2))

    if cvg_criterion == "abs_rec_error":
        rule = lambda d: abs(d) < tol                                 # noqa: E731
    elif cvg_criterion == "rec_error":
        rule = lambda d: d < tol                                      # noqa: E731
    else:
        raise ValueError("Unknown convergence criterion")
    comb = lambda e: (w1 * e[0] + w2 * e[1]) / (w1 + w2)              # noqa: E731

    graph = None
    check_every = 1 if (tol and tol > 0) else n_it
    it = 0
    hist = np.zeros((0, 2))
    if int(n_iter_max) < 1:
        return s1, s2, [], []
    if dev.type == "cuda" and use_graph and int(n_iter_max) >= 3:
        iteration()                                   # eager first iteration: cold caches, lazy initialisations
        hist_dev[0].copy_(err_dev)
        it = 1
        torch.cuda.synchronize(dev)
        state = (s1.T_ok, s1.U_ok, s2.T_ok, s2.U_ok)
        try:
            side = torch.cuda.Stream(device=dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                iteration()
            # the cache state at the end of an iteration depends only on the update order, so the captured iteration starts
            # and ends in the same state and can be replayed
            assert state == (s1.T_ok, s1.U_ok, s2.T_ok, s2.U_ok)
        except Exception:
            graph = None
            torch.cuda.synchronize(dev)
            s1.T_ok = s1.U_ok = s2.T_ok = s2.U_ok = False
    stop = False
    if it:
        hist = hist_dev[:it].cpu().numpy()
    while it < int(n_iter_max) and not stop:
        if graph is not None:
            graph.replay()
        else:
            iteration()
        hist_dev[it].copy_(err_dev)
        it += 1
        if graph is None and dev.type != "cuda" or it % check_every == 0 or it == int(n_iter_max):
            new = hist_dev[:it].cpu().numpy()
            for k in range(max(len(hist), 1), it):
                if tol and rule(comb(new[k - 1]) - comb(new[k])):     # :437-458, from the second iteration on
                    new = new[:k + 1]
                    stop = True
                    break
            hist = new
    return s1, s2, [np.float64(e) for e in hist[:, 0]], [np.float64(e) for e in hist[:, 1]]


def coupled_non_negative_parafac(tensor1, tensor2, rank, mode_mapping, mask1=None, mask2=None, n_iter_max=100, init="svd",
                                 svd="truncated_svd", tol=10e-7, random_state=None, verbose=0, normalize_factors=False,
                                 return_errors=False, cvg_criterion="abs_rec_error", balance_errors=True,
                                 separate_weights=True, ops=None):
    """Coupled non-negative CP of two tensors sharing the mode pairs ``mode_mapping['shared']``
    (coupled_factorization.py:123-482).  Returns ``(cp_tensor1, cp_tensor2[, (errors1, errors2)])``.

    ``separate_weights`` only matters together with ``normalize_factors=True`` (weights are all ones otherwise), which --
    as in the single-tensor path -- is not supported here (cell2cell never sets it).  tensorly records no errors when
    ``tol`` is falsy; they are always recorded here."""
    if normalize_factors:
        raise NotImplementedError("normalize_factors=True is not supported by the B200 path (cell2cell never sets it)")
    rank = int(rank)
    X1 = as_device_tensor(_raw(tensor1))
    X2 = as_device_tensor(_raw(tensor2), X1.device)
    mode_mapping = _process_mode_mapping(X1, X2, mode_mapping)
    shared_pairs = [(int(a), int(b)) for a, b in mode_mapping.get("shared", [])]
    _validate_tensors(X1, X2, {"shared": shared_pairs})
    if len(set(a for a, _ in shared_pairs)) != len(shared_pairs) or len(set(b for _, b in shared_pairs)) != len(shared_pairs):
        raise ValueError("A mode can be shared only once")
    M1 = as_device_mask(mask1, X1)
    M2 = as_device_mask(mask2, X2)
    # the reference multiplies by the mask before every use (:342-343): values under a zero mask byte never matter.  The
    # device path needs zeros there (the error's imputed-entry term is accumulated with x taken as 0)
    if M1 is not None:
        X1 = X1 * M1
    if M2 is not None:
        X2 = X2 * M2
    w1, w2 = _balance_weights(X1.shape, X2.shape, shared_pairs, balance_errors)
    if not isinstance(init, str):
        raise ValueError("coupled factorization takes init='svd' or init='random'")
    init1 = initialize_factors(X1, rank, init=init, svd=svd, random_state=random_state)        # :272-280: same random_state
    init2 = initialize_factors(X2, rank, init=init, svd=svd, random_state=random_state)
    init1 = [f.astype(np.float32) for f in init1]
    init2 = [f.astype(np.float32) for f in init2]
    tol_run = float(tol) if tol else -1.0
    passes = (None, None)
    if ops is None and rank <= 32 and os.environ.get("C2C_B200_COUPLED_HOST_DRIVER") != "1":
        # the whole run inside the library (c2c_coupled_ncp_run): update chain, contraction caches, error scalars and the
        # combined stop rule on the device, one captured iteration replayed
        A1 = [engine.pack_factor(f, rank, X1.device) for f in init1]
        A2 = [engine.pack_factor(f, rank, X1.device) for f in init2]
        for a, b in shared_pairs:                                   # :294-295 shared factors start as tensor 1's
            A2[b].copy_(A1[a])
        cached = [getattr(X, "_c2c_norm_x2", None) for X in (X1, X2)]
        errs, n_iter, _ = engine.coupled_ncp_run(X1, M1, A1, X2, M2, A2, rank, shared_pairs, n_iter_max=int(n_iter_max),
                                                 tol=tol_run, cvg_criterion=cvg_criterion, weights=(w1, w2), check_every=10,
                                                 norm_x2=cached if all(c is not None for c in cached) else None)
        errors1 = [np.float64(e) for e in errs[:, 0]]
        errors2 = [np.float64(e) for e in errs[:, 1]]
        recon = lambda X, A: engine.recon_sums(X, None, A, rank)      # noqa: E731
    else:
        # host-driven iteration over the single-step entry points (any rank up to 128; also what the CPU tests drive with
        # numpy stand-ins for the kernels)
        if ops is None:
            from ..sharded import DeviceOps
            ops = DeviceOps(rank)
        s1, s2, errors1, errors2 = _coupled_mu(X1, X2, M1, M2, init1, init2, rank, shared_pairs, int(n_iter_max), tol_run,
                                               cvg_criterion, w1, w2, ops)
        A1, A2 = s1.A, s2.A
        passes = (s1.passes, s2.passes)
        recon = lambda X, A: ops.recon_sums(X, None, A, None)         # noqa: E731
    cps = []
    for X, M, A, errors, npass in ((X1, M1, A1, errors1, passes[0]), (X2, M2, A2, errors2, passes[1])):
        sums = None
        if M is None and len(errors):
            # as in the single-tensor path the kept (last) error is the directly accumulated ||X - rec|| / ||X|| (fp64 sums
            # of one fused pass): the Gram form cancels in fp32 once the error is small
            sums = recon(X, A)
            errors[-1] = np.float64(np.sqrt(sums[1]) / np.sqrt(sums[3]))
        cp = CPTensor(np.ones(rank, dtype=np.float32), [engine.unpack_factor(a, rank).cpu().numpy() for a in A], device=X1.device)
        if sums is not None:
            cp.recon_sums_ = sums
        cp.passes_ = npass
        cps.append(cp)
    if verbose:
        for it, (e1, e2) in enumerate(zip(errors1, errors2)):
            print("Iteration {}: rec_error1={:.6f}, rec_error2={:.6f}, combined={:.6f}".format(
                it, e1, e2, (w1 * e1 + w2 * e2) / (w1 + w2)))
    if return_errors:
        return cps[0], cps[1], (errors1, errors2)
    return cps[0], cps[1]


def _compute_coupled_tensor_factorization(tensor1, tensor2, rank, mode_mapping, mask1=None, mask2=None,
                                          tf_type="coupled_non_negative_cp", init="svd", svd="truncated_svd",
                                          random_state=None, n_iter_max=100, tol=10e-7, verbose=False, balance_errors=True,
                                          **kwargs):
    """Same contract as the reference's ``_compute_coupled_tensor_factorization`` (coupled_factorization.py:485-518)."""
    kwargs = dict(kwargs or {})
    kwargs.setdefault("return_errors", False)
    if tf_type != "coupled_non_negative_cp":
        raise ValueError("Not a valid tf_type for coupled factorization.")
    return coupled_non_negative_parafac(tensor1=tensor1, tensor2=tensor2, rank=rank, mode_mapping=mode_mapping, mask1=mask1,
                                        mask2=mask2, init="random" if (mask1 is not None or mask2 is not None) else init,
                                        svd=svd, random_state=random_state, n_iter_max=n_iter_max, tol=tol, verbose=verbose,
                                        balance_errors=balance_errors, **kwargs)


def _coupled_loss(X1, X2, M1, M2, mode_mapping, rank, balance_errors, **fact_kwargs):
    """Combined error of one coupled factorization (coupled_factorization.py:598-641)."""
    cp1, cp2, (errors1, errors2) = _compute_coupled_tensor_factorization(X1, X2, rank, mode_mapping, mask1=M1, mask2=M2,
                                                                         balance_errors=balance_errors, **fact_kwargs)
    error1 = np.float64(errors1[-1]) if M1 is None else _compute_norm_error(X1, cp1, M1)
    error2 = np.float64(errors2[-1]) if M2 is None else _compute_norm_error(X2, cp2, M2)
    return float(_combined_error(error1, error2, X1.shape, X2.shape, mode_mapping, balance_errors))


def _run_coupled_elbow_analysis(tensor1, tensor2, mode_mapping, upper_rank=50, tf_type="coupled_non_negative_cp", init="svd",
                                svd="truncated_svd", random_state=None, mask1=None, mask2=None, n_iter_max=100, tol=10e-7,
                                verbose=False, balance_errors=True, disable_pbar=False, **kwargs):
    """One coupled factorization per rank 1..upper_rank; list of (rank, combined error) (coupled_factorization.py:520-643)."""
    kwargs = dict(kwargs or {})
    kwargs["return_errors"] = True
    X1 = as_device_tensor(_raw(tensor1))
    X2 = as_device_tensor(_raw(tensor2), X1.device)
    M1, M2 = as_device_mask(mask1, X1), as_device_mask(mask2, X2)
    mode_mapping = _process_mode_mapping(X1, X2, mode_mapping)
    loss = []
    for r in tqdm(range(1, upper_rank + 1), disable=disable_pbar):
        loss.append((r, np.float64(_coupled_loss(X1, X2, M1, M2, mode_mapping, r, balance_errors, tf_type=tf_type, init=init,
                                                 svd=svd, random_state=random_state, n_iter_max=n_iter_max, tol=tol,
                                                 verbose=verbose, **kwargs))))
    return loss


def _multiple_runs_coupled_elbow_analysis(tensor1, tensor2, mode_mapping, upper_rank=50, runs=10,
                                          tf_type="coupled_non_negative_cp", init="svd", svd="truncated_svd", metric="error",
                                          random_state=None, mask1=None, mask2=None, n_iter_max=100, tol=10e-7, verbose=False,
                                          balance_errors=True, **kwargs):
    """``runs`` coupled factorizations per rank; ndarray (runs x upper_rank) of combined errors
    (coupled_factorization.py:646-810).  The ``upper_rank * runs`` units are independent: with ``torch.distributed``
    initialised they are spread over the ranks (``cell2cell_b200.parallel``), both tensors replicated per GPU."""
    assert isinstance(runs, int), "runs must be an integer"
    if metric == "similarity":
        raise NotImplementedError("Similarity metric for coupled tensors not yet implemented")
    from ..parallel import run_units
    kwargs = dict(kwargs or {})
    kwargs["return_errors"] = True
    X1 = as_device_tensor(_raw(tensor1))
    X2 = as_device_tensor(_raw(tensor2), X1.device)
    M1, M2 = as_device_mask(mask1, X1), as_device_mask(mask2, X2)
    mode_mapping = _process_mode_mapping(X1, X2, mode_mapping)
    units = [(r, run) for r in range(1, upper_rank + 1) for run in range(runs)]

    def one(unit):
        r, run = unit
        seed = None if random_state is None else random_state + run
        return _coupled_loss(X1, X2, M1, M2, mode_mapping, r, balance_errors, tf_type=tf_type, init=init, svd=svd,
                             random_state=seed, n_iter_max=n_iter_max, tol=tol, verbose=verbose, **kwargs)

    vals = run_units(units, one, cost=lambda u: float(u[0]), gather="float")
    return np.asarray(vals, dtype=np.float64).reshape(upper_rank, runs).T


_ = engine

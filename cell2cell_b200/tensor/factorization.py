"""Factorization front end: same functions and signatures as ``cell2cell/tensor/factorization.py``, with the tensorly
calls replaced by the device driver (``cell2cell_b200.engine`` -> ``libc2c_b200.so``).

* ``non_negative_parafac`` / ``non_negative_parafac_hals`` keep tensorly's signatures (the reference forwards ``**kwargs``
  verbatim, factorization.py:101) and semantics: host-seeded or device-svd init, Gauss-Seidel mode order, eps clipping,
  the Gram-form reconstruction error, ``abs(err[-2] - err[-1]) < tol`` stop rule, masked imputation with the current
  factors before every mode.
* ``_compute_tensor_factorization`` (factorization.py:14-143), ``_run_elbow_analysis`` (:186-269),
  ``_multiple_runs_elbow_analysis`` (:272-373), ``_compute_norm_error`` (:376-432), ``normalized_error`` (:435-458) and
  ``_compute_elbow`` (:146-183, Kneedle restated in ``elbow.py`` because kneed is not installable here).
"""
import os

import numpy as np
import torch
from tqdm.auto import tqdm

from .. import engine
from .._lib import C2CError
from .cp import CPTensor, initialize_factors
from .elbow import kneedle_elbow

__all__ = ["non_negative_parafac", "non_negative_parafac_hals", "parafac", "constrained_parafac", "_compute_tensor_factorization", "_compute_elbow",
           "_run_elbow_analysis", "_multiple_runs_elbow_analysis", "_compute_norm_error", "normalized_error",
           "as_device_tensor", "as_device_mask"]


def _shared_upload(host, dev, group=None):
    """The SAME host tensor on every rank of ``group`` -> an fp32 replica on every GPU with ONE host-to-device transfer of
    it in total: rank k uploads the k-th of ``world`` near-equal pieces, then an NCCL all-gather over NVLink assembles the
    replicas.  N simultaneous full uploads share the host's memory / PCIe root (measured: 8 x 768 MB copies take 2.6 x one
    copy); the all-gather moves the same bytes at NVLink speed instead."""
    import torch.distributed as dist
    world, me = dist.get_world_size(group), dist.get_rank(group)
    flat = host.reshape(-1)
    n = flat.numel()
    piece = -(-n // world)
    piece = -(-piece // 4) * 4                      # keeps every piece, and the replica's base, 16-byte aligned
    full = torch.empty(piece * world, dtype=torch.float32, device=dev)
    lo, hi = min(me * piece, n), min((me + 1) * piece, n)
    mine = full[me * piece:(me + 1) * piece]
    if hi > lo:
        mine[:hi - lo].copy_(flat[lo:hi], non_blocking=True)
    dist.all_gather_into_tensor(full, mine, group=group)
    return full[:n].view(host.shape)


def as_device_tensor(tensor, device=None, shared_upload=False, group=None):
    """fp32 contiguous CUDA view / copy of ``tensor`` (numpy or torch, any float or int dtype).

    ``shared_upload=True`` (with ``torch.distributed`` initialised on NCCL): the caller states that every rank of ``group``
    passes the same host tensor -- the replicas of an elbow sweep or of best-of-runs -- and the upload is split over the ranks
    (``_shared_upload``)."""
    if isinstance(tensor, torch.Tensor) and tensor.is_cuda and device is None:
        dev = tensor.device
    else:
        dev = engine.require_cuda(device)
    if shared_upload:
        import torch.distributed as dist
        host = tensor if isinstance(tensor, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(np.asarray(tensor)))
        if (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and not host.is_cuda
                and host.dtype == torch.float32 and dist.get_backend(group) == "nccl"):
            with torch.cuda.device(dev):
                return _shared_upload(host.contiguous(), dev, group)
    if isinstance(tensor, torch.Tensor):
        t = tensor.to(device=dev, dtype=torch.float32)
    else:
        arr = np.asarray(tensor)
        t = torch.as_tensor(np.ascontiguousarray(arr)).to(device=dev, dtype=torch.float32)
    return t.contiguous()


def as_device_mask(mask, like):
    """uint8 contiguous CUDA mask on ``like``'s device (non-zero = observed), or None."""
    if mask is None:
        return None
    if isinstance(mask, torch.Tensor):
        m = mask.to(device=like.device)
    else:
        m = torch.as_tensor(np.ascontiguousarray(np.asarray(mask))).to(like.device)
    if m.shape != like.shape:
        raise ValueError("mask and tensor must have the same shape")
    if m.dtype != torch.uint8:
        m = (m != 0).to(torch.uint8)
    return m.contiguous()


def _factorize(tensor, rank, algo, n_iter_max, init, svd, tol, random_state, verbose, return_errors, mask,
               cvg_criterion, normalize_factors, fixed_modes, check_every=10):
    if normalize_factors:
        raise NotImplementedError("normalize_factors=True is not supported by the B200 path (cell2cell never sets it)")
    if fixed_modes:
        raise NotImplementedError("fixed_modes is not supported by the B200 path (cell2cell never sets it)")
    rank = int(rank)
    X = as_device_tensor(tensor)
    M = as_device_mask(mask, X)
    host_factors = initialize_factors(X, rank, init=init, svd=svd, random_state=random_state)
    factors = [engine.pack_factor(f.astype(np.float32), rank, X.device) for f in host_factors]
    # tensorly records no errors when tol is falsy; the reference would then fail on errors[-1] -- we always record them.
    errors, n_iter, converged = engine.ncp_run(X, M, factors, rank, tf_type=algo, n_iter_max=int(n_iter_max),
                                               tol=float(tol) if tol else -1.0, cvg_criterion=cvg_criterion,
                                               check_every=check_every, norm_x2=getattr(X, "_c2c_norm_x2", None))
    sums = None
    if M is None and n_iter > 0:
        # The per-iteration errors are tensorly's Gram form sqrt(|nx2 + rn2 - 2 ip|) / sqrt(nx2): in fp32 it cancels once
        # the error is small.  The last one -- the value cell2cell keeps (tensor.py:311-315, factorization.py:262,356) --
        # is replaced by the directly accumulated ||X - rec|| / ||X|| (one extra fused pass, fp64 sums), which is what the
        # fp64 reference's Gram form equals.
        s = sums = engine.recon_sums(X, None, factors, rank)
        errors = np.array(errors, dtype=np.float64)
        errors[-1] = np.sqrt(s[1]) / np.sqrt(s[3])
    if verbose:
        for it, e in enumerate(errors):
            print("iteration {}, reconstruction error: {}".format(it, e))
        if converged:
            print("PARAFAC converged after {} iterations".format(n_iter))
    out = [engine.unpack_factor(f, rank).cpu().numpy() for f in factors]
    cp = CPTensor(np.ones(rank, dtype=np.float32), out, device=X.device)
    if sums is not None:
        cp.recon_sums_ = sums          # [sum d, sum d^2, sum x, sum x^2, n] of this tensor and these factors (explained_variance)
    if return_errors:
        return cp, [np.float64(e) for e in errors]
    return cp


def _factorize_batch(X, runs, n_iter_max, tol, cvg_criterion="abs_rec_error", check_every=10):
    """The factorizations ``runs = [(rank, seed), ...]`` (ranks adding up to <= 32) as ONE batched device factorization
    (``c2c_ncp_run_batched``): one read of the tensor per pass serves every run.  Same results as the corresponding calls of
    ``non_negative_parafac(init='random')`` (the runs do not interact; each has its own error history and stop iteration).
    Returns ``[(CPTensor, errors), ...]`` in the order of ``runs``."""
    ranks = [int(r) for r, _ in runs]
    total = sum(ranks)
    host = [initialize_factors(X, r, init="random", random_state=sd) for r, sd in runs]            # [run][mode]
    packed = [engine.pack_factor(np.concatenate([host[b][m].astype(np.float32) for b in range(len(runs))], axis=1), total,
                                 X.device) for m in range(X.dim())]
    errors, iters = engine.ncp_run_batched(X, packed, ranks, n_iter_max=int(n_iter_max), tol=float(tol) if tol else -1.0,
                                           cvg_criterion=cvg_criterion, check_every=check_every,
                                           norm_x2=getattr(X, "_c2c_norm_x2", None))
    out = []
    c0 = 0
    for b, rank in enumerate(ranks):
        fs = [engine.pack_factor(p[:, c0:c0 + rank], rank, X.device) for p in packed]
        c0 += rank
        errs = np.array(errors[b, :int(iters[b])], dtype=np.float64)
        sums = None
        if len(errs):
            sums = engine.recon_sums(X, None, fs, rank)                  # the kept error is the directly accumulated one
            errs[-1] = np.sqrt(sums[1]) / np.sqrt(sums[3])
        cp = CPTensor(np.ones(rank, dtype=np.float32), [engine.unpack_factor(f, rank).cpu().numpy() for f in fs], device=X.device)
        if sums is not None:
            cp.recon_sums_ = sums
        out.append((cp, [np.float64(e) for e in errs]))
    return out


def _multi_eligible(X, ranks, n_iter_max):
    """Small tensor, several runs: the one-CTA-per-run kernel beats whole-GPU kernels run after run (engine.multi_run_cost)."""
    if len(ranks) < 2 or max(ranks) > 32 or X.dim() < 3 or max(int(X.shape[-1]), int(X.shape[-2])) > 128:
        return False
    if X.numel() * 4 > 96e6:                                 # the grid of CTAs streams the tensor from L2
        return False
    multi, whole = engine.multi_run_cost(X.shape, ranks, n_iter_max)
    return multi < whole


def _factorize_multi(X, runs, n_iter_max, tol, cvg_criterion="abs_rec_error"):
    """The factorizations ``runs = [(rank, seed), ...]`` as ONE launch of the one-CTA-per-run kernel (``c2c_ncp_run_multi``).
    Same results as the corresponding ``non_negative_parafac(init='random')`` calls, to rounding.  Returns
    ``[(CPTensor, errors), ...]`` in the order of ``runs``; raises ``C2CError`` (code -61) if the shape is outside that path."""
    inits = [initialize_factors(X, r, init="random", random_state=sd) for r, sd in runs]
    res = engine.ncp_run_multi(X, inits, n_iter_max=int(n_iter_max), tol=float(tol) if tol else -1.0,
                               cvg_criterion=cvg_criterion, norm_x2=getattr(X, "_c2c_norm_x2", None))
    out = []
    for (rank, _), (fs, errs, n_iter, _conv, sums) in zip(runs, res):
        errs = np.array(errs, dtype=np.float64)
        if len(errs):
            errs[-1] = np.sqrt(sums[1]) / np.sqrt(sums[3])    # the kept error: directly accumulated (as _factorize does)
        cp = CPTensor(np.ones(rank, dtype=np.float32), fs, device=X.device)
        if len(errs):
            cp.recon_sums_ = sums
        out.append((cp, [np.float64(e) for e in errs]))
    return out


def _pack_runs(runs, capacity=28):
    """First-fit-decreasing packing of ``(rank, run)`` units into bins of at most ``capacity`` factor columns (a pass over
    the tensor costs about the same for 17 columns as for 28, so fuller bins mean fewer passes; past 28 columns the fiber pass
    loses its shared-memory running sums and a bin costs more per column)."""
    bins = []
    for unit in sorted(runs, key=lambda u: (-u[0], u[1])):
        for b in bins:
            if b[0] + unit[0] <= capacity and len(b[1]) < 32:
                b[0] += unit[0]
                b[1].append(unit)
                break
        else:
            bins.append([unit[0], [unit]])
    return [b[1] for b in bins]


def non_negative_parafac(tensor, rank, n_iter_max=100, init="svd", svd="truncated_svd", tol=10e-7, random_state=None,
                         verbose=0, normalize_factors=False, return_errors=False, mask=None,
                         cvg_criterion="abs_rec_error", fixed_modes=None):
    """tensorly ``non_negative_parafac`` (multiplicative updates) on the device."""
    return _factorize(tensor, rank, "non_negative_cp", n_iter_max, init, svd, tol, random_state, verbose, return_errors,
                      mask, cvg_criterion, normalize_factors, fixed_modes)


def non_negative_parafac_hals(tensor, rank, n_iter_max=100, init="svd", svd="truncated_svd", tol=10e-8, random_state=None,
                              sparsity_coefficients=None, fixed_modes=None, nn_modes="all", exact=False,
                              normalize_factors=False, verbose=False, return_errors=False, cvg_criterion="abs_rec_error"):
    """tensorly ``non_negative_parafac_hals`` (HALS column sweeps, exact=False) on the device."""
    if sparsity_coefficients is not None and any(s for s in sparsity_coefficients):
        raise NotImplementedError("sparsity_coefficients are not supported by the B200 path")
    if nn_modes not in ("all", None) and set(nn_modes) != set(range(np.ndim(tensor))):
        raise NotImplementedError("nn_modes other than 'all' is not supported by the B200 path")
    if exact:
        raise NotImplementedError("exact=True (active-set NNLS) is not supported by the B200 path")
    # tensorly's HALS variant always returns (cp, errors) to cell2cell (factorization.py:107)
    cp, errors = _factorize(tensor, rank, "non_negative_cp_hals", n_iter_max, init, svd, tol, random_state, verbose, True,
                            None, cvg_criterion, normalize_factors, fixed_modes)
    return (cp, errors) if return_errors else cp


def _khatri_rao_rows(factors, rank):
    """[prod I_n, rank] Khatri-Rao rows of packed ``factors`` (first factor varies slowest, the tensor's C order)."""
    kr = factors[0][:, :rank]
    for f in factors[1:]:
        kr = (kr[:, None, :] * f[None, :, :rank]).reshape(-1, rank)
    return kr


def parafac(tensor, rank, n_iter_max=100, init="svd", svd="truncated_svd", normalize_factors=False, orthogonalise=False,
            tol=1e-8, random_state=None, verbose=0, return_errors=False, sparsity=None, l2_reg=0, mask=None,
            cvg_criterion="abs_rec_error", fixed_modes=None, svd_mask_repeats=5, linesearch=False, callback=None):
    """tensorly ``parafac`` (alternating least squares, factors of any sign) on the device -- reached from cell2cell with
    ``tf_type='parafac'`` (factorization.py:116-126).

    The two streaming passes of the NCP path do the tensor work: ``T`` (every plane against the two trailing factors) gives
    the MTTKRPs of the leading modes, ``U`` (every in-plane position against the leading Khatri-Rao rows) those of the two
    trailing modes -- two reads of the tensor per iteration.  The ``R x R`` normal equations are solved in fp64 on the device
    (``torch.linalg.solve``, cuSOLVER).  With a mask, tensorly re-imputes the missing entries once per iteration from the
    factors of that iteration (``error_calc``); that imputed copy is materialised here (a cuBLAS product of the leading and
    trailing Khatri-Rao rows and one extra write of the tensor per iteration).  Options tensorly offers beyond cell2cell's call are rejected loudly."""
    for name, val in (("normalize_factors", normalize_factors), ("orthogonalise", orthogonalise), ("sparsity", sparsity),
                      ("fixed_modes", fixed_modes), ("linesearch", linesearch), ("callback", callback)):
        if val:
            raise NotImplementedError("parafac(%s=%r) is not supported by the B200 path (cell2cell never sets it)" % (name, val))
    if cvg_criterion not in ("abs_rec_error", "rec_error"):
        raise TypeError("Unknown convergence criterion")
    rank = int(rank)
    X = as_device_tensor(tensor)
    M = as_device_mask(mask, X)
    if M is not None and isinstance(init, str) and init == "svd":
        raise NotImplementedError("parafac(init='svd') with a mask (svd_mask_repeats) is not supported by the B200 path; "
                                  "cell2cell forces init='random' for masked tensors (factorization.py:118)")
    if X.dim() < 3:
        raise C2CError("the B200 path factorizes tensors of order >= 3")
    host = initialize_factors(X, rank, init=init, svd=svd, random_state=random_state, non_negative=False)
    factors = [engine.pack_factor(f.astype(np.float32), rank, X.device) for f in host]
    nd, shape = X.dim(), tuple(X.shape)
    ws = engine.Workspace(shape, rank, False, X.device)
    gws = engine.Workspace(shape, rank, False, X.device)   # keeps the factor Grams between the calls below
    nx2 = torch.tensor(getattr(X, "_c2c_norm_x2", None) or engine.sumsq(X), dtype=torch.float64, device=X.device)
    ridge = float(l2_reg) * torch.eye(rank, dtype=torch.float64, device=X.device)
    Xc = X                                                   # masked: the re-imputed copy of the current iteration
    keep = None if M is None else M.bool()
    errors = []
    converged = False
    # one factor changes per update: only its Gram is recomputed, and the same call yields the next mode's Hadamard product
    G, rn2 = engine.gram_hadamard(factors, shape, rank, skip_mode=0, ws=gws, as_tensor=True)
    for it in range(int(n_iter_max)):
        TU = None
        for mode in range(nd):
            if mode == 0:
                TU = engine.plane_contract(Xc, None, factors, rank, ws=ws)
            elif mode == nd - 2:
                TU = engine.fiber_contract(Xc, None, factors, rank, ws=ws)
            Mk = engine.mttkrp_from_contraction(TU, shape, factors, rank, mode)
            new = torch.linalg.solve(G[:rank, :rank].to(torch.float64).T + ridge, Mk[:, :rank].to(torch.float64).T).T
            factors[mode][:, :rank] = new.to(torch.float32)
            G, rn2 = engine.gram_hadamard(factors, shape, rank, skip_mode=(mode + 1) % nd, ws=gws, as_tensor=True,
                                          changed_mode=mode)
        if keep is not None:
            rec = (_khatri_rao_rows(factors[:-2], rank) @ _khatri_rao_rows(factors[-2:], rank).T).view(shape)
            Xc = torch.where(keep, X, rec)
            del rec
            nx2 = torch.tensor(engine.sumsq(Xc), dtype=torch.float64, device=X.device)
        iprod = (Mk[:, :rank].double() * factors[-1][:, :rank].double()).sum()
        errors.append(float((torch.sqrt(torch.abs(nx2 + rn2[0] - 2.0 * iprod)) / torch.sqrt(nx2)).item()))
        if verbose:
            print("iteration {}, reconstruction error: {}".format(it, errors[-1]))
        if tol and it >= 1:
            dec = errors[-2] - errors[-1]
            if (abs(dec) < tol) if cvg_criterion == "abs_rec_error" else (dec < tol):
                converged = True
                break
    if verbose and converged:
        print("PARAFAC converged after {} iterations".format(len(errors)))
    sums = None
    if M is None and errors:
        # as for the NCP path: the kept (last) error is the directly accumulated ||X - rec|| / ||X|| (fp64 sums)
        sums = engine.recon_sums(X, None, factors, rank)
        errors[-1] = float(np.sqrt(sums[1]) / np.sqrt(sums[3]))
    cp = CPTensor(np.ones(rank, dtype=np.float32), [engine.unpack_factor(f, rank).cpu().numpy() for f in factors],
                  device=X.device)
    if sums is not None:
        cp.recon_sums_ = sums
    if return_errors:
        return cp, [np.float64(e) for e in errors]
    return cp


_ADMM_UNSUPPORTED = ("l2_reg", "unimodality", "normalize", "simplex", "normalized_sparsity", "soft_sparsity", "smoothness",
                     "monotonicity", "hard_sparsity")


def _mode_constraint(mode, non_negative, l1_reg, l2_square_reg):
    """tensorly ``proximal_operator``'s per-mode selection: scalar = every mode, dict / list = the modes named; one per mode."""
    found = []
    for name, val in (("non_negative", non_negative), ("l1_reg", l1_reg), ("l2_square_reg", l2_square_reg)):
        if val is None:
            continue
        v = val.get(mode) if isinstance(val, dict) else (val[mode] if isinstance(val, (list, tuple)) else val)
        if v is None or v is False:
            continue
        found.append((name, 0.0 if name == "non_negative" else float(v)))
    if len(found) > 1:
        raise ValueError("You selected two constraints for the same mode. Consider to check your input")
    return found[0] if found else (None, 0.0)


def constrained_parafac(tensor, rank, n_iter_max=100, n_iter_max_inner=10, init="svd", svd="truncated_svd", tol_outer=1e-8,
                        tol_inner=1e-6, random_state=None, verbose=0, return_errors=False, cvg_criterion="abs_rec_error",
                        fixed_modes=None, non_negative=None, l1_reg=None, l2_reg=None, l2_square_reg=None, unimodality=None,
                        normalize=None, simplex=None, normalized_sparsity=None, soft_sparsity=None, smoothness=None,
                        monotonicity=None, hard_sparsity=None):
    """tensorly ``constrained_parafac`` (AO-ADMM) on the device -- reached from cell2cell with
    ``tf_type='constrained_parafac'`` (factorization.py:128-136; no mask and no ``tol`` are passed on that path).

    The tensor work is the two streaming passes of the NCP path (``T`` serves the leading modes, ``U`` the two trailing
    ones: two reads of the tensor per outer iteration); a mode's whole inner ADMM loop -- Cholesky of ``G + rho I``, the
    row solves, the proximal step, the dual update and the stop rule -- is ONE single-CTA kernel (``c2c_admm_update``,
    csrc/c2c_admm.cu) with the factor, dual and split variables kept in fp64 on the device.  Constraints: ``non_negative``,
    ``l1_reg``, ``l2_square_reg`` (scalar, dict or list per mode, one per mode) or none; the rest of tensorly's proximal
    family raises ``NotImplementedError``."""
    given = dict(l2_reg=l2_reg, unimodality=unimodality, normalize=normalize, simplex=simplex,
                 normalized_sparsity=normalized_sparsity, soft_sparsity=soft_sparsity, smoothness=smoothness,
                 monotonicity=monotonicity, hard_sparsity=hard_sparsity)
    for name in _ADMM_UNSUPPORTED:
        if given[name] is not None:
            raise NotImplementedError("constrained_parafac(%s=...) is not supported by the B200 path (non_negative, l1_reg, "
                                      "l2_square_reg are)" % name)
    if fixed_modes:
        raise NotImplementedError("constrained_parafac(fixed_modes=...) is not supported by the B200 path")
    if cvg_criterion not in ("abs_rec_error", "rec_error"):
        raise TypeError("Unknown convergence criterion")
    rank = int(rank)
    X = as_device_tensor(tensor)
    if X.dim() < 3:
        raise C2CError("the B200 path factorizes tensors of order >= 3")
    nd, shape, dev = X.dim(), tuple(X.shape), X.device
    cons = [_mode_constraint(m, non_negative, l1_reg, l2_square_reg) for m in range(nd)]
    if isinstance(init, str) and init == "svd":
        # initialize_constrained_parafac: left singular vectors, mode 0 scaled by the singular values -- what svd_factors
        # returns with non_negative=False
        host = initialize_factors(X, rank, init="svd", svd=svd, random_state=random_state, non_negative=False)
    else:
        host = initialize_factors(X, rank, init=init, svd=svd, random_state=random_state, non_negative=False)

    def prox_host(f, name, param):
        if name == "non_negative":
            return np.clip(f, 0, None)
        if name == "l1_reg":
            return np.sign(f) * np.maximum(np.abs(f) - param, 0)
        if name == "l2_square_reg":
            return f / (1 + 2 * param)
        return f
    host = [prox_host(np.asarray(f, dtype=np.float64), *cons[m]) for m, f in enumerate(host)]
    factors = [engine.pack_factor(f.astype(np.float32), rank, dev) for f in host]
    x64 = [torch.as_tensor(np.ascontiguousarray(f)).to(dev) for f in host]
    duals = [torch.zeros_like(f) for f in x64]
    auxs = [torch.zeros_like(f) for f in x64]
    ws = engine.Workspace(shape, rank, False, dev)
    gws = engine.Workspace(shape, rank, False, dev)
    nx2 = torch.tensor(getattr(X, "_c2c_norm_x2", None) or engine.sumsq(X), dtype=torch.float64, device=dev)
    cerr = torch.zeros((nd, 2), dtype=torch.float64, device=dev)
    errors = []
    G, rn2 = engine.gram_hadamard(factors, shape, rank, skip_mode=0, ws=gws, as_tensor=True)
    for it in range(int(n_iter_max)):
        TU = Mk = None
        for mode in range(nd):
            if mode == 0:
                TU = engine.plane_contract(X, None, factors, rank, ws=ws)
            elif mode == nd - 2:
                TU = engine.fiber_contract(X, None, factors, rank, ws=ws)
            Mk = engine.mttkrp_from_contraction(TU, shape, factors, rank, mode)
            engine.admm_update(factors[mode], x64[mode], duals[mode], auxs[mode], Mk, G, rank, constraint=cons[mode][0],
                               param=cons[mode][1], n_inner=n_iter_max_inner, tol=tol_inner, out=cerr[mode])
            G, rn2 = engine.gram_hadamard(factors, shape, rank, skip_mode=(mode + 1) % nd, ws=gws, as_tensor=True,
                                          changed_mode=mode)
        iprod = (Mk[:, :rank].double() * x64[-1]).sum()
        both = torch.stack([torch.sqrt(torch.abs(nx2 + rn2[0] - 2.0 * iprod)) / torch.sqrt(nx2), cerr[:, 0].sum()]).cpu()
        errors.append(float(both[0]))
        if verbose:
            print("iteration {}, reconstruction error: {}, constraints error: {}".format(it, errors[-1], float(both[1])))
        if tol_outer and it >= 1:
            dec = errors[-2] - errors[-1]
            if float(both[1]) < tol_outer:
                break
            if (abs(dec) < tol_outer) if cvg_criterion == "abs_rec_error" else (dec < tol_outer):
                break
    cp = CPTensor(np.ones(rank, dtype=np.float32), [f.cpu().numpy() for f in x64], device=dev)
    if return_errors:
        return cp, [np.float64(e) for e in errors]
    return cp


def _compute_tensor_factorization(tensor, rank, tf_type="non_negative_cp", init="svd", svd="truncated_svd",
                                  random_state=None, mask=None, n_iter_max=100, tol=10e-7, verbose=False, **kwargs):
    """Same contract as the reference's ``_compute_tensor_factorization`` (factorization.py:14-143)."""
    return_errors = bool(kwargs.get("return_errors", False)) if kwargs else False
    kwargs = dict(kwargs or {})
    if tf_type == "non_negative_cp":
        kwargs.pop("return_errors", None)
        cp_tf = non_negative_parafac(tensor=tensor, rank=rank, init="random" if mask is not None else init, svd=svd,
                                     random_state=random_state, mask=mask, n_iter_max=n_iter_max, tol=tol,
                                     verbose=verbose, return_errors=True, **kwargs)
        cp_tf, errors = cp_tf
    elif tf_type == "non_negative_cp_hals":
        kwargs.pop("return_errors", None)
        cp_tf, errors = non_negative_parafac_hals(tensor=tensor, rank=rank, init=init, svd=svd, random_state=random_state,
                                                  n_iter_max=n_iter_max, tol=tol, verbose=verbose, return_errors=True,
                                                  **kwargs)
    elif tf_type == "parafac":
        kwargs.pop("return_errors", None)
        cp_tf, errors = parafac(tensor=tensor, rank=rank, init="random" if mask is not None else init, svd=svd,
                                random_state=random_state, mask=mask, n_iter_max=n_iter_max, tol=tol, verbose=verbose,
                                return_errors=True, **kwargs)
    elif tf_type == "constrained_parafac":
        # factorization.py:128-136: neither `mask` nor `tol` reach tensorly on this path (tol_outer keeps its default)
        kwargs.pop("return_errors", None)
        cp_tf, errors = constrained_parafac(tensor=tensor, rank=rank, init="random" if mask is not None else init, svd=svd,
                                            random_state=random_state, n_iter_max=n_iter_max, verbose=verbose,
                                            return_errors=True, **kwargs)
    else:
        raise ValueError("Not a valid tf_type.")
    if return_errors:
        return cp_tf, errors
    return cp_tf


def _compute_elbow(loss, S=1.0, curve="convex", direction="decreasing", interp_method="interp1d"):
    """x coordinate of the elbow of ``loss`` = [(x, y), ...] (factorization.py:146-183, kneed's KneeLocator)."""
    x, y = zip(*loss)
    return kneedle_elbow(x, y, S=S, curve=curve, direction=direction, interp_method=interp_method)


def _loss_of(tensor, tl_object, errors, mask):
    if mask is None:
        return np.float64(errors[-1])
    return _compute_norm_error(tensor, tl_object, mask)


def _run_elbow_analysis(tensor, upper_rank=50, tf_type="non_negative_cp", init="svd", svd="truncated_svd",
                        random_state=None, mask=None, n_iter_max=100, tol=10e-7, verbose=False, disable_pbar=False,
                        **kwargs):
    """One factorization per rank 1..upper_rank; list of (rank, loss) (factorization.py:186-269)."""
    kwargs = dict(kwargs or {})
    kwargs["return_errors"] = True
    X = as_device_tensor(tensor)
    M = as_device_mask(mask, X)
    loss = []
    for r in tqdm(range(1, upper_rank + 1), disable=disable_pbar):
        tl_object, errors = _compute_tensor_factorization(X, rank=r, tf_type=tf_type, init=init, svd=svd,
                                                          random_state=random_state, mask=M, n_iter_max=n_iter_max,
                                                          tol=tol, verbose=verbose, **kwargs)
        loss.append((r, _loss_of(X, tl_object, errors, M)))
    return loss


def _multiple_runs_elbow_analysis(tensor, upper_rank=50, runs=10, tf_type="non_negative_cp", init="svd",
                                  svd="truncated_svd", metric="error", random_state=None, mask=None, n_iter_max=100,
                                  tol=10e-7, verbose=False, batch_runs=True, host_threads=None, **kwargs):
    """``runs`` factorizations per rank; ndarray (runs x upper_rank) of losses (factorization.py:272-373).

    The ``upper_rank * runs`` units are independent; when ``torch.distributed`` is initialised they are spread over the
    ranks (one process per GPU, no data-path collective -- see ``cell2cell_b200.parallel``)."""
    assert isinstance(runs, int), "runs must be an integer"
    assert metric in ("similarity", "error"), "`metric` must be either 'similarity' or 'error'"
    from ..parallel import run_units

    kwargs = dict(kwargs or {})
    kwargs["return_errors"] = True
    X = as_device_tensor(tensor)
    M = as_device_mask(mask, X)
    # Runs of one rank are independent; as many as fit 32 factor columns are factorized side by side, so one read of the
    # tensor per pass serves all of them (MU, unmasked, host-seeded random init: what the elbow analysis uses by default).
    batchable = (tf_type == "non_negative_cp" and M is None and init == "random" and random_state is not None and
                 set(kwargs) <= {"return_errors", "cvg_criterion"} and X.dim() >= 3 and batch_runs)
    all_units = [(r, run) for r in range(1, upper_rank + 1) for run in range(runs)]
    multi = batchable and upper_rank <= 32 and _multi_eligible(X, [r for r, _ in all_units], n_iter_max)
    if multi:
        # small tensor: every run is one CTA of a single launch; with several GPUs each takes a cost-balanced share
        from ..parallel import world, lpt_schedule
        _, size = world()
        owner = lpt_schedule([engine.ldr_of(r) for r, _ in all_units], size)
        units = [[u for u, w in zip(all_units, owner) if w == k] for k in range(size)]
        units = [u for u in units if u]
    elif batchable:
        units = _pack_runs([u for u in all_units if u[0] <= 32]) + [[u] for u in all_units if u[0] > 32]
    else:
        units = [[u] for u in all_units]

    def single(r, run):
        seed = None if random_state is None else random_state + run
        return _compute_tensor_factorization(X, rank=r, tf_type=tf_type, init=init, svd=svd, random_state=seed, mask=M,
                                             n_iter_max=n_iter_max, tol=tol, verbose=verbose, **kwargs)

    def one(unit):
        res = None
        if multi:
            try:
                res = _factorize_multi(X, [(r, random_state + run) for r, run in unit], n_iter_max, tol,
                                       cvg_criterion=kwargs.get("cvg_criterion", "abs_rec_error"))
            except C2CError as err:
                if getattr(err, "code", 0) != -61:
                    raise
        elif len(unit) > 1:
            try:
                res = _factorize_batch(X, [(r, random_state + run) for r, run in unit], n_iter_max, tol,
                                       cvg_criterion=kwargs.get("cvg_criterion", "abs_rec_error"))
            except C2CError:
                res = None                       # shape outside the fused update kernels: run by run on the device instead
        if res is None:
            res = [single(r, run) for r, run in unit]
        if metric == "error":
            return [float(_loss_of(X, tl, errs, M)) for tl, errs in res]
        return [tl for tl, errs in res]

    def cost(unit):
        cols = sum(r for r, _ in unit)
        if multi:
            return float(sum(engine.ldr_of(r) for r, _ in unit))
        return (1.0 if cols <= 12 else (1.2 if cols <= 16 else 1.45)) if len(unit) > 1 or cols <= 32 else cols / 16.0

    if host_threads is None:
        # host threads issuing this rank's units, each on a CUDA stream of its own: the host part of one bin of runs (factor
        # init, upload, graph capture, read-back) overlaps the device part of another (parallel._run_local).  Measured on the
        # atlas sweep (117 bins): 5.71 s with 1 thread, 5.47 with 2, 5.40 with 3 (profiles/r02_elbow_host_threads.jsonl); the
        # one-launch small-tensor path has nothing to overlap.
        host_threads = int(os.environ.get("C2C_B200_HOST_THREADS", "1" if multi else "2"))
    parts = run_units(units, one, cost=cost, gather="object", threads=host_threads)
    flat = {}
    for unit, vals in zip(units, parts):
        for key, v in zip(unit, vals):
            flat[key] = v
    if metric == "error":
        all_loss = np.asarray([[flat[(r, run)] for run in range(runs)] for r in range(1, upper_rank + 1)], dtype=np.float64)
    else:
        from .metrics import pairwise_correlation_index
        from scipy.spatial.distance import squareform
        rows = []
        for r in range(1, upper_rank + 1):
            fs = [flat[(r, run)].factors for run in range(runs)]
            rows.append((1.0 - squareform(pairwise_correlation_index(fs))).tolist())
        all_loss = np.asarray(rows, dtype=np.float64)
    return all_loss.T


def _compute_norm_error(tensor, tl_object, mask=None):
    """``||M*X - M*rec|| / ||M*X||`` (factorization.py:376-432) in one fused pass, without materialising ``rec``."""
    X = as_device_tensor(tensor)
    M = as_device_mask(mask, X)
    weights, factors = tl_object
    rank = int(np.asarray(factors[0]).shape[1])
    fs = []
    for i, f in enumerate(factors):
        f = np.asarray(f, dtype=np.float64)
        if i == 0 and weights is not None:
            f = f * np.asarray(weights, dtype=np.float64)[None, :]
        fs.append(engine.pack_factor(f.astype(np.float32), rank, X.device))
    s = engine.recon_sums(X, M, fs, rank)
    return np.float64(np.sqrt(s[1]) / np.sqrt(s[3]))


def normalized_error(reference_tensor, reconstructed_tensor):
    """``||ref - rec|| / ||ref||`` of two full tensors (factorization.py:435-458); plain device arithmetic."""
    a = as_device_tensor(reference_tensor)
    b = as_device_tensor(reconstructed_tensor, a.device)
    return float(torch.linalg.vector_norm((a - b).double()) / torch.linalg.vector_norm(a.double()))


def require_device():
    """Raises unless a CUDA device and the native library are usable (no CPU fallback)."""
    from .._lib import lib
    lib()
    return engine.require_cuda()


_ = C2CError

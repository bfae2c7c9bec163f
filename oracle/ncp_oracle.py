"""CPU oracle for the factorization stage: a numpy (float64) restatement of tensorly 0.8.x semantics.

TEST INFRASTRUCTURE ONLY -- imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, never by the product package ``cell2cell_b200``.

PARITY UNPINNED: the arithmetic of this stage lives in **tensorly** (PyPI, un-vendored and un-pinned in the reference:
``setup.py:101``; the release-note trail ``release/0.6.7-notes.md:10`` / ``release/0.8.0-notes.md:9-10`` and the import of
``tensorly.decomposition._cp.error_calc`` at ``cell2cell/tensor/coupled_factorization.py:6`` put it at tensorly >= 0.8.0).
tensorly is not installed in this image and is not in /opt/wheelhouse, and the reference repository holds no test,
golden vector or fixture for factorization outputs.  What anchors this restatement:

* the reference's call sites: ``cell2cell/tensor/factorization.py:91-104`` (non_negative_parafac; ``init='random'`` forced
  whenever a mask is given, ``:94``), ``:106-115`` (non_negative_parafac_hals), ``cell2cell/tensor/tensor.py:294-362``
  (best-of-runs, cp_normalize, ordering) and ``cell2cell/tensor/factorization.py:376-458`` (_compute_norm_error);
* the reference's **in-repo, line-by-line mirror of tensorly's multiplicative update**:
  ``cell2cell/tensor/coupled_factorization.py:336-348`` (Gram-Hadamard ``accum`` with ``w w^T`` scaling, masked imputation
  ``tensor*mask + cp_to_tensor(..., mask=1-mask)``, ``unfolding_dot_khatri_rao``, ``clip(., eps)``, ``A*num/den``) and of
  its stop rule ``:437-458``;
* the published algorithm of tensorly 0.8.1 ``tensorly/decomposition/_nn_cp.py`` (non_negative_parafac,
  non_negative_parafac_hals), ``_cp.py`` (initialize_cp), ``cp_tensor.py`` (cp_normalize, cp_norm, cp_to_tensor),
  ``tenalg/svd.py`` (svd_flip, make_svd_non_negative 'nndsvd'), ``solvers/nnls.py`` (hals_nnls), restated from the
  published source (not available offline -- re-verify if a tensorly wheel ever appears).

Self-consistency properties checked in ``tests/test_ncp_oracle.py``: MU error sequence is monotone non-increasing
(unmasked), factors stay >= 0, the Gram-form error equals the directly computed ||X-rec||/||X|| to 1e-12, a fixed
point (exact low-rank tensor, true factors as init) stays a fixed point, ``unfolding_dot_khatri_rao`` equals an einsum.
"""
import numpy as np

__all__ = [
    "CPResult", "random_init", "svd_init", "initialize_cp", "mttkrp", "cp_to_tensor", "cp_norm", "cp_normalize",
    "non_negative_parafac", "non_negative_parafac_hals", "hals_nnls", "masked_norm_error", "unfold", "parafac",
    "constrained_parafac", "admm",
]


class CPResult(tuple):
    """(weights, factors) pair with a ``to_tensor`` method -- the part of tensorly's CPTensor the reference uses
    (``tensor.py:355`` unpacks it, ``tensor.py:670`` / ``factorization.py:408`` call ``.to_tensor()``)."""

    def __new__(cls, weights, factors):
        return super().__new__(cls, (weights, factors))

    @property
    def weights(self):
        return self[0]

    @property
    def factors(self):
        return self[1]

    def to_tensor(self):
        return cp_to_tensor(self[0], self[1])


# --------------------------------------------------------------------------------------------------------------------
# basic tensor algebra (tensorly.base.unfold, tenalg.khatri_rao / unfolding_dot_khatri_rao, cp_tensor.*)
# --------------------------------------------------------------------------------------------------------------------
def unfold(x, mode):
    """tensorly.base.unfold: mode-n unfolding, remaining modes in C order."""
    return np.reshape(np.moveaxis(x, mode, 0), (x.shape[mode], -1))


def cp_to_tensor(weights, factors, mask=None):
    """tensorly.cp_tensor.cp_to_tensor; ``mask`` multiplies the reconstruction element-wise (tensorly passes it into
    the Khatri-Rao product, which is the same thing)."""
    letters = "abcdefgh"[: len(factors)]
    w = np.ones(factors[0].shape[1]) if weights is None else weights
    expr = ",".join(l + "r" for l in letters) + ",r->" + letters
    full = np.einsum(expr, *factors, w, optimize=True)
    if mask is not None:
        full = full * mask
    return full


def mttkrp(x, weights, factors, mode):
    """tensorly.cp_tensor.unfolding_dot_khatri_rao: for r in range(R): multi_mode_dot(x, [f[:, r]], skip=mode),
    stacked on axis 1, times the weights.  Restated as R sequential chains of mode-vector products (that is what makes
    the reference read the tensor R times per mode -- kept on purpose: this function is also the timed CPU baseline)."""
    rank = factors[0].shape[1]
    cols = []
    for r in range(rank):
        y = x
        # contract the last mode first: every mode below k is still present, so mode k still sits at axis k
        for k in range(x.ndim - 1, -1, -1):
            if k != mode:
                y = np.tensordot(y, factors[k][:, r], axes=([k], [0]))
        cols.append(y)
    out = np.stack(cols, axis=1)
    if weights is not None:
        out = out * weights.reshape(1, -1)
    return out


def cp_norm(weights, factors):
    """tensorly.cp_tensor.cp_norm: sqrt(sum_{r,s} w_r w_s prod_n (A_n^T A_n)[r,s])."""
    rank = factors[0].shape[1]
    g = np.ones((rank, rank))
    for f in factors:
        g = g * (f.T @ f)
    if weights is not None:
        g = g * (weights.reshape(-1, 1) * weights.reshape(1, -1))
    return np.sqrt(np.sum(g))


def cp_normalize(weights, factors):
    """tensorly.cp_tensor.cp_normalize (called at ``tensor.py:327``): fold the weights into factor 0, then give every
    column unit l2 norm, collecting the norms in the weights; a zero column keeps scale 1."""
    rank = factors[0].shape[1]
    w = np.ones(rank) if weights is None else np.array(weights, dtype=float)
    out = []
    for i, f in enumerate(factors):
        f = np.array(f, dtype=float)
        if i == 0:
            f = f * w
            w = np.ones(rank)
        s = np.sqrt(np.sum(f * f, axis=0))
        w = w * s
        out.append(f / np.where(s == 0, 1.0, s))
    return CPResult(w, out)


# --------------------------------------------------------------------------------------------------------------------
# initialisation (tensorly.decomposition._cp.initialize_cp, tensorly.random.random_cp, tenalg.svd)
# --------------------------------------------------------------------------------------------------------------------
def _rng(random_state):
    """tensorly.check_random_state: None -> numpy's global RandomState, int -> RandomState(seed)."""
    if random_state is None:
        return np.random.mtrand._rand
    if isinstance(random_state, np.random.RandomState):
        return random_state
    return np.random.RandomState(int(random_state))


def random_init(shape, rank, random_state=None):
    """tensorly.random.random_cp(normalise_factors=False): one ``random_sample((I_n, R))`` per mode, in mode order, from
    the same RandomState; weights = ones."""
    rng = _rng(random_state)
    return [rng.random_sample((int(s), int(rank))) for s in shape]


def _svd_flip(u, v):
    """tensorly.tenalg.svd.svd_flip (u_based_decision=True): make the largest-|.| entry of every column of U positive."""
    idx = np.argmax(np.abs(u), axis=0)
    signs = np.sign(u[idx, np.arange(u.shape[1])])
    signs = np.where(signs == 0, 1.0, signs)
    return u * signs, v * signs[:, None]


def _nndsvd(u, s, v, eps):
    """tensorly.tenalg.svd.make_svd_non_negative(nntype='nndsvd') (Boutsidis & Gallopoulos 2008): returns W only."""
    w = np.zeros_like(u)
    h = np.zeros_like(v)
    w[:, 0] = np.sqrt(s[0]) * np.abs(u[:, 0])
    h[0, :] = np.sqrt(s[0]) * np.abs(v[0, :])
    for j in range(1, len(s)):
        a, b = u[:, j], v[j, :]
        a_p, b_p = np.maximum(a, 0), np.maximum(b, 0)
        a_n, b_n = np.abs(np.minimum(a, 0)), np.abs(np.minimum(b, 0))
        a_p_nrm, b_p_nrm = np.linalg.norm(a_p), np.linalg.norm(b_p)
        a_n_nrm, b_n_nrm = np.linalg.norm(a_n), np.linalg.norm(b_n)
        m_p, m_n = a_p_nrm * b_p_nrm, a_n_nrm * b_n_nrm
        if m_p > m_n:
            uu, vv, sigma = a_p / a_p_nrm, b_p / b_p_nrm, m_p
        else:
            uu, vv, sigma = a_n / a_n_nrm, b_n / b_n_nrm, m_n
        lbd = np.sqrt(s[j] * sigma)
        w[:, j] = lbd * uu
        h[j, :] = lbd * vv
    return np.sign(w) * np.maximum(np.abs(w) - eps, 0.0)  # tensorly soft_thresholding(W, eps)


def svd_init(x, rank, random_state=None, non_negative=True):
    """initialize_cp(init='svd', svd='truncated_svd', non_negative=True): per mode, truncated SVD of the unfolding,
    svd_flip, NNDSVD, scale mode 0 by the singular values, pad with random columns when I_n < R.  ``non_negative=False``
    (plain ``parafac``) keeps the flipped left singular vectors as they are."""
    rng = _rng(random_state)
    eps = np.finfo(x.dtype).eps
    factors = []
    for mode in range(x.ndim):
        m = unfold(x, mode)
        k = min(rank, min(m.shape))
        u, s, vt = np.linalg.svd(m, full_matrices=False)
        u, s, vt = u[:, :k], s[:k], vt[:k, :]
        u, vt = _svd_flip(u, vt)
        w = _nndsvd(u, s, vt, eps) if non_negative else u
        if mode == 0:
            w = w * s[None, :k]
        if x.shape[mode] < rank:
            w = np.concatenate([w, rng.random_sample((w.shape[0], rank - x.shape[mode]))], axis=1)
        factors.append(w[:, :rank])
    return factors


def initialize_cp(x, rank, init="svd", random_state=None):
    """tensorly initialize_cp(..., non_negative=True, normalize_factors=False) -> list of |factors| (weights = ones)."""
    if isinstance(init, str):
        if init == "random":
            factors = random_init(x.shape, rank, random_state)
        elif init == "svd":
            factors = svd_init(x, rank, random_state)
        else:
            raise ValueError('Initialization method "%s" not recognized' % init)
    else:
        w, fs = init
        fs = [np.array(f, dtype=float) for f in fs]
        if w is not None and not np.all(np.asarray(w) == 1):
            w = np.asarray(w, dtype=float)
            w_avg = np.prod(w) ** (1.0 / w.shape[0])
            fs = [f * w_avg for f in fs]
        factors = fs
    return [np.abs(f) for f in factors]


# --------------------------------------------------------------------------------------------------------------------
# non_negative_parafac (multiplicative updates) -- tensorly/decomposition/_nn_cp.py, mirrored in the reference at
# cell2cell/tensor/coupled_factorization.py:336-348 (update) and :437-458 (stop rule)
# --------------------------------------------------------------------------------------------------------------------
def non_negative_parafac(tensor, rank, n_iter_max=100, init="svd", tol=10e-7, random_state=None, mask=None,
                         cvg_criterion="abs_rec_error", dtype=np.float64, return_errors=True, return_iters=False):
    x = np.array(tensor, dtype=dtype)
    eps = np.finfo(x.dtype).eps
    factors = [f.astype(dtype) for f in initialize_cp(x, rank, init=init, random_state=random_state)]
    weights = np.ones(rank, dtype=dtype)
    norm_x = np.sqrt(np.sum(x.astype(np.float64) ** 2)).astype(dtype)  # of the tensor as passed (missing entries = 0)
    if mask is not None:
        mask = np.asarray(mask).astype(dtype)
    errors = []
    n_modes = x.ndim
    it_done = 0
    for it in range(n_iter_max):
        for mode in range(n_modes):
            accum = np.ones((rank, rank), dtype=dtype)
            for k in range(n_modes):
                if k != mode:
                    accum = accum * (factors[k].T @ factors[k])
            accum = weights.reshape(-1, 1) * accum * weights.reshape(1, -1)
            if mask is not None:
                # the tensor is overwritten: observed entries persist, missing ones are re-imputed with the CURRENT factors
                x = x * mask + cp_to_tensor(weights, factors, mask=1 - mask)
            m = mttkrp(x, weights, factors, mode)
            num = np.clip(m, eps, None)
            den = np.clip(factors[mode] @ accum, eps, None)
            factors[mode] = factors[mode] * num / den
        it_done = it + 1
        if tol:
            rec_norm = cp_norm(weights, factors)
            iprod = np.sum(m * factors[n_modes - 1])  # mttkrp of the last mode (pre-update) x updated last factor
            err = np.sqrt(np.abs(norm_x ** 2 + rec_norm ** 2 - 2 * iprod)) / norm_x
            errors.append(err)
            if it >= 1:
                dec = errors[-2] - errors[-1]
                stop = (abs(dec) < tol) if cvg_criterion == "abs_rec_error" else (dec < tol)
                if stop:
                    break
    res = CPResult(weights, factors)
    out = (res, errors) if return_errors else (res,)
    if return_iters:
        out = out + (it_done,)
    return out if len(out) > 1 else out[0]


def masked_norm_error(tensor, cp, mask=None):
    """``_compute_norm_error`` / ``normalized_error`` (cell2cell/tensor/factorization.py:376-458):
    ||M*X - M*rec|| / ||M*X||."""
    rec = cp_to_tensor(cp[0], cp[1])
    x = np.asarray(tensor, dtype=float)
    if mask is not None:
        x = x * mask
        rec = rec * mask
    return np.sqrt(np.sum((x - rec) ** 2)) / np.sqrt(np.sum(x ** 2))


# --------------------------------------------------------------------------------------------------------------------
# non_negative_parafac_hals -- tensorly/decomposition/_nn_cp.py + tensorly/solvers/nnls.py (hals_nnls)
# --------------------------------------------------------------------------------------------------------------------
def hals_nnls(utm, utu, v, n_iter_max=500, tol=1e-8):
    """tensorly.solvers.nnls.hals_nnls (exact=False, no sparsity / ridge): cyclic non-negative coordinate descent on
    the rows of V (rank x I).  Per sweep, for every k with UtU[k,k] != 0:
    ``delta = max((UtM[k,:] - UtU[k,:] @ V) / UtU[k,k], -V[k,:])``, ``V[k,:] += delta``, ``rec += delta.delta``.
    Stop after a sweep when ``rec < tol * rec_first_sweep`` or ``sweep_index > 1 + 0.5 * complexity_ratio`` with
    ``complexity_ratio = 1 + (|V| + I*rank) / (rank*rank + rank)`` (tensorly's numerator / denominator)."""
    rank, n_col = utm.shape
    v = v.copy()
    rec0 = 0.0
    ratio = 1.0 + (rank * n_col + n_col * rank) / float(rank * rank + rank)
    for it in range(n_iter_max):
        rec = 0.0
        for k in range(rank):
            if utu[k, k] != 0:
                delta = np.maximum((utm[k, :] - utu[k, :] @ v) / utu[k, k], -v[k, :])
                v[k, :] = v[k, :] + delta
                rec += float(delta @ delta)
        if it == 0:
            rec0 = rec
        if rec < tol * rec0 or it > 1 + 0.5 * ratio:
            break
    return v


def non_negative_parafac_hals(tensor, rank, n_iter_max=100, init="svd", tol=10e-8, random_state=None,
                              cvg_criterion="abs_rec_error", dtype=np.float64, inner_iter_max=100):
    """tensorly non_negative_parafac_hals (exact=False, no sparsity): per mode G = Hadamard of the other Grams,
    M = MTTKRP, then ``hals_nnls(M^T, G, A_mode^T, n_iter_max=100)``; error / stop rule as in the MU version."""
    x = np.array(tensor, dtype=dtype)
    factors = [f.astype(dtype) for f in initialize_cp(x, rank, init=init, random_state=random_state)]
    weights = np.ones(rank, dtype=dtype)
    norm_x = np.sqrt(np.sum(x.astype(np.float64) ** 2))
    errors = []
    n_modes = x.ndim
    for it in range(n_iter_max):
        for mode in range(n_modes):
            g = np.ones((rank, rank), dtype=dtype)
            for k in range(n_modes):
                if k != mode:
                    g = g * (factors[k].T @ factors[k])
            m = mttkrp(x, None, factors, mode)
            factors[mode] = hals_nnls(m.T, g, factors[mode].T, n_iter_max=inner_iter_max).T
        if tol:
            rec_norm = cp_norm(weights, factors)
            iprod = np.sum(m * factors[n_modes - 1])
            err = np.sqrt(np.abs(norm_x ** 2 + rec_norm ** 2 - 2 * iprod)) / norm_x
            errors.append(err)
            if it >= 1:
                dec = errors[-2] - errors[-1]
                stop = (abs(dec) < tol) if cvg_criterion == "abs_rec_error" else (dec < tol)
                if stop:
                    break
    return CPResult(weights, factors), errors


# --------------------------------------------------------------------------------------------------------------------
# parafac (alternating least squares) -- tensorly/decomposition/_cp.py, reached from the reference at
# cell2cell/tensor/factorization.py:116-126 (tf_type='parafac'; init='random' forced when a mask is given)
# --------------------------------------------------------------------------------------------------------------------
def parafac(tensor, rank, n_iter_max=100, init="svd", tol=1e-8, random_state=None, mask=None, l2_reg=0.0,
            cvg_criterion="abs_rec_error", dtype=np.float64, return_iters=False):
    """tensorly 0.8.x ``parafac`` with its defaults (no normalisation, orthogonalisation, sparsity, line search, fixed
    modes): per mode ``A_n = solve(G^T, M^T)^T`` with ``G`` the Hadamard product of the other Grams (+ ``l2_reg`` I) and
    ``M`` the MTTKRP; after the last mode ``error_calc``: with a mask the tensor is re-imputed ONCE per iteration
    (``tensor*mask + rec*(1-mask)``, norm recomputed) and the Gram-form error uses the last mode's MTTKRP (taken before
    the re-imputation) -- PARITY UNPINNED like the rest of this file.  Returns ``(CPResult, errors)``."""
    x = np.array(tensor, dtype=dtype)
    n_modes = x.ndim
    if isinstance(init, str):
        factors = random_init(x.shape, rank, random_state) if init == "random" else \
            svd_init(x, rank, random_state, non_negative=False)
    else:
        w0, fs = init
        factors = [np.array(f, dtype=float) for f in fs]
        if w0 is not None and not np.all(np.asarray(w0) == 1):
            w_avg = np.prod(np.asarray(w0, dtype=float)) ** (1.0 / len(w0))
            factors = [f * w_avg for f in factors]
    factors = [f.astype(dtype) for f in factors]
    weights = np.ones(rank, dtype=dtype)
    if mask is not None:
        mask = np.asarray(mask).astype(dtype)
    norm_x = np.sqrt(np.sum(x ** 2))
    eye = np.eye(rank, dtype=dtype) * l2_reg
    errors = []
    it_done = 0
    for it in range(n_iter_max):
        for mode in range(n_modes):
            g = np.ones((rank, rank), dtype=dtype)
            for k in range(n_modes):
                if k != mode:
                    g = g * (factors[k].T @ factors[k])
            g = g + eye
            g = weights.reshape(-1, 1) * g * weights.reshape(1, -1)
            m = mttkrp(x, weights, factors, mode)
            factors[mode] = np.linalg.solve(g.T, m.T).T
        it_done = it + 1
        # error_calc(tensor, norm_tensor, weights, factors, sparsity=None, mask, mttkrp)
        if mask is not None:
            x = x * mask + cp_to_tensor(weights, factors) * (1 - mask)
            norm_x = np.sqrt(np.sum(x ** 2))
        iprod = np.sum(m * factors[n_modes - 1])
        err = np.sqrt(np.abs(norm_x ** 2 + cp_norm(weights, factors) ** 2 - 2 * iprod)) / norm_x
        errors.append(err)
        if tol and it >= 1:
            dec = errors[-2] - errors[-1]
            if (abs(dec) < tol) if cvg_criterion == "abs_rec_error" else (dec < tol):
                break
    out = (CPResult(weights, factors), errors)
    return out + (it_done,) if return_iters else out


# --------------------------------------------------------------------------------------------------------------------
# tensorly.decomposition.constrained_parafac (AO-ADMM) -- reached from cell2cell/tensor/factorization.py:128-136
# --------------------------------------------------------------------------------------------------------------------
def _constraint_of(mode, n_modes, non_negative=None, l1_reg=None, l2_square_reg=None):
    """tensorly ``proximal_operator`` / ``validate_constraints``: a constraint given as a scalar applies to every mode, as a
    dict ``{mode: value}`` or a list (one entry per mode) to the modes it names; at most one constraint per mode."""
    found = []
    for name, val in (("non_negative", non_negative), ("l1_reg", l1_reg), ("l2_square_reg", l2_square_reg)):
        if val is None:
            continue
        if isinstance(val, dict):
            v = val.get(mode)
        elif isinstance(val, (list, tuple)):
            v = val[mode]
        else:
            v = val
        if v is None or v is False:
            continue
        found.append((name, v))
    if len(found) > 1:
        raise ValueError("You selected two constraints for the same mode. Consider to check your input")
    return found[0] if found else (None, None)


def _prox(x, name, param):
    """tensorly.tenalg.proximal: non_negative -> clip(x, 0, max(x)); l1_reg -> soft thresholding; l2_square_reg -> x / (1 + 2 reg)."""
    if name == "non_negative":
        return np.clip(x, 0, None)
    if name == "l1_reg":
        return np.sign(x) * np.maximum(np.abs(x) - param, 0)
    if name == "l2_square_reg":
        return x / (1 + 2 * param)
    return x


def admm(utm, utu, x, dual, n_iter_max, constraint, param, tol):
    """tensorly/solvers/admm.py ``admm`` with ``n_const`` given (the constrained_parafac call)."""
    rho = np.trace(utu) / x.shape[1]
    x_split = None
    for _ in range(n_iter_max):
        x_old = x.copy()
        x_split = np.linalg.solve((utu + rho * np.eye(utu.shape[1])).T, (utm + rho * (x + dual)).T)
        x = _prox(x_split.T - dual, constraint, param)
        dual = dual + x - x_split.T
        dual_residual = x - x_split.T
        primal_residual = x - x_old
        if np.linalg.norm(dual_residual) < tol * np.linalg.norm(x) and np.linalg.norm(primal_residual) < tol * np.linalg.norm(dual):
            break
    return x, x_split, dual


def constrained_parafac(tensor, rank, n_iter_max=100, n_iter_max_inner=10, init="svd", tol_outer=1e-8, tol_inner=1e-6,
                        random_state=None, cvg_criterion="abs_rec_error", non_negative=None, l1_reg=None, l2_square_reg=None):
    """tensorly 0.8.x ``constrained_parafac`` (AO-ADMM) for the constraints the product supports -- PARITY UNPINNED like the
    rest of this file, and restated from memory of the published source: ``initialize_constrained_parafac`` (random: seeded
    ``random_cp``; svd: left singular vectors, mode 0 scaled by the singular values; then the proximal operator of every
    mode), per outer iteration and mode ``admm(mttkrp, hadamard of the other Grams, ...)``, the Gram-form error from the last
    mode's MTTKRP, stop on the constraint error or on the error decrease.  Returns ``(CPResult, errors)``."""
    x = np.asarray(tensor, dtype=np.float64)
    n = x.ndim
    cons = [_constraint_of(m, n, non_negative, l1_reg, l2_square_reg) for m in range(n)]
    if isinstance(init, str):
        if init == "random":
            factors = random_init(x.shape, rank, random_state)
        else:
            rng = _rng(random_state)
            factors = []
            for mode in range(n):
                u, s, vt = np.linalg.svd(unfold(x, mode), full_matrices=False)
                u, _ = _svd_flip(u, vt)
                u = u[:, :rank].copy()
                if mode == 0:
                    k = min(rank, s.shape[0])
                    u[:, :k] = u[:, :k] * s[:k]
                if x.shape[mode] < rank:
                    u = np.concatenate([u, rng.random_sample((u.shape[0], rank - x.shape[mode]))], axis=1)
                factors.append(u[:, :rank])
    else:
        factors = [np.array(f, dtype=np.float64) for f in init[1]]
    factors = [_prox(np.asarray(f, dtype=np.float64), *cons[m]) for m, f in enumerate(factors)]
    norm_x = np.sqrt(np.sum(x ** 2))
    duals = [np.zeros_like(f) for f in factors]
    aux = [np.zeros_like(f).T for f in factors]
    errors = []
    m = None
    for it in range(n_iter_max):
        for mode in range(n):
            g = np.ones((rank, rank))
            for k in range(n):
                if k != mode:
                    g = g * (factors[k].T @ factors[k])
            m = mttkrp(x, None, factors, mode)
            factors[mode], aux[mode], duals[mode] = admm(m, g, factors[mode], duals[mode], n_iter_max_inner, cons[mode][0],
                                                         cons[mode][1], tol_inner)
        iprod = np.sum(m * factors[-1])
        err = np.sqrt(np.abs(norm_x ** 2 + cp_norm(None, factors) ** 2 - 2 * iprod)) / norm_x
        errors.append(err)
        c_err = sum(np.linalg.norm(factors[k] - aux[k].T) / np.linalg.norm(factors[k]) for k in range(n))
        if tol_outer and it >= 1:
            dec = errors[-2] - errors[-1]
            if c_err < tol_outer:
                break
            if (abs(dec) < tol_outer) if cvg_criterion == "abs_rec_error" else (dec < tol_outer):
                break
    return CPResult(np.ones(rank), factors), errors

"""CP result type and factor initialisation (the parts of tensorly the reference's hot path touches).

``CPTensor`` stands in for ``tensorly.cp_tensor.CPTensor``: the reference unpacks it as ``(weights, factors)``
(cell2cell/tensor/tensor.py:343,355) and calls ``.to_tensor()`` (tensor.py:670, factorization.py:408).
``initialize_factors`` follows tensorly ``initialize_cp(..., non_negative=True)`` as called from
``non_negative_parafac`` (reached at cell2cell/tensor/factorization.py:92-101): 'random' draws are host-seeded with
numpy's MT19937 ``RandomState`` so they are bit-identical to the reference's; 'svd' runs on the device (cuSOLVER /
cuBLAS through torch.linalg).
"""
import numpy as np
import torch

__all__ = ["CPTensor", "check_random_state", "random_factors", "svd_factors", "initialize_factors"]


class CPTensor:
    """(weights, factors) with host numpy arrays; iterable like a 2-tuple."""

    def __init__(self, weights, factors, device=None):
        self.weights = np.asarray(weights)
        self.factors = [np.asarray(f) for f in factors]
        self.device = device
        self.rank = int(self.factors[0].shape[1])
        self.shape = tuple(int(f.shape[0]) for f in self.factors)

    def __iter__(self):
        yield self.weights
        yield self.factors

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.weights, self.factors)[i]

    def to_tensor(self):
        """Full reconstruction as an fp32 CUDA tensor (tensorly ``cp_to_tensor``).  An accessor for downstream code: the
        factorization itself never materialises it."""
        from ..engine import require_cuda
        dev = require_cuda(self.device)
        fs = [torch.as_tensor(np.ascontiguousarray(f), dtype=torch.float32, device=dev) for f in self.factors]
        w = torch.as_tensor(np.ascontiguousarray(self.weights), dtype=torch.float32, device=dev)
        letters = "abcdefgh"[: len(fs)]
        expr = ",".join(l + "r" for l in letters) + ",r->" + letters
        return torch.einsum(expr, *fs, w)

    def copy(self):
        return CPTensor(self.weights.copy(), [f.copy() for f in self.factors], self.device)


def check_random_state(seed):
    """tensorly ``check_random_state``: None -> numpy's global RandomState, int -> ``RandomState(seed)``."""
    if seed is None:
        return np.random.mtrand._rand
    if isinstance(seed, np.random.RandomState):
        return seed
    if isinstance(seed, (int, np.integer)):
        return np.random.RandomState(int(seed))
    raise ValueError("Seed should be None, int or np.random.RandomState")


def random_factors(shape, rank, random_state=None):
    """tensorly ``random_cp(..., normalise_factors=False)``: one ``random_sample((I_n, rank))`` per mode, in mode order."""
    rng = check_random_state(random_state)
    return [rng.random_sample((int(s), int(rank))) for s in shape]


def _nndsvd_columns(U, S, Vt, eps):
    """tensorly ``make_svd_non_negative(..., nntype='nndsvd')`` (Boutsidis & Gallopoulos 2008), W only; on the device."""
    W = torch.zeros_like(U)
    W[:, 0] = torch.sqrt(S[0]) * U[:, 0].abs()
    for j in range(1, S.shape[0]):
        a, b = U[:, j], Vt[j, :]
        a_p, b_p = a.clamp(min=0), b.clamp(min=0)
        a_n, b_n = (-a).clamp(min=0), (-b).clamp(min=0)
        a_p_n, b_p_n, a_n_n, b_n_n = a_p.norm(), b_p.norm(), a_n.norm(), b_n.norm()
        m_p, m_n = a_p_n * b_p_n, a_n_n * b_n_n
        if bool(m_p > m_n):
            u, sigma = a_p / a_p_n, m_p
        else:
            u, sigma = a_n / a_n_n, m_n
        W[:, j] = torch.sqrt(S[j] * sigma) * u
    return torch.sign(W) * (W.abs() - eps).clamp(min=0)


def svd_factors(X, rank, random_state=None, non_negative=True):
    """tensorly ``initialize_cp(init='svd', svd='truncated_svd', non_negative=True)`` on a CUDA tensor.

    Per mode: left singular vectors of the unfolding from the eigen-decomposition of its fp64 Gram matrix
    ``X_(n) X_(n)^T`` (I_n x I_n, cuBLAS + cuSOLVER ``syevd``), right vectors as ``X_(n)^T U / S``, ``svd_flip``, NNDSVD,
    mode-0 columns scaled by the singular values, random padding columns when ``I_n < rank``.  ``non_negative=False``
    (plain ``parafac``) keeps the flipped left singular vectors instead of their NNDSVD."""
    rng = check_random_state(random_state)
    eps = float(np.finfo(np.float32).eps)
    out = []
    nd = X.dim()
    for mode in range(nd):
        Xn = X.movedim(mode, 0).reshape(X.shape[mode], -1).to(torch.float64)
        k = min(int(rank), Xn.shape[0], Xn.shape[1])
        evals, evecs = torch.linalg.eigh(Xn @ Xn.T)
        order = torch.argsort(evals, descending=True)[:k]
        S = evals[order].clamp(min=0).sqrt()
        U = evecs[:, order]
        Vt = (U.T @ Xn) / torch.where(S > 0, S, torch.ones_like(S))[:, None]
        # svd_flip, u-based: the largest-|.| entry of every column of U is made positive
        idx = U.abs().argmax(dim=0)
        signs = torch.sign(U[idx, torch.arange(k, device=U.device)])
        signs = torch.where(signs == 0, torch.ones_like(signs), signs)
        U, Vt = U * signs, Vt * signs[:, None]
        W = _nndsvd_columns(U, S, Vt, eps) if non_negative else U
        if mode == 0:
            W = W * S[None, :]
        W = W.cpu().numpy()
        if X.shape[mode] < rank:
            W = np.concatenate([W, rng.random_sample((W.shape[0], int(rank) - int(X.shape[mode])))], axis=1)
        out.append(W[:, :rank])
        del Xn
    return out


def initialize_factors(X, rank, init="svd", svd="truncated_svd", random_state=None, non_negative=True):
    """List of ``[I_n, rank]`` float64 host factors (weights are ones); non-negative (``abs``) unless ``non_negative=False``."""
    if isinstance(init, str):
        if init == "random":
            factors = random_factors(X.shape, rank, random_state)
        elif init == "svd":
            if svd not in ("truncated_svd", "numpy_svd"):
                raise ValueError("svd=%r is not supported (only 'truncated_svd')" % (svd,))
            factors = svd_factors(X, rank, random_state, non_negative=non_negative)
        else:
            raise ValueError('Initialization method "%s" not recognized' % init)
    else:
        try:
            weights, fs = init
        except Exception:
            raise ValueError("If initialization method is a mapping, then it must be possible to convert it to a CPTensor instance")
        fs = [np.array(f, dtype=np.float64) for f in fs]
        if len(fs) != X.dim() or any(f.shape != (X.shape[i], rank) for i, f in enumerate(fs)):
            raise ValueError("init factors do not match the tensor shape / rank")
        if weights is not None and not np.all(np.asarray(weights) == 1):
            w = np.asarray(weights, dtype=np.float64)
            w_avg = np.prod(w) ** (1.0 / w.shape[0])
            fs = [f * w_avg for f in fs]
        factors = fs
    factors = [np.asarray(f, dtype=np.float64) for f in factors]
    return [np.abs(f) for f in factors] if non_negative else factors

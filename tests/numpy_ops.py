"""fp64 numpy versions of the ``ops`` interface of ``cell2cell_b200.sharded.DeviceOps`` (test infrastructure): lets the
CPU-only suite run the host-side orchestration of the sharded and the coupled factorization without a GPU."""
import numpy as np
import torch

from oracle import ncp_oracle as o


class NumpyOps:
    """fp64 numpy versions of the ops interface (test infrastructure)."""

    def __init__(self, rank):
        self.R = rank
        self.calls = []                 # streaming passes issued, in order

    def pack(self, f, device):
        return torch.as_tensor(np.array(f, dtype=np.float64))

    def unpack(self, a):
        return a

    def workspace(self, shape, device, masked=False):
        return None

    def sumsq(self, X):
        return float((X.numpy().astype(np.float64) ** 2).sum())

    @staticmethod
    def _kr(factors):
        w = factors[0].numpy()
        for f in factors[1:]:
            w = (w[:, None, :] * f.numpy()[None, :, :]).reshape(-1, w.shape[1])
        return w

    @staticmethod
    def _ret(val, out):
        val = torch.as_tensor(val)
        if out is None:
            return val
        out.copy_(val)
        return out

    def plane(self, X, factors, ws, out=None):
        self.calls.append("plane")
        x = X.numpy().astype(np.float64)
        x3 = x.reshape(-1, x.shape[-2], x.shape[-1])
        return self._ret(np.einsum("psv,sr,vr->pr", x3, factors[-2].numpy(), factors[-1].numpy()), out)

    def fiber(self, X, factors, ws, out=None):
        self.calls.append("fiber")
        x = X.numpy().astype(np.float64)
        x3 = x.reshape(-1, x.shape[-2], x.shape[-1])
        w = self._kr(factors[:-2])
        return self._ret(np.einsum("psv,pr->svr", x3, w).reshape(-1, w.shape[1]), out)

    def m_from(self, TU, shape, factors, mode):
        nd = len(shape)
        tu = TU.numpy()
        R = tu.shape[1]
        if mode < nd - 2:
            lead = shape[:nd - 2]
            t = tu.reshape(tuple(lead) + (R,))
            letters = "abcdef"[: len(lead)]
            ops = [letters + "r"]
            args = [t]
            for j, l in enumerate(letters):
                if j != mode:
                    ops.append(l + "r")
                    args.append(factors[j].numpy())
            return torch.as_tensor(np.einsum(",".join(ops) + "->" + letters[mode] + "r", *args))
        u = tu.reshape(shape[-2], shape[-1], R)
        if mode == nd - 2:
            return torch.as_tensor(np.einsum("svr,vr->sr", u, factors[-1].numpy()))
        return torch.as_tensor(np.einsum("svr,sr->vr", u, factors[-2].numpy()))

    def gram_hadamard(self, factors, shape, skip, changed=None, out=None):
        """Like the device op, the per-factor Grams are cached per shape and only the Gram of ``changed`` is recomputed -- so a
        caller that names the wrong factor gets wrong numbers here too."""
        R = factors[0].shape[1]
        key = tuple(int(s) for s in shape)
        if not hasattr(self, "_grams"):
            self._grams = {}
        if changed is None or key not in self._grams:
            self._grams[key] = [f.numpy().T @ f.numpy() for f in factors]
        else:
            self._grams[key][changed] = factors[changed].numpy().T @ factors[changed].numpy()
        g = np.ones((R, R))
        allg = np.ones((R, R))
        for k, gk in enumerate(self._grams[key]):
            allg = allg * gk
            if k != skip:
                g = g * gk
        if out is not None:
            out.copy_(torch.as_tensor(g))
            return out, torch.as_tensor(np.array([allg.sum()]))
        return torch.as_tensor(g), torch.as_tensor(np.array([allg.sum()]))

    def mu_update(self, A, M, G):
        eps = np.finfo(np.float64).eps
        a = A.numpy()
        a *= np.clip(M.numpy(), eps, None) / np.clip(a @ G.numpy(), eps, None)
        return A

    # ---- coupled factorization ----
    def masked_mttkrp(self, X, mask, factors, mode, ws):
        self.calls.append("masked_mttkrp")
        fs = [f.numpy() for f in factors]
        x = X.numpy().astype(np.float64)
        m = mask.numpy().astype(np.float64)
        x = x * m + o.cp_to_tensor(None, fs, mask=1 - m)
        return torch.as_tensor(o.mttkrp(x, None, fs, mode))

    def mu_update_coupled(self, A1, A2, M1, M2, G1, G2, iprod):
        eps = np.finfo(np.float64).eps
        a1, a2 = A1.numpy(), A2.numpy()
        num = np.clip((M1.numpy() + M2.numpy()) / 2.0, eps, None)
        den = np.clip((a1 @ G1.numpy() + a2 @ G2.numpy()) / 2.0, eps, None)
        new = a1 * num / den
        a1[...] = new
        a2[...] = new
        if iprod is not None:
            iprod[0] = float((M1.numpy() * new).sum())
            iprod[1] = float((M2.numpy() * new).sum())
        return A1

    def missing_rec_sumsq(self, X, mask, factors, ws, out=None):
        rec = o.cp_to_tensor(None, [f.numpy() for f in factors], mask=1 - mask.numpy().astype(np.float64))
        val = torch.tensor([float((rec * rec).sum())], dtype=torch.float64)
        if out is not None:
            out.copy_(val)
            return out
        return val

    def recon_sums(self, X, mask, factors, ws):
        x = X.numpy().astype(np.float64)
        d = x - o.cp_to_tensor(None, [f.numpy() for f in factors])
        return np.array([d.sum(), (d * d).sum(), x.sum(), (x * x).sum(), x.size])

"""Where the end-to-end time of PreBuiltTensor(host array).compute_tensor_factorization(...) goes (atlas shape, rank 10).

    python scripts/e2e_breakdown.py [--iters 100]
Wall-clock sections with a device synchronize on both sides; prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cell2cell_b200 import engine  # noqa: E402
from cell2cell_b200.synthetic import synthetic_lowrank_tensor  # noqa: E402
from cell2cell_b200.tensor import PreBuiltTensor  # noqa: E402
from cell2cell_b200.tensor import factorization as fz  # noqa: E402
from cell2cell_b200.tensor import tensor as tz  # noqa: E402


class Timer:
    def __init__(self):
        self.t = {}

    def section(self, name):
        timer = self

        class _S:
            def __enter__(self):
                torch.cuda.synchronize()
                self.t0 = time.perf_counter()

            def __exit__(self, *a):
                torch.cuda.synchronize()
                timer.t[name] = timer.t.get(name, 0.0) + (time.perf_counter() - self.t0) * 1e3
        return _S()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--shape", type=int, nargs="+", default=[60, 2000, 40, 40])
    ap.add_argument("--rank", type=int, default=10)
    args = ap.parse_args()
    shape, rank = tuple(args.shape), args.rank
    x = torch.from_numpy(synthetic_lowrank_tensor(shape, rank=5, noise=0.05, seed=1)).pin_memory()
    dev = torch.device("cuda", 0)
    names = [list(range(s)) for s in shape]
    tm = Timer()
    for rep in range(3):
        tm.t = {}
        with tm.section("total_api"):
            t = PreBuiltTensor(x, order_names=names, device=dev)
            t.compute_tensor_factorization(rank=rank, init="random", random_state=rep, n_iter_max=args.iters, tol=-1.0)
        del t
    total = dict(tm.t)
    tm.t = {}
    for rep in range(3):
        with tm.section("h2d_768MB"):
            X = x.to(dev, non_blocking=True)
        with tm.section("prebuilt_ctor_device_input"):
            t = PreBuiltTensor(X, order_names=names)
        with tm.section("initialize_factors_host"):
            hf = fz.initialize_factors(t.tensor, rank, init="random", svd="truncated_svd", random_state=rep)
        with tm.section("pack_factors"):
            fs = [engine.pack_factor(f.astype(np.float32), rank, dev) for f in hf]
        with tm.section("ncp_run"):
            errors, n_iter, conv = engine.ncp_run(t.tensor, None, fs, rank, n_iter_max=args.iters, tol=-1.0, check_every=10)
        with tm.section("recon_sums"):
            engine.recon_sums(t.tensor, None, fs, rank)
        with tm.section("unpack_to_host"):
            out = [engine.unpack_factor(f, rank).cpu().numpy() for f in fs]
        cp = fz.CPTensor(np.ones(rank, dtype=np.float32), out, device=dev)
        t.tl_object = cp
        with tm.section("cp_normalize"):
            tz._cp_normalize(cp)
        with tm.section("explained_variance"):
            t.explained_variance()
        del t, X
    out = {k: round(v / 3, 3) for k, v in tm.t.items()}
    out["total_api_ms"] = round(total["total_api"], 3)
    out["h2d_gbs"] = round(x.numel() * 4 / (out["h2d_768MB"] * 1e-3) / 1e9, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

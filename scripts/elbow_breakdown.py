"""Where the wall time of the elbow sweep goes (atlas shape, ranks 1-25 x 10 runs x 100 fixed iterations).

    python scripts/elbow_breakdown.py [--upper-rank 25] [--runs 10]
Wraps the pieces of ``_factorize_batch`` with wall-clock timers (no extra synchronisation: a section that ends in a
device->host read contains the device time it waited for) and prints one JSON line; ``--cprofile`` adds the top
cumulative entries of a cProfile of the same sweep.
"""
import argparse
import cProfile
import io
import json
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cell2cell_b200 import engine  # noqa: E402
from cell2cell_b200.synthetic import synthetic_lowrank_tensor  # noqa: E402
from cell2cell_b200.tensor import PreBuiltTensor  # noqa: E402
from cell2cell_b200.tensor import factorization as fz  # noqa: E402

ACC = {}


def timed(mod, name, label=None):
    fn = getattr(mod, name)
    label = label or name

    def wrap(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            ACC[label] = ACC.get(label, 0.0) + (time.perf_counter() - t0)
            ACC[label + "_calls"] = ACC.get(label + "_calls", 0) + 1
    setattr(mod, name, wrap)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs="+", default=[60, 2000, 40, 40])
    ap.add_argument("--upper-rank", type=int, default=25)
    ap.add_argument("--runs", type=int, default=10)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--cprofile", action="store_true")
    args = ap.parse_args()
    shape = tuple(args.shape)
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(synthetic_lowrank_tensor(shape, rank=5, noise=0.05, seed=1)).to(dev)
    t = PreBuiltTensor(x, order_names=[list(range(s)) for s in shape], device=dev)

    def sweep():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t.elbow_rank_selection(upper_rank=args.upper_rank, runs=args.runs, init="random", random_state=888,
                               n_iter_max=args.iters, tol=-1.0, automatic_elbow=False, output_fig=False)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    warm = sweep()
    plain = sweep()
    timed(fz, "_factorize_batch")
    timed(fz, "initialize_factors")
    timed(engine, "ncp_run_batched")
    timed(engine, "recon_sums")
    timed(engine, "pack_factor")
    timed(engine, "unpack_factor")
    ACC.clear()
    instrumented = sweep()
    out = {"shape": shape, "upper_rank": args.upper_rank, "runs": args.runs, "iters": args.iters, "warm_s": warm,
           "sweep_s": plain, "instrumented_s": instrumented, "sections_s": dict(ACC)}
    if args.cprofile:
        pr = cProfile.Profile()
        pr.enable()
        sweep()
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
        out["cprofile"] = s.getvalue().splitlines()[:70]
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Coupled factorization micro-benchmark: two synthetic atlas-size tensors sharing contexts / senders / receivers,
fixed iteration counts through the public ``coupled_non_negative_parafac``; the per-iteration time is the difference of
two iteration counts (so init, graph capture and the final error pass cancel).  One JSON line per rank.

    python scripts/coupled_bench.py [--shape1 60 2000 40 40] [--shape2 60 1500 40 40] [--ranks 10 16]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cell2cell_b200 import engine  # noqa: E402
from cell2cell_b200.tensor.coupled_factorization import coupled_non_negative_parafac  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape1", type=int, nargs="+", default=[60, 2000, 40, 40])
    ap.add_argument("--shape2", type=int, nargs="+", default=[60, 1500, 40, 40])
    ap.add_argument("--non-shared", type=int, default=1)
    ap.add_argument("--ranks", type=int, nargs="+", default=[10, 16])
    ap.add_argument("--iters", type=int, nargs=2, default=[20, 120])
    args = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(0)
    X1 = torch.rand(tuple(args.shape1), generator=g, device="cuda")
    X2 = torch.rand(tuple(args.shape2), generator=g, device="cuda")
    nbytes = 4 * (X1.numel() + X2.numel())
    for R in args.ranks:
        times = []
        for n in args.iters + args.iters:
            torch.cuda.synchronize()
            l0 = engine.launch_count()
            t0 = time.perf_counter()
            cp1, cp2, (e1, e2) = coupled_non_negative_parafac(X1, X2, R, args.non_shared, n_iter_max=n, init="random", tol=-1.0,
                                                              random_state=0, return_errors=True)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
            launches = engine.launch_count() - l0
        lo, hi = args.iters
        per_it = (min(times[1], times[3]) - min(times[0], times[2])) / (hi - lo)
        print(json.dumps({"shape1": args.shape1, "shape2": args.shape2, "rank": R, "iter_ms": round(per_it * 1e3, 4),
                          "it_per_s": round(1.0 / per_it, 1), "call_ms": {str(n): round(t * 1e3, 2) for n, t in zip(args.iters, times[2:])},
                          "reads_per_iteration": 2, "stream_gbs": round(2 * nbytes / per_it / 1e9, 1),
                          "host_issued_launches_last_call": int(launches), "final_errors": [float(e1[-1]), float(e2[-1])]}), flush=True)


if __name__ == "__main__":
    main()

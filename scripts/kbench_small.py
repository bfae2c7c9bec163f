"""Small-tensor benchmark (BASELINE configs 1-2): the one-CTA-per-run kernel (c2c_ncp_run_multi) against the whole-GPU path.

    python scripts/kbench_small.py [--workload balf|toy|asd] [--upper-rank 25] [--runs 10] [--iters 100]
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
from cell2cell_b200.tensor.cp import random_factors  # noqa: E402

SHAPES = {"balf": (12, 1054, 6, 6), "toy": (4, 300, 6, 6), "asd": (13, 1900, 17, 17)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="balf", choices=sorted(SHAPES))
    ap.add_argument("--upper-rank", type=int, default=25)
    ap.add_argument("--runs", type=int, default=10)
    ap.add_argument("--iters", type=int, default=100)
    args = ap.parse_args()
    shape = SHAPES[args.workload]
    x = synthetic_lowrank_tensor(shape, rank=5, noise=0.05, seed=1)
    X = torch.from_numpy(x).cuda()
    nx2 = engine.sumsq(X)
    units = [(r, run) for r in range(1, args.upper_rank + 1) for run in range(args.runs)]
    inits = [random_factors(shape, r, 888 + run) for r, run in units]
    engine.ncp_run_multi(X, inits[:4], n_iter_max=3, tol=-1.0, norm_x2=nx2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = engine.ncp_run_multi(X, inits, n_iter_max=args.iters, tol=-1.0, norm_x2=nx2)
    torch.cuda.synchronize()
    t_multi = time.perf_counter() - t0
    # one rank-10 run on the whole GPU (the path a single factorization takes)
    f = [engine.pack_factor(a.astype(np.float32), 10, X.device) for a in random_factors(shape, 10, 1)]
    ws = engine.Workspace(shape, 10, False, X.device)
    engine.ncp_run(X, None, f, 10, n_iter_max=args.iters, tol=-1.0, check_every=0, norm_x2=nx2, ws=ws)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    engine.ncp_run(X, None, f, 10, n_iter_max=args.iters, tol=-1.0, check_every=0, norm_x2=nx2, ws=ws)
    torch.cuda.synchronize()
    t_single = time.perf_counter() - t0
    # the elbow analysis through the public API
    names = [list(range(s)) for s in shape]
    t = PreBuiltTensor(X, order_names=names)
    t.elbow_rank_selection(upper_rank=2, runs=2, init="random", random_state=1, n_iter_max=5, tol=-1.0, output_fig=False,
                           automatic_elbow=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, loss = t.elbow_rank_selection(upper_rank=args.upper_rank, runs=args.runs, init="random", random_state=888,
                                     n_iter_max=args.iters, tol=-1.0, output_fig=False, automatic_elbow=False)
    torch.cuda.synchronize()
    t_elbow = time.perf_counter() - t0
    t0 = time.perf_counter()
    t2 = PreBuiltTensor(X, order_names=names)
    t2.elbow_rank_selection(upper_rank=args.upper_rank, runs=args.runs, init="random", random_state=888,
                            n_iter_max=args.iters, tol=-1.0, output_fig=False, automatic_elbow=False, batch_runs=False) \
        if False else None
    print(json.dumps({"workload": args.workload, "shape": shape, "units": len(units), "iters": args.iters,
                      "multi_kernel_wall_s": round(t_multi, 4),
                      "multi_aggregate_it_per_s": round(len(units) * args.iters / t_multi, 1),
                      "single_run_rank10_it_per_s": round(args.iters / t_single, 1),
                      "elbow_api_wall_s": round(t_elbow, 4), "loss_rank1": float(loss[0][1]), "loss_last": float(loss[-1][1]),
                      "final_err_run0": float(res[0][1][-1])}), flush=True)


if __name__ == "__main__":
    main()

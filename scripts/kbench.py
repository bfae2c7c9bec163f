"""Kernel micro-benchmark: times the two streaming passes (plane_contract / fiber_contract) and one full NCP iteration on
a synthetic tensor, checks them against torch einsum, prints one JSON line per (shape, rank).

    python scripts/kbench.py [--shape 60 2000 40 40] [--ranks 10 20] [--masked]
    C2C_B200_LIB=/path/to/variant.so python scripts/kbench.py ...
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cell2cell_b200 import engine  # noqa: E402


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs="+", default=[60, 2000, 40, 40])
    ap.add_argument("--ranks", type=int, nargs="+", default=[10])
    ap.add_argument("--masked", action="store_true")
    ap.add_argument("--structured", action="store_true", help="masked: missing rows / columns / planes instead of 20 %% random entries")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--tag", default=os.environ.get("C2C_B200_LIB", "product"))
    ap.add_argument("--quick", action="store_true", help="only time the two streaming passes (no reference check, no iteration)")
    args = ap.parse_args()
    shape = tuple(args.shape)
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.rand(shape, generator=g, device="cuda")
    mask = (torch.rand(shape, generator=g, device="cuda") > 0.2).to(torch.uint8) if args.masked else None
    if args.structured and mask is not None:          # what how='outer' produces: a cell type missing from every 3rd context
        mask.fill_(1)                                 # (its sender rows and receiver columns) and 2 % of the planes missing
        mask[2::3, ..., 5, :] = 0
        mask[2::3, ..., :, 5] = 0
        gone = torch.rand(shape[:-2], generator=g, device="cuda") < 0.02
        mask[gone] = 0
    if mask is not None:
        X = X * mask                                  # cell2cell's tensors hold zeros at their missing entries (nan_to_num)
    n_x = X.numel()
    letters = "abcdefgh"[: len(shape)]
    for R in args.ranks:
        fs = [engine.pack_factor(torch.rand((s, R), generator=g, device="cuda"), R, X.device) for s in shape]
        ws = engine.Workspace(shape, R, args.masked, X.device)
        T = engine.plane_contract(X, mask, fs, R, ws=ws)
        U = engine.fiber_contract(X, mask, fs, R, ws=ws)
        if args.quick:
            tp = timeit(lambda: engine.plane_contract(X, mask, fs, R, ws=ws, out=T), args.reps)
            tf = timeit(lambda: engine.fiber_contract(X, mask, fs, R, ws=ws, out=U), args.reps)
            print(json.dumps({"tag": args.tag, "rank": R, "plane_ms": round(tp, 4), "fiber_ms": round(tf, 4),
                              "plane_gbs": round(n_x * 4 / tp / 1e6, 1), "fiber_gbs": round(n_x * 4 / tf / 1e6, 1)}), flush=True)
            continue
        Xh = X.double()
        if args.masked:
            rec = torch.einsum(",".join(l + "r" for l in letters) + "->" + letters, *[f[:, :R].double() for f in fs])
            Xh = torch.where(mask.bool(), Xh, rec)
        lead, tr = letters[:-2], letters[-2:]
        Tw = torch.einsum(letters + "," + tr[0] + "r," + tr[1] + "r->" + lead + "r", Xh, fs[-2][:, :R].double(), fs[-1][:, :R].double())
        Uw = torch.einsum(letters + "," + ",".join(l + "r" for l in lead) + "->" + tr + "r", Xh, *[f[:, :R].double() for f in fs[:-2]])
        eT = float((T[:, :R].double() - Tw.reshape(-1, R)).norm() / Tw.norm())
        eU = float((U[:, :R].double() - Uw.reshape(-1, R)).norm() / Uw.norm())
        tp = timeit(lambda: engine.plane_contract(X, mask, fs, R, ws=ws, out=T), args.reps)
        tf = timeit(lambda: engine.fiber_contract(X, mask, fs, R, ws=ws, out=U), args.reps)
        nx2 = engine.sumsq(X)
        fs2 = [f.clone() for f in fs]
        engine.ncp_run(X, mask, fs2, R, n_iter_max=5, tol=-1.0, check_every=0, norm_x2=nx2, ws=ws)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        engine.ncp_run(X, mask, fs2, R, n_iter_max=args.iters, tol=-1.0, check_every=0, norm_x2=nx2, ws=ws, sync=False)
        b.record()
        torch.cuda.synchronize()
        ti = a.elapsed_time(b) / args.iters
        bpe = 5 if args.masked else 4
        print(json.dumps({"tag": args.tag, "shape": shape, "rank": R, "masked": args.masked, "plane_ms": round(tp, 4),
                          "fiber_ms": round(tf, 4), "plane_gbs": round(n_x * bpe / tp / 1e6, 1),
                          "fiber_gbs": round(n_x * bpe / tf / 1e6, 1), "iter_ms": round(ti, 4), "it_per_s": round(1e3 / ti, 1),
                          "errT": eT, "errU": eU}), flush=True)
        assert eT < 1e-5 and eU < 1e-5, (eT, eU)


if __name__ == "__main__":
    main()

"""One bin of batched runs (the unit of the elbow sweep) on the atlas tensor: wall / device time of the whole call and of its
parts; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel list.
    python scripts/bin_probe.py [--ranks 25 3] [--iters 100]"""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cell2cell_b200 import engine  # noqa: E402
from cell2cell_b200.synthetic import synthetic_lowrank_tensor  # noqa: E402
from cell2cell_b200.tensor import factorization as fz  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", type=int, nargs="+", default=[25, 3])
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
shape = (60, 2000, 40, 40)
dev = torch.device("cuda", 0)
X = torch.from_numpy(synthetic_lowrank_tensor(shape, rank=5, noise=0.05, seed=1)).to(dev)
X._c2c_norm_x2 = engine.sumsq(X)
runs = [(r, 100 + i) for i, r in enumerate(args.ranks)]
fz._factorize_batch(X, runs, 5, -1.0)
torch.cuda.synchronize()
out = []
for _ in range(args.reps):
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fz._factorize_batch(X, runs, args.iters, -1.0)
    b.record()
    torch.cuda.synchronize()
    out.append({"wall_ms": 1e3 * (time.perf_counter() - t0), "device_ms": a.elapsed_time(b)})
print(json.dumps({"ranks": args.ranks, "iters": args.iters, "calls": out}))

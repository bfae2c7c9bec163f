"""Smallest run that exercises every kernel family once (for compute-sanitizer): tensor-pipe passes (stacked, unstacked and
ragged in-plane size), FFMA streaming passes, fused updates, batched runs, masked passes, tensor build."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cell2cell_b200 import engine  # noqa: E402
from cell2cell_b200.tensor.factorization import _factorize_batch, non_negative_parafac  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
engine.set_mma_min_rank(1)
for shape, rank in [((40, 500, 16, 8), 5), ((40, 500, 16, 8), 20), ((20000, 12, 12), 10)]:
    X = torch.rand(shape, generator=g, device="cuda")
    fs = [engine.pack_factor(torch.rand((s, rank), generator=g, device="cuda"), rank, X.device) for s in shape]
    engine.ncp_run(X, None, fs, rank, n_iter_max=6, tol=-1.0)
    print("mma", shape, rank, "ok", flush=True)
engine.set_mma_min_rank(13)
X = torch.rand((6, 400, 16, 16), generator=g, device="cuda")
non_negative_parafac(X, 6, n_iter_max=6, init="random", tol=-1.0, random_state=1)
_factorize_batch(X, 3, [1, 2, 3, 4], 6, -1.0)
mask = (torch.rand(X.shape, generator=g, device="cuda") > 0.2).to(torch.uint8)
non_negative_parafac(X, 4, n_iter_max=4, init="random", tol=-1.0, random_state=1, mask=mask)
print("ffma / batched / masked ok", flush=True)

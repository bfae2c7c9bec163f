"""Tensor-build micro-benchmark (K1): times c2c_build_ccc alone on resident inputs, the whole InteractionTensor(...) call
(host name logic + uploads + kernel), for a BASELINE config shape.

    python scripts/kbench_build.py [--config atlas] [--how outer] [--score expression_gmean]
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
from cell2cell_b200._lib import AGG, SCORE, check, lib  # noqa: E402
from cell2cell_b200.synthetic import CONFIGS, synthetic_contexts, synthetic_ppi  # noqa: E402
from cell2cell_b200.tensor import InteractionTensor  # noqa: E402
from cell2cell_b200.tensor import build_index as bi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="atlas")
    ap.add_argument("--how", default="outer")
    ap.add_argument("--score", default="expression_gmean")
    ap.add_argument("--agg", default="min")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    C, L, K = cfg["C"], cfg["L"], cfg["K"]
    ppi = synthetic_ppi(L, 2 * L)
    mats = synthetic_contexts(C, 2 * L, K, outer=(args.how != "inner" and args.config == "asd"))
    t0 = time.perf_counter()
    t = InteractionTensor(mats, ppi, how=args.how, communication_score=args.score, complex_sep="&",
                          complex_agg_method=args.agg, verbose=False)
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    t = InteractionTensor(mats, ppi, how=args.how, communication_score=args.score, complex_sep="&",
                          complex_agg_method=args.agg, verbose=False)
    torch.cuda.synchronize()
    t_api = time.perf_counter() - t0
    shape = tuple(t.tensor.shape)
    n_x = t.tensor.numel()
    # kernel alone on resident inputs
    complexes = bi.complex_table(ppi, "&")[4]
    tables = [bi.context_rows(list(df.index), complexes, True) for df in mats]
    names = [tb.names for tb in tables]
    cols = [list(df.columns) for df in mats]
    genes = bi.ordered(names[0], bi.select_elements(names, args.how in ("outer", "outer_genes"), 0.0))
    cells = bi.ordered(cols[0], bi.select_elements(cols, args.how in ("outer", "outer_cells"), 0.0))
    f = bi.filter_lr_pairs(ppi, genes, "&", True)
    idx = bi.resolve(tables, cols, genes, cells, f["A"].tolist(), f["B"].tolist())
    m32 = [np.ascontiguousarray(df.to_numpy(dtype=np.float32)) for df in mats]
    off = np.cumsum([0] + [m.size for m in m32[:-1]]).astype(np.int64)
    ld = np.array([m.shape[1] for m in m32], dtype=np.int32)
    dev = t.tensor.device
    up = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).to(dev)
    d_expr = up(np.concatenate([m.reshape(-1) for m in m32]), np.float32)
    d = [up(off, np.int64), up(ld, np.int32), up(idx["cell_col"], np.int32), up(idx["a_first"], np.int32),
         up(idx["a_count"], np.int32), up(idx["b_first"], np.int32), up(idx["b_count"], np.int32), up(idx["sub_rows"], np.int32)]
    X = torch.empty(shape, dtype=torch.float32, device=dev)
    with_mask = args.how != "inner"
    M = torch.empty(shape, dtype=torch.uint8, device=dev) if with_mask else None
    vp = lambda x: engine._ptr(x)

    def run():
        check(lib().c2c_build_ccc(vp(d_expr), *[vp(x) for x in d], shape[0], shape[1], shape[2], SCORE[args.score], AGG[args.agg],
                                  vp(X), vp(M), engine._stream()), "c2c_build_ccc")
    run()
    torch.cuda.synchronize()
    assert torch.equal(X, t.tensor)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.reps):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.reps
    a.record()
    for _ in range(args.reps):
        X.fill_(1.0)
    b.record()
    torch.cuda.synchronize()
    fill_ms = a.elapsed_time(b) / args.reps
    bytes_alg = n_x * (5 if with_mask else 4)
    print(json.dumps({"config": args.config, "shape": shape, "how": args.how, "score": args.score, "kernel_ms": round(ms, 4),
                      "algorithmic_bytes": bytes_alg, "write_gbs": round(bytes_alg / ms / 1e6, 1),
                      "frac_of_6561.6": round(bytes_alg / ms / 1e6 / 6561.6, 3), "torch_fill_ms": round(fill_ms, 4), "torch_fill_gbs": round(n_x * 4 / fill_ms / 1e6, 1), "api_s_first": round(t_first, 3),
                      "api_s": round(t_api, 3), "elements_per_s_api": round(n_x / t_api / 1e6, 1)}))


if __name__ == "__main__":
    main()

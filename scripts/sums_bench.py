"""Times c2c_recon_sums (the final-error / explained-variance pass) on a synthetic tensor; one JSON line per rank.

    python scripts/sums_bench.py [--shape 60 2000 40 40] [--ranks 4 10 12 20]
    C2C_B200_NO_STREAM_SUMS=1 python scripts/sums_bench.py ...      # the general kernel
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cell2cell_b200 import engine  # noqa: E402
from cell2cell_b200._lib import check, lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs="+", default=[60, 2000, 40, 40])
    ap.add_argument("--ranks", type=int, nargs="+", default=[4, 10, 12, 20])
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    shape = tuple(args.shape)
    assert len(shape) == 4, "the fp64 check below is written for 4-way tensors"
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.rand(shape, generator=g, device="cuda")
    for R in args.ranks:
        fs = [engine.pack_factor(torch.rand((s, R), generator=g, device="cuda") * 0.3, R, X.device) for s in shape]
        ws = engine.Workspace(shape, R, False, X.device)
        out = torch.zeros(5, dtype=torch.float64, device=X.device)

        def run():
            check(lib().c2c_recon_sums(engine._ptr(X), engine._ptr(None), X.dim(), engine._shape_arr(X.shape), engine._ptr_arr(fs),
                                       engine.ldr_of(R), R, engine._ptr(out), ws.ptr, ws.nbytes, engine._stream()), "c2c_recon_sums")
        run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps):
            run()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.reps
        # reference sums in fp64 on the device (chunked over the first mode)
        want = torch.zeros(4, dtype=torch.float64, device=X.device)
        for c in range(shape[0]):
            lead = fs[0][c, :R].double()
            for f in fs[1:-2]:
                lead = (lead[None, :] * f[:, :R].double()).reshape(-1, R) if lead.dim() == 2 else lead[None, :] * f[:, :R].double()
            tr = (fs[-2][:, None, :R].double() * fs[-1][None, :, :R].double()).reshape(-1, R)
            rec = (lead.reshape(-1, R) @ tr.T).reshape(X[c].shape)
            d = X[c].double() - rec
            want += torch.stack([d.sum(), (d * d).sum(), X[c].double().sum(), (X[c].double() ** 2).sum()])
        got = out.cpu().numpy()
        rel = [abs(float(got[i]) - float(want[i])) / max(abs(float(want[i])), 1e-30) for i in range(4)]
        print(json.dumps({"shape": shape, "rank": R, "recon_sums_ms": round(ms, 4), "gbs": round(X.numel() * 4 / ms / 1e6, 1),
                          "rel_err": rel, "count_ok": bool(got[4] == X.numel()),
                          "stream": os.environ.get("C2C_B200_NO_STREAM_SUMS", "0") != "1"}), flush=True)


if __name__ == "__main__":
    main()

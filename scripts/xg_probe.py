"""Peer-memory plumbing check (2+ GPUs, under torchrun): the IPC-mapped communication buffers and one small fused sharded run."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from cell2cell_b200 import engine
    from cell2cell_b200.sharded import PeerComm, shard_bounds, sharded_non_negative_parafac
    from cell2cell_b200.tensor.cp import random_factors
    t0 = time.time()
    comm = PeerComm.get(1 << 20, dev)
    print("rank %d: PeerComm ok in %.2f s, ptrs %s" % (rank, time.time() - t0, [hex(p) for p in comm.ptrs]), flush=True)
    shape, R = (16, 200, 16, 16), 6
    rs = np.random.RandomState(0)
    x = rs.random_sample(shape).astype(np.float32)
    lo, hi = shard_bounds(shape[0], world, rank)
    init = random_factors(shape, R, 11)
    Xl = torch.as_tensor(x[lo:hi].copy()).cuda()
    cp, errs = sharded_non_negative_parafac(Xl, shape, R, init, n_iter_max=20, tol=-1.0)
    cp2, errs2 = sharded_non_negative_parafac(Xl, shape, R, init, n_iter_max=20, tol=-1.0, fused=False)
    d = max(float(np.abs(np.asarray(a) - np.asarray(b)).max()) for a, b in zip(cp.factors, cp2.factors))
    print("rank %d: fused vs NCCL-driven max abs diff %.3g, errors %.6f %.6f" % (rank, d, errs[-1], errs2[-1]), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

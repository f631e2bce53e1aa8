"""Worker of tests/test_gpu_dp_nccl.py (one process per GPU, launched by torch.distributed.run): the data-parallel
training step of dxvae_b200/train.py over NCCL on a fixed global batch.  Rank 0 saves what every rank must now hold:
the all-reduced gradient blob, the 5 loss terms of the GLOBAL batch, and the weights after two optimiser steps."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out, n_global, precision):
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from dxvae_b200 import DXVAE
    from dxvae_b200.dxdata import DXGraphBatch, voices_to_batch
    from dxvae_b200.synth import random_voices
    from dxvae_b200.train import Trainer
    rank = dist.get_rank()
    torch.manual_seed(1000 + rank)                      # DIFFERENT seeds per rank: the Trainer must broadcast rank 0's weights
    m = DXVAE(); m.verbose = False; m.precision = precision
    tr = Trainer(m, lr=1e-3)
    w0 = m._flat.clone()
    pool = voices_to_batch(random_voices(n_global, seed=3))
    eps = torch.randn(n_global, 128, generator=torch.Generator().manual_seed(5))
    lo, hi = tr.shard(n_global)
    d = m._prepare(DXGraphBatch(pool.X[lo:hi], pool.params[lo:hi], pool.adj[lo:hi]))
    loss5 = tr.grad_step(d, eps[lo:hi].cuda(), n_global).clone()
    g = tr.g.clone()
    # every rank holds the same reduced gradient
    gmax = g.clone(); dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
    same = bool((gmax == g).all())
    # two full optimiser steps through the public step() with injected noise
    idx = list(range(n_global))
    for k in range(2):
        tr.step(pool, idx, eps=eps)
    # ... and one with the shared-generator noise: replicas must stay bit-identical
    tr.step(pool, idx)
    wmax = m._flat.clone(); dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
    wmin = m._flat.clone(); dist.all_reduce(wmin, op=dist.ReduceOp.MIN)
    if rank == 0:
        torch.save({"w0": w0.cpu(), "g": g.cpu(), "loss5": loss5.cpu(), "same_on_all_ranks": same,
                    "replicas_identical": bool((wmax == wmin).all()), "world": dist.get_world_size()}, out)
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3])

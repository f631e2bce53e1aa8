"""Data-parallel path on CPU: world_size=2 over gloo, compute through the emulation build.
Checks what train.py relies on: sharding a global batch in equal contiguous slices with
inv_batch = 1/global_batch and SUM all-reducing the flat gradient / loss terms reproduces the
single-process result on the concatenated batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dxvae_oracle as O
from tests import util


def _worker(rank, world, port, idx, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dxvae_b200.train import Trainer
    from tests.emu.emu import Emu
    X, P, E, A = util.dataset_graphs(idx)
    o = O.make_weights(0, 3.0)
    emu = Emu(o.state_dict())
    torch.manual_seed(77)
    eps = torch.randn(len(idx), 128)
    tr = Trainer.__new__(Trainer)
    tr.world, tr.rank = world, rank
    lo, hi = tr.shard(len(idx))
    bt = emu.batch(X[lo:hi].numpy(), P[lo:hi].numpy(), E[lo:hi])
    loss5, _, _, g = emu.elbo(bt, eps[lo:hi].numpy(), inv_batch=1.0 / len(idx))
    g = torch.from_numpy(g); l5 = torch.from_numpy(loss5)
    dist.all_reduce(g); dist.all_reduce(l5)
    if rank == 0:
        np.save(os.path.join(out_dir, "g.npy"), g.numpy()); np.save(os.path.join(out_dir, "l.npy"), l5.numpy())
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process(tmp_path):
    idx = util.pick_by_alg([3, 5, 18, 31])
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, idx, str(tmp_path)), nprocs=2, join=True)
    from tests.emu.emu import Emu
    X, P, E, A = util.dataset_graphs(idx)
    o = O.make_weights(0, 3.0)
    emu = Emu(o.state_dict())
    torch.manual_seed(77)
    eps = torch.randn(len(idx), 128)
    loss5, _, _, g = emu.elbo(emu.batch(X.numpy(), P.numpy(), E), eps.numpy())
    g2 = np.load(tmp_path / "g.npy"); l2 = np.load(tmp_path / "l.npy")
    assert np.allclose(l2, loss5, rtol=1e-5)
    assert np.abs(g2 - g).max() <= 1e-5 * np.abs(g).max()
    with pytest.raises(ValueError):
        from dxvae_b200.train import Trainer
        t = Trainer.__new__(Trainer); t.world, t.rank = 2, 0
        t.shard(5)

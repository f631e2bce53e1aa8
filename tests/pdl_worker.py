"""Worker of tests/test_gpu_launch.py: one small-batch and one mid-size ELBO train step plus a greedy decode, results saved
for comparison between launch modes (DX_NO_PDL in the environment is read once per process)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out):
    from dxvae_b200 import DXVAE
    from dxvae_b200.dxdata import voices_to_batch
    from dxvae_b200.synth import random_voices
    from dxvae_b200.train import Trainer
    torch.manual_seed(11)
    m = DXVAE(); m.verbose = False
    tr = Trainer(m, lr=1e-3)
    res = {}
    for n in (128, 3000):
        pool = voices_to_batch(random_voices(n, seed=4))
        eps = torch.randn(n, 128, generator=torch.Generator().manual_seed(6)).cuda()
        d = m._prepare(pool)
        res["loss5_%d" % n] = tr.grad_step(d, eps, n).cpu()
        res["g_%d" % n] = tr.g.cpu().clone()
    z = torch.randn(512, 128, generator=torch.Generator().manual_seed(8))
    gb = m.decode(z)
    res["adj"] = gb.adj.cpu(); res["params"] = gb.params.cpu()
    torch.save(res, out)


if __name__ == "__main__":
    main(sys.argv[1])

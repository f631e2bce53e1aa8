"""Data-parallel ELBO training driver behind DXVAE.train (model.py:374-391).

One process per GPU.  Every loss term of the reference is a batch mean (model.py:303-365),
so a global batch shards into equal contiguous slices: each rank runs the fused native step
(encode + teacher-forced loss + hand-written backward) on its slice with inv_batch =
1/global_batch, then ONE NCCL all-reduce (sum) of the flat 12.08 M-float gradient blob makes
every rank hold the single-GPU gradient; the AdamW update (torch.optim.AdamW defaults, fused
over the flat blob) is replicated.  No other collective is on the path."""
import torch
import torch.distributed as dist

from . import _lib
from .dxdata import DXGraphBatch


class Trainer:
    def __init__(self, model, lr=1e-3, w=(2.0, 5.0, 0.01), betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.model = model
        model._ensure_flat()
        n = model._total
        self.lr, self.w, self.betas, self.eps, self.wd = float(lr), tuple(float(x) for x in w), betas, eps, weight_decay
        self.g = torch.zeros(n, device="cuda")
        self.m = torch.zeros(n, device="cuda")
        self.v = torch.zeros(n, device="cuda")
        self.t = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0

    @staticmethod
    def upload(G):
        gb = DXGraphBatch.from_graphs(G)
        return DXGraphBatch(gb.X.to("cuda", torch.float32), gb.params.to("cuda", torch.float32),
                            gb.adj.to("cuda", torch.int64))

    def shard(self, n):
        if n % self.world:
            raise ValueError("global batch %d is not divisible by world size %d" % (n, self.world))
        per = n // self.world
        return self.rank * per, (self.rank + 1) * per

    def grad_step(self, d, eps, global_batch):
        """Fused fwd+bwd on this rank's prepared slice; leaves the (all-reduced) gradient in self.g."""
        self.g.zero_()
        loss5 = self.model.elbo_step(d, eps, self.w, grads=self.g, inv_batch=1.0 / global_batch)
        if self.world > 1:
            dist.all_reduce(self.g)
            dist.all_reduce(loss5)
        return loss5

    def apply(self):
        L = _lib.lib()
        self.t += 1
        m = self.model
        _lib.check(L.dxvae_adamw_step(m._total, m._flat.data_ptr(), self.g.data_ptr(), self.m.data_ptr(),
                                      self.v.data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                      self.t, 1.0, torch.cuda.current_stream().cuda_stream), "dxvae_adamw_step")

    def step(self, data, idx, eps=None):
        """One optimiser step on the global batch data[idx] (idx: list of ints, same on every rank)."""
        lo, hi = self.shard(len(idx))
        ii = torch.as_tensor(idx[lo:hi], device="cuda", dtype=torch.int64)
        sub = DXGraphBatch(data.X[ii], data.params[ii], data.adj[ii])
        d = self.model._prepare(sub)
        if eps is None:
            e = torch.empty(hi - lo, 128, device="cuda").normal_()
        else:
            e = torch.as_tensor(eps)[lo:hi].to("cuda", torch.float32).contiguous()
        loss5 = self.grad_step(d, e, len(idx))
        self.apply()
        return loss5

"""Data-parallel ELBO training driver behind DXVAE.train (model.py:374-391).

One process per GPU.  Every loss term of the reference is a batch mean (model.py:303-365),
so a global batch shards into equal contiguous slices: each rank runs the fused native step
(encode + teacher-forced loss + hand-written backward) on its slice with inv_batch =
1/global_batch, then ONE NCCL all-reduce (sum) of the flat 12.08 M-float gradient blob makes
every rank hold the single-GPU gradient; the AdamW update (torch.optim.AdamW defaults, fused
over the flat blob) is replicated.  No other collective is on the path.

Small batches (the reference's own regime: size_batch 32..128, BASELINE config 2) are bound by the
latency of ~800 dependent kernels.  Optionally (graph_max_batch > 0) the step is captured ONCE into a
CUDA graph over static buffers and replayed: the schedule is made batch-independent (encoder levels =
the reference's node order 6..1, every teacher-forcing step on every graph — both are valid schedules
of the same function), so one graph serves every batch of that size."""
import numpy as np
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
        # per-rank batches up to this size replay a captured CUDA graph.  Off by default: measured on B200 the
        # B=128 step is bound by ~10 us of GPU-side latency per dependent kernel, not by CPU launches (8.4 ms
        # eager with the compacted schedule vs 9.9 ms replayed with the batch-independent dense one)
        self.graph_max_batch = 0
        self._graphs = {}

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

    # ------------------------------------------------------------------ CUDA-graph step (small batches)
    def _static_step(self, B):
        """Captured (pack -> zero grads -> fused fwd+bwd) for per-rank batch B; returns the static buffers."""
        from . import _abi
        m = self.model
        key = (B, m.precision, self.w, m._flat.data_ptr())
        st = self._graphs.get(key)
        if st is not None:
            return st
        L = _lib.lib()
        dev = "cuda"
        st = {"X": torch.zeros(B, 7, 27, device=dev), "P": torch.zeros(B, 7, 21, device=dev),
              "adj": torch.zeros(B, dtype=torch.int64, device=dev), "eps": torch.zeros(B, 128, device=dev),
              "Xn": torch.zeros(7, B, 32, device=dev), "cls": torch.zeros(14, B, dtype=torch.int32, device=dev),
              "loss5": torch.zeros(5, device=dev), "inv": None}
        # node-order level schedule: level k = node 6-k over all graphs (rows v*B + b)
        rows = np.concatenate([np.arange(v * B, (v + 1) * B, dtype=np.int32) for v in range(6, 0, -1)])
        st["level_rows"] = torch.from_numpy(rows).to(dev)
        st["level_ptr"] = np.array([0, B, 2 * B, 3 * B, 4 * B, 5 * B, 6 * B, 6 * B], np.int32)   # (no rare-first order: both halves)
        st["ws"] = m._workspace(_abi.OP_TRAIN, B, fresh=True)

        def body(inv_batch):
            s = torch.cuda.current_stream().cuda_stream
            _lib.check(L.dxvae_pack_graphs(B, st["X"].data_ptr(), st["P"].data_ptr(), st["Xn"].data_ptr(),
                                           st["cls"].data_ptr(), s), "dxvae_pack_graphs")
            self.g.zero_()
            _lib.check(L.dxvae_elbo_step(
                m._flat.data_ptr(), B, st["Xn"].data_ptr(), st["cls"].data_ptr(), st["adj"].data_ptr(), 6,
                st["level_ptr"].ctypes.data, st["level_rows"].data_ptr(), None, st["eps"].data_ptr(), self.w[0], self.w[1],
                self.w[2], inv_batch, st["loss5"].data_ptr(), None, None, self.g.data_ptr(), st["ws"].data_ptr(),
                st["ws"].numel(), m._prec(), None, None, s), "dxvae_elbo_step")

        st["body"] = body
        st["graph"] = None
        self._graphs[key] = st
        return st

    def _replay(self, st, inv_batch):
        if st["graph"] is None or st["inv"] != inv_batch:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):              # warm-up outside capture (one-time attribute / table set-up)
                st["body"](inv_batch)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st["body"](inv_batch)
            st["graph"], st["inv"] = g, inv_batch
        st["graph"].replay()

    def step(self, data, idx, eps=None):
        """One optimiser step on the global batch data[idx] (idx: list of ints, same on every rank)."""
        lo, hi = self.shard(len(idx))
        ii = torch.as_tensor(idx[lo:hi], device="cuda", dtype=torch.int64)
        if self.world == 1 and 0 < hi - lo <= self.graph_max_batch:
            st = self._static_step(hi - lo)
            torch.index_select(data.X, 0, ii, out=st["X"])
            torch.index_select(data.params, 0, ii, out=st["P"])
            torch.index_select(data.adj, 0, ii, out=st["adj"])
            if eps is None:
                st["eps"].normal_()
            else:
                st["eps"].copy_(torch.as_tensor(eps)[lo:hi])
            self._replay(st, 1.0 / len(idx))
            self.apply()
            return st["loss5"].clone()
        sub = DXGraphBatch(data.X[ii], data.params[ii], data.adj[ii])
        d = self.model._prepare(sub)
        if eps is None:
            e = torch.empty(hi - lo, 128, device="cuda").normal_()
        else:
            e = torch.as_tensor(eps)[lo:hi].to("cuda", torch.float32).contiguous()
        loss5 = self.grad_step(d, e, len(idx))
        self.apply()
        return loss5

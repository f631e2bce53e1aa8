"""Data-parallel ELBO training driver behind DXVAE.train (model.py:374-391).

One process per GPU.  Every loss term of the reference is a batch mean (model.py:303-365),
so a global batch shards into equal contiguous slices: each rank runs the fused native step
(encode + teacher-forced loss + hand-written backward) on its slice with inv_batch =
1/global_batch, then ONE NCCL all-reduce (sum) of the flat 12.08 M-float gradient blob makes
every rank hold the single-GPU gradient; the AdamW update (torch.optim.AdamW defaults, fused
over the flat blob) is replicated.  No other collective is on the path.

Small batches (the reference's own regime: size_batch 32..128, BASELINE config 2) are bound by the
latency of several hundred dependent kernels; the native side answers with programmatic dependent launch,
cluster split-K products and one pass over all six decoder nodes for the work that does not depend on the
nodes in front (DESIGN.md §5 "Small batches": 4.49 ms per batch-128 step).  Optionally
(graph_max_batch > 0) the step is captured ONCE into a CUDA graph over static buffers and replayed:
the schedule is made batch-independent (encoder levels = the reference's node order 6..1, every
teacher-forcing step on every graph — both are valid schedules of the same function), so one graph
serves every batch of that size.  Measured on B200 the replay is SLOWER than the eager step
(7.3 vs 4.5 ms at batch 128: the dense schedule does more work and the graph serialises what the
stream overlaps), so it stays off by default."""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .dxdata import DXGraphBatch


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class FusedAdamW:
    """torch.optim.AdamW(lr, betas, eps, weight_decay) over the model's flat parameter blob as ONE kernel
    (dxvae_adamw_step), for the reference's own step idiom (model.py:383-386):

        opt.zero_grad(); loss, *_ = model(G); loss.backward(); opt.step()

    DXVAE.forward()'s backward hands every parameter a .grad that is a view of one flat blob, so step() needs no gather.
    Under torch.distributed it all-reduces that blob (sum) and averages: each rank's forward() is a mean over its own
    shard, which makes the update the mean over the global batch (equal shards, model.py:377 drop-last)."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.model = model
        model._ensure_flat()
        n = model._total
        self.lr, self.betas, self.eps, self.wd = float(lr), betas, eps, weight_decay
        self.m = torch.zeros(n, device="cuda")
        self.v = torch.zeros(n, device="cuda")
        self.t = 0
        self.world = _world()
        if self.world > 1:
            dist.broadcast(model._flat, 0)        # replicas must start from the same weights

    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None
        self.model._last_gflat = None

    def _flat_grad(self):
        m = self.model
        g = m._last_gflat
        named = dict(m.named_parameters())
        if g is not None:
            base = g.data_ptr()
            if all(named[n].grad is not None and named[n].grad.data_ptr() == base + 4 * off for n, off, _ in m._table):
                return g
        g = torch.zeros(m._total, device="cuda")   # gradients that did not come from the fused step: gather them
        for n, off, shape in m._table:
            gr = named[n].grad
            if gr is not None:
                g[off:off + gr.numel()].copy_(gr.reshape(-1))
        return g

    @torch.no_grad()
    def step(self):
        L = _lib.lib()
        m = self.model
        g = self._flat_grad()
        if self.world > 1:
            dist.all_reduce(g)
        self.t += 1
        _lib.check(L.dxvae_adamw_step(m._total, m._flat.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                      self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.t, 1.0 / self.world,
                                      torch.cuda.current_stream().cuda_stream), "dxvae_adamw_step")


class Trainer:
    def __init__(self, model, lr=1e-3, w=(2.0, 5.0, 0.01), betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.model = model
        model._ensure_flat()
        n = model._total
        self.lr, self.w, self.betas, self.eps, self.wd = float(lr), tuple(float(x) for x in w), betas, eps, weight_decay
        # gradient blob + 8 trailing floats: the 5 loss terms ride in the same buffer, so data-parallel training issues
        # ONE collective per step
        self.gbuf = torch.zeros(n + 8, device="cuda")
        self.g = self.gbuf[:n]
        self.loss5 = self.gbuf[n:n + 5]
        self.m = torch.zeros(n, device="cuda")
        self.v = torch.zeros(n, device="cuda")
        self.t = 0
        self.world = _world()
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.eps_gen = None
        self.comm_stream = None
        if self.world > 1:
            # replicas start from rank 0's weights whatever each process was seeded with, and draw the reparameterisation
            # noise of a GLOBAL batch from one shared generator (each rank keeps its slice): an N-rank step is then the
            # 1-rank step on the concatenated batch
            dist.broadcast(model._flat, 0)
            seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device="cuda")
            dist.broadcast(seed, 0)
            self.eps_gen = torch.Generator(device="cuda")
            self.eps_gen.manual_seed(int(seed.item()))
            self.comm_stream = torch.cuda.Stream()
            o = {n_: off for n_, off, _ in model._table}
            # [decoder-only tensors) = combin_decode.weight_ih .. gate.0.weight: their gradient is final when the decoder
            # backward returns, before the encoder backward starts
            self.dec_lo, self.dec_hi = o["combin_decode.weight_ih"], o["gate.0.weight"]
            self.dec_done = torch.cuda.Event()
            self.dec_done.record()                      # (creates the underlying cudaEvent_t handle)
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

    def draw_eps(self, lo, hi, global_batch):
        """N(0,1) noise for rows [lo, hi) of a global batch (model.py:284 rsample).  One process: the global torch generator,
        as the reference.  Data parallel: every rank draws the SAME global matrix from the shared generator and keeps its
        slice (33 M normals for 8 x 32768 graphs: ~0.1 ms)."""
        if self.world == 1:
            return torch.empty(hi - lo, 128, device="cuda").normal_()
        return torch.empty(global_batch, 128, device="cuda").normal_(generator=self.eps_gen)[lo:hi].contiguous()

    def grad_step(self, d, eps, global_batch):
        """Fused fwd+bwd on this rank's prepared slice; leaves the (all-reduced) gradient in self.g and returns the 5 loss
        terms of the global batch.  Data parallel: the decoder-only gradient range is all-reduced on a side stream as
        soon as the decoder backward has been issued, overlapping the encoder backward; the rest (+ the loss terms) follows
        in one more collective."""
        self.gbuf.zero_()
        overlap = self.world > 1 and self.comm_stream is not None
        self.model.elbo_step(d, eps, self.w, grads=self.g, inv_batch=1.0 / global_batch, loss5=self.loss5,
                             decoder_done=self.dec_done if overlap else None)
        if self.world > 1:
            if overlap:
                cur = torch.cuda.current_stream()
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(self.dec_done)
                    dist.all_reduce(self.gbuf[self.dec_lo:self.dec_hi])
                dist.all_reduce(self.gbuf[:self.dec_lo])
                dist.all_reduce(self.gbuf[self.dec_hi:])
                cur.wait_stream(self.comm_stream)
            else:
                dist.all_reduce(self.gbuf)
        return self.loss5

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
                st["ws"].numel(), m._prec(), None, None, None, s), "dxvae_elbo_step")

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
            e = self.draw_eps(lo, hi, len(idx))
        else:
            e = torch.as_tensor(eps)[lo:hi].to("cuda", torch.float32).contiguous()
        loss5 = self.grad_step(d, e, len(idx)).clone()
        self.apply()
        return loss5

"""DXVAE — host-side mirror of the reference class (model.py:10-391) over the CUDA library.

Same constructor, same public methods (encode / decode / encode_decode / generate / loss /
forward / train, plus the `reparameterize` seam for injected noise), same 53-tensor
state_dict, same graph format in and out.  All arithmetic happens in hand-written sm_100a
kernels behind the C ABI of include/dxvae_b200.h; torch is used for device memory, streams,
autograd plumbing and torch.distributed.  There is no CPU path: compute methods raise
without the CUDA extension or without a GPU.
"""
import ctypes
import random

import numpy as np
import torch
import torch.nn as nn
from torch.distributions.normal import Normal

from . import _abi, _lib
from .dxdata import DXGraphBatch, IndexedBatch
from .params import param_table

_FIXED = dict(n_nodes=7, n_params=21, size_X=27, size_X0=23, size_H=512, size_Z=128)


class _DevBatch:
    """Device-resident batch in kernel layout (see include/dxvae_b200.h)."""
    __slots__ = ("B", "Xn", "cls", "adj", "n_levels", "level_ptr", "level_rows", "level", "csr", "step_ptr", "step_rows")


def _stream():
    return torch.cuda.current_stream().cuda_stream


class DXVAE(nn.Module):
    def __init__(self, n_nodes=7, n_params=21, size_X=27, size_X0=23, size_H=512, size_Z=128, checkpoint=None):
        super().__init__()
        given = dict(n_nodes=n_nodes, n_params=n_params, size_X=size_X, size_X0=size_X0, size_H=size_H, size_Z=size_Z)
        if given != _FIXED:
            raise ValueError("dxvae_b200 kernels are specialised on %r; got %r" % (_FIXED, given))
        self.device = "cuda" if torch.cuda.is_available() else "cpu"     # model.py:13
        self.n_nodes, self.n_params, self.size_X, self.size_Z = n_nodes, n_params, size_X, size_Z
        self.zero_X0 = torch.zeros(size_X0, device=self.device)
        self.zero_H = torch.zeros(size_H, device=self.device)
        self.hidden = None
        self.p_dist = Normal(0., 1.)
        H, Z, X, X0 = size_H, size_Z, size_X, size_X0
        # Parameter containers only (never called): identical registration order to
        # model.py:24-72, hence identical state_dict keys and identical seeded init.
        self.combin_encode = nn.GRUCell(X, H)
        self.loop_encode = nn.GRUCell(X, H)
        self.root_encode = nn.GRUCell(X0, H)
        self.h_to_mu = nn.Linear(H, Z)
        self.h_to_std = nn.Sequential(nn.Linear(H, Z), nn.Softplus())
        self.combin_decode = nn.GRUCell(X, H)
        self.loop_decode = nn.GRUCell(X, H)
        self.root_decode = nn.GRUCell(X0, H)
        self.z_to_h = nn.Sequential(nn.Linear(Z, H), nn.Tanh())
        self.h_to_x0 = nn.Sequential(nn.Linear(H, 2 * H), nn.ReLU(), nn.Linear(2 * H, 2 * H), nn.ReLU(),
                                     nn.Linear(2 * H, X0 + 32))
        self.h_to_x = nn.Sequential(nn.Linear(H, 2 * H), nn.ReLU(), nn.Linear(2 * H, 2 * H), nn.ReLU(),
                                    nn.Linear(2 * H, X))
        self.h_to_edge_self = nn.Sequential(nn.Linear(H, 2 * H), nn.ReLU(), nn.Linear(2 * H, 1))
        self.h_to_edge = nn.Sequential(nn.Linear(2 * H, 4 * H), nn.ReLU(), nn.Linear(4 * H, 2))
        self.gate = nn.Sequential(nn.Linear(2 * H, H), nn.Sigmoid())
        self.mapper = nn.Sequential(nn.Linear(2 * H, H, bias=False))
        self._flat = None           # fp32 blob all parameters are views of (device)
        self._table = None
        self._ws = {}
        self.max_chunk = 32768      # graphs per kernel pass for encode / decode
        self.host_batcher_max = 2048   # larger host batches are scheduled on the device
        self.verbose = True
        self.last_margins = None
        self.last_quant_margins = None
        self.last_loss5 = None      # device tensor (total, x0, xi, e, kld*w) of the latest forward()
        self._prep_tag = None       # workspace tag of the batcher (set while it runs on the upload stream)
        self._upload_stream = None
        self._last_gflat = None     # flat gradient blob the latest backward() handed out views of
        # arithmetic of the dense products, per entry point: "3xtf32" (default: tcgen05 tensor cores with error-compensated
        # hi/lo operand splits and chunked FP32 accumulation — FP32-accurate, meets the reference tolerances of
        # tests/test_gpu_parity.py), "fp32" (FFMA kernels) or, for training and encode only, "tf32" (plain tensor-core
        # TF32 under the looser stated bounds of tests/test_gpu_tf32.py)
        self.precision = "3xtf32"          # forward / loss / backward (training)
        self.encode_precision = "3xtf32"   # inference encode
        self.decode_precision = "3xtf32"   # greedy decode ("fp32" or "3xtf32": its discrete outputs need FP32 accuracy)
        # skip teacher-forced re-propagates that add no edge (exact; see dxvae_batch_steps)
        self.compact_steps = True
        if checkpoint is not None:
            self.load_state_dict(torch.load(checkpoint, map_location=self.device))

    # ------------------------------------------------------------------ plumbing
    def _ensure_flat(self):
        """Parameters live as views of one flat CUDA blob laid out per dxvae_param_entry()."""
        L = _lib.require_cuda()
        if self._table is None:
            self._table = param_table(L)
            self._total = int(L.dxvae_param_blob_floats())
        named = dict(self.named_parameters())
        ok = self._flat is not None and self._flat.is_cuda
        if ok:
            base = self._flat.data_ptr()
            ok = all(named[n].data_ptr() == base + 4 * off for n, off, _ in self._table)
        if not ok:
            flat = torch.zeros(self._total, dtype=torch.float32, device="cuda")
            for n, off, shape in self._table:
                p = named[n]
                k = p.numel()
                flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = flat[off:off + k].view(shape)
                p.grad = None
            self._flat = flat
            self.device = "cuda"
        return L

    _PREC = {"fp32": _abi.PREC_FP32, "tf32": _abi.PREC_TF32, "3xtf32": _abi.PREC_3XTF32}

    def _prec(self):
        if self.precision not in self._PREC:
            raise ValueError("precision must be 'fp32', '3xtf32' or 'tf32'")
        return self._PREC[self.precision]

    def _workspace(self, op, B, fresh=False, d=None, tag=None):
        """Caller-owned scratch for one native call.  d (a prepared batch): sized for ITS schedules — the largest encoder
        level, the active rows of the teacher-forced steps (dxvae_workspace_bytes_sched) — instead of the worst case."""
        L = _lib.lib()
        if d is not None:
            n = int(L.dxvae_workspace_bytes_sched(op, B, d.n_levels, d.level_ptr.ctypes.data,
                                                  None if d.step_ptr is None else d.step_ptr.ctypes.data))
            ws = self._ws.get((op,))
            if not fresh and (ws is None or ws.numel() < n):
                # the need follows the batch's topologies: leave 6 % of headroom (never above the worst case) so that
                # the next batches of the same size do not reallocate tens of GB
                n = max(n, min(int(L.dxvae_workspace_bytes(op, B)), n + n // 16))
        else:
            n = int(L.dxvae_workspace_bytes(op, B))
        if fresh:
            return torch.empty(n, dtype=torch.uint8, device="cuda")
        key = (op,) if tag is None else (op, tag)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < n:
            self._ws[key] = None
            ws = torch.empty(n, dtype=torch.uint8, device="cuda")
            self._ws[key] = ws
        return ws

    def _prepare(self, G, need_cls=True, host_batcher=None):
        """list of graphs / DXGraphBatch -> _DevBatch (the batcher).

        Host lists go through the C++ host batcher (flat CSR + level schedule + feedback
        marks); device-resident batches through the device scheduler."""
        L = _lib.require_cuda()
        gb = DXGraphBatch.from_graphs(G, staging=True)
        B = len(gb)
        if B == 0:
            raise ValueError("empty batch")
        st = _stream()
        d = _DevBatch()
        d.B = B
        d.Xn = torch.empty(7, B, 32, device="cuda")
        d.cls = torch.empty(14, B, dtype=torch.int32, device="cuda")
        indexed = isinstance(gb, IndexedBatch)
        if indexed and B <= self.host_batcher_max and host_batcher is not False:
            gb, indexed = gb.materialise(), False        # small batches take the host batcher (explicit CSR): gather on the host
        if indexed:
            # rows of a pinned host batch: upload the index list, the device gathers X / params / adjacency straight out of
            # the pinned memory into the kernels' layout (this IS the step's host-to-device transfer)
            base = gb.base
            ii = torch.from_numpy(gb.index).to("cuda", non_blocking=True)
            d.adj = torch.empty(B, dtype=torch.int64, device="cuda")
            _lib.check(L.dxvae_pack_graphs_indexed(B, ii.data_ptr(), base.X.data_ptr(), base.params.data_ptr(), base.adj.data_ptr(),
                                                   d.Xn.data_ptr(), d.cls.data_ptr(), d.adj.data_ptr(), st),
                       "dxvae_pack_graphs_indexed")
        else:
            Xg = gb.X.to("cuda", torch.float32, non_blocking=True).contiguous()
            Pg = gb.params.to("cuda", torch.float32, non_blocking=True).contiguous()
            _lib.check(L.dxvae_pack_graphs(B, Xg.data_ptr(), Pg.data_ptr(), d.Xn.data_ptr(), d.cls.data_ptr(), st),
                       "dxvae_pack_graphs")
        # host-resident batches of a few thousand graphs go through the C++ host batcher (explicit CSR, no device round
        # trip); larger ones are uploaded and scheduled on the device (same result, bit for bit)
        use_host = False if indexed else ((not gb.adj.is_cuda and B <= self.host_batcher_max) if host_batcher is None else host_batcher)
        d.level_ptr = np.zeros(16, np.int32)       # 8 level offsets, then per-level counts of back-edge-target rows
        d.csr = None
        if use_host:
            edges = gb.edge_lists()
            eptr = np.zeros(B + 1, np.int32)
            eptr[1:] = np.cumsum([len(e[0]) for e in edges])
            src = np.fromiter((s for e in edges for s in e[0]), np.int8, count=int(eptr[-1]))
            dst = np.fromiter((t for e in edges for t in e[1]), np.int8, count=int(eptr[-1]))
            ne = max(1, int(eptr[-1]))
            adj = np.zeros(B, np.uint64); indptr = np.zeros(7 * B + 1, np.int32)
            indices = np.zeros(ne, np.int32); eflags = np.zeros(ne, np.uint8)
            level = np.zeros((B, 7), np.uint8); rows = np.zeros(6 * B, np.int32)
            nl = _abi.I32(0)
            pv = lambda a: a.ctypes.data
            _lib.check(L.dxvae_batch_build_host(B, pv(eptr), pv(src), pv(dst), pv(adj), pv(indptr), pv(indices),
                                                pv(eflags), pv(level), pv(d.level_ptr), pv(rows), ctypes.byref(nl)),
                       "dxvae_batch_build_host")
            d.n_levels = int(nl.value)
            d.adj = torch.from_numpy(adj.view(np.int64)).to("cuda")
            d.level_rows = torch.from_numpy(rows).to("cuda")
            d.level = level
            d.csr = (indptr, indices[:int(eptr[-1])], eflags[:int(eptr[-1])])
            d.step_ptr = d.step_rows = None
            if need_cls and self.compact_steps:
                d.step_ptr = np.zeros(34, np.int32)
                srows = np.zeros(33 * B, np.int32)
                _lib.check(L.dxvae_batch_steps_host(B, pv(adj), pv(d.step_ptr), pv(srows)), "dxvae_batch_steps_host")
                d.step_rows = torch.from_numpy(srows[:max(1, int(d.step_ptr[33]))]).to("cuda")
        else:
            if not indexed:
                d.adj = gb.adj.to("cuda", torch.int64, non_blocking=True).contiguous()
            self._schedule(d)
            d.step_ptr = d.step_rows = None
            if need_cls and self.compact_steps:
                d.step_ptr = np.zeros(34, np.int32)
                d.step_rows = torch.empty(33 * B, dtype=torch.int32, device="cuda")
                sp_dev = torch.empty(34, dtype=torch.int32, device="cuda")
                ws = self._workspace(_abi.OP_SCHEDULE, B, tag=self._prep_tag)
                _lib.check(L.dxvae_batch_steps(B, d.adj.data_ptr(), sp_dev.data_ptr(), d.step_rows.data_ptr(),
                                               d.step_ptr.ctypes.data, ws.data_ptr(), ws.numel(), st), "dxvae_batch_steps")
        return d

    def _prepare_uploading(self, G):
        """_prepare for HOST-resident graphs on a side stream: the batcher's H2D copies, its pack / schedule kernels and the
        two small read-backs that size the schedule (which synchronise their stream) then overlap the compute the caller
        has already queued — typically the previous training step — instead of waiting behind it.  The compute stream
        picks the batch up through an event."""
        if isinstance(G, DXGraphBatch) and G.X.is_cuda:
            return self._prepare(G, need_cls=True)       # device-resident input: stay ordered with its producer
        cur = torch.cuda.current_stream()
        if self._upload_stream is None:
            self._upload_stream = torch.cuda.Stream()
        side = self._upload_stream
        self._prep_tag = "upload"
        try:
            with torch.cuda.stream(side):
                d = self._prepare(G, need_cls=True)
        finally:
            self._prep_tag = None
        for t in (d.Xn, d.cls, d.adj, d.level_rows, d.step_rows, d.level):
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(cur)                     # allocated on the upload stream, consumed on the compute stream
        cur.wait_stream(side)
        return d

    def _schedule(self, d):
        L = _lib.lib()
        B = d.B
        d.level = torch.empty(B, 7, dtype=torch.uint8, device="cuda")
        d.level_rows = torch.empty(6 * B, dtype=torch.int32, device="cuda")
        lp_dev = torch.empty(16, dtype=torch.int32, device="cuda")     # 8 offsets + 8 per-level prefix sizes (ABI: 16 ints)
        ws = self._workspace(_abi.OP_SCHEDULE, B, tag=self._prep_tag)
        _lib.check(L.dxvae_batch_schedule(B, d.adj.data_ptr(), d.level.data_ptr(), lp_dev.data_ptr(),
                                          d.level_rows.data_ptr(), d.level_ptr.ctypes.data, ws.data_ptr(), ws.numel(),
                                          _stream()), "dxvae_batch_schedule")
        nz = [l for l in range(7) if d.level_ptr[l + 1] > d.level_ptr[l]]
        d.n_levels = (max(nz) + 1) if nz else 1

    def _params_tuple(self):
        return tuple(dict(self.named_parameters())[n] for n, _, _ in self._table)

    def _grad_views(self, gflat, scale=None):
        out = []
        for _, off, shape in self._table:
            k = int(np.prod(shape))
            g = gflat[off:off + k].view(shape)
            out.append(g if scale is None else g * scale)
        return tuple(out)

    # ------------------------------------------------------------------ encode
    def encode(self, G):
        """model.py:200-212.  Returns Normal(mu, std) over the batch (device tensors).  With
        autograd enabled the result is differentiable w.r.t. the encoder parameters."""
        L = self._ensure_flat()
        gb = DXGraphBatch.from_graphs(G)
        B = len(gb)
        self.hidden = B                                   # model.py:201 sizes the scratch state here
        if torch.is_grad_enabled() and B > self.max_chunk and any(p.requires_grad for p in self.parameters()):
            raise ValueError("encode() with autograd enabled takes at most max_chunk=%d graphs per call (got %d); "
                             "use torch.no_grad() for inference or raise max_chunk" % (self.max_chunk, B))
        if torch.is_grad_enabled():
            d = self._prepare(gb, need_cls=True)
            mu, sd = _EncodeFn.apply(self, d, *self._params_tuple())
            q = Normal(mu, sd, validate_args=False)
            q._dx_batch = (d, G)                          # reused by loss(q, G) when it is handed the same G
            return q
        mu = torch.empty(B, 128, device="cuda"); sd = torch.empty(B, 128, device="cuda")
        for lo in range(0, B, self.max_chunk):
            hi = min(B, lo + self.max_chunk)
            d = self._prepare(gb[lo:hi], need_cls=False)
            ws = self._workspace(_abi.OP_ENCODE, d.B)
            _lib.check(L.dxvae_encode_fwd(self._flat.data_ptr(), d.B, d.Xn.data_ptr(), d.adj.data_ptr(), d.n_levels,
                                          d.level_ptr.ctypes.data, d.level_rows.data_ptr(), d.level_ptr[8:].ctypes.data, mu[lo:hi].data_ptr(),
                                          sd[lo:hi].data_ptr(), ws.data_ptr(), ws.numel(), 0,
                                          {"fp32": _abi.PREC_FP32, "tf32": _abi.PREC_TF32, "3xtf32": _abi.PREC_3XTF32}[self.encode_precision],
                                          _stream()),
                       "dxvae_encode_fwd")
        return Normal(mu, sd, validate_args=False)

    # ------------------------------------------------------------------ reparameterize
    def reparameterize(self, q_dist, eps=None):
        """z = mu + std * eps (what q_dist.rsample() computes, model.py:284).  eps=None draws
        from torch's global generator exactly as Normal.rsample does on this device."""
        L = _lib.require_cuda()
        mu, sd = (q_dist.loc, q_dist.scale) if isinstance(q_dist, Normal) else q_dist
        mu = mu.detach().contiguous(); sd = sd.detach().contiguous()
        if eps is None:
            eps = torch.empty_like(mu).normal_()
        eps = eps.to(mu.device, torch.float32).contiguous()
        z = torch.empty_like(mu)
        _lib.check(L.dxvae_reparameterize(mu.numel(), mu.data_ptr(), sd.data_ptr(), eps.data_ptr(), z.data_ptr(),
                                          _stream()), "dxvae_reparameterize")
        return z

    # ------------------------------------------------------------------ decode
    @torch.no_grad()
    def decode(self, z):
        """model.py:214-253: greedy generation.  Returns a DXGraphBatch (list-like of graphs with
        ndata['X'], ndata['params'], edges() in the reference's insertion order)."""
        L = self._ensure_flat()
        z = torch.as_tensor(z).to("cuda", torch.float32).contiguous()
        B = z.shape[0]
        Xg = torch.empty(B, 7, 27, device="cuda"); Pg = torch.empty(B, 7, 21, device="cuda")
        adj = torch.empty(B, dtype=torch.int64, device="cuda"); mg = torch.empty(B, 2, device="cuda")
        for lo in range(0, B, self.max_chunk):
            hi = min(B, lo + self.max_chunk)
            ws = self._workspace(_abi.OP_DECODE, hi - lo)
            _lib.check(L.dxvae_decode_greedy(self._flat.data_ptr(), hi - lo, z[lo:hi].data_ptr(), Xg[lo:hi].data_ptr(),
                                             Pg[lo:hi].data_ptr(), adj[lo:hi].data_ptr(), mg[lo:hi].data_ptr(),
                                             ws.data_ptr(), ws.numel(),
                                             {"fp32": _abi.PREC_FP32, "3xtf32": _abi.PREC_3XTF32}[self.decode_precision],
                                             _stream()), "dxvae_decode_greedy")
        self.last_margins = mg[:, 0]          # min |edge logit| per graph
        self.last_quant_margins = mg[:, 1]    # min distance of a parameter logit to a rounding / arg-max tie
        return DXGraphBatch(Xg, Pg, adj)

    def encode_decode(self, G_true, stochastic=False):
        """model.py:255-262."""
        with torch.no_grad():
            q_dist = self.encode(G_true)
            z = q_dist.sample() if stochastic else q_dist.loc
        return self.decode(z)

    def generate(self, n):
        """model.py:264-268 (prior sample drawn on the CPU generator, as the reference does)."""
        self.hidden = n
        sample = self.p_dist.sample((n, self.size_Z)).to("cuda")
        return self.decode(sample)

    # ------------------------------------------------------------------ loss / forward
    def loss(self, q_dist, G_true, w_env=2, w_frq=5, w_kld=0.01, eps=None):
        """model.py:270-367.  `eps` injects the reparameterisation noise; None draws it the way
        q_dist.rsample() would (the reference always samples: App. C.1)."""
        self._ensure_flat()
        cached = getattr(q_dist, "_dx_batch", None)
        if cached is not None and cached[1] is G_true:   # the targets must come from G_true itself (model.py:272-280)
            d = cached[0]
        else:
            d = self._prepare(G_true, need_cls=True)
        mu, sd = q_dist.loc, q_dist.scale
        if eps is None:
            eps = torch.empty(mu.shape, device="cuda").normal_()
        eps = torch.as_tensor(eps).to("cuda", torch.float32).contiguous()
        need = torch.is_grad_enabled()
        out = _LossFn.apply(self, d, eps, (float(w_env), float(w_frq), float(w_kld), need), mu, sd,
                            *self._params_tuple())
        total, rest = out[0], out[1]
        return total, rest[0], rest[1], rest[2], rest[3]

    def forward(self, G_true, w_env=2, w_frq=5, w_kld=0.01, eps=None):
        """model.py:369-372: encode + loss, fused into one native call (dxvae_elbo_step)."""
        self._ensure_flat()
        d = self._prepare_uploading(G_true)
        self.hidden = d.B
        if eps is None:
            eps = torch.empty(d.B, 128, device="cuda").normal_()
        eps = torch.as_tensor(eps).to("cuda", torch.float32).contiguous()
        need = torch.is_grad_enabled()
        out = _ElboFn.apply(self, d, eps, (float(w_env), float(w_frq), float(w_kld), need), *self._params_tuple())
        total, rest = out[0], out[1]
        return total, rest[0], rest[1], rest[2], rest[3]

    def elbo_step(self, d, eps, w, grads=None, inv_batch=None, mu_out=None, std_out=None, loss5=None, decoder_done=None):
        """Raw fused step on a prepared batch: returns loss5 (device tensor of 5, or the caller's); accumulates
        into `grads` (flat blob) when given.  decoder_done (torch.cuda.Event): recorded on the stream once the decoder's
        backward has been issued, i.e. when the decoder-only gradient range is final (data-parallel overlap)."""
        L = _lib.lib()
        if loss5 is None:
            loss5 = torch.empty(5, device="cuda")
        ws = self._workspace(_abi.OP_TRAIN, d.B, d=d)
        _lib.check(L.dxvae_elbo_step(
            self._flat.data_ptr(), d.B, d.Xn.data_ptr(), d.cls.data_ptr(), d.adj.data_ptr(), d.n_levels,
            d.level_ptr.ctypes.data, d.level_rows.data_ptr(), d.level_ptr[8:].ctypes.data, eps.data_ptr(), w[0], w[1], w[2],
            (1.0 / d.B) if inv_batch is None else inv_batch, loss5.data_ptr(),
            None if mu_out is None else mu_out.data_ptr(), None if std_out is None else std_out.data_ptr(),
            None if grads is None else grads.data_ptr(), ws.data_ptr(), ws.numel(), self._prec(),
            None if d.step_ptr is None else d.step_ptr.ctypes.data,
            None if d.step_ptr is None else d.step_rows.data_ptr(),
            None if decoder_done is None else decoder_done.cuda_event, _stream()), "dxvae_elbo_step")
        return loss5

    # ------------------------------------------------------------------ train
    def train(self, G_true, epochs, size_batch=32, lr=0.001, checkpoint=None, w_env=2, w_frq=5, w_kld=0.01):
        """model.py:374-391: AdamW(lr), epochs+1 passes, in-place shuffle with Python's `random`,
        drop-last batching, checkpoint per epoch.  (Shadows nn.Module.train as the reference
        does.)  Under torch.distributed every global batch is sharded in equal contiguous
        slices across ranks and the flat gradient is all-reduced (sum) over NCCL."""
        from .train import Trainer
        import torch.distributed as dist
        t = Trainer(self, lr=lr, w=(float(w_env), float(w_frq), float(w_kld)))   # (data parallel: broadcasts rank 0's weights)
        n_samples = len(G_true)
        n_iters = int(n_samples / size_batch)
        data = t.upload(G_true)
        order = list(range(n_samples))
        for epoch in range(epochs + 1):
            if self.verbose:
                print(f'Epoch: {epoch}')
            perm = list(range(n_samples))
            random.shuffle(perm)                      # same RNG consumption as random.shuffle(G_true)
            if t.world > 1:                           # one permutation for all ranks (rank 0's), whatever their seeds
                box = [perm]
                dist.broadcast_object_list(box, src=0)
                perm = box[0]
            order = [order[i] for i in perm]
            if isinstance(G_true, list):
                G_true[:] = [G_true[i] for i in perm]  # the reference shuffles the caller's list in place
            for i in range(n_iters):
                idx = order[i * size_batch:(i + 1) * size_batch]
                loss5 = t.step(data, idx)
                if self.verbose:
                    l = loss5.tolist()
                    print(f'batch: {i}\tloss: {l[0]:.4f}\tx0: {l[1]:.4f}\txi: {l[2]:.4f}\te: {l[3]:.4f}\tkld: {l[4]:.4f}')
            if checkpoint is not None:
                if t.rank == 0:                       # replicas hold identical weights: one writer
                    torch.save(self.state_dict(), checkpoint)
                if t.world > 1:
                    dist.barrier()
                if self.verbose:
                    print(f'\nCheckpoint [{checkpoint}] saved\n')
        if self.verbose:
            print('Finished Training')


# =============================================================================================
# autograd glue: gradients are produced by the native backward kernels, never by autograd
# =============================================================================================
class _ElboFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, d, eps, w, *params):
        need = w[3] and any(p.requires_grad for p in params)   # grad mode is off inside forward(): flag from caller
        g = torch.zeros(model._total, device="cuda") if need else None
        loss5 = model.elbo_step(d, eps, w, grads=g)
        ctx.model, ctx.g = model, g
        model.last_loss5 = loss5
        total, rest = loss5[0].clone(), loss5[1:].clone()
        ctx.mark_non_differentiable(rest)
        return total, rest

    @staticmethod
    def backward(ctx, gtotal, _grest):
        # one in-place scale of the flat blob (the chain rule's upstream factor), then views of it: the parameters'
        # .grad tensors alias ONE buffer, which FusedAdamW / the data-parallel all-reduce consume without a gather
        if getattr(ctx, "consumed", False):
            raise RuntimeError("DXVAE.forward(): backward() through the fused step can run once (its gradient blob is scaled in place)")
        ctx.consumed = True
        ctx.g.mul_(gtotal)
        ctx.model._last_gflat = ctx.g
        return (None, None, None, None) + ctx.model._grad_views(ctx.g)


class _EncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, d, *params):
        L = _lib.lib()
        ws = model._workspace(_abi.OP_ENCODE_TRAIN, d.B, fresh=True)
        mu = torch.empty(d.B, 128, device="cuda"); sd = torch.empty(d.B, 128, device="cuda")
        _lib.check(L.dxvae_encode_fwd(model._flat.data_ptr(), d.B, d.Xn.data_ptr(), d.adj.data_ptr(), d.n_levels,
                                      d.level_ptr.ctypes.data, d.level_rows.data_ptr(), d.level_ptr[8:].ctypes.data, mu.data_ptr(), sd.data_ptr(),
                                      ws.data_ptr(), ws.numel(), 1, model._prec(), _stream()), "dxvae_encode_fwd")
        ctx.model, ctx.d, ctx.ws, ctx.sd = model, d, ws, sd
        return mu, sd

    @staticmethod
    def backward(ctx, dmu, dsd):
        model, d = ctx.model, ctx.d
        L = _lib.lib()
        g = torch.zeros(model._total, device="cuda")
        dmu = torch.zeros(d.B, 128, device="cuda") if dmu is None else dmu.contiguous()
        dsd = torch.zeros(d.B, 128, device="cuda") if dsd is None else dsd.contiguous()
        _lib.check(L.dxvae_encode_bwd(model._flat.data_ptr(), d.B, d.Xn.data_ptr(), d.adj.data_ptr(), d.n_levels,
                                      d.level_ptr.ctypes.data, d.level_rows.data_ptr(), d.level_ptr[8:].ctypes.data, ctx.sd.data_ptr(),
                                      dmu.data_ptr(), dsd.data_ptr(), g.data_ptr(), ctx.ws.data_ptr(), ctx.ws.numel(),
                                      model._prec(), _stream()), "dxvae_encode_bwd")
        return (None, None) + model._grad_views(g)


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, d, eps, w, mu, sd, *params):
        L = _lib.lib()
        mu = mu.detach().contiguous(); sd = sd.detach().contiguous()
        need = w[3]
        g = torch.zeros(model._total, device="cuda") if need else None
        dmu = torch.empty(d.B, 128, device="cuda") if need else None
        dsd = torch.empty(d.B, 128, device="cuda") if need else None
        loss5 = torch.empty(5, device="cuda")
        ws = model._workspace(_abi.OP_LOSS, d.B, d=d)
        _lib.check(L.dxvae_loss_step(
            model._flat.data_ptr(), d.B, d.Xn.data_ptr(), d.cls.data_ptr(), d.adj.data_ptr(), mu.data_ptr(),
            sd.data_ptr(), eps.data_ptr(), w[0], w[1], w[2], 1.0 / d.B, loss5.data_ptr(),
            None if g is None else g.data_ptr(), None if dmu is None else dmu.data_ptr(),
            None if dsd is None else dsd.data_ptr(), ws.data_ptr(), ws.numel(), model._prec(),
            None if d.step_ptr is None else d.step_ptr.ctypes.data,
            None if d.step_ptr is None else d.step_rows.data_ptr(), _stream()), "dxvae_loss_step")
        ctx.model, ctx.g, ctx.dmu, ctx.dsd = model, g, dmu, dsd
        total, rest = loss5[0].clone(), loss5[1:].clone()
        ctx.mark_non_differentiable(rest)
        return total, rest

    @staticmethod
    def backward(ctx, gtotal, _grest):
        if getattr(ctx, "consumed", False):
            raise RuntimeError("DXVAE.loss(): backward() through the fused step can run once (its gradient blob is scaled in place)")
        ctx.consumed = True
        ctx.g.mul_(gtotal)
        return (None, None, None, None, ctx.dmu * gtotal, ctx.dsd * gtotal) + ctx.model._grad_views(ctx.g)

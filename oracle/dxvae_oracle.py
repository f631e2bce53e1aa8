"""CPU oracle for the DX-VAE hot path.  TEST INFRASTRUCTURE ONLY.

A dense-tensor restatement (torch, fp32, CPU) of what the reference computes on
lists of DGL graphs.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product package
(dxvae_b200/) never does and fails loudly without its CUDA extension.

Parity status: the reference ships no tests or golden vectors for the model
arithmetic ("parity unpinned" by the reference itself).  This file
is therefore pinned against the reference *executed here* (oracle/ref_loader.py +
oracle/shim): oracle/make_golden.py checks every function below against the
unmodified /root/reference/model.py and writes tests/golden/*.npz; the data
format functions are pinned by the reference's own artefacts
(DX_data/DXDataset.bin, generated/gen_patch.syx).

Tensor conventions: X (B,7,27) f32, params (B,7,21) f32, A (B,7,7) with
A[b,src,dst]=1 (dgl adj(): row=src, col=dst, model.py:279).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

N_NODES, N_PARAMS, SIZE_X, SIZE_X0, SIZE_H, SIZE_Z = 7, 21, 27, 23, 512, 128


class OracleDXVAE(nn.Module):
    """Same 46-tensor parameter layout / registration order as model.py:24-72,
    so torch.manual_seed(k) initialises it identically to the reference and
    state_dicts are interchangeable."""

    def __init__(self):
        super().__init__()
        H, Z, X, X0 = SIZE_H, SIZE_Z, SIZE_X, SIZE_X0
        self.combin_encode = nn.GRUCell(X, H)      # model.py:24
        self.loop_encode = nn.GRUCell(X, H)        # model.py:25
        self.root_encode = nn.GRUCell(X0, H)       # model.py:26
        self.h_to_mu = nn.Linear(H, Z)             # model.py:27
        self.h_to_std = nn.Sequential(nn.Linear(H, Z), nn.Softplus())  # model.py:28-30
        self.combin_decode = nn.GRUCell(X, H)      # model.py:33
        self.loop_decode = nn.GRUCell(X, H)        # model.py:34
        self.root_decode = nn.GRUCell(X0, H)       # model.py:35
        self.z_to_h = nn.Sequential(nn.Linear(Z, H), nn.Tanh())  # model.py:36-39
        self.h_to_x0 = nn.Sequential(nn.Linear(H, 2 * H), nn.ReLU(), nn.Linear(2 * H, 2 * H), nn.ReLU(),
                                     nn.Linear(2 * H, X0 + 32))  # model.py:40-46
        self.h_to_x = nn.Sequential(nn.Linear(H, 2 * H), nn.ReLU(), nn.Linear(2 * H, 2 * H), nn.ReLU(),
                                    nn.Linear(2 * H, X))  # model.py:47-53
        self.h_to_edge_self = nn.Sequential(nn.Linear(H, 2 * H), nn.ReLU(), nn.Linear(2 * H, 1))  # model.py:54-58
        self.h_to_edge = nn.Sequential(nn.Linear(2 * H, 4 * H), nn.ReLU(), nn.Linear(4 * H, 2))  # model.py:59-63
        self.gate = nn.Sequential(nn.Linear(2 * H, H), nn.Sigmoid())  # model.py:66-69
        self.mapper = nn.Sequential(nn.Linear(2 * H, H, bias=False))  # model.py:70-72

    # ------------------------------------------------------------------ propagate
    def _message_sum(self, hs, i_flags, o_flags):
        """model.py:163-181.  hs: list of (B,512) neighbour states in slot order;
        i_flags[k] (B,) = neighbour k is a predecessor, o_flags[k] = successor.
        Zero-padded slots are run through gate/mapper exactly as the reference does."""
        forth = torch.stack([h * i.unsqueeze(1) for h, i in zip(hs, i_flags)], 1)
        back = torch.stack([h * o.unsqueeze(1) for h, o in zip(hs, o_flags)], 1)
        h_in = torch.cat([forth, back], 2)
        return (self.gate(h_in) * self.mapper(h_in)).sum(1)

    def _node_update(self, x, h_in, self_loop, v, encode):
        """model.py:183-193."""
        if encode:
            rooter, combiner, looper = self.root_encode, self.combin_encode, self.loop_encode
        else:
            rooter, combiner, looper = self.root_decode, self.combin_decode, self.loop_decode
        if v == 0:
            return rooter(x[:, :SIZE_X0], h_in)
        x_loop = x * self_loop.unsqueeze(1)
        return looper(x_loop, combiner(x, h_in))

    # ------------------------------------------------------------------ encode
    def encode(self, X, A):
        """model.py:200-212 -> (mu, std)."""
        B = X.shape[0]
        A = A.to(X.dtype)
        hid = [None] * N_NODES
        for v in range(N_NODES - 1, -1, -1):
            if v == N_NODES - 1:
                h_in = X.new_zeros(B, SIZE_H)
            else:
                nb = list(range(v + 1, N_NODES))
                h_in = self._message_sum([hid[x] for x in nb], [A[:, x, v] for x in nb], [A[:, v, x] for x in nb])
            hid[v] = self._node_update(X[:, v], h_in, A[:, v, v], v, True)
        self.enc_hidden = hid
        return self.h_to_mu(hid[0]), self.h_to_std(hid[0])

    # ------------------------------------------------------------------ loss
    def loss(self, mu, std, X, params, A, eps, w_env=2, w_frq=5, w_kld=0.01):
        """model.py:270-367 with z = mu + std*eps (what rsample draws, :284)."""
        B = X.shape[0]
        A = A.to(X.dtype)
        P = params.long()
        bce = lambda a, t: F.binary_cross_entropy_with_logits(a, t, reduction="none")
        ce = lambda a, t: F.cross_entropy(a, t, reduction="none")
        mse = lambda a, t: F.mse_loss(a, t, reduction="none")

        z = mu + std * eps
        h_init = self.z_to_h(z)
        L0 = self.h_to_x0(h_init)
        X0t = X[:, 0]
        hid = [None] * N_NODES
        hid[0] = self._node_update(X0t, h_init, A[:, 0, 0], 0, False)
        lx0 = mse(L0[:, :8] * w_env, X0t[:, :8] * w_env).mean(0).sum()
        lx0 = lx0 + mse(L0[:, 8] * w_frq, X0t[:, 8] * w_frq).mean(0).sum()
        lx0 = lx0 + mse(L0[:, 9:15], X0t[:, 9:15]).mean(0).sum()
        lx0 = lx0 + bce(L0[:, 15:17], X0t[:, 15:17]).mean(0).sum()
        lx0 = lx0 + ce(L0[:, 17:23], P[:, 0, 17]).mean()
        lx0 = lx0 + ce(L0[:, 23:], P[:, 0, 18]).mean()
        lxi = 0
        le = 0
        zero = X.new_zeros(B)
        for vi in range(1, N_NODES):
            Li = self.h_to_x(hid[vi - 1])
            Xt = X[:, vi]
            Hi = self._node_update(Xt, X.new_zeros(B, SIZE_H), zero, vi, False)   # :320 (no edges yet)
            lxi = lxi + mse(Li[:, :9] * w_env, Xt[:, :9] * w_env).mean(0).sum()
            lxi = lxi + mse(Li[:, 9] * w_frq, Xt[:, 9] * w_frq).mean(0).sum()
            lxi = lxi + mse(Li[:, 10:18], Xt[:, 10:18]).mean(0).sum()
            lxi = lxi + bce(Li[:, 18], Xt[:, 18]).mean()
            lxi = lxi + ce(Li[:, 19:23], P[:, vi, 19]).mean()
            lxi = lxi + ce(Li[:, 23:27], P[:, vi, 20]).mean()
            ls = self.h_to_edge_self(Hi)
            s = A[:, vi, vi]
            Hi = self._node_update(Xt, X.new_zeros(B, SIZE_H), s, vi, False)      # :337
            le = le + bce(ls, s.unsqueeze(1)).mean()
            hs, fi, fo, Ei = [], [], [], []
            for vj in range(vi - 1, -1, -1):
                Ei.append(self.h_to_edge(torch.cat([Hi, hid[vj]], -1)).unsqueeze(1))
                hs.append(hid[vj]); fi.append(A[:, vj, vi]); fo.append(A[:, vi, vj])
                # the reference pads the not-yet-visited slots with zeros: they add exact 0
                h_in = self._message_sum(hs, fi, fo)
                Hi = self._node_update(Xt, h_in, s, vi, False)                    # :358
            Ei.reverse()
            Ei = torch.cat(Ei, 1)
            Et = torch.stack([A[:, :vi, vi], A[:, vi, :vi]], 2)
            le = le + bce(Ei, Et).mean(0).sum()
            hid[vi] = Hi
        # KL(prior || posterior), model.py:365 (closed form of _kl_normal_normal(p, q))
        var_ratio = (1.0 / std) ** 2
        t1 = (mu / std) ** 2
        kld = (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).mean(0).sum()
        return lx0 + lxi + le + kld * w_kld, lx0, lxi, le, kld * w_kld

    # ------------------------------------------------------------------ decode
    @staticmethod
    def _q_lin(x, scale):                      # model.py:87-91
        p = (x * scale).round().clamp(0, scale)
        return p / scale, p

    @staticmethod
    def _q_log(x, scale):                      # model.py:93-98
        ls = torch.tensor(scale + 1).log()
        p = ((x * ls).exp() - 1).round().clamp(0, scale)
        return (p + 1).log() / ls, p

    def _reg_x0(self, L0):                     # model.py:109-125
        B = L0.shape[0]
        X0 = torch.zeros(B, SIZE_X); p0 = torch.zeros(B, N_PARAMS)
        X0[:, :8], p0[:, :8] = self._q_lin(L0[:, :8], 99)
        X0[:, 8], p0[:, 8] = self._q_lin(L0[:, 8], 48)
        X0[:, 9:13], p0[:, 9:13] = self._q_lin(L0[:, 9:13], 99)
        X0[:, 13:15], p0[:, 13:15] = self._q_lin(L0[:, 13:15], 7)
        b = L0[:, 15:17].sigmoid().round()
        X0[:, 15:17], p0[:, 15:17] = b, b
        lfw = L0[:, 17:23].argmax(1)
        X0[:, 17:23], p0[:, 17] = F.one_hot(lfw, 6).float(), lfw.float()
        p0[:, 18] = L0[:, 23:55].argmax(1).float()
        return X0, p0

    def _reg_xi(self, Li):                     # model.py:127-149
        B = Li.shape[0]
        Xi = Li.clone(); pi = torch.zeros(B, N_PARAMS)
        Xi[:, :9], pi[:, :9] = self._q_lin(Li[:, :9], 99)
        Xi[:, 11], pi[:, 11] = self._q_lin(Li[:, 11], 14)
        Xi[:, 12:15], pi[:, 12:15] = self._q_lin(Li[:, 12:15], 99)
        Xi[:, 15], pi[:, 15] = self._q_lin(Li[:, 15], 3)
        Xi[:, 16:18], pi[:, 16:18] = self._q_lin(Li[:, 16:18], 7)
        m = Li[:, 18].sigmoid().round()
        Xi[:, 18], pi[:, 18] = m, m
        lc = Li[:, 19:23].argmax(1)
        Xi[:, 19:23], pi[:, 19] = F.one_hot(lc, 4).float(), lc.float()
        rc = Li[:, 23:26].argmax(1)            # the 23:26 quirk, model.py:139
        Xi[:, 23:27], pi[:, 20] = F.one_hot(rc, 4).float(), rc.float()
        xl9, pl9 = self._q_log(Li[:, 9], 31); xq9, pq9 = self._q_lin(Li[:, 9], 3)
        xl10, pl10 = self._q_log(Li[:, 10], 99); xq10, pq10 = self._q_lin(Li[:, 10], 99)
        ratio = m == 0
        Xi[:, 9] = torch.where(ratio, xl9, xq9); pi[:, 9] = torch.where(ratio, pl9, pq9)
        Xi[:, 10] = torch.where(ratio, xl10, xq10); pi[:, 10] = torch.where(ratio, pl10, pq10)
        return Xi, pi

    @torch.no_grad()
    def decode(self, z, return_margins=False):
        """model.py:214-253 -> X (B,7,27), params (B,7,21), A (B,7,7) uint8.
        Edge insertion order is canonical given A: for vi=1..6: (vi,vi)?, then
        for vj=vi-1..0: (vj,vi)?, (vi,vj)?  (see edges_from_adj)."""
        B = z.shape[0]
        h_init = self.z_to_h(z)
        L0 = self.h_to_x0(h_init)
        X0, p0 = self._reg_x0(L0)
        Xs = [X0]; Ps = [p0]
        A = torch.zeros(B, N_NODES, N_NODES)
        hid = [None] * N_NODES
        hid[0] = self._node_update(X0, h_init, A[:, 0, 0], 0, False)
        margins = {"edge": [], "self": [], "logits": [L0]}
        zero = torch.zeros(B)
        for vi in range(1, N_NODES):
            Li = self.h_to_x(hid[vi - 1])
            margins["logits"].append(Li)
            Xi, pi = self._reg_xi(Li)
            Xs.append(Xi); Ps.append(pi)
            Hi = self._node_update(Xi, torch.zeros(B, SIZE_H), zero, vi, False)
            ls = self.h_to_edge_self(Hi)
            margins["self"].append(ls)
            s = (ls.sigmoid() > 0.5).float().squeeze(1)
            A[:, vi, vi] = s
            Hi = self._node_update(Xi, torch.zeros(B, SIZE_H), s, vi, False)
            hs, fi, fo = [], [], []
            for vj in range(vi - 1, -1, -1):
                lg = self.h_to_edge(torch.cat([Hi, hid[vj]], -1))
                margins["edge"].append(lg)
                e = (lg.sigmoid() > 0.5).float()
                A[:, vj, vi] = e[:, 0]; A[:, vi, vj] = e[:, 1]
                hs.append(hid[vj]); fi.append(e[:, 0]); fo.append(e[:, 1])
                Hi = self._node_update(Xi, self._message_sum(hs, fi, fo), s, vi, False)
            hid[vi] = Hi
        out = (torch.stack(Xs, 1), torch.stack(Ps, 1), A.to(torch.uint8))
        return out + (margins,) if return_margins else out


def edges_from_adj(A):
    """Edge list of one decoded graph in the reference's insertion order
    (model.py:237-250).  A: (7,7) array-like, A[src,dst]."""
    src, dst = [], []
    for vi in range(1, N_NODES):
        if A[vi][vi]:
            src.append(vi); dst.append(vi)
        for vj in range(vi - 1, -1, -1):
            if A[vj][vi]:
                src.append(vj); dst.append(vi)
            if A[vi][vj]:
                src.append(vi); dst.append(vj)
    return src, dst


# =============================================================================
# dxdata.py restatements (integer/byte work: numpy)
# =============================================================================
# dxdata.py:140-171  DX_ALGO: alg -> (src[], dst[])
DX_ALGO = {
    0: ([1, 2, 3, 4, 5, 6, 6], [0, 1, 0, 3, 4, 5, 6]), 1: ([1, 2, 2, 3, 4, 5, 6], [0, 1, 2, 0, 3, 4, 5]),
    2: ([1, 2, 3, 4, 5, 6, 6], [0, 1, 2, 0, 4, 5, 6]), 3: ([1, 2, 3, 4, 4, 5, 6], [0, 1, 2, 0, 6, 4, 5]),
    4: ([1, 2, 3, 4, 5, 6, 6], [0, 1, 0, 3, 0, 5, 6]), 5: ([1, 2, 3, 4, 5, 5, 6], [0, 1, 0, 3, 0, 6, 5]),
    6: ([1, 2, 3, 4, 5, 6, 6], [0, 1, 0, 3, 3, 5, 6]), 7: ([1, 2, 3, 4, 4, 5, 6], [0, 1, 0, 3, 4, 3, 5]),
    8: ([1, 2, 2, 3, 4, 5, 6], [0, 1, 2, 0, 3, 3, 5]), 9: ([1, 2, 3, 3, 4, 5, 6], [0, 1, 2, 3, 0, 4, 4]),
    10: ([1, 2, 3, 4, 5, 6, 6], [0, 1, 2, 0, 4, 4, 6]), 11: ([1, 2, 2, 3, 4, 5, 6], [0, 1, 2, 0, 3, 3, 3]),
    12: ([1, 2, 3, 4, 5, 6, 6], [0, 1, 0, 3, 3, 3, 6]), 13: ([1, 2, 3, 4, 5, 6, 6], [0, 1, 0, 3, 4, 4, 6]),
    14: ([1, 2, 2, 3, 4, 5, 6], [0, 1, 2, 0, 3, 4, 4]), 15: ([1, 2, 3, 4, 5, 6, 6], [0, 1, 1, 3, 1, 5, 6]),
    16: ([1, 2, 2, 3, 4, 5, 6], [0, 1, 2, 1, 3, 1, 5]), 17: ([1, 2, 3, 3, 4, 5, 6], [0, 1, 1, 3, 1, 4, 5]),
    18: ([1, 2, 3, 4, 5, 6, 6, 6], [0, 1, 2, 0, 0, 4, 5, 6]), 19: ([1, 2, 3, 3, 3, 4, 5, 6], [0, 0, 1, 2, 3, 0, 4, 4]),
    20: ([1, 2, 3, 3, 3, 4, 5, 6, 6], [0, 0, 1, 2, 3, 0, 0, 4, 5]),
    21: ([1, 2, 3, 4, 5, 6, 6, 6, 6], [0, 1, 0, 0, 0, 3, 4, 5, 6]),
    22: ([1, 2, 3, 4, 5, 6, 6, 6], [0, 0, 2, 0, 0, 4, 5, 6]),
    23: ([1, 2, 3, 4, 5, 6, 6, 6, 6], [0, 0, 0, 0, 0, 3, 4, 5, 6]),
    24: ([1, 2, 3, 4, 5, 6, 6, 6], [0, 0, 0, 0, 0, 4, 5, 6]), 25: ([1, 2, 4, 3, 5, 6, 6], [0, 0, 0, 2, 4, 4, 6]),
    26: ([1, 2, 3, 3, 4, 5, 6], [0, 0, 2, 3, 0, 4, 4]), 27: ([1, 2, 3, 4, 5, 5, 6], [0, 1, 0, 3, 4, 5, 0]),
    28: ([1, 2, 3, 4, 5, 6, 6], [0, 0, 0, 3, 0, 5, 6]), 29: ([1, 2, 3, 4, 5, 5, 6], [0, 0, 0, 3, 4, 5, 0]),
    30: ([1, 2, 3, 4, 5, 6, 6], [0, 0, 0, 0, 0, 5, 6]), 31: ([1, 2, 3, 4, 5, 6, 6], [0, 0, 0, 0, 0, 0, 6]),
}


def make_graph(pz):
    """dxdata.py:174-312 for one packed 128-byte voice (array of ints).
    Returns (X (7,27) f32, params (7,21) f32, src list, dst list).  Float steps use
    torch scalar ops so the f32 rounding is the reference's."""
    pz = [int(b) for b in pz]
    f = lambda v: torch.tensor(v, dtype=torch.int64)
    clampi = lambda v, lo, hi: max(lo, min(hi, v))
    X = torch.zeros(7, 27); Pm = torch.zeros(7, 21)
    for k in range(1, 7):                                   # parse_op, :175-244
        i = (6 - k) * 17
        env = [clampi(b, 0, 99) for b in pz[i:i + 8]]
        bp, ld, rd = (clampi(pz[i + j], 0, 99) for j in (8, 9, 10))
        rc, lc = (pz[i + 11] // 4) % 4, pz[i + 11] % 4
        det, rs = clampi(pz[i + 12] // 8, 0, 14), pz[i + 12] % 8
        kvs, ams = (pz[i + 13] // 4) % 8, pz[i + 13] % 4
        lev = clampi(pz[i + 14], 0, 99)
        fc, mode = (pz[i + 15] // 2) % 32, pz[i + 15] % 2
        ff = clampi(pz[i + 16], 0, 99)
        if mode == 0:
            fc_x = (f(fc) + 1).log() / torch.tensor(32.).log()
            ff_x = (f(ff) + 1).log() / torch.tensor(100.).log()
        else:
            fc = fc % 4
            fc_x = f(fc) / 3
            ff_x = f(ff) / 99
        Pm[k] = torch.tensor([lev] + env + [fc, ff, det, bp, ld, rd, ams, kvs, rs, mode, lc, rc], dtype=torch.float32)
        row = [f(lev) / 99] + [f(e) / 99 for e in env] + [fc_x, ff_x, f(det) / 14, f(bp) / 99, f(ld) / 99, f(rd) / 99,
                                                           f(ams) / 3, f(kvs) / 7, f(rs) / 7, f(mode).float()]
        X[k, :19] = torch.stack([r.float() for r in row])
        X[k, 19 + lc] = 1.0
        X[k, 23 + rc] = 1.0
    peg = [clampi(b, 0, 99) for b in pz[102:110]]           # parse_global, :246-300
    alg = pz[110] % 32
    oks, fb = (pz[111] // 8) % 2, pz[111] % 8
    lfs, lfd, lpmd, lamd = (clampi(pz[j], 0, 99) for j in (112, 113, 114, 115))
    lpms = pz[116] // 16
    lfw = clampi((pz[116] // 2) % 8, 0, 5)
    lks = pz[116] % 2
    tsp = clampi(pz[117], 0, 48)
    Pm[0] = torch.tensor(peg + [tsp, lfs, lfd, lpmd, lamd, fb, lpms, oks, lks, lfw, alg, 0, 0], dtype=torch.float32)
    row = [f(e) / 99 for e in peg] + [f(tsp) / 48, f(lfs) / 99, f(lfd) / 99, f(lpmd) / 99, f(lamd) / 99,
                                      f(fb) / 7, f(lpms) / 7, f(oks).float(), f(lks).float()]
    X[0, :17] = torch.stack([r.float() for r in row])
    X[0, 17 + lfw] = 1.0
    src, dst = DX_ALGO[pz[110]]                             # :308 (un-modded key)
    return X, Pm, list(src), list(dst)


def graph_to_syx_bytes(params):
    """dxdata.py:341-397: params (G,7,21) ints -> the full file image
    F0 + [67,0,9,32,0] + G*128 voice bytes + [88] + F7."""
    P = np.asarray(params).astype(np.int64)
    name = [68, 88, 45, 86, 65, 69, 46, 46, 46, 46]
    body = []
    for pg in P:
        for idx in range(6, 0, -1):
            pi = pg[idx]
            lev, env, fc, ff, det, bp, ld, rd, ams, kvs, rs, mode, lc, rc = (
                pi[0], list(pi[1:9]), pi[9], pi[10], pi[11], pi[12], pi[13], pi[14], pi[15], pi[16], pi[17], pi[18],
                pi[19], pi[20])
            body += env + [bp, ld, rd, rc * 4 + lc, det * 8 + rs, kvs * 4 + ams, lev, fc * 2 + mode, ff]
        p0 = pg[0]
        body += list(p0[0:8]) + [p0[18], p0[15] * 8 + p0[13], p0[9], p0[10], p0[11], p0[12],
                                 p0[14] * 16 + p0[17] * 2 + p0[16], p0[8]] + name
    data = [67, 0, 9, 32, 0] + [int(b) for b in body] + [88]
    return bytes([0xF0] + data + [0xF7])


# =============================================================================
# Batcher oracle (pure-Python set logic).  The reference has no batcher: it
# queries each DGL graph from Python loops (model.py:164-177, 189-191).  This is
# the specification of the flat structures the CUDA path consumes instead.
# =============================================================================
EDGE_FWD, EDGE_BACK, EDGE_SELF = 0, 1, 2     # src>dst ; src<dst (feedback back-edge) ; src==dst (feedback self-loop)


def batch_oracle(edge_lists):
    """edge_lists: list over graphs of (src list, dst list).

    Returns dict of numpy arrays:
      adj      u64[B]     bit (src*7+dst) set per edge
      indptr   i32[7B+1]  CSR by destination over flat node ids b*7+v
      indices  i32[E]     flat source ids, ascending within a destination
      eflags   u8[E]      EDGE_FWD / EDGE_BACK / EDGE_SELF (feedback edges marked)
      level    u8[B,7]    encode level: 0 if no adjacent x>v, else 1+max level(x)  
      level_ptr i32[nl+1], level_rows i32[6B]
                          operator nodes (v>=1) grouped by level; row id = v*B+b
                          (node-major).  Inside a level the rows a feedback back-edge ARRIVES
                          at (an edge u->v with u < v) come first, then the rest; ascending in
                          each group.  Node 0 rows are not listed: the root step always runs
                          last over all B graphs.
      level_rare i32[nl]  size of the first group of every level
    """
    B = len(edge_lists)
    adj = np.zeros(B, np.uint64)
    per_dst = [[] for _ in range(7 * B)]
    level = np.zeros((B, 7), np.uint8)
    for b, (src, dst) in enumerate(edge_lists):
        es = set(zip([int(s) for s in src], [int(d) for d in dst]))
        m = 0
        for s, d in es:
            m |= 1 << (s * 7 + d)
            per_dst[b * 7 + d].append((b * 7 + s, EDGE_SELF if s == d else (EDGE_FWD if s > d else EDGE_BACK)))
        adj[b] = m
        for v in range(6, -1, -1):
            nb = [x for x in range(v + 1, 7) if (x, v) in es or (v, x) in es]
            level[b, v] = 0 if not nb else 1 + max(level[b, x] for x in nb)
    indptr = np.zeros(7 * B + 1, np.int32)
    indices, eflags = [], []
    for n, lst in enumerate(per_dst):
        lst.sort()
        indices += [s for s, _ in lst]
        eflags += [fl for _, fl in lst]
        indptr[n + 1] = len(indices)
    nl = int(level[:, 1:].max()) + 1 if B else 0
    rows, ptr, rare = [], [0], []
    sets = [set(zip([int(s) for s in src], [int(d) for d in dst])) for src, dst in edge_lists]
    back_in = lambda b, v: any((u, v) in sets[b] for u in range(v))
    for L in range(nl):
        r1 = sorted(v * B + b for b in range(B) for v in range(1, 7) if level[b, v] == L and back_in(b, v))
        r2 = sorted(v * B + b for b in range(B) for v in range(1, 7) if level[b, v] == L and not back_in(b, v))
        rows += r1 + r2
        rare.append(len(r1))
        ptr.append(len(rows))
    return dict(adj=adj, indptr=indptr, indices=np.array(indices, np.int32), eflags=np.array(eflags, np.uint8),
                level=level, level_ptr=np.array(ptr, np.int32), level_rows=np.array(rows, np.int32),
                level_rare=np.array(rare, np.int32))


# =============================================================================
# Deterministic test weights.  dx_1024.chk is absent from the reference tree and
# a plain seeded init decodes every z to nearly the same topology, which makes a
# weak parity test.  recipe: seeded init, then input-side / head-output weights
# scaled by `gain` and the two edge-head output biases zeroed, so decisions
# depend strongly on z (gain=3: ~135 distinct topologies in 256 samples).
# Reproducible anywhere torch 2.11 CPU is (no 48 MB checkpoint to ship).
# =============================================================================
_GAIN_KEYS = ("z_to_h.0.weight", "h_to_edge.2.weight", "h_to_edge_self.2.weight", "h_to_x.4.weight",
              "h_to_x0.4.weight")


def make_weights(seed=0, gain=1.0):
    torch.manual_seed(seed)
    o = OracleDXVAE()
    if gain != 1.0:
        with torch.no_grad():
            for n, p in o.named_parameters():
                if n in _GAIN_KEYS or n.endswith("weight_ih"):
                    p.mul_(gain)
                if n in ("h_to_edge.2.bias", "h_to_edge_self.2.bias"):
                    p.zero_()
    return o

"""CPU emulation build of the native library — TEST INFRASTRUCTURE ONLY.

The build container has no GPU.  To check the host orchestration (level schedule,
buffer carving, the hand-written gradient chain) before spending GPU time, the SAME
sources are compiled with g++ -DDX_EMU: element-wise functors run as serial loops and
the GEMM is a naive triple loop.  The product package never loads this library; it is
not a fallback (dxvae_b200/_lib.py only ever opens the CUDA build).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from dxvae_b200 import _abi
from dxvae_b200.params import flatten_state_dict, param_table

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "dxvae_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libdxvae_emu.so")
FILES = ["dx_gemm.cu", "dx_tc_gemm.cu", "dx_encoder.cu", "dx_decoder.cu", "dx_data.cu", "dx_api.cu"]


def build():
    srcs = [os.path.join(SRC, f) for f in FILES]
    deps = srcs + [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith(".h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # DX_EMU_ASAN=1 (with LD_PRELOAD=$(g++ -print-file-name=libasan.so)): the same orchestration and functors under
    # AddressSanitizer — the nearest thing to compute-sanitizer for the host-side indexing of the hot path
    san = ["-fsanitize=address", "-fno-omit-frame-pointer", "-g"] if os.environ.get("DX_EMU_ASAN") else []
    cmd = ["g++", "-std=c++17", "-O2", "-DDX_EMU", "-fPIC", "-shared"] + san + ["-x", "c++"] + srcs + ["-o", OUT]
    subprocess.check_call(cmd)
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _abi.bind(C.CDLL(build()))
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Emu:
    """numpy-level driver of the emulated C ABI."""

    def __init__(self, state_dict):
        self.lib = lib()
        self.table = param_table(self.lib)
        self.total = int(self.lib.dxvae_param_blob_floats())
        self.blob = flatten_state_dict(state_dict, self.table, self.total)

    def batch(self, X, P, edge_lists):
        L = self.lib
        B = len(edge_lists)
        eptr = np.zeros(B + 1, np.int32)
        src = np.array([s for e in edge_lists for s in e[0]], np.int8)
        dst = np.array([d for e in edge_lists for d in e[1]], np.int8)
        eptr[1:] = np.cumsum([len(e[0]) for e in edge_lists])
        out = dict(adj=np.zeros(B, np.uint64), indptr=np.zeros(7 * B + 1, np.int32),
                   indices=np.zeros(max(1, len(src)), np.int32), eflags=np.zeros(max(1, len(src)), np.uint8),
                   level=np.zeros((B, 7), np.uint8), level_ptr=np.zeros(16, np.int32),
                   level_rows=np.zeros(6 * B, np.int32))
        nl = C.c_int32(0)
        _abi.check(L, L.dxvae_batch_build_host(B, ptr(eptr), ptr(src), ptr(dst), ptr(out["adj"]), ptr(out["indptr"]),
                                               ptr(out["indices"]), ptr(out["eflags"]), ptr(out["level"]),
                                               ptr(out["level_ptr"]), ptr(out["level_rows"]), C.byref(nl)), "batch")
        out["n_levels"] = nl.value
        out["indices"] = out["indices"][:len(src)]; out["eflags"] = out["eflags"][:len(src)]
        if X is not None:
            Xg = np.ascontiguousarray(X, np.float32); Pg = np.ascontiguousarray(P, np.float32)
            out["Xn"] = np.zeros((7, B, 32), np.float32); out["cls"] = np.zeros((14, B), np.int32)
            _abi.check(L, L.dxvae_pack_graphs(B, ptr(Xg), ptr(Pg), ptr(out["Xn"]), ptr(out["cls"]), None), "pack")
        out["B"] = B
        return out

    def encode(self, bt):
        L = self.lib
        B = bt["B"]
        ws = np.full(L.dxvae_workspace_bytes(_abi.OP_ENCODE, B), 0xFF, np.uint8)   # NaN-poisoned: a read of workspace that was never written shows up in the outputs
        mu = np.zeros((B, 128), np.float32); sd = np.zeros((B, 128), np.float32)
        _abi.check(L, L.dxvae_encode_fwd(ptr(self.blob), B, ptr(bt["Xn"]), ptr(bt["adj"]), bt["n_levels"],
                                         ptr(bt["level_ptr"]), ptr(bt["level_rows"]), ptr(bt["level_ptr"][8:]), ptr(mu), ptr(sd), ptr(ws),
                                         ws.nbytes, 0, 0, None), "encode")
        return mu, sd

    def elbo(self, bt, eps, w=(2, 5, 0.01), inv_batch=None, grads=True, compact=False):
        L = self.lib
        B = bt["B"]
        sp = sr = None
        if compact:
            sp = np.zeros(34, np.int32); sr = np.zeros(33 * B, np.int32)
            _abi.check(L, L.dxvae_batch_steps_host(B, ptr(bt["adj"]), ptr(sp), ptr(sr)), "steps")
        # NaN-poisoned: a read of workspace that was never written shows up in the outputs.  Compacted steps: the
        # schedule-sized workspace (per-step buffers hold the active rows only), exactly as large as the library asks for
        nws = L.dxvae_workspace_bytes_sched(_abi.OP_TRAIN, B, bt["n_levels"], ptr(bt["level_ptr"]), ptr(sp)) if compact else L.dxvae_workspace_bytes(_abi.OP_TRAIN, B)
        assert nws <= L.dxvae_workspace_bytes(_abi.OP_TRAIN, B)
        ws = np.full(nws, 0xFF, np.uint8)
        loss5 = np.zeros(5, np.float32)
        mu = np.zeros((B, 128), np.float32); sd = np.zeros((B, 128), np.float32)
        g = np.zeros(self.total, np.float32) if grads else None
        eps = np.ascontiguousarray(eps, np.float32)
        _abi.check(L, L.dxvae_elbo_step(ptr(self.blob), B, ptr(bt["Xn"]), ptr(bt["cls"]), ptr(bt["adj"]),
                                        bt["n_levels"], ptr(bt["level_ptr"]), ptr(bt["level_rows"]), ptr(bt["level_ptr"][8:]), ptr(eps),
                                        w[0], w[1], w[2], inv_batch or 1.0 / B, ptr(loss5), ptr(mu), ptr(sd), ptr(g),
                                        ptr(ws), ws.nbytes, 0, ptr(sp), ptr(sr), None, None), "elbo")
        return loss5, mu, sd, g

    def decode(self, z):
        L = self.lib
        z = np.ascontiguousarray(z, np.float32)
        B = z.shape[0]
        ws = np.full(L.dxvae_workspace_bytes(_abi.OP_DECODE, B), 0xFF, np.uint8)   # NaN-poisoned: a read of workspace that was never written shows up in the outputs
        Xg = np.zeros((B, 7, 27), np.float32); Pg = np.zeros((B, 7, 21), np.float32)
        adj = np.zeros(B, np.uint64); mg = np.zeros((B, 2), np.float32)
        _abi.check(L, L.dxvae_decode_greedy(ptr(self.blob), B, ptr(z), ptr(Xg), ptr(Pg), ptr(adj), ptr(mg), ptr(ws),
                                            ws.nbytes, 0, None), "decode")
        return Xg, Pg, adj, mg

"""Data-format kernels (dxdata.py restatements) through the CPU emulation build, against the
oracle and the reference's own artefacts (DXDataset.bin via golden voices, gen_patch.syx)."""
import hashlib
import os

import numpy as np

import dxvae_oracle as O
from dxvae_b200.synth import random_voices
from tests import util
from tests.emu import emu


def _make_graphs(v):
    L = emu.lib(); p = emu.ptr
    B = len(v)
    out = dict(Xn=np.zeros((7, B, 32), np.float32), cls=np.zeros((14, B), np.int32), adj=np.zeros(B, np.uint64),
               Xg=np.zeros((B, 7, 27), np.float32), Pg=np.zeros((B, 7, 21), np.float32))
    v = np.ascontiguousarray(v, np.uint8)
    assert L.dxvae_voices_to_graphs(B, p(v), p(out["Xn"]), p(out["cls"]), p(out["adj"]), p(out["Xg"]), p(out["Pg"]),
                                    None) == 0
    return out


def test_make_graph_reproduces_dataset_bin():
    g = util.voices()
    out = _make_graphs(g["voices"])
    assert hashlib.sha256(out["Xg"].tobytes()).hexdigest() == str(g["X_sha256"])
    assert hashlib.sha256(out["Pg"].tobytes()).hexdigest() == str(g["params_sha256"])
    ptr, es, ed = g["edge_ptr"], g["edge_src"], g["edge_dst"]
    for i in range(1024):
        m = 0
        for s, d in zip(es[ptr[i]:ptr[i + 1]], ed[ptr[i]:ptr[i + 1]]):
            m |= 1 << (int(s) * 7 + int(d))
        assert int(out["adj"][i]) == m


def test_make_graph_synthetic_vs_oracle_and_layouts():
    v = random_voices(300, 3)
    out = _make_graphs(v)
    for i in range(0, 300, 7):
        X, P, s, d = O.make_graph(v[i])
        assert np.array_equal(X.numpy().view(np.uint32), out["Xg"][i].view(np.uint32))
        assert np.array_equal(P.numpy(), out["Pg"][i])
    assert np.array_equal(out["Xn"][:, :, :27].transpose(1, 0, 2), out["Xg"]) and not out["Xn"][:, :, 27:].any()
    assert np.array_equal(out["cls"][0], out["Pg"][:, 0, 17]) and np.array_equal(out["cls"][1], out["Pg"][:, 0, 18])
    assert np.array_equal(out["cls"][2:8].T, out["Pg"][:, 1:, 19]) and np.array_equal(out["cls"][8:14].T, out["Pg"][:, 1:, 20])
    # pack_graphs produces the same node-major tensors from the graph-major ones
    L = emu.lib(); p = emu.ptr
    Xn = np.zeros_like(out["Xn"]); cls = np.zeros_like(out["cls"])
    assert L.dxvae_pack_graphs(300, p(out["Xg"]), p(out["Pg"]), p(Xn), p(cls), None) == 0
    assert np.array_equal(Xn, out["Xn"]) and np.array_equal(cls, out["cls"])
    Xg = np.zeros_like(out["Xg"]); Pg = np.zeros_like(out["Pg"])
    Pn = np.zeros((7, 300, 32), np.float32); Pn[:, :, :21] = out["Pg"].transpose(1, 0, 2)
    assert L.dxvae_unpack_graphs(300, p(Xn), p(Pn), p(Xg), p(Pg), None) == 0
    assert np.array_equal(Xg, out["Xg"]) and np.array_equal(Pg, out["Pg"])


def test_pack_syx_matches_reference_file_and_oracle():
    L = emu.lib(); p = emu.ptr
    gen = open(os.path.join(util.GOLDEN, "gen_patch.syx"), "rb").read()
    gv = np.frombuffer(gen[6:6 + 4096], np.uint8).reshape(32, 128).copy()
    Pg = _make_graphs(gv)["Pg"]
    o = np.zeros(32 * 128, np.uint8)
    assert L.dxvae_pack_syx(32, p(Pg), p(o), None) == 0
    assert bytes([0xF0, 67, 0, 9, 32, 0]) + o.tobytes() + bytes([88, 0xF7]) == gen
    v = random_voices(100, 5)
    Pg = _make_graphs(v)["Pg"]
    o = np.zeros(100 * 128, np.uint8)
    assert L.dxvae_pack_syx(100, p(Pg), p(o), None) == 0
    assert O.graph_to_syx_bytes(Pg)[6:-2] == o.tobytes()


def test_abi_exports_every_declared_symbol():
    """The CUDA library loads without a GPU and exports every symbol include/dxvae_b200.h declares."""
    import re
    from dxvae_b200 import _abi, _lib
    lib = _lib.lib()
    hdr = open(os.path.join(os.path.dirname(util.GOLDEN), "..", "include", "dxvae_b200.h")).read()
    declared = set(re.findall(r"\b(dxvae_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("dxvae_param_entry_t")
    assert declared == set(_abi.SIGNATURES), declared ^ set(_abi.SIGNATURES)
    for name in declared:
        getattr(lib, name)
    assert lib.dxvae_param_count() == 12083541
    from dxvae_b200.params import param_table
    t = param_table(lib)
    sd = O.OracleDXVAE().state_dict()
    assert [n for n, _, _ in t] == list(sd.keys())
    assert all(tuple(sd[n].shape) == s for n, _, s in t) and all(off % 64 == 0 for _, off, _ in t)

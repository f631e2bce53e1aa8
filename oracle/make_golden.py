"""Validate oracle/dxvae_oracle.py against the UNMODIFIED reference and write the
fixtures under tests/golden/.  Run in the build container only (needs
/root/reference):   python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.

Fixtures written
  synprez_voices.npz  the 1024 packed voices behind DX_data/DXDataset.bin (file order,
                      SURVEY App. B.1), the bin's edge lists, SHA-256 of its X / params
                      tensors, full X/params of the first 64 graphs; Dexed_01 voices.
  gen_patch.syx       the reference's generated/gen_patch.syx (pins graph_to_syx layout).
  model_golden.npz    outputs of the reference model (model.py, run under oracle/shim)
                      for the weight recipes of dxvae_oracle.make_weights on a fixed
                      subset of dataset graphs: mu/std, 5 loss terms for a given eps,
                      gradient fingerprints of all 46 tensors, greedy-decode outputs.
"""
import hashlib
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dxvae_oracle as O  # noqa: E402
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
BANK_ORDER = [29, 1, 15, 14, 28, 16, 2, 3, 17, 13, 7, 6, 12, 4, 10, 11, 5, 8, 20, 21, 9, 23, 22, 32, 26, 27, 25, 31,
              19, 18, 30, 24]
SUBSET = list(range(0, 1024, 16))          # 64 dataset graphs
GRAD_SAMPLES = 48


def read_bank(path):
    raw = open(path, "rb").read()
    assert raw[0] == 0xF0 and raw[-1] == 0xF7 and len(raw) == 4104
    return np.frombuffer(raw[6:6 + 4096], np.uint8).reshape(32, 128).copy()


def data_fixtures(dxdata, G):
    root = ref_loader.REF_ROOT
    voices = np.concatenate([read_bank(os.path.join(root, "DX_data", "SynprezFM", "SynprezFM_%02d.syx" % b))
                             for b in BANK_ORDER])
    dexed = read_bank(os.path.join(root, "DX_data", "Dexed_01.syx"))
    X = torch.stack([g.ndata["X"] for g in G]); P = torch.stack([g.ndata["params"] for g in G])
    # 1. the oracle's make_graph reproduces the bin bit-exactly from the voices
    src_all, dst_all, eptr = [], [], [0]
    for i, g in enumerate(G):
        Xo, Po, s, d = O.make_graph(voices[i])
        assert torch.equal(Xo.view(torch.int32), X[i].view(torch.int32)), i
        assert torch.equal(Po, P[i]), i
        es, ed = g.edges()
        assert (es.tolist(), ed.tolist()) == (s, d), i
        src_all += s; dst_all += d; eptr.append(len(src_all))
    # 2. ... and the reference's own _make_graph agrees on a sample (it is slow)
    ds = dxdata.DXDataset.__new__(dxdata.DXDataset)
    ds.DX_ALGO = O.DX_ALGO
    for i in range(0, 1024, 37):
        g = ds._make_graph(torch.tensor(voices[i].astype(np.int64)))
        assert torch.equal(g.ndata["X"].view(torch.int32), X[i].view(torch.int32))
    # 3. graph_to_syx layout: gen_patch.syx -> graphs -> bytes round trip
    gen = open(os.path.join(root, "generated", "gen_patch.syx"), "rb").read()
    gv = np.frombuffer(gen[6:6 + 4096], np.uint8).reshape(32, 128)
    gp = np.stack([O.make_graph(v)[1].numpy() for v in gv])
    assert O.graph_to_syx_bytes(gp) == gen, "graph_to_syx restatement differs from gen_patch.syx"
    shutil.copyfile(os.path.join(root, "generated", "gen_patch.syx"), os.path.join(OUT, "gen_patch.syx"))
    np.savez_compressed(
        os.path.join(OUT, "synprez_voices.npz"), voices=voices, dexed01=dexed,
        edge_src=np.array(src_all, np.int8), edge_dst=np.array(dst_all, np.int8), edge_ptr=np.array(eptr, np.int32),
        X_sha256=hashlib.sha256(X.numpy().tobytes()).hexdigest(),
        params_sha256=hashlib.sha256(P.numpy().tobytes()).hexdigest(),
        X_first64=X[:64].numpy(), params_first64=P[:64].numpy())
    print("data fixtures ok: 1024 voices reproduce DXDataset.bin bit-exactly; gen_patch.syx round-trips")


def grad_fingerprint(model):
    g = torch.Generator().manual_seed(99)
    names, norms, sums, idxs, vals = [], [], [], [], []
    for n, p in model.named_parameters():
        gr = p.grad.detach().flatten()
        ix = torch.randint(0, gr.numel(), (GRAD_SAMPLES,), generator=g)
        names.append(n); norms.append(gr.double().norm().item()); sums.append(gr.double().sum().item())
        idxs.append(ix.numpy()); vals.append(gr[ix].numpy())
    return dict(names=np.array(names), norms=np.array(norms), sums=np.array(sums), idx=np.stack(idxs),
                vals=np.stack(vals))


def model_fixtures(model_mod, G):
    Gs = [G[i] for i in SUBSET]
    X = torch.stack([g.ndata["X"] for g in Gs]); P = torch.stack([g.ndata["params"] for g in Gs])
    A = torch.stack([g.adj().to_dense() for g in Gs])
    out = dict(subset=np.array(SUBSET, np.int32))
    for tag, (seed, gain) in {"init": (0, 1.0), "stress": (0, 3.0)}.items():
        o = O.make_weights(seed, gain)
        m = model_mod.DXVAE()
        m.load_state_dict(o.state_dict())
        # ---- encode
        q = m.encode(Gs)
        mu_o, std_o = o.encode(X, A)
        assert (q.loc - mu_o).abs().max() < 1e-6 and (q.scale - std_o).abs().max() < 1e-6
        # ---- loss with injected noise (rsample after manual_seed == randn after manual_seed)
        for w in ((2, 5, 0.01), (3, 6, 0.002)):
            torch.manual_seed(1234)
            lr = m.loss(q, Gs, *w)
            torch.manual_seed(1234)
            eps = torch.randn(len(Gs), 128)
            lo = o.loss(mu_o, std_o, X, P, A, eps, *w)
            for a, b in zip(lr, lo):
                assert abs(a.item() - b.item()) <= 1e-5 * abs(a.item()), (tag, w, a.item(), b.item())
            out["%s_loss_w%d" % (tag, w[0])] = np.array([t.item() for t in lr], np.float64)
        # gradients of the (3,6,0.002) call
        m.zero_grad(); o.zero_grad()
        lr[0].backward(); lo[0].backward()
        fr, fo = grad_fingerprint(m), grad_fingerprint(o)
        rel = np.abs(fr["vals"] - fo["vals"]).max(1) / (np.abs(fr["vals"]).max(1) + 1e-30)
        assert rel.max() < 1e-4, rel.max()
        for k, v in fr.items():
            out["%s_grad_%s" % (tag, k)] = v
        out[tag + "_eps"] = eps.numpy()
        out[tag + "_mu"] = q.loc.detach().numpy(); out[tag + "_std"] = q.scale.detach().numpy()
        # ---- greedy decode from mu and from prior samples
        torch.manual_seed(4321)
        zs = {"mu": q.loc.detach(), "prior": torch.randn(len(Gs), 128)}
        for zt, z in zs.items():
            with torch.no_grad():
                m.hidden = [[None] * 7 for _ in range(len(z))]
                D = m.decode(z)
            Xr = torch.stack([g.ndata["X"] for g in D]); Pr = torch.stack([g.ndata["params"] for g in D])
            Ar = torch.stack([g.adj().to_dense() for g in D]).to(torch.uint8)
            Xo, Po, Ao, mg = o.decode(z, return_margins=True)
            assert torch.equal(Pr.int(), Po.int()) and torch.equal(Ar, Ao), (tag, zt)
            assert (Xr - Xo).abs().max() <= 1e-6
            for i, g in enumerate(D):
                es, ed = g.edges()
                assert (es.tolist(), ed.tolist()) == O.edges_from_adj(Ar[i].tolist())
            lg = torch.cat([l.flatten(1) for l in mg["edge"] + mg["self"]], 1)
            out["%s_dec_%s_z" % (tag, zt)] = z.numpy()
            out["%s_dec_%s_X" % (tag, zt)] = Xr.numpy()
            out["%s_dec_%s_params" % (tag, zt)] = Pr.numpy().astype(np.int16)
            out["%s_dec_%s_adj" % (tag, zt)] = Ar.numpy()
            out["%s_dec_%s_minmargin" % (tag, zt)] = lg.abs().min(1).values.numpy()
        print(tag, "ok: losses", out[tag + "_loss_w3"], "decoded topologies",
              len({bytes(a.numpy().tobytes()) for a in Ar}))
    np.savez_compressed(os.path.join(OUT, "model_golden.npz"), **out)


if __name__ == "__main__":
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    model_mod, dxdata = ref_loader.load_reference()
    G = ref_loader.load_dataset_graphs()
    data_fixtures(dxdata, G)
    model_fixtures(model_mod, G)
    print("golden fixtures written to", OUT)

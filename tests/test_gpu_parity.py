"""GPU parity tests: the CUDA library, called through the C ABI / the DXVAE mirror, against
the CPU oracle (oracle/dxvae_oracle.py) and against the golden fixtures generated from the
UNMODIFIED reference (tests/golden, oracle/make_golden.py).

Tolerances (fp32 path; SURVEY §8c, from the reference's own fp32-vs-fp64 noise floor):
  latents |d| <= 1e-5 abs; each loss term rel <= 1e-5; gradients max-norm-relative <= 1e-4 per
  tensor; decoded params / topology exact on graphs whose decision margins exceed 1e-4
  (tie-aware); decoded X <= 1e-6 abs; batcher / data-format outputs bit-exact."""
import ctypes
import hashlib
import os

import numpy as np
import pytest
import torch

import dxvae_oracle as O
from tests import util

pytestmark = pytest.mark.gpu

TOL_LAT, TOL_LOSS, TOL_GRAD, MARGIN = 1e-5, 1e-5, 1e-4, 2e-5


@pytest.fixture(scope="module")
def lib():
    from dxvae_b200 import _lib
    return _lib.require_cuda()


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(util.GOLDEN, "model_golden.npz"))


def make_model(seed, gain, prec="fp32"):
    """The tests of this file pin the FFMA arithmetic unless they say otherwise (the class default is the FP32-accurate
    tensor-core mode "3xtf32": covered here where a test is parametrised over `prec`, by tests/test_gpu_tf32.py,
    tests/test_cfg1_trained.py, the 1 M-patch tests of tests/test_gpu_scale.py and smoke())."""
    from dxvae_b200 import DXVAE
    o = O.make_weights(seed, gain)
    m = DXVAE()
    m.load_state_dict(o.state_dict())
    m.verbose = False
    m.precision = m.encode_precision = m.decode_precision = prec
    return m, o


def st():
    return torch.cuda.current_stream().cuda_stream


# --------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 1536, 512), (1000, 1536, 27), (513, 55, 1024), (300, 2, 2048), (77, 1, 1024),
                                   (4096, 2048, 512), (2048, 1024, 1024), (64, 512, 128), (1, 27, 1024)])
def test_gemm_forward(lib, M, N, K):
    from dxvae_b200 import _lib
    g = torch.Generator().manual_seed(M * 7 + N)
    lda = 32 if K == 27 else K
    A = torch.randn(M, lda, generator=g).cuda(); W = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    for act, fn in ((0, lambda t: t), (1, torch.relu), (2, torch.tanh), (4, torch.nn.functional.softplus)):
        C = torch.empty(M, N, device="cuda")
        _lib.check(lib.dxvae_test_gemm(0, M, N, K, A.data_ptr(), lda, W.data_ptr(), K, C.data_ptr(), N, b.data_ptr(),
                                       act, 0, st()), "gemm")
        ref = fn(A[:, :K].double() @ W.double().t() + b.double())
        err = (C.double() - ref).abs().max().item()
        assert err <= 2e-4 * max(1.0, ref.abs().max().item()), (act, err)


@pytest.mark.parametrize("M,N,K", [(128, 1536, 512), (1000, 55, 1024), (4096, 2048, 512), (333, 2, 2048)])
def test_gemm_dgrad_wgrad(lib, M, N, K):
    from dxvae_b200 import _lib
    g = torch.Generator().manual_seed(M + N + K)
    dY = torch.randn(M, N, generator=g).cuda(); W = torch.randn(N, K, generator=g).cuda()
    X = torch.randn(M, K, generator=g).cuda()
    dX = torch.zeros(M, K, device="cuda")
    _lib.check(lib.dxvae_test_gemm(1, M, N, K, dY.data_ptr(), N, W.data_ptr(), K, dX.data_ptr(), K, None, 0, 0, st()), "dgrad")
    ref = dY.double() @ W.double()
    assert (dX.double() - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()
    _lib.check(lib.dxvae_test_gemm(1, M, N, K, dY.data_ptr(), N, W.data_ptr(), K, dX.data_ptr(), K, None, 0, 1, st()), "dgrad+")
    assert (dX.double() - 2 * ref).abs().max().item() <= 4e-4 * ref.abs().max().item()
    dW = torch.zeros(N, K, device="cuda")
    _lib.check(lib.dxvae_test_gemm(2, M, N, K, dY.data_ptr(), N, X.data_ptr(), K, dW.data_ptr(), K, None, 0, 0, st()), "wgrad")
    refw = dY.double().t() @ X.double()
    assert (dW.double() - refw).abs().max().item() <= 2e-4 * refw.abs().max().item()


# --------------------------------------------------------------------------- batcher / data formats
def test_device_schedule_matches_set_logic(lib):
    from dxvae_b200 import DXVAE
    from dxvae_b200.dxdata import DXGraphBatch, mask_from_edges
    m = DXVAE()
    for n, p, seed in ((5000, 0.2, 1), (1, 0.5, 2), (1025, 0.0, 3), (3000, 1.0, 4), (2048, 0.08, 5)):
        E = util.random_edge_lists(n, p, seed)
        ob = O.batch_oracle(E)
        adj = torch.tensor([mask_from_edges(*e) for e in E], dtype=torch.int64, device="cuda")
        d = type("D", (), {})()
        d.B, d.adj, d.level_ptr = n, adj, np.zeros(16, np.int32)
        m._schedule(d)
        assert np.array_equal(d.level.cpu().numpy(), ob["level"])
        assert d.n_levels == len(ob["level_ptr"]) - 1
        assert np.array_equal(d.level_ptr[:d.n_levels + 1], ob["level_ptr"])
        assert np.array_equal(d.level_rows.cpu().numpy(), ob["level_rows"])
        assert np.array_equal(d.level_ptr[8:8 + d.n_levels], ob["level_rare"])


def test_host_batcher_matches_set_logic(lib):
    from dxvae_b200 import DXVAE
    from dxvae_b200.dxdata import DXGraph
    E = util.random_edge_lists(200, 0.2, 9)
    G = [DXGraph(torch.zeros(7, 27), torch.zeros(7, 21), *e) for e in E]
    d = DXVAE()._prepare(G)
    ob = O.batch_oracle(E)
    assert np.array_equal(d.adj.cpu().numpy().view(np.uint64), ob["adj"])
    assert np.array_equal(d.csr[0], ob["indptr"]) and np.array_equal(d.csr[1], ob["indices"])
    assert np.array_equal(d.csr[2], ob["eflags"]) and np.array_equal(d.level, ob["level"])
    assert np.array_equal(d.level_rows.cpu().numpy(), ob["level_rows"])


def test_shuffled_views_of_a_pinned_batch_are_gathered_by_the_device(lib):
    """A training loop hands forward() a shuffled Python list of graph objects that are row views of one pinned host batch
    (what DXDataset / DXGraphBatch iteration give out).  The batcher then ships an index list and the device gathers the
    rows out of the pinned memory (dxvae_pack_graphs_indexed): same batch, bit for bit, as stacking the graphs on the host;
    a modified or foreign graph in the list falls back to stacking."""
    import random
    from dxvae_b200.dxdata import DXGraph, DXGraphBatch, IndexedBatch
    idx = list(range(0, 1024, 2))
    X, P, E, A = util.dataset_graphs(idx)
    m, _ = make_model(0, 1.0, "3xtf32")
    m.host_batcher_max = 0                                    # (small batches would otherwise take the host batcher)
    host = DXGraphBatch.from_graphs(_graphs(X, P, E)).pin_memory()
    views = list(host)
    random.Random(4).shuffle(views)
    gb = DXGraphBatch.from_graphs(views, staging=True)
    assert isinstance(gb, IndexedBatch) and len(gb) == len(idx)
    order = [int(i) for i in gb.index]
    plain = DXGraphBatch(host.X[order], host.params[order], host.adj[order])
    d1, d2 = m._prepare(views), m._prepare(plain)
    assert torch.equal(d1.Xn, d2.Xn) and torch.equal(d1.cls, d2.cls) and torch.equal(d1.adj, d2.adj)
    assert torch.equal(d1.level_rows, d2.level_rows) and np.array_equal(d1.level_ptr, d2.level_ptr)
    eps = torch.randn(len(idx), 128, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        a = m.forward(views, eps=eps); b = m.forward(plain, eps=eps)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    views[5].ndata["X"] = views[5].ndata["X"].clone()         # touching a view detaches it: the list is stacked instead
    gb2 = DXGraphBatch.from_graphs(views, staging=True)
    assert not isinstance(gb2, IndexedBatch) and torch.equal(gb2.X, plain.X)
    views[6] = DXGraph(X[0], P[0], *E[0])                      # ... and so does a foreign graph
    assert not isinstance(DXGraphBatch.from_graphs(views, staging=True), IndexedBatch)


def test_make_graph_matches_dataset_bin_bit_exact(lib):
    """dxdata.py:_make_graph on the device == DX_data/DXDataset.bin (golden: SHA-256 of the
    bin's X / params tensors + its edge lists, written by oracle/make_golden.py)."""
    from dxvae_b200.dxdata import voices_to_batch
    v = util.voices()
    gb = voices_to_batch(v["voices"]).cpu()
    assert hashlib.sha256(gb.X.numpy().tobytes()).hexdigest() == str(v["X_sha256"])
    assert hashlib.sha256(gb.params.numpy().tobytes()).hexdigest() == str(v["params_sha256"])
    assert np.array_equal(gb.X[:64].numpy().view(np.uint32), v["X_first64"].view(np.uint32))
    ptr, es, ed = v["edge_ptr"], v["edge_src"], v["edge_dst"]
    A = util.adj_from_masks(gb.adj.numpy().view(np.uint64))
    for i in range(1024):
        ref = np.zeros((7, 7), np.uint8)
        ref[es[ptr[i]:ptr[i + 1]], ed[ptr[i]:ptr[i + 1]]] = 1
        assert np.array_equal(A[i], ref), i


def test_graph_to_syx_matches_reference_file(lib):
    """dxdata.py:graph_to_syx layout pinned by the reference's generated/gen_patch.syx."""
    from dxvae_b200.dxdata import graph_to_syx_bytes, read_syx, voices_to_batch
    path = os.path.join(util.GOLDEN, "gen_patch.syx")
    gb = voices_to_batch(read_syx(path))
    assert graph_to_syx_bytes(gb) == open(path, "rb").read()
    # round trip on every dataset voice: pack(make_graph(v)) reproduces the legal bits of v
    v = util.voices()["voices"]
    gb = voices_to_batch(v)
    out = np.frombuffer(graph_to_syx_bytes(gb)[6:-2], np.uint8).reshape(-1, 128)
    gb2 = voices_to_batch(out)
    assert torch.equal(gb.params, gb2.params) and torch.equal(gb.X, gb2.X)


# --------------------------------------------------------------------------- encode
def _encode_cases():
    idx = list(range(0, 1024, 16))
    X, P, E, A = util.dataset_graphs(idx)
    return idx, X, P, E, A


def _graphs(X, P, E):
    from dxvae_b200.dxdata import DXGraph
    return [DXGraph(X[i], P[i], *E[i]) for i in range(len(E))]


def test_class_defaults_are_the_fp32_accurate_tensor_core_mode(lib):
    from dxvae_b200 import DXVAE
    m = DXVAE()
    assert (m.precision, m.encode_precision, m.decode_precision) == ("3xtf32", "3xtf32", "3xtf32")


@pytest.mark.parametrize("prec", ["fp32", "3xtf32"])
@pytest.mark.parametrize("tag,gain", [("init", 1.0), ("stress", 3.0)])
def test_encode_matches_oracle_and_reference_golden(lib, golden, tag, gain, prec):
    idx, X, P, E, A = _encode_cases()
    assert list(golden["subset"]) == idx
    m, o = make_model(0, gain, prec)
    with torch.no_grad():
        q = m.encode(_graphs(X, P, E))
        mu_o, sd_o = o.encode(X, A)
    mu, sd = q.loc.cpu(), q.scale.cpu()
    assert (mu - mu_o).abs().max() <= TOL_LAT and (sd - sd_o).abs().max() <= TOL_LAT
    assert np.abs(mu.numpy() - golden[tag + "_mu"]).max() <= TOL_LAT
    assert np.abs(sd.numpy() - golden[tag + "_std"]).max() <= TOL_LAT


def test_encode_arbitrary_topologies_and_batch_paths(lib):
    m, o = make_model(1, 3.0)
    n = 700
    idx = list(np.random.default_rng(0).integers(0, 1024, n))
    X, P, _, _ = util.dataset_graphs(idx)
    E = util.random_edge_lists(n, 0.3, 21)
    with torch.no_grad():
        mu_o, sd_o = o.encode(X, util.adj_dense(E))
        q = m.encode(_graphs(X, P, E))                       # host batcher path
        from dxvae_b200.dxdata import DXGraphBatch
        gb = DXGraphBatch.from_graphs(_graphs(X, P, E))
        gbd = DXGraphBatch(gb.X.cuda(), gb.params.cuda(), gb.adj.cuda())
        m.max_chunk = 256                                     # chunked, device scheduler path
        q2 = m.encode(gbd)
    assert (q.loc.cpu() - mu_o).abs().max() <= TOL_LAT and (q.scale.cpu() - sd_o).abs().max() <= TOL_LAT
    assert torch.equal(q.loc, q2.loc) and torch.equal(q.scale, q2.scale)


# --------------------------------------------------------------------------- ELBO + gradients
def _check_grads(m, o, scale=1.0):
    worst = 0.0
    named = dict(m.named_parameters())
    for n, p in o.named_parameters():
        ref = p.grad
        got = named[n].grad.cpu() * scale
        rel = (ref - got).abs().max().item() / (ref.abs().max().item() + 1e-30)
        worst = max(worst, rel)
        assert rel <= TOL_GRAD, (n, rel)
    return worst


@pytest.mark.parametrize("tag,gain", [("init", 1.0), ("stress", 3.0)])
def test_elbo_forward_backward_matches_oracle_and_reference_golden(lib, golden, tag, gain):
    idx, X, P, E, A = _encode_cases()
    m, o = make_model(0, gain)
    G = _graphs(X, P, E)
    eps = torch.from_numpy(golden[tag + "_eps"])
    for w, key in (((2, 5, 0.01), "_loss_w2"), ((3, 6, 0.002), "_loss_w3")):
        m.zero_grad(); o.zero_grad()
        out = m.forward(G, *w, eps=eps)
        mu_o, sd_o = o.encode(X, A)
        lo = o.loss(mu_o, sd_o, X, P, A, eps, *w)
        for a, b, c in zip(out, lo, golden[tag + key]):
            assert abs(a.item() - b.item()) <= TOL_LOSS * abs(b.item()) + 1e-7
            assert abs(a.item() - c) <= TOL_LOSS * abs(c) + 1e-7        # the reference's own numbers
    out[0].backward(); lo[0].backward()
    _check_grads(m, o)
    # gradient fingerprints of the reference run (sampled entries, norms)
    named = dict(m.named_parameters())
    for k, n in enumerate(golden[tag + "_grad_names"]):
        g = named[str(n)].grad.cpu().flatten()
        vals = g[torch.from_numpy(golden[tag + "_grad_idx"][k])].numpy()
        ref = golden[tag + "_grad_vals"][k]
        scale = np.abs(g.numpy()).max() + 1e-30
        assert np.abs(vals - ref).max() / scale <= TOL_GRAD, n
        assert abs(g.double().norm().item() - golden[tag + "_grad_norms"][k]) <= 1e-4 * golden[tag + "_grad_norms"][k] + 1e-12


def test_split_encode_then_loss_equals_fused_forward(lib):
    idx = util.pick_by_alg([3, 5, 18, 20, 0, 31, 7, 16])
    X, P, E, A = util.dataset_graphs(idx)
    m, o = make_model(2, 3.0)
    G = _graphs(X, P, E)
    torch.manual_seed(3)
    eps = torch.randn(len(G), 128)
    m.zero_grad()
    total, *_ = m.forward(G, eps=eps)
    total.backward()
    g1 = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad()
    q = m.encode(G)
    total2, *_ = m.loss(q, G, eps=eps)
    total2.backward()
    assert abs(total.item() - total2.item()) <= 1e-6 * abs(total.item())
    for n, p in m.named_parameters():
        ref = g1[n]
        assert (p.grad - ref).abs().max().item() <= 1e-5 * (ref.abs().max().item() + 1e-30), n
    # and against the oracle on arbitrary topologies
    E2 = util.random_edge_lists(len(idx), 0.35, 5)
    A2 = util.adj_dense(E2)
    G2 = _graphs(X, P, E2)
    m.zero_grad(); o.zero_grad()
    out = m.forward(G2, eps=eps)
    mu_o, sd_o = o.encode(X, A2)
    lo = o.loss(mu_o, sd_o, X, P, A2, eps)
    assert abs(out[0].item() - lo[0].item()) <= TOL_LOSS * abs(lo[0].item())
    out[0].backward(); lo[0].backward()
    _check_grads(m, o)


def test_adamw_kernel_matches_torch_optimizer(lib):
    """Same gradients in -> same weights out as torch.optim.AdamW (model.py:375 defaults)."""
    from dxvae_b200 import _lib
    g = torch.Generator().manual_seed(0)
    n = 100003
    w0 = torch.randn(n, generator=g)
    p = torch.nn.Parameter(w0.clone())
    opt = torch.optim.AdamW([p], lr=1e-3)
    w = w0.clone().cuda(); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    for step in range(1, 6):
        gr = torch.randn(n, generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g)))
        p.grad = gr.clone()
        opt.step()
        gd = gr.cuda()
        _lib.check(lib.dxvae_adamw_step(n, w.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), 1e-3, 0.9, 0.999,
                                        1e-8, 0.01, step, 1.0, st()), "adamw")
        assert (w.cpu() - p.data).abs().max().item() <= 2e-6


def test_trainer_trajectory_matches_oracle(lib):
    from dxvae_b200.train import Trainer
    idx = list(range(0, 1024, 8))
    X, P, E, A = util.dataset_graphs(idx)
    m, o = make_model(0, 1.0)
    t = Trainer(m, lr=1e-3, w=(2, 5, 0.01))
    data = t.upload(_graphs(X, P, E))
    opt = torch.optim.AdamW(o.parameters(), lr=1e-3)
    torch.manual_seed(11)
    first = None
    for step in range(4):
        eps = torch.randn(len(idx), 128)
        loss5 = t.step(data, list(range(len(idx))), eps=eps)
        opt.zero_grad()
        mu_o, sd_o = o.encode(X, A)
        lo = o.loss(mu_o, sd_o, X, P, A, eps)
        lo[0].backward()
        opt.step()
        assert abs(loss5[0].item() - lo[0].item()) <= 2e-3 * abs(lo[0].item()), step
        first = first if first is not None else loss5[0].item()
    assert loss5[0].item() < first


# --------------------------------------------------------------------------- greedy decode
@pytest.mark.parametrize("prec", ["fp32", "3xtf32"])
@pytest.mark.parametrize("tag,gain", [("init", 1.0), ("stress", 3.0)])
def test_decode_matches_oracle_and_reference_golden(lib, golden, tag, gain, prec):
    """Greedy decode against the reference's own outputs, in both FP32-accurate arithmetics (FFMA and 3xTF32 on the
    tensor cores).  Tie-aware on EVERY discrete decision: a graph is compared when its smallest edge-logit margin (from
    the reference run) and its smallest quantiser margin (reported by the kernel: distance of a parameter logit to a
    rounding / sigmoid / arg-max tie) both exceed MARGIN."""
    m, o = make_model(0, gain, prec)
    for zt in ("mu", "prior"):
        z = torch.from_numpy(golden["%s_dec_%s_z" % (tag, zt)])
        gb = m.decode(z)
        Xo, Po, Ao, mg = o.decode(z, return_margins=True)
        qm = m.last_quant_margins.cpu().numpy()
        assert (qm > 0).all() and (qm < 10).all()
        ok = (golden["%s_dec_%s_minmargin" % (tag, zt)] > MARGIN) & (qm > MARGIN)          # tie-aware
        print(tag, zt, prec, "compared %d of %d graphs; min quantiser margin %.2e" % (ok.sum(), len(ok), qm.min()))
        assert ok.sum() >= 0.5 * len(ok)
        A = util.adj_from_masks(gb.adj.cpu().numpy().view(np.uint64))
        Pd = gb.params.cpu().numpy().astype(np.int32)
        assert np.array_equal(A[ok], golden["%s_dec_%s_adj" % (tag, zt)][ok])             # reference topology
        assert np.array_equal(Pd[ok], golden["%s_dec_%s_params" % (tag, zt)].astype(np.int32)[ok])
        assert np.abs(gb.X.cpu().numpy() - golden["%s_dec_%s_X" % (tag, zt)])[ok].max() <= 1e-6
        assert np.array_equal(A[ok], Ao.numpy()[ok])
        # edge insertion order of the returned graph objects (model.py:237-250)
        g0 = gb[int(np.nonzero(ok)[0][0])]
        s, d = g0.edges()
        assert (s.tolist(), d.tolist()) == O.edges_from_adj(A[int(np.nonzero(ok)[0][0])].tolist())
        assert np.allclose(m.last_margins.cpu().numpy()[ok],
                           golden["%s_dec_%s_minmargin" % (tag, zt)][ok], rtol=1e-2, atol=1e-5)


def test_api_compositions_and_rng_seams(lib):
    """model.py:255-268 (`encode_decode`, `generate`) and the `rsample` seam (model.py:284): with the same
    torch seed the class draws the same noise the reference's calls would, and the compositions equal
    their parts."""
    from torch.distributions import Normal
    idx = util.pick_by_alg([0, 3, 5, 7, 16, 18, 20, 31])
    X, P, E, A = util.dataset_graphs(idx)
    m, o = make_model(0, 3.0)
    G = _graphs(X, P, E)
    with torch.no_grad():
        q = m.encode(G)
    # reparameterize: explicit eps, and eps=None == Normal.rsample() on the same device generator
    eps = torch.randn(len(G), 128, generator=torch.Generator().manual_seed(2))
    z = m.reparameterize(q, eps)
    assert torch.equal(z, q.loc + q.scale * eps.cuda()) or (z - (q.loc + q.scale * eps.cuda())).abs().max() <= 1e-6
    torch.manual_seed(9); z1 = m.reparameterize(q)
    torch.manual_seed(9); z2 = Normal(q.loc, q.scale).rsample()
    assert (z1 - z2).abs().max().item() <= 1e-6
    # loss(eps=None) consumes the generator exactly like rsample: same value as injecting that draw
    torch.manual_seed(4)
    with torch.no_grad():
        l_none = m.loss(q, G)
    torch.manual_seed(4)
    e4 = torch.empty(len(G), 128, device="cuda").normal_()
    with torch.no_grad():
        l_inj = m.loss(q, G, eps=e4)
    assert all(abs(a.item() - b.item()) <= 1e-6 * abs(b.item()) + 1e-9 for a, b in zip(l_none, l_inj))
    # encode_decode(G) == decode(mu)  (deterministic branch), stochastic branch == decode(q.sample()) under the seed
    a = m.encode_decode(G)
    b = m.decode(q.loc)
    assert torch.equal(a.params, b.params) and torch.equal(a.adj, b.adj)
    torch.manual_seed(6); c = m.encode_decode(G, stochastic=True)
    torch.manual_seed(6); d = m.decode(Normal(q.loc, q.scale).sample())
    assert torch.equal(c.params, d.params) and torch.equal(c.adj, d.adj)
    # generate(n): prior draw on the CPU generator as the reference does (p_dist = Normal(0., 1.)), then decode
    torch.manual_seed(8); g1 = m.generate(64)
    torch.manual_seed(8); zp = Normal(0., 1.).sample((64, 128))
    Xo, Po, Ao, mg = o.decode(zp, return_margins=True)
    margins = torch.cat([l.flatten(1) for l in mg["edge"] + mg["self"]], 1).abs().min(1).values.numpy()
    ok = margins > 1e-4
    assert ok.sum() >= 32 and len(g1) == 64 and m.hidden == 64
    assert np.array_equal(util.adj_from_masks(g1.adj.cpu().numpy().view(np.uint64))[ok], Ao.numpy()[ok])
    assert np.array_equal(g1.params.cpu().numpy().astype(np.int32)[ok], Po.numpy().astype(np.int32)[ok])
    # empty input is an error, not a silent no-op
    with pytest.raises(ValueError):
        m.encode([])


def test_decode_ragged_batch_sizes_are_position_independent(lib):
    """Batches of 1, 3, 129 and 1000 graphs decode to exactly the rows of the 1000-graph batch (the greedy path compacts
    the graphs that gain an edge at every step: row lists of any length, including empty ones, must work)."""
    m, _ = make_model(0, 3.0)
    z = torch.randn(1000, 128, generator=torch.Generator().manual_seed(9))
    full = m.decode(z)
    for n in (1, 3, 129):
        part = m.decode(z[:n])
        assert torch.equal(part.params, full.params[:n]) and torch.equal(part.adj, full.adj[:n]) and torch.equal(part.X, full.X[:n])
    last = m.decode(z[-1:])
    assert torch.equal(last.params, full.params[-1:]) and torch.equal(last.adj, full.adj[-1:])
    # a model that decides no edge at all (every step's active list is empty) and one that decides all of them
    named = dict(m.named_parameters())
    for shift in (-50.0, 50.0):
        eb, sb = named["h_to_edge.2.bias"], named["h_to_edge_self.2.bias"]
        eb0, sb0 = eb.data.clone(), sb.data.clone()
        eb.data.add_(shift); sb.data.add_(shift)
        g = m.decode(z[:65])
        nbits = [bin(int(a) & ((1 << 49) - 1)).count("1") for a in g.adj.cpu().numpy().view(np.uint64)]
        assert set(nbits) == ({0} if shift < 0 else {48})          # 6 self-loops + 21 pairs x 2 directions
        eb.data.copy_(eb0); sb.data.copy_(sb0)


def test_decode_large_batch_properties(lib):
    """Full-size style checks that need no oracle: chunking invariance, legal parameter ranges,
    decode -> .syx -> _make_graph round trip."""
    from dxvae_b200.dxdata import graph_to_syx_bytes, voices_to_batch
    m, _ = make_model(0, 3.0)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(20000, 128, generator=g)
    m.max_chunk = 32768
    a = m.decode(z)
    m.max_chunk = 4096
    b = m.decode(z)
    assert torch.equal(a.params, b.params) and torch.equal(a.adj, b.adj) and torch.equal(a.X, b.X)
    P = a.params.cpu().numpy()
    assert P.min() >= 0 and P[:, 1:, 0:9].max() <= 99 and P[:, 0, 18].max() <= 31 and P[:, 1:, 20].max() <= 2
    voices = np.frombuffer(graph_to_syx_bytes(a)[6:-2], np.uint8).reshape(-1, 128)
    back = voices_to_batch(voices)
    # fc in fixed mode is stored mod 4 by _make_graph but the quantiser already emits 0..3: exact round trip
    assert torch.equal(back.params.cpu(), a.params.cpu().abs())
    assert np.abs(back.X.cpu().numpy() - a.X.cpu().numpy()).max() <= 1e-6


def test_graph_replayed_step_equals_eager_step(lib):
    """Small batches replay a captured CUDA graph over a batch-independent schedule (node-order levels,
    every teacher-forcing step on every graph): same losses and same weights as the eager, compacted step."""
    from dxvae_b200.train import Trainer
    idx = list(range(0, 1024, 4))
    X, P, E, A = util.dataset_graphs(idx)
    res = {}
    for mode, gmax in (("graph", 1024), ("eager", 0)):
        m, _ = make_model(0, 1.0)
        t = Trainer(m, lr=1e-3, w=(2, 5, 0.01))
        t.graph_max_batch = gmax
        data = t.upload(_graphs(X, P, E))
        g = torch.Generator().manual_seed(5)
        losses = []
        for step in range(3):
            pick = torch.randperm(len(idx), generator=g)[:96].tolist()       # a different batch every step
            eps = torch.randn(96, 128, generator=g)
            losses.append(t.step(data, pick, eps=eps).cpu())
        assert (len(t._graphs) == 1) == (mode == "graph")
        res[mode] = (torch.stack(losses), m._flat.clone())
    assert (res["graph"][0] - res["eager"][0]).abs().max().item() <= 1e-5 * res["eager"][0].abs().max().item()
    # AdamW normalises each gradient element, so elements whose gradient is pure round-off may move by +-lr
    # either way; everything else must land on the same weight
    diff = (res["graph"][1] - res["eager"][1]).abs()
    assert (diff > 2e-6).float().mean().item() <= 1e-4 and diff.max().item() <= 3.1e-3


# --------------------------------------------------------------------------- compacted teacher forcing
def test_device_step_schedule_matches_set_logic(lib):
    from dxvae_b200 import DXVAE, _abi, _lib
    from dxvae_b200.dxdata import mask_from_edges
    m = DXVAE()
    for n, p, seed in ((3000, 0.2, 1), (1, 1.0, 2), (1025, 0.0, 3), (2500, 1.0, 4)):
        E = util.random_edge_lists(n, p, seed)
        adj = torch.tensor([mask_from_edges(*e) for e in E], dtype=torch.int64, device="cuda")
        sp = np.zeros(34, np.int32)
        sr = torch.full((33 * n,), -1, dtype=torch.int32, device="cuda")
        spd = torch.empty(34, dtype=torch.int32, device="cuda")
        ws = m._workspace(_abi.OP_SCHEDULE, n)
        _lib.check(lib.dxvae_batch_steps(n, adj.data_ptr(), spd.data_ptr(), sr.data_ptr(), sp.ctypes.data, ws.data_ptr(),
                                         ws.numel(), st()), "steps")
        sr = sr.cpu().numpy()
        sets = [set(zip(s, d)) for s, d in E]
        t = 0
        for vi in range(1, 7):
            for vj in range(vi - 1, -1, -1):
                want = [b for b in range(n) if (vj, vi) in sets[b] or (vi, vj) in sets[b]]
                assert list(sr[sp[t]:sp[t + 1]]) == want, (vi, vj)
                t += 1
        for vi in range(1, 7):
            want = [b for b in range(n) if (vi, vi) in sets[b]]
            assert list(sr[sp[t]:sp[t + 1]]) == want, ("self", vi)
            t += 1
        for x in range(6):
            want = [b for b in range(n) if any(s == x and d > x for s, d in sets[b])]
            assert list(sr[sp[t]:sp[t + 1]]) == want, ("back-edge source", x)
            t += 1


def test_compacted_steps_equal_dense_replay(lib):
    """Skipping identity re-propagates must not change the function: same losses and gradients as
    replaying all 21 steps on every graph (fp32 path), on dataset and on arbitrary topologies."""
    idx = list(range(0, 1024, 8))
    X, P, E, A = util.dataset_graphs(idx)
    m, o = make_model(3, 3.0)
    torch.manual_seed(21)
    eps = torch.randn(len(idx), 128)
    for edges in (E, util.random_edge_lists(len(idx), 0.3, 17)):
        G = _graphs(X, P, edges)
        res = {}
        for mode in (True, False):
            m.compact_steps = mode
            m.zero_grad()
            out = m.forward(G, eps=eps)
            out[0].backward()
            res[mode] = ([t.item() for t in out], {n: p.grad.clone() for n, p in m.named_parameters()})
        for a, b in zip(res[True][0], res[False][0]):
            assert abs(a - b) <= 2e-6 * abs(b) + 1e-7
        for n, g in res[False][1].items():
            assert (res[True][1][n] - g).abs().max().item() <= 2e-5 * (g.abs().max().item() + 1e-30), n
    m.compact_steps = True


# --------------------------------------------------------------------------- DXVAE.train (model.py:374-391)
def test_train_method_semantics(lib, tmp_path):
    """epochs+1 passes, drop-last batching, in-place shuffle of the caller's list driven by Python's
    `random`, checkpoint per epoch that the reference layout can load, loss going down."""
    import random
    from dxvae_b200 import DXVAE
    idx = list(range(0, 1024, 4))[:200]
    X, P, E, A = util.dataset_graphs(idx)
    G = _graphs(X, P, E)
    ids = {id(g): i for i, g in enumerate(G)}
    torch.manual_seed(0)
    m = DXVAE(); m.verbose = False
    chk = str(tmp_path / "t.chk")
    random.seed(3)
    expect = list(range(len(G)))
    rnd = random.Random(3)
    for _ in range(3):                                   # epochs + 1 shuffles of the same list
        rnd.shuffle(expect)
    with torch.no_grad():
        before = m.forward(G, eps=torch.zeros(len(G), 128))[0].item()
    m.train(G, epochs=2, size_batch=64, lr=1e-3, checkpoint=chk)
    assert [ids[id(g)] for g in G] == expect             # same permutation random.shuffle(G) x3 would produce
    with torch.no_grad():
        after = m.forward(G, eps=torch.zeros(len(G), 128))[0].item()
    assert after < before                                 # 9 AdamW steps (3 per pass, remainder dropped)
    sd = torch.load(chk, map_location="cpu")
    o = O.OracleDXVAE()
    o.load_state_dict(sd)                                 # same 53 keys / shapes as the reference module
    for n, p in m.state_dict().items():
        assert torch.equal(sd[n], p.cpu()), n
    m2 = DXVAE(checkpoint=chk)                            # model.py:80-81
    m2.verbose = False
    with torch.no_grad():
        again = m2.forward(G, eps=torch.zeros(len(G), 128))[0].item()
    assert abs(again - after) <= 1e-5 * abs(after)

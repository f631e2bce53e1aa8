"""Full-size checks (BASELINE.json configs 3, 4 and 5) through size-independent properties —
the oracle cannot run a million patches, so these use what the domain offers:

  * patches are independent: a patch's latents / decoded voice do not depend on its position in
    the batch, on the chunking, or on which other patches share the launch (bit-exact in FP32);
  * decode -> .syx bulk dump -> _make_graph is a round trip on the decoded parameters;
  * every ELBO term is a batch mean: loss(batch) = mean of the losses of its halves and the
    gradient is the sum of the halves' gradients (the property the data-parallel shards rely on).

Sizes: 1,048,576 synthetic patches for encode / decode, the benchmark micro-batch (32768) for the
training step."""
import hashlib

import numpy as np
import pytest
import torch

import dxvae_oracle as O

pytestmark = pytest.mark.gpu

N_FULL = 1 << 20


@pytest.fixture(scope="module")
def lib():
    from dxvae_b200 import _lib
    return _lib.require_cuda()


def _model(gain=3.0):
    from dxvae_b200 import DXVAE
    o = O.make_weights(0, gain)
    m = DXVAE()
    m.load_state_dict(o.state_dict())
    m.verbose = False
    return m, o


def test_encode_one_million_patches_position_independent(lib):
    """cfg 3: 1 M synthetic patch graphs -> latents; a patch's (mu, std) is bit-identical wherever it
    sits (full run in 32768-chunks vs. a shuffled 50 000-patch subset in 8192-chunks), and a small
    sample matches the CPU oracle within the FP32 tolerance."""
    from dxvae_b200.dxdata import DXGraphBatch, voices_to_batch
    from dxvae_b200.synth import random_voices
    m, o = _model()
    voices = random_voices(N_FULL, seed=0)
    gb = voices_to_batch(voices)
    m.max_chunk = 32768
    with torch.no_grad():
        q = m.encode(gb)
    mu, sd = q.loc, q.scale
    assert mu.shape == (N_FULL, 128) and bool(torch.isfinite(mu).all()) and bool((sd > 0).all())
    g = torch.Generator().manual_seed(1)
    pick = torch.randperm(N_FULL, generator=g)[:50000].cuda()
    sub = DXGraphBatch(gb.X[pick], gb.params[pick], gb.adj[pick])
    m.max_chunk = 8192
    with torch.no_grad():
        q2 = m.encode(sub)
    assert torch.equal(q2.loc, mu[pick]) and torch.equal(q2.scale, sd[pick])
    # oracle on a sample taken across the whole range
    idx = torch.arange(0, N_FULL, N_FULL // 48)[:48]
    Xs, A = gb.X[idx.cuda()].cpu(), torch.zeros(len(idx), 7, 7)
    for k, i in enumerate(idx.tolist()):
        _, _, s, d = O.make_graph(voices[i])
        A[k, s, d] = 1.0
    mu_o, sd_o = o.encode(Xs, A)
    assert (mu[idx.cuda()].cpu() - mu_o).abs().max().item() <= 1e-5
    assert (sd[idx.cuda()].cpu() - sd_o).abs().max().item() <= 1e-5


def test_decode_one_million_patches_to_syx_round_trip(lib):
    """cfg 4: z ~ N(0,1) (torch.Generator seed 0, shape (1 M, 128)) -> greedy decode -> .syx bytes.
    Chunking / position invariance (bit-exact), legal parameter ranges, and the bulk dump read back
    through _make_graph semantics reproduces the decoded parameters for all 1 M voices."""
    from dxvae_b200.dxdata import graph_to_syx_bytes, voices_to_batch
    m, o = _model()
    z = torch.randn(N_FULL, 128, generator=torch.Generator().manual_seed(0))
    m.max_chunk = 32768
    a = m.decode(z)
    P = a.params
    assert P.shape == (N_FULL, 7, 21)
    assert float(P.abs().min()) >= 0 and float(P[:, 1:, 0:9].max()) <= 99 and float(P[:, 0, 18].max()) <= 31
    assert float(P[:, 1:, 20].max()) <= 2 and float(P[:, 0, 8].max()) <= 48          # rc quirk, transpose
    # the CPU oracle on a sample taken across the whole range: same topology and parameters wherever every discrete
    # decision clears its tie margin (edge logits from the oracle run, quantiser margins reported by the kernel)
    idx = torch.arange(0, N_FULL, N_FULL // 96)[:96]
    Xo, Po, Ao, mgo = o.decode(z[idx], return_margins=True)
    em = torch.cat([l.flatten(1) for l in mgo["edge"] + mgo["self"]], 1).abs().min(1).values.numpy()
    ok = (em > 2e-5) & (m.last_quant_margins[idx.cuda()].cpu().numpy() > 2e-5)
    assert ok.sum() >= 48
    from tests import util
    Ad = util.adj_from_masks(a.adj[idx.cuda()].cpu().numpy().view(np.uint64))
    assert np.array_equal(Ad[ok], Ao.numpy()[ok])
    assert np.array_equal(P[idx.cuda()].cpu().numpy().astype(np.int32)[ok], Po.numpy().astype(np.int32)[ok])
    # position / chunk invariance on a shuffled subset
    pick = torch.randperm(N_FULL, generator=torch.Generator().manual_seed(2))[:30000]
    m.max_chunk = 4096
    b = m.decode(z[pick])
    pc = pick.cuda()
    assert torch.equal(b.params, P[pc]) and torch.equal(b.adj, a.adj[pc]) and torch.equal(b.X, a.X[pc])
    # .syx round trip for every voice
    raw = graph_to_syx_bytes(a)
    assert len(raw) == 6 + 128 * N_FULL + 2 and raw[:6] == bytes([0xF0, 67, 0, 9, 32, 0]) and raw[-2:] == bytes([88, 0xF7])
    voices = np.frombuffer(raw[6:-2], np.uint8).reshape(-1, 128)
    assert int(voices[:, :118].max()) < 128                                         # 7-bit sysex payload
    back = voices_to_batch(voices)
    assert torch.equal(back.params, P.abs())
    assert float((back.X - a.X).abs().max()) <= 1e-6
    # host restatement of graph_to_syx on a slice: same bytes (checksum of the slice)
    sl = slice(123456, 123456 + 4096)
    want = O.graph_to_syx_bytes(P[sl].cpu().numpy())
    got = bytes([0xF0, 67, 0, 9, 32, 0]) + voices[sl].tobytes() + bytes([88, 0xF7])
    assert hashlib.sha256(got).hexdigest() == hashlib.sha256(want).hexdigest()


@pytest.mark.parametrize("precision,tol_l,tol_g", [("fp32", 2e-6, 2e-5), ("3xtf32", 2e-6, 2e-5), ("tf32", 1e-5, 2e-3)])
def test_train_step_is_a_batch_mean_at_benchmark_size(lib, precision, tol_l, tol_g):
    """cfg 5 micro-batch (32768 patches): the five loss terms equal the mean over the two halves and the
    gradient equals the sum of the halves' gradients computed with inv_batch = 1/32768 — exactly what
    the data-parallel ranks do before the all-reduce."""
    from dxvae_b200.dxdata import DXGraphBatch, voices_to_batch
    from dxvae_b200.synth import random_voices
    m, _ = _model(1.0)
    m.precision = precision
    m._ensure_flat()
    B = 32768
    gb = voices_to_batch(random_voices(B, seed=5))
    eps = torch.randn(B, 128, generator=torch.Generator().manual_seed(7)).cuda()
    w = (2.0, 5.0, 0.01)

    def run(lo, hi):
        d = m._prepare(DXGraphBatch(gb.X[lo:hi], gb.params[lo:hi], gb.adj[lo:hi]))
        g = torch.zeros_like(m._flat)
        loss5 = m.elbo_step(d, eps[lo:hi], w, grads=g, inv_batch=1.0 / B)
        return loss5.clone(), g

    l_all, g_all = run(0, B)
    l_a, g_a = run(0, B // 2)
    l_b, g_b = run(B // 2, B)
    assert bool(torch.isfinite(l_all).all()) and bool(torch.isfinite(g_all).all())
    for k in range(5):
        want = l_all[k].item()
        assert abs((l_a[k] + l_b[k]).item() - want) <= tol_l * abs(want) + 1e-9, (k, want)
    gs = g_a + g_b
    # per-tensor max-norm-relative comparison
    for n, p in m.named_parameters():
        lo = p.data_ptr() - m._flat.data_ptr()
        assert lo % 4 == 0
        lo //= 4
        a, b = g_all[lo:lo + p.numel()], gs[lo:lo + p.numel()]
        rel = (a - b).abs().max().item() / (a.abs().max().item() + 1e-30)
        assert rel <= tol_g, (n, rel)


def test_tf32_and_fp32_training_trajectories_agree(lib):
    """Twelve AdamW steps on 8192 synthetic patches from the same initial weights and the same injected noise: the
    tensor-core path (TF32 products, compacted schedules) must follow the FP32 FFMA path — total loss within 2e-3
    relative at every step — and both must descend."""
    from dxvae_b200 import DXVAE
    from dxvae_b200.dxdata import voices_to_batch
    from dxvae_b200.synth import random_voices
    from dxvae_b200.train import Trainer
    B = 8192
    pool = voices_to_batch(random_voices(B, seed=11))
    idx = list(range(B))
    traj = {}
    for prec in ("fp32", "tf32", "3xtf32"):
        torch.manual_seed(0)
        m = DXVAE(); m.verbose = False; m.precision = prec
        tr = Trainer(m, lr=1e-3)
        g = torch.Generator(device="cuda").manual_seed(5)
        losses = []
        for step in range(12):
            eps = torch.randn(B, 128, device="cuda", generator=g)
            losses.append(tr.step(pool, idx, eps=eps)[0].item())
        traj[prec] = losses
        assert all(np.isfinite(losses)) and losses[-1] < 0.9 * losses[0], (prec, losses)
    for a, b, c in zip(traj["fp32"], traj["tf32"], traj["3xtf32"]):
        assert abs(a - b) <= 2e-3 * abs(a), (traj["fp32"], traj["tf32"])
        assert abs(a - c) <= 2e-5 * abs(a), (traj["fp32"], traj["3xtf32"])      # the FP32-accurate mode tracks the FFMA path

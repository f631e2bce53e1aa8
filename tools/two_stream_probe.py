"""Does running two independent half-batches on two streams overlap the tensor-bound GEMMs of one with the
HBM-bound element-wise kernels of the other?  Compares 1 x M with 2 x M/2 (same total work per step)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dxvae_b200 import DXVAE, _abi, _lib
from dxvae_b200.dxdata import DXGraphBatch, voices_to_batch
from dxvae_b200.synth import random_voices

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
m = DXVAE(); m.verbose = False; m.precision = "tf32"; m._ensure_flat()
L = _lib.lib()
pool = voices_to_batch(random_voices(2 * M, seed=1))
w = (2.0, 5.0, 0.01)

def one_stream(i):
    lo = (i % 2) * M
    d = m._prepare(DXGraphBatch(pool.X[lo:lo + M], pool.params[lo:lo + M], pool.adj[lo:lo + M]))
    eps = torch.empty(M, 128, device="cuda").normal_()
    g = torch.zeros_like(m._flat)
    return m.elbo_step(d, eps, w, grads=g, inv_batch=1.0 / M), g

streams = [torch.cuda.Stream() for _ in range(NS)]
wss = [torch.empty(int(L.dxvae_workspace_bytes(_abi.OP_TRAIN, M // NS)), dtype=torch.uint8, device="cuda") for _ in range(NS)]
gs = [torch.zeros_like(m._flat) for _ in range(NS)]
l5 = [torch.empty(5, device="cuda") for _ in range(NS)]

def multi_stream(i):
    lo = (i % 2) * M
    h = M // NS
    ds, es = [], []
    for k in range(NS):
        a = lo + k * h
        ds.append(m._prepare(DXGraphBatch(pool.X[a:a + h], pool.params[a:a + h], pool.adj[a:a + h])))
        es.append(torch.empty(h, 128, device="cuda").normal_())
    cur = torch.cuda.current_stream()
    for k in range(NS):
        streams[k].wait_stream(cur)
        with torch.cuda.stream(streams[k]):
            d = ds[k]
            gs[k].zero_()
            _lib.check(L.dxvae_elbo_step(
                m._flat.data_ptr(), d.B, d.Xn.data_ptr(), d.cls.data_ptr(), d.adj.data_ptr(), d.n_levels,
                d.level_ptr.ctypes.data, d.level_rows.data_ptr(), es[k].data_ptr(), w[0], w[1], w[2], 1.0 / M,
                l5[k].data_ptr(), None, None, gs[k].data_ptr(), wss[k].data_ptr(), wss[k].numel(), m._prec(),
                d.step_ptr.ctypes.data, d.step_rows.data_ptr(), torch.cuda.current_stream().cuda_stream), "elbo")
    for k in range(NS):
        cur.wait_stream(streams[k])
    g = gs[0]
    for k in range(1, NS):
        g = g + gs[k]
    return sum(l5), g

for name, fn in (("1 stream", one_stream), ("%d streams" % NS, multi_stream), ("1 stream", one_stream), ("%d streams" % NS, multi_stream)):
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(4):
        out = fn(i)
    e1.record(); torch.cuda.synchronize()
    print("%s: %.2f ms/step (M=%d)  loss %.5f gnorm %.5f" % (name, e0.elapsed_time(e1) / 4, M, out[0][0].item(), out[1].norm().item()), flush=True)

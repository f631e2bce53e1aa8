"""How long does the HOST take to enqueue one training step vs how long the GPU takes to run it?"""
import sys, time, torch
sys.path.insert(0, '.')
from dxvae_b200 import DXVAE
from dxvae_b200.dxdata import DXGraphBatch, voices_to_batch
from dxvae_b200.synth import random_voices
from dxvae_b200.train import Trainer
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
m = DXVAE(); m.verbose = False; m.precision = sys.argv[2] if len(sys.argv) > 2 else "tf32"; m._ensure_flat()
tr = Trainer(m)
pool = voices_to_batch(random_voices(M, 1))
d = m._prepare(pool)
eps = torch.randn(M, 128, device="cuda")
for _ in range(3):
    tr.grad_step(d, eps, M); tr.apply()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    tr.grad_step(d, eps, M); tr.apply()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("M=%d host enqueue %.2f ms, total %.2f ms" % (M, (t1 - t0) * 1e3, (t2 - t0) * 1e3))
t0 = time.perf_counter(); d = m._prepare(pool); torch.cuda.synchronize(); print("prepare %.2f ms" % ((time.perf_counter() - t0) * 1e3))

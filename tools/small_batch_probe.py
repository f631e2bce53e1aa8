"""B=128 train step (BASELINE config 2 shape): eager vs CUDA-graph replay in every arithmetic."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dxvae_b200 import DXVAE
from dxvae_b200.dxdata import voices_to_batch
from dxvae_b200.synth import random_voices
from dxvae_b200.train import Trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pool = voices_to_batch(random_voices(max(1024, B), seed=3))
for prec in (sys.argv[2:] or ["3xtf32", "tf32", "fp32"]):
    for gmax in (0, 1024):
        torch.manual_seed(0)
        m = DXVAE(); m.verbose = False; m.precision = prec
        tr = Trainer(m); tr.graph_max_batch = gmax
        idx = list(range(B))
        for _ in range(3):
            tr.step(pool, idx)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20):
            tr.step(pool, idx)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20
        print("B=%d %s %s: %.2f ms/step, %.0f patches/s" % (B, prec, "graph" if gmax else "eager", dt * 1e3, B / dt), flush=True)

"""Host-side time of one public-API training step (opt.zero_grad(); model(G)[0].backward(); opt.step()) split by stage,
at the benchmark's micro-batch, with shuffled lists of host graph objects: what the CPU does while the GPU runs the
previous step.  python tools/e2e_host_probe.py [micro_batch]"""
import os
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dxvae_b200 import DXVAE  # noqa: E402
from dxvae_b200.dxdata import DXGraphBatch, voices_to_batch  # noqa: E402
from dxvae_b200.synth import random_voices  # noqa: E402
from dxvae_b200.train import FusedAdamW  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
torch.manual_seed(0)
m = DXVAE(); m.verbose = False
host = voices_to_batch(random_voices(2 * M, seed=1)).cpu().pin_memory()
lists = [list(host[k * M:(k + 1) * M]) for k in range(2)]
for k, gl in enumerate(lists):
    random.Random(k).shuffle(gl)
opt = FusedAdamW(m)
T = {"from_graphs": 0.0, "prepare_rest": 0.0, "forward_rest": 0.0, "backward": 0.0, "opt_step": 0.0}
orig_fg = DXGraphBatch.from_graphs.__func__
orig_prep = DXVAE._prepare


def fg(cls, graphs, staging=False):
    t = time.perf_counter(); r = orig_fg(cls, graphs, staging); T["from_graphs"] += time.perf_counter() - t; return r


def prep(self, G, *a, **k):
    t = time.perf_counter(); f0 = T["from_graphs"]; r = orig_prep(self, G, *a, **k)
    T["prepare_rest"] += time.perf_counter() - t - (T["from_graphs"] - f0); return r


DXGraphBatch.from_graphs = classmethod(fg)
DXVAE._prepare = prep
N = 6
for it in range(2 + N):
    if it == 2:
        torch.cuda.synchronize(); T = {k: 0.0 for k in T}; t_all = time.perf_counter()
    G = lists[it % 2]
    opt.zero_grad()
    t = time.perf_counter(); p0 = T["from_graphs"] + T["prepare_rest"]
    loss = m(G)[0]
    T["forward_rest"] += time.perf_counter() - t - (T["from_graphs"] + T["prepare_rest"] - p0)
    t = time.perf_counter(); loss.backward(); T["backward"] += time.perf_counter() - t
    t = time.perf_counter(); opt.step(); T["opt_step"] += time.perf_counter() - t
cpu = time.perf_counter() - t_all
torch.cuda.synchronize()
wall = time.perf_counter() - t_all
print("micro-batch %d: host time per step %.1f ms (wall incl. GPU drain %.1f ms)" % (M, 1e3 * cpu / N, 1e3 * wall / N))
for k, v in T.items():
    print("   %-14s %.1f ms" % (k, 1e3 * v / N))

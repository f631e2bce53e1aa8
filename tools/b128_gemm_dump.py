"""B = 128 train step (BASELINE config 2): per-launch CUDA-event times of every product of ONE eager step, grouped per
shape (DX_PROF_DUMP lines -> tools/gemm_dump_summary.py), plus the step time with and without the events.
usage: DX_PROF_DUMP=1 python tools/b128_gemm_dump.py 2> dump.err ; python tools/gemm_dump_summary.py dump.err"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dxvae_b200 import DXVAE, _lib
from dxvae_b200.dxdata import voices_to_batch
from dxvae_b200.synth import random_voices
from dxvae_b200.train import Trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = _lib.require_cuda()
pool = voices_to_batch(random_voices(1024, seed=3))
m = DXVAE(); m.verbose = False
tr = Trainer(m); tr.graph_max_batch = 0
idx = list(range(B))
for _ in range(3):
    tr.step(pool, idx)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    tr.step(pool, idx)
torch.cuda.synchronize()
print("B=%d eager: %.2f ms/step" % (B, (time.perf_counter() - t0) / 20 * 1e3))
L.dxvae_prof_begin(4096)
n0 = _lib.launch_count()
tr.step(pool, idx)
ms = (ctypes.c_double * 3)(); fl = (ctypes.c_double * 3)(); nv = (ctypes.c_longlong * 3)()
L.dxvae_prof_end(ms, fl, nv)
print("launches %d; products: %s launches, %s ms (event-timed, one step)" % (_lib.launch_count() - n0, list(nv), ["%.2f" % x for x in ms]))

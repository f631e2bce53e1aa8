#!/bin/bash
# per-kernel durations of greedy decode (fp32) and inference encode (fp32) at 16384 / 32768 patches
mkdir -p gpurun_out
cat > /tmp/dec.py <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from dxvae_b200 import DXVAE
from dxvae_b200.dxdata import voices_to_batch
from dxvae_b200.synth import random_voices
m = DXVAE(); m.verbose = False
z = torch.randn(16384, 128, device="cuda")
gb = voices_to_batch(random_voices(32768, seed=3))
for _ in range(2):
    m.decode(z)
    with torch.no_grad():
        m.encode(gb)
torch.cuda.synchronize(); t0 = time.perf_counter(); m.decode(z); torch.cuda.synchronize(); print("decode", 16384 / (time.perf_counter() - t0))
t0 = time.perf_counter()
with torch.no_grad():
    m.encode(gb)
torch.cuda.synchronize(); print("encode", 32768 / (time.perf_counter() - t0))
PY
python /tmp/dec.py && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 500 --csv --log-file gpurun_out/launches_dec.csv python /tmp/dec.py > gpurun_out/ncu_dec.log 2>&1
echo rc=$?

#!/bin/bash
# per-kernel durations of the B=128 train step (eager, tf32)
mkdir -p gpurun_out
cat > /tmp/sb.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from dxvae_b200 import DXVAE
from dxvae_b200.dxdata import voices_to_batch
from dxvae_b200.synth import random_voices
from dxvae_b200.train import Trainer
pool = voices_to_batch(random_voices(1024, seed=3))
m = DXVAE(); m.verbose = False; m.precision = "tf32"
tr = Trainer(m); tr.graph_max_batch = 0
for _ in range(4):
    tr.step(pool, list(range(128)))
torch.cuda.synchronize()
PY
python /tmp/sb.py && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2800 -c 900 --csv --log-file gpurun_out/launches_b128.csv python /tmp/sb.py > gpurun_out/ncu_b128.log 2>&1
echo rc=$?

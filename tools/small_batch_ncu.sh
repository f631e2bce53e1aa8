#!/bin/bash
# B = 128 train step (BASELINE config 2): eager vs CUDA-graph replay in every arithmetic, then the ncu launch list of one
# eager step in the default arithmetic (the script prints its launch count; the summary takes it from the end of the capture)
mkdir -p gpurun_out
python tools/small_batch_probe.py 128 | tee gpurun_out/b128_probe.log
cat > /tmp/sb.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from dxvae_b200 import DXVAE, _lib
from dxvae_b200.dxdata import voices_to_batch
from dxvae_b200.synth import random_voices
from dxvae_b200.train import Trainer
pool = voices_to_batch(random_voices(1024, seed=3))
m = DXVAE(); m.verbose = False
tr = Trainer(m); tr.graph_max_batch = 0
for _ in range(3):
    tr.step(pool, list(range(128)))
torch.cuda.synchronize(); n0 = _lib.launch_count()
tr.step(pool, list(range(128)))
torch.cuda.synchronize(); print("%d launches" % (_lib.launch_count() - n0))
PY
python /tmp/sb.py | tee gpurun_out/b128_plain.log && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b128.csv python /tmp/sb.py > gpurun_out/ncu_b128.log 2>&1
echo rc=$?
python - <<'PY'
import csv, re, collections
n = int(re.findall(r"(\d+) launches", open("gpurun_out/b128_plain.log").read())[0])
rows = list(csv.DictReader(l for l in open("gpurun_out/launches_b128.csv") if not l.startswith("==")))
def dur(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    return v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
part = rows[-(n + 40):]           # (torch's own index_select / normal_ kernels are in the list too)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in part:
    k = re.sub(r"\(.*", "", r["Kernel Name"])[:100]; agg[k][0] += 1; agg[k][1] += dur(r)
tot = sum(a[1] for a in agg.values())
print("# B=128 train step: %d kernels, %.2f ms of kernel time" % (len(part), tot))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%8.3f ms %5.1f%% n=%4d  avg %5.1f us  %s" % (t, 100 * t / tot, c, 1e3 * t / c, k))
PY

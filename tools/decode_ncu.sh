#!/bin/bash
# per-kernel durations of greedy decode and inference encode (default arithmetic) at 16384 / 32768 patches: the ncu
# launch list of ONE decode followed by ONE encode (the script prints how many kernels each launched; the summary below
# takes them from the end of the capture)
mkdir -p gpurun_out
cat > /tmp/dec.py <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from dxvae_b200 import DXVAE, _lib
from dxvae_b200.dxdata import voices_to_batch
from dxvae_b200.synth import random_voices
m = DXVAE(); m.verbose = False
z = torch.randn(16384, 128, device="cuda")
gb = voices_to_batch(random_voices(32768, seed=3))
m.decode(z)
with torch.no_grad():
    m.encode(gb)
torch.cuda.synchronize()
n0 = _lib.launch_count(); t0 = time.perf_counter(); m.decode(z); torch.cuda.synchronize(); t1 = time.perf_counter(); n1 = _lib.launch_count()
with torch.no_grad():
    m.encode(gb)
torch.cuda.synchronize(); t2 = time.perf_counter(); n2 = _lib.launch_count()
print("decode %.0f patches/s, %d launches; encode %.0f patches/s, %d launches" % (16384 / (t1 - t0), n1 - n0, 32768 / (t2 - t1), n2 - n1))
PY
python /tmp/dec.py | tee gpurun_out/dec_plain.log && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dec.csv python /tmp/dec.py > gpurun_out/ncu_dec.log 2>&1
echo rc=$?
python - <<'PY'
import csv, re, collections
plain = open("gpurun_out/dec_plain.log").read()
nd, ne = (int(x) for x in re.findall(r"(\d+) launches", plain))
rows = list(csv.DictReader(l for l in open("gpurun_out/launches_dec.csv") if not l.startswith("==")))
def dur(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    return v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
for name, part in (("greedy decode, 16384 patches", rows[-(nd + ne):-ne]), ("encode, 32768 patches", rows[-ne:])):
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in part:
        k = re.sub(r"\(.*", "", r["Kernel Name"])[:100]; agg[k][0] += 1; agg[k][1] += dur(r)
    tot = sum(a[1] for a in agg.values())
    print("# %s: %d launches, %.2f ms of kernel time" % (name, len(part), tot))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print("%8.3f ms %5.1f%% n=%4d  %s" % (t, 100 * t / tot, n, k))
PY

"""DRAM traffic of the designated GEMM launches against their algorithmic bytes, from an `ncu --set full` capture of
`python tools/x3_probe.py one M N K` (forward, dgrad, wgrad of one shape, each launched twice: the SECOND launch of each
form is read).  Writes profiles/<prefix>_tc_gemm_traffic.json, which bench.py reports as roofline.traffic(+_detail).
usage: python tools/traffic_from_ncu.py <file.ncu-rep | raw-page .csv> <profiles prefix> <precision: 3xtf32|tf32> M N K"""
import csv
import json
import os
import subprocess
import sys

rep, prefix, prec = sys.argv[1], sys.argv[2], sys.argv[3]
M, N, K = (int(a) for a in sys.argv[4:7])
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, body = rows[0], rows[1], rows[2:]


def val(r, name, scale=True):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "usecond": 1e-6,
            "msecond": 1e-3, "nsecond": 1e-9}.get(u, 1)
    return v * mult if scale else v


gemm = [r for r in body if "k_tc_gemm" in r[hdr.index("Kernel Name")]]
assert len(gemm) >= 3, "expected the forward, dgrad, wgrad launches of the second round, got %d" % len(gemm)
gemm = gemm[-3:]
forms = ["forward  y = x W^T", "dgrad    dx = dy W", "wgrad    dW += dy^T x"]
alg = [4.0 * (M * K + N * K + M * N), 4.0 * (M * N + N * K + M * K), 4.0 * (M * N + M * K + N * K)]
out = {"shape": [M, N, K], "precision": prec, "source": "ncu --set full --clock-control none of tools/x3_probe.py one %d %d %d" % (M, N, K),
       "launches": []}
for f in range(3):
    r = gemm[f]                                            # second round (captured with -s 3 -c 3)
    rd, wr, t = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")
    out["launches"].append({"form": forms[f], "kernel": r[hdr.index("Kernel Name")][:60], "dram_read_bytes": rd, "dram_write_bytes": wr,
                            "algorithmic_bytes": alg[f], "dram_over_algorithmic": (rd + wr) / alg[f], "duration_us_under_ncu": t * 1e6,
                            "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False)})
out["dram_bytes_per_launch"] = out["launches"][0]["dram_read_bytes"] + out["launches"][0]["dram_write_bytes"]
out["algorithmic_bytes_per_launch"] = alg[0]
path = os.path.join("profiles", "%s_tc_gemm_traffic.json" % prefix)
old = {}
if os.path.exists(path):
    old = json.load(open(path))
old[prec] = out
json.dump(old, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))

"""Turns gpurun_out/{launches_<tag>.csv, prof_tc_<tag>.ncu-rep} into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <gpurun tag> <profiles prefix, e.g. r01b> [note]"""
import collections
import csv
import re
import subprocess
import sys

tag, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
lines = [l for l in open("gpurun_out/launches_%s.csv" % tag) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    key = re.sub(r"\(.*", "", row["Kernel Name"])[:120]
    agg[key][0] += 1; agg[key][1] += v; tot += v
with open("profiles/%s_launches_summary.txt" % out, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none over `python bench.py --steps 1 --warmup 1 "
            "--no-cpu --no-extra --micro-batch 32768` (cold-cache, serialised: compare SHARES, not absolutes)\n# %s\n" % note)
    f.write("# captured launches: %d, total %.2f ms\n" % (sum(a[0] for a in agg.values()), tot))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%9.3f ms %5.1f%% n=%5d  %s\n" % (t, 100 * t / tot, n, k))
import os
rawcsv = "gpurun_out/prof_tc_%s_raw.csv" % tag
raw = open(rawcsv).read() if os.path.exists(rawcsv) else subprocess.run(
    ["ncu", "-i", "gpurun_out/prof_tc_%s.ncu-rep" % tag, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
with open("profiles/%s_tc_gemm_ncu.txt" % out, "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm (same command)\n# %s\n" % note)
    for r in rows[2:]:
        f.write("---\n")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                f.write("%s = %s %s\n" % (w, r[i][:140], units[i]))
print(open("profiles/%s_launches_summary.txt" % out).read()[:2400])

"""Source-level stall view of one captured launch: the SASS rows with the most warp-stall samples.
usage: python tools/ncu_stall_rows.py <file.ncu-rep> <launch index among the captured ones> <out.txt> [rows=48]"""
import csv
import subprocess
import sys

rep, which, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 48
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(which), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur); continue
    if cur is not None:
        cur["rows"].append(r)
s = secs[0]
hdr, body = s["rows"][0], s["rows"][1:]
i_s, i_src, i_ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(b[i_s]) for b in body)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for b in body:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(b[i])
with open(out, "w") as f:
    f.write("# %s\n# %d SASS rows, %d warp-stall samples; by reason: %s\n" %
            (s["name"][:110], len(body), tot, ", ".join("%s %d" % kv for kv in sorted(agg.items(), key=lambda kv: -kv[1])[:8])))
    f.write("# row  samples  executed  instruction  top stall reasons\n")
    top = sorted(range(len(body)), key=lambda k: -int(body[k][i_s]))[:top_n]
    for k in sorted(top):
        b = body[k]
        st = sorted(((hdr[i], int(b[i])) for i in stall_cols if int(b[i]) > 0), key=lambda kv: -kv[1])[:3]
        f.write("%6d %6s %9s  %-72s %s\n" % (k, b[i_s], b[i_ex], b[i_src].strip()[:72], st))
print(open(out).read()[:1500])

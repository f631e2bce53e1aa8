"""Selected raw counters of every captured launch of an .ncu-rep whose kernel name matches a regex: L2 -> SM feed
(lts2xbar bytes, L2 throughput, hit rate), shared-memory pipe, tensor pipe, issue / stall figures.
usage: python tools/ncu_counters.py <file.ncu-rep | raw-page .csv> <kernel regex> > profiles/<name>.txt"""
import csv
import re
import subprocess
import sys

rep, pat = sys.argv[1], re.compile(sys.argv[2])
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, body = rows[0], rows[1], rows[2:]
want = [r"^Kernel Name$", r"^launch__grid_size$", r"^launch__registers_per_thread$", r"^gpu__time_duration\.sum$",
        r"^sm__cycles_elapsed\.avg$", r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$",
        r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active$", r"^sm__inst_executed_pipe_tensor.*\.sum$",
        r"^sm__inst_executed_pipe_uniform\.sum$",
        r"^lts__t_bytes\.sum$", r"^lts__t_sector_hit_rate\.pct$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^derived__lts__lts2xbar_bytes\.sum\.per_second$", r"^lts__t_sectors_srcunit_tex.*\.sum$",
        r"^l1tex__m_xbar2l1tex_read_bytes\.sum(\.per_second)?$", r"^l1tex__m_xbar2l1tex_read_bytes_mem_.*tma.*\.sum$",
        r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?$",
        r"^smsp__inst_executed\.sum$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio$", r"^smsp__average_warp_latency_issue_stalled_.*\.ratio$"]
cols = [i for i, h in enumerate(hdr) if any(re.search(w, h) for w in want)]
for r in body:
    if not pat.search(r[hdr.index("Kernel Name")]):
        continue
    print("---")
    for i in cols:
        v = r[i]
        if v in ("", "0", "n/a") and not hdr[i].startswith(("Kernel", "launch")):
            continue
        print("%s = %s %s" % (hdr[i], v[:120], units[i]))

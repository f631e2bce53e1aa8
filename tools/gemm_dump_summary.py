"""Groups the DX_PROF_DUMP lines of a bench run ([gemm] cls M N K form us tflops) per shape: launches, total ms, TFLOP/s.
usage: python tools/gemm_dump_summary.py <bench stderr file>"""
import collections
import re
import sys

agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for ln in open(sys.argv[1]):
    m = re.match(r"\[gemm\] cls=(\d+) M=(\d+) N=(\d+) K=(\d+) form=(\d+) us=([\d.]+)", ln)
    if not m:
        continue
    c, M, N, K, f, us = int(m[1]), int(m[2]), int(m[3]), int(m[4]), int(m[5]), float(m[6])
    Mb = M if M >= 30000 or M < 600 else (M // 2000) * 2000     # bucket the compacted-step row counts
    a = agg[(c, f % 4, Mb, N, K)]
    a[0] += 1; a[1] += us; a[2] += 2.0 * M * N * K
tot = sum(a[1] for a in agg.values())
print("form: 0 fwd (y = x W^T), 1 dgrad (dx = dy W), 3 wgrad (dW += dy^T x); M bucketed for the compacted steps; total %.2f ms" % (tot / 1e3))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("cls %d form %d M~%6d N %5d K %6d  n=%4d  %8.2f ms %5.1f%%  %7.1f TF/s" % (*k, a[0], a[1] / 1e3, 100 * a[1] / tot, a[2] / (a[1] * 1e-6) / 1e12))

#!/bin/bash
# Quick iteration pass (1 GPU): GPU tests, the 3xTF32 GEMM probe, and a short bench with the per-shape GEMM dump.
mkdir -p gpurun_out
TAG=${TAG:-it}
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests_$TAG.log
if [ -n "$PROBE" ]; then echo "== probe"; timeout 600 python tools/x3_probe.py > gpurun_out/probe_$TAG.log 2>&1; echo "probe rc=$?"; cat gpurun_out/probe_$TAG.log; fi
echo "== bench"; DX_PROF_DUMP=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu ${BENCH_ARGS:---no-extra} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
j = json.load(open("gpurun_out/bench_$TAG.json"))
r = j["roofline"]
print("value %.0f  ms/step %.2f  e2e %.0f  gemm %.1f TF/s  gemm ms/step %.2f  share %.2f  clocks %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], r["achieved"], r["classes"][2]["ms_per_step"], r["share_of_step"], j["clocks"]))
for k, v in j.get("extra", {}).items(): print("  ", k, v)
PY
python tools/gemm_dump_summary.py gpurun_out/bench_$TAG.err | head -70

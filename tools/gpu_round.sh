#!/bin/bash
# Full evidence pass for the current tree (1 GPU): smoke, GPU parity tests, default bench, the ncu launch
# list of the benchmark's step at the benchmark micro-batch and one --set full capture of the top GEMM.
# Every ncu pass runs only after the same command exited 0 without ncu.  Outputs in gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-cur}
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/tests_$TAG.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
if [ -n "$PROFILE" ]; then
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --micro-batch ${MB:-32768}"
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1100} -c ${COUNT:-1000} --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
# .ncu-rep files of these kernels are ~14 MB per launch (18 k SASS rows with source) and gpurun copies back at most 64 MiB:
# the summaries are extracted HERE and the reports deleted
$CMD > gpurun_out/prof_plain2_$TAG.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none -k regex:k_tc_gemm -s ${GSKIP:-600} -c ${GCOUNT:-10} \
    -o gpurun_out/prof_tc_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/prof_tc_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_tc_${TAG}_raw.csv 2>/dev/null; rm -f gpurun_out/prof_tc_$TAG.ncu-rep
# designated products (forward / dgrad / wgrad of M=32768 N=1536 K=512, cold L2) for the traffic-vs-algorithmic-bytes figure,
# the L2-feed counters and the source-level stall view: 3xTF32 and plain TF32
for VB in 32 16; do
  ONE="python tools/x3_probe.py one 32768 1536 512 $VB"
  $ONE > gpurun_out/one_plain_${VB}_$TAG.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 3 -c 3 -o gpurun_out/prof_one_${VB}_$TAG -f $ONE > gpurun_out/ncu_one_${VB}_$TAG.log 2>&1
  echo "one $VB rc=$?"
  ncu -i gpurun_out/prof_one_${VB}_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_one_${VB}_${TAG}_raw.csv 2>/dev/null
  for W in 0 1 2; do python tools/ncu_stall_rows.py gpurun_out/prof_one_${VB}_$TAG.ncu-rep $W gpurun_out/stall_${VB}_${W}_$TAG.txt > /dev/null 2>&1; done
  rm -f gpurun_out/prof_one_${VB}_$TAG.ncu-rep
done
fi
du -sh gpurun_out | tail -1   # (gpurun copies back at most 64 MiB)

#!/bin/bash
# ncu evidence for bench.py's step (1 GPU).  Each ncu pass runs only after the same command
# exited 0 without ncu (B200_PROFILING.md).  Outputs in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --micro-batch ${MB:-8192}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1500} -c ${COUNT:-1400} --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gemm -s ${GSKIP:-700} -c 3 \
    -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
ls -la gpurun_out

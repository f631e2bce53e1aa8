#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --precision tf32 --micro-batch ${MB:-8192}"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1500} -c ${COUNT:-1400} --csv \
    --log-file gpurun_out/launches_tf32.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s ${GSKIP:-400} -c 8 \
    -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"

#!/bin/bash
# B = 128 (BASELINE config 2) iteration pass: host enqueue time vs GPU time, eager vs CUDA-graph replay, and the effect of
# the few-row tile rule (DX_X3_FEW_ROWS); optionally (NCU=1) the per-kernel launch list of one eager step.
mkdir -p gpurun_out
TAG=${TAG:-b128}
for FR in ${FEW_ROWS_LIST:-0 256}; do
  echo "== DX_X3_FEW_ROWS=$FR"
  DX_X3_FEW_ROWS=$FR python tools/host_bound.py 128 3xtf32 2>&1 | tail -4
  DX_X3_FEW_ROWS=$FR python tools/small_batch_probe.py 128 3xtf32 2>&1 | tail -2
done | tee gpurun_out/b128_iter_$TAG.log
if [ -n "$NCU" ]; then DX_X3_FEW_ROWS=${NCU_FEW_ROWS:-0} bash tools/small_batch_ncu.sh > gpurun_out/b128_ncu_$TAG.log 2>&1; tail -30 gpurun_out/b128_ncu_$TAG.log; fi

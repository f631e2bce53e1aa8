#!/bin/bash
# ncu --set full captures of the HBM-bound kernels of the train step (one GPU).  Runs after the same command exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --micro-batch 32768"
$CMD > gpurun_out/prof_plain_elem.log 2>&1 || exit 1
for K in k_cell_bwd k_edge_head_fwd; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 4 -o gpurun_out/prof_$K -f $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "$K rc=$?"
done
for K in cell_fwd head_sum msg_bwd; do
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$K -s 40 -c 4 -o gpurun_out/prof_$K -f $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "$K rc=$?"
done

#!/bin/bash
# N-GPU data-parallel bench (torchrun, NCCL) + the reference arm.
N=${1:-2}
mkdir -p gpurun_out
timeout ${DP_TIMEOUT:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
echo "dp rc=$?"; cat gpurun_out/bench_dp$N.json; tail -5 gpurun_out/bench_dp$N.err

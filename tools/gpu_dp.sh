#!/bin/bash
# N-GPU data-parallel pass (run under `gpurun --gpus N`): the NCCL N-rank == 1-rank test, then the bench on N GPUs
# (torchrun, NCCL; NCCL_DEBUG=INFO goes to stderr so stdout stays the one JSON line).
N=${1:-2}
mkdir -p gpurun_out
echo "== dp test"; timeout 900 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_dp$N.log 2>&1; echo "dp test rc=$?"; tail -5 gpurun_out/tests_dp$N.log
echo "== dp bench"
NCCL_DEBUG=INFO timeout ${DP_TIMEOUT:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
echo "dp rc=$?"; cat gpurun_out/bench_dp$N.json; grep -E "NVLS|nranks|Connected all" gpurun_out/bench_dp$N.err | head -5

#!/bin/bash
# Tensor-core path bring-up: unit tests first (bounded by timeout: a wrong descriptor can hang).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tf32.py -q --tb=line -p no:cacheprovider -s > gpurun_out/tc_tests.log 2>&1
echo "tc tests rc=$?"; tail -40 gpurun_out/tc_tests.log

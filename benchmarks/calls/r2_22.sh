#!/bin/bash
# Round 2, GPU call 22 (2 GPUs): multi-rank tests on the final code (authentication now signs SHA-256 tree digests).
mkdir -p gpurun_out
export AGB_FLAG_TIMEOUT_S=60
timeout 800 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r2c22_pytest_multigpu.log 2>&1; echo "exit $?" >> gpurun_out/r2c22_pytest_multigpu.log; tail -n 5 gpurun_out/r2c22_pytest_multigpu.log | cut -c1-400

#!/bin/bash
# Round 2, GPU call 14 (1 GPU): launch-ordering check of the overlapped backward pass (eager / graph x PDL / weight-gradient stream) and
# the CTA-pair GEMM after the leader-only barrier arming.
mkdir -p gpurun_out
O=gpurun_out/r2c14
timeout 300 python benchmarks/overlap_check.py > ${O}_overlap_check.log 2>&1; echo "exit $?" >> ${O}_overlap_check.log; tail -n 12 ${O}_overlap_check.log | cut -c1-300
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q > ${O}_pytest_gemm.log 2>&1; tail -n 2 ${O}_pytest_gemm.log
timeout 300 python benchmarks/gemm_pair_bench.py > ${O}_gemm_pair_bench.log 2>&1; tail -n 14 ${O}_gemm_pair_bench.log | cut -c1-250

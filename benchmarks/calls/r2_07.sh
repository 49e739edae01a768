#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c7
timeout 120 python benchmarks/tf32_layer_diag.py > ${O}_tf32_diag.log 2>&1; tail -n 20 ${O}_tf32_diag.log | cut -c1-200
timeout 300 python -m pytest tests/test_layers_gpu.py -q -s -k "resnet_step_tf32" > ${O}_pytest_tf32_step.log 2>&1; grep -a "worst\|passed\|failed" ${O}_pytest_tf32_step.log | head -5 | cut -c1-900
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x -k "pair" > ${O}_pytest_pair.log 2>&1; echo "exit $?" >> ${O}_pytest_pair.log; tail -n 12 ${O}_pytest_pair.log | cut -c1-250
timeout 600 python benchmarks/gemm_pair_bench.py > ${O}_gemm_pair_bench.log 2>&1; tail -n 16 ${O}_gemm_pair_bench.log | cut -c1-420

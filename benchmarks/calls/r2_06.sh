#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c6
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -k "tf32" > ${O}_pytest_gemm_tf32.log 2>&1; echo "exit $?" >> ${O}_pytest_gemm_tf32.log; tail -n 6 ${O}_pytest_gemm_tf32.log | cut -c1-250
timeout 600 python -m pytest tests/test_layers_gpu.py -q -k "tf32" > ${O}_pytest_layers_tf32.log 2>&1; echo "exit $?" >> ${O}_pytest_layers_tf32.log; grep -a "^E   \|passed\|failed" ${O}_pytest_layers_tf32.log | head -20 | cut -c1-250
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest_all.log 2>&1; echo "exit $?" >> ${O}_pytest_all.log; tail -n 6 ${O}_pytest_all.log | cut -c1-250
timeout 900 python bench.py --dtype tf32 --steps 10 --warmup 3 > ${O}_bench_tf32.log 2>&1; grep '^{"metric' ${O}_bench_tf32.log | cut -c1-500

#!/bin/bash
# Round 2, GPU call 5 (1 GPU): TF32 path (GEMM, convolution, templated fp32 layer kernels), regression of the bf16 path, tf32 bench.
mkdir -p gpurun_out
O=gpurun_out/r2c5
export AGB_FLAG_TIMEOUT_S=30
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -k "tf32" > ${O}_pytest_gemm_tf32.log 2>&1; echo "exit $?" >> ${O}_pytest_gemm_tf32.log; tail -n 25 ${O}_pytest_gemm_tf32.log | cut -c1-250
timeout 600 python -m pytest tests/test_layers_gpu.py -q -k "tf32" > ${O}_pytest_layers_tf32.log 2>&1; echo "exit $?" >> ${O}_pytest_layers_tf32.log; tail -n 30 ${O}_pytest_layers_tf32.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest_all.log 2>&1; echo "exit $?" >> ${O}_pytest_all.log; tail -n 12 ${O}_pytest_all.log | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 --no-baseline --skip-e2e > ${O}_bench_bf16.log 2>&1; grep -o '"ms_per_step": [0-9.]*' ${O}_bench_bf16.log | head -1
timeout 900 python bench.py --dtype tf32 --steps 10 --warmup 3 > ${O}_bench_tf32.log 2>&1; grep '^{"metric' ${O}_bench_tf32.log | cut -c1-700; tail -n 3 ${O}_bench_tf32.log | cut -c1-300
timeout 300 python benchmarks/gar_bench.py --gar-iters 10 --gar-rules krum,average --gar-out gpurun_out/r2c5_gar > ${O}_gar_bench.log 2>&1; grep -a "^krum\|^average" ${O}_gar_bench.log | cut -c1-200

#!/bin/bash
# Round 2, GPU call 4 (1 GPU): aggregation kernels v2 (n <= 32, bucketed phase A, device-resident step state), whole-step CUDA graph.
mkdir -p gpurun_out
O=gpurun_out/r2c4
export AGB_FLAG_TIMEOUT_S=30
timeout 600 python -m pytest tests/test_gar_gpu.py -x -q > ${O}_pytest_gar.log 2>&1; echo "exit $?" >> ${O}_pytest_gar.log; tail -n 30 ${O}_pytest_gar.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest_all.log 2>&1; echo "exit $?" >> ${O}_pytest_all.log; tail -n 15 ${O}_pytest_all.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-baseline > ${O}_bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*' ${O}_bench.log | head -2; grep -a "graph\]\|fused\]" ${O}_bench.log | head -5; tail -n 3 ${O}_bench.log | cut -c1-400
AGB_BATCH_WORKERS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-baseline --skip-e2e > ${O}_bench_b32.log 2>&1; grep -o '"ms_per_step": [0-9.]*' ${O}_bench_b32.log | head -1
AGB_OVERLAP=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-baseline --skip-e2e > ${O}_bench_nooverlap.log 2>&1; grep -o '"ms_per_step": [0-9.]*' ${O}_bench_nooverlap.log | head -1
timeout 600 python benchmarks/gar_bench.py --gar-iters 10 --gar-out gpurun_out/r2c4_gar > ${O}_gar_bench.log 2>&1; tail -n 8 ${O}_gar_bench.log | cut -c1-400

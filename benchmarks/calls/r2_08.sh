#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c8
timeout 300 python -m pytest tests/test_layers_gpu.py -q -s -k "resnet_step_tf32" > ${O}_pytest_tf32_step.log 2>&1; grep -a "cos(\|passed\|failed" ${O}_pytest_tf32_step.log | head -5 | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest_all.log 2>&1; echo "exit $?" >> ${O}_pytest_all.log; tail -n 6 ${O}_pytest_all.log | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 --no-baseline --skip-e2e > ${O}_bench_bf16.log 2>&1; grep -o '"ms_per_step": [0-9.]*' ${O}_bench_bf16.log | head -1
AGB_BATCH_WORKERS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-baseline --skip-e2e > ${O}_bench_b32.log 2>&1; grep -o '"ms_per_step": [0-9.]*' ${O}_bench_b32.log | head -1
timeout 600 python bench.py --dtype tf32 --steps 10 --warmup 3 --no-baseline --skip-e2e > ${O}_bench_tf32.log 2>&1; grep -o '"ms_per_step": [0-9.]*' ${O}_bench_tf32.log | head -1

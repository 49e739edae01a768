#!/bin/bash
# Round 2, GPU call 3 (1 GPU): stride-2 implicit convolution, rectangular filters, preprocessing kernel, factory nets, headline bench.
mkdir -p gpurun_out
O=gpurun_out/r2c3
timeout 600 python -m pytest tests/test_layers_gpu.py -q -k "test_conv or rectangular" > ${O}_pytest_conv.log 2>&1; echo "exit $?" >> ${O}_pytest_conv.log; tail -n 25 ${O}_pytest_conv.log | cut -c1-300
timeout 300 python -m pytest tests/test_input_pipeline.py -m gpu -q > ${O}_pytest_input.log 2>&1; echo "exit $?" >> ${O}_pytest_input.log; tail -n 8 ${O}_pytest_input.log | cut -c1-300
timeout 900 python -m pytest tests/test_layers_gpu.py -q -s -k "every_factory_net" > ${O}_pytest_nets.log 2>&1; echo "exit $?" >> ${O}_pytest_nets.log
grep -a "aten fallbacks\|passed\|failed" ${O}_pytest_nets.log | grep -v print | tail -n 40
timeout 600 python -m pytest tests -m gpu -x -q > ${O}_pytest_all.log 2>&1; echo "exit $?" >> ${O}_pytest_all.log; tail -n 4 ${O}_pytest_all.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-baseline --skip-e2e > ${O}_bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*' ${O}_bench.log | head -1
AGB_BATCH_WORKERS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-baseline --skip-e2e > ${O}_bench_b32.log 2>&1; grep -o '"ms_per_step": [0-9.]*' ${O}_bench_b32.log | head -1

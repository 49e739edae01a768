#!/bin/bash
# Round 2, GPU call 24 (1 GPU): the input-pipeline GPU tests after the device-index / producer-failure fixes.
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_input_pipeline.py -x -q -m gpu > gpurun_out/r2c24_pytest_input.log 2>&1; echo "exit $?" >> gpurun_out/r2c24_pytest_input.log; tail -n 6 gpurun_out/r2c24_pytest_input.log | cut -c1-300

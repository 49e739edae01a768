#!/bin/bash
# Round 2, GPU call 23 (1 GPU): grouped input streams (one pinned slab / H2D copy / static-input copy per step for all local workers).
mkdir -p gpurun_out
O=gpurun_out/r2c23
timeout 300 python -m pytest tests/test_input_pipeline.py -x -q -m gpu > ${O}_pytest_input.log 2>&1; echo "exit $?" >> ${O}_pytest_input.log; tail -n 3 ${O}_pytest_input.log | cut -c1-300
for mode in 1 0; do
  AGB_GROUP_STREAMS=$mode timeout 400 python bench.py --steps 20 --warmup 5 --no-baseline > ${O}_bench_group$mode.log 2>&1
  echo "AGB_GROUP_STREAMS=$mode: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_bench_group$mode.log | tr '\n' ' ') $(grep -ao '"last_loss": [a-zA-Z0-9.+-]*' ${O}_bench_group$mode.log)"
done

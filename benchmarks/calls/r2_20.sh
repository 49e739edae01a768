#!/bin/bash
# Round 2, GPU call 20 (1 GPU): final validation — max-pool backward variants, the whole GPU test-suite, smoke(), the default bench line.
mkdir -p gpurun_out
O=gpurun_out/r2c20
echo "specialised: $(timeout 120 python benchmarks/maxpool_bench.py 32 2>&1 | tail -n 1) | $(timeout 120 python benchmarks/maxpool_bench.py 256 2>&1 | tail -n 1)"
echo "generic loop: $(AGB_MAXPOOL_LOOP=1 timeout 120 python benchmarks/maxpool_bench.py 32 2>&1 | tail -n 1) | $(AGB_MAXPOOL_LOOP=1 timeout 120 python benchmarks/maxpool_bench.py 256 2>&1 | tail -n 1)"
timeout 1200 python -m pytest tests -x -q -m gpu > ${O}_pytest_gpu.log 2>&1; echo "exit $?" >> ${O}_pytest_gpu.log; tail -n 4 ${O}_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${O}_smoke.log 2>&1; tail -n 2 ${O}_smoke.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > ${O}_bench_1gpu.log 2>&1; grep -a '^{"metric' ${O}_bench_1gpu.log | cut -c1-1400
AGB_MAXPOOL_LOOP=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-baseline --skip-e2e > ${O}_bench_loop.log 2>&1; echo "generic max-pool loop: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_bench_loop.log | head -1)"

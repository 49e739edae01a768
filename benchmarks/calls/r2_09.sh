#!/bin/bash
# Round 2, GPU call 9 (2 GPUs): multi-rank tests (whole-step graph + overlapped distance pass, replicas identical), aggregation bench, headline bench at N = 2.
mkdir -p gpurun_out
O=gpurun_out/r2c9
export AGB_FLAG_TIMEOUT_S=60
timeout 900 python -m pytest tests/test_multigpu.py -x -q > ${O}_pytest_multigpu.log 2>&1; echo "exit $?" >> ${O}_pytest_multigpu.log; tail -n 25 ${O}_pytest_multigpu.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 benchmarks/gar_bench.py --gar-iters 10 --gar-out gpurun_out/r2c9_gar > ${O}_gar_bench.log 2>&1
grep -a "^krum\|^bulyan\|^average \|^median\|measured peer" ${O}_gar_bench.log | cut -c1-700
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > ${O}_bench_2gpu.log 2>&1
grep -a '^{"metric' ${O}_bench_2gpu.log | cut -c1-900; grep -a "graph\]\|fused\]" ${O}_bench_2gpu.log | head -4 | cut -c1-250; tail -n 5 ${O}_bench_2gpu.log | cut -c1-300
AGB_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 --no-baseline --skip-e2e > ${O}_bench_2gpu_nooverlap.log 2>&1
grep -ao '"ms_per_step": [0-9.]*' ${O}_bench_2gpu_nooverlap.log | head -1

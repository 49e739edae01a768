#!/bin/bash
# Round 2, GPU call 25 (2 GPUs): the bench line at 2 ranks x 4 workers on the final build (grouped input streams on every rank).
mkdir -p gpurun_out
export AGB_FLAG_TIMEOUT_S=30
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 6 --warmup 3 --no-baseline > gpurun_out/r2c25_bench_2gpu.log 2>&1
echo "exit $?"; grep -a '^{"metric' gpurun_out/r2c25_bench_2gpu.log | cut -c1-1300; tail -n 3 gpurun_out/r2c25_bench_2gpu.log | cut -c1-300

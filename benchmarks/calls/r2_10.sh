#!/bin/bash
# 2 GPUs: overlap variants (phase A launch shape) vs no overlap, headline config
mkdir -p gpurun_out
O=gpurun_out/r2c10
export AGB_FLAG_TIMEOUT_S=60
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29620 + RANDOM % 50)) bench.py --gpus 2 --steps 10 --warmup 3 --no-baseline --skip-e2e > ${O}_${tag}.log 2>&1; echo "$tag: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_${tag}.log | head -1)"; }
run nooverlap AGB_OVERLAP=0
run c148t64 AGB_PHASE_A_CTAS=148 AGB_PHASE_A_THREADS=64
run c74t64 AGB_PHASE_A_CTAS=74 AGB_PHASE_A_THREADS=64
run c32t64 AGB_PHASE_A_CTAS=32 AGB_PHASE_A_THREADS=64
run c148t128 AGB_PHASE_A_CTAS=148 AGB_PHASE_A_THREADS=128
run nooverlap_nographagg AGB_OVERLAP=0 AGB_GRAPH_AGGREGATION=0

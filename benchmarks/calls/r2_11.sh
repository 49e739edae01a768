#!/bin/bash
# Round 2, GPU call 11 (8 GPUs): headline at N = 8 with the in-run baseline, overlap / launch-overlap A/B, aggregation bench with the
# measured peer bandwidth, BASELINE.json configurations 2-5, resilience trajectories, multi-rank tests on all 8 GPUs.
mkdir -p gpurun_out
O=gpurun_out/r2c11
export AGB_FLAG_TIMEOUT_S=60
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 > ${O}_bench_8gpu.log 2>&1
grep -a '^{"metric' ${O}_bench_8gpu.log | cut -c1-1200; grep -a "graph\]\|fused\]" ${O}_bench_8gpu.log | head -3 | cut -c1-250
run() { tag=$1; shift; env "$@" timeout 400 $TR --master-port $((29710 + RANDOM % 80)) bench.py --gpus 8 --steps 20 --warmup 5 --no-baseline --skip-e2e > ${O}_${tag}.log 2>&1; echo "$tag: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_${tag}.log | head -1)"; }
run default_repeat AGB_X=1
run nooverlap AGB_OVERLAP=0
run nolaunchoverlap AGB_LAUNCH_OVERLAP=0
run neither AGB_OVERLAP=0 AGB_LAUNCH_OVERLAP=0
run phaseA_148x64 AGB_PHASE_A_THREADS=64
run phaseA_74x128 AGB_PHASE_A_CTAS=74
timeout 600 $TR --master-port 29702 benchmarks/gar_bench.py --gar-iters 10 --gar-out gpurun_out/r2c11_gar > ${O}_gar_bench.log 2>&1
grep -a "^krum\|^bulyan\|^average\|^median\|measured peer" ${O}_gar_bench.log | cut -c1-650
bash benchmarks/baseline_configs.sh 8 gpurun_out/r2c11_baseline_configs.jsonl
# resilience: ResNet-50, 8 workers of which 2 flip their gradients (x -10), Krum vs plain averaging, same data and seed
for rule in krum average; do
  timeout 400 $TR --master-port $((29790 + RANDOM % 9)) runner.py --server '{"ps": ["127.0.0.1:7000"], "workers": ["127.0.0.1:7001","127.0.0.1:7002","127.0.0.1:7003","127.0.0.1:7004","127.0.0.1:7005","127.0.0.1:7006","127.0.0.1:7007","127.0.0.1:7008"], "eval": ["127.0.0.1:7000"]}' --no-wait \
    --experiment slim-resnet_v1_50-imagenet --experiment-args batch-size:32 synthetic-samples:4096 image-size:64 --aggregator $rule --nb-workers 8 --nb-decl-byz-workers 2 --nb-real-byz-workers 2 \
    --attack flip --attack-args factor:-10 --learning-rate-args initial-rate:0.02 --max-step 150 --use-gpu --evaluation-delta 50 --evaluation-period -1 --checkpoint-dir /tmp/res_$rule --checkpoint-delta -1 --checkpoint-period -1 --summary-dir - > ${O}_resilience_${rule}.log 2>&1
  echo "resilience $rule: $(grep -a 'total loss' ${O}_resilience_${rule}.log | sed -n '1p;50p;100p;150p' | sed 's/.*total loss = //' | tr '\n' ' ') | $(grep -a 'top1-X-acc' ${O}_resilience_${rule}.log | tail -n 1 | cut -c1-120)"
done
timeout 900 python -m pytest tests/test_multigpu.py -x -q > ${O}_pytest_multigpu.log 2>&1; echo "exit $?" >> ${O}_pytest_multigpu.log; tail -n 4 ${O}_pytest_multigpu.log | cut -c1-300

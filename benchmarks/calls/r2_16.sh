#!/bin/bash
# Round 2, GPU call 16 (8 GPUs): headline after the launch-ordering fix (both arms), the same without launch overlap, and the loss
# of the median / Krum + attack configurations (they diverged in call 11, where the batch norm kernel could start before its producer).
mkdir -p gpurun_out
O=gpurun_out/r2c16
export AGB_FLAG_TIMEOUT_S=60
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 5 > ${O}_bench_8gpu.log 2>&1
grep -a '^{"metric' ${O}_bench_8gpu.log | cut -c1-1500
AGB_LAUNCH_OVERLAP=0 timeout 300 $TR --master-port 29712 bench.py --gpus 8 --steps 20 --warmup 5 --no-baseline > ${O}_serial_launches.log 2>&1
echo "serial launches: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_serial_launches.log | head -1) $(grep -ao '"last_loss": [a-zA-Z0-9.+-]*' ${O}_serial_launches.log)"
timeout 300 $TR --master-port 29713 bench.py --gpus 8 --steps 20 --warmup 5 --no-baseline --aggregator median > ${O}_median.log 2>&1
echo "median: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_median.log | head -1) $(grep -ao '"last_loss": [a-zA-Z0-9.+-]*' ${O}_median.log)"
timeout 300 $TR --master-port 29714 bench.py --gpus 8 --steps 20 --warmup 5 --no-baseline --nb-real-byz-workers 2 --attack flip --attack-args factor:-10 > ${O}_krum_flip.log 2>&1
echo "krum + flip: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_krum_flip.log | head -1) $(grep -ao '"last_loss": [a-zA-Z0-9.+-]*' ${O}_krum_flip.log)"

#!/bin/bash
# Round 2, GPU call 1 (1 GPU): regression tests after the advisor fixes, the new in-run baseline arm, batch-32 launch-overlap A/B
# experiments (what each GPU runs at N = 8), validation of the depthwise / SAME-avg-pool / ReLU6 kernels.
mkdir -p gpurun_out
O=gpurun_out/r2c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${O}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -n 3 ${O}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > ${O}_bench.log 2>&1; echo "bench exit $?" >> ${O}_bench.log; grep '^{"metric' ${O}_bench.log | cut -c1-600
timeout 300 python bench.py --impl baseline --dtype tf32 --steps 10 --warmup 3 --skip-e2e > ${O}_bench_base_tf32.log 2>&1; grep -o '"ms_per_step": [0-9.]*' ${O}_bench_base_tf32.log | head -1
# what one GPU does at N = 8: eight sequential batch-32 passes (AGB_BATCH_WORKERS=0); stream / PDL variants
for variant in "" "AGB_WGRAD_STREAM=1" "AGB_PDL=1" "AGB_WGRAD_STREAM=1 AGB_PDL=1"; do
  tag=$(echo "b32_${variant}" | tr ' =' '__')
  env AGB_BATCH_WORKERS=0 $variant timeout 300 python bench.py --steps 10 --warmup 4 --skip-e2e --no-baseline > ${O}_${tag}.log 2>&1
  echo "unbatched [$variant]: $(grep -o '"ms_per_step": [0-9.]*' ${O}_${tag}.log | head -1)"
done
AGB_NATIVE_PREVIEW=1 timeout 600 python -m pytest tests/test_layers_gpu.py -x -q -k "depthwise_native or avgpool2d_and_relu6 or searched_and_inception" > ${O}_preview_pytest.log 2>&1
echo "preview pytest exit $?" >> ${O}_preview_pytest.log; tail -n 3 ${O}_preview_pytest.log

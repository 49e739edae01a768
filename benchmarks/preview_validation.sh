#!/bin/bash
# First GPU call of the next round: validate the kernels written after round 1's GPU budget was spent (depthwise convolution, SAME
# average pooling, ReLU6 — AGB_NATIVE_PREVIEW=1) and measure what they are worth on the networks that use them.
# Usage (1 GPU): bash benchmarks/preview_validation.sh   -> gpurun_out/preview_*.log
mkdir -p gpurun_out
export AGB_NATIVE_PREVIEW=1
timeout 600 python -m pytest tests/test_layers_gpu.py -x -q -k "depthwise_native or avgpool2d_and_relu6 or searched_and_inception" > gpurun_out/preview_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/preview_pytest.log
tail -3 gpurun_out/preview_pytest.log
for model in mobilenet_v2 nasnet_mobile inception_v3; do
  for preview in 0 1; do
    AGB_NATIVE_PREVIEW=$preview timeout 300 python bench.py --model $model --aggregator krum --nb-workers 8 --nb-decl-byz-workers 2 --batch-size 16 --steps 10 --warmup 4 --skip-e2e \
      > gpurun_out/preview_bench_${model}_${preview}.log 2>&1
    echo "$model preview=$preview: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/preview_bench_${model}_${preview}.log)"
  done
done

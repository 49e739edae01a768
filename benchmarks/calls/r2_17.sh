#!/bin/bash
# Round 2, GPU call 17 (1 GPU): staged stem im2col + channels_last pooling outputs: tests, launch list, step times; resilience
# trajectories (Krum vs averaging under 2 flipping workers) on the final kernels.
mkdir -p gpurun_out
O=gpurun_out/r2c17
timeout 600 python -m pytest tests/test_layers_gpu.py -x -q -k "stem_im2col or test_conv or pools or launch_overlap" > ${O}_pytest_stem.log 2>&1; echo "exit $?" >> ${O}_pytest_stem.log; tail -n 4 ${O}_pytest_stem.log | cut -c1-300
AGB_NATIVE_STRICT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${O}_launches_b32.csv python benchmarks/profile_step.py --batch-size 32 > ${O}_profile_step.log 2>&1
tail -n 2 ${O}_profile_step.log; python benchmarks/summarize_launches.py ${O}_launches_b32.csv 45 > ${O}_launches_b32.txt 2>&1; head -n 4 ${O}_launches_b32.txt; grep -a "im2col\|direct_copy\|maxpool" ${O}_launches_b32.txt
run() {
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --no-baseline > ${O}_$name.log 2>&1
  echo "$name: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_$name.log | head -1) $(grep -ao '"last_loss": [a-zA-Z0-9.+-]*' ${O}_$name.log)"
}
run default AGB_X=0
run unbatched_overlap AGB_BATCH_WORKERS=0 AGB_PDL=1 AGB_WGRAD_STREAM=1
for rule in krum average; do
  timeout 300 python runner.py --server '{"local": ["127.0.0.1:7000"]}' --ps-job-name local --wk-job-name local --ev-job-name local --no-wait \
    --experiment slim-resnet_v1_50-imagenet --experiment-args batch-size:32 synthetic-samples:4096 image-size:64 --aggregator $rule --nb-workers 8 --nb-decl-byz-workers 2 --nb-real-byz-workers 2 \
    --attack flip --attack-args factor:-10 --learning-rate-args initial-rate:0.02 --max-step 150 --use-gpu --reuse-gpu --evaluation-delta 50 --evaluation-period -1 --checkpoint-dir /tmp/res_$rule --checkpoint-delta -1 --checkpoint-period -1 --summary-dir - > ${O}_resilience_${rule}.log 2>&1
  echo "resilience $rule: $(grep -a 'total loss' ${O}_resilience_${rule}.log | sed -n '1p;50p;100p;150p' | sed 's/.*total loss = //' | tr '\n' ' ')"
done

#!/bin/bash
# Round 2, GPU call 15 (1 GPU): after the batch-norm changes (PDL wait, batched sum loads): layer tests, launch-ordering check,
# batch-norm micro-benchmark, launch list of a batch-32 step, step time of 8 sequential batch-32 passes and of the default configuration.
mkdir -p gpurun_out
O=gpurun_out/r2c15
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_gemm_gpu.py -x -q > ${O}_pytest_layers.log 2>&1; echo "exit $?" >> ${O}_pytest_layers.log; tail -n 5 ${O}_pytest_layers.log | cut -c1-300
timeout 300 python benchmarks/overlap_check.py > ${O}_overlap_check.log 2>&1; echo "exit $?" >> ${O}_overlap_check.log; tail -n 12 ${O}_overlap_check.log | cut -c1-200
timeout 300 python benchmarks/bn_bench.py --batch 32 --out ${O}_bn_bench_b32.json > ${O}_bn_bench.log 2>&1; tail -n 16 ${O}_bn_bench.log | cut -c1-200
AGB_NATIVE_STRICT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${O}_launches_b32.csv python benchmarks/profile_step.py --batch-size 32 > ${O}_profile_step.log 2>&1
tail -n 2 ${O}_profile_step.log; python benchmarks/summarize_launches.py ${O}_launches_b32.csv 45 > ${O}_launches_b32.txt 2>&1; head -n 8 ${O}_launches_b32.txt
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:bn_fused -s 4 -c 2 -o ${O}_ncu_bn python benchmarks/ncu_targets.py bn > ${O}_ncu_bn.log 2>&1; tail -n 1 ${O}_ncu_bn.log
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:gemm_tcgen05_pair -s 2 -c 1 -o ${O}_ncu_gemm_pair python benchmarks/ncu_targets.py gemm_pair > ${O}_ncu_gemm_pair.log 2>&1; tail -n 1 ${O}_ncu_gemm_pair.log
for rep in ${O}_ncu_*.ncu-rep; do python benchmarks/ncu_summary.py $rep > ${rep%.ncu-rep}_summary.txt 2>&1; python benchmarks/ncu_source_top.py $rep 25 > ${rep%.ncu-rep}_hotspots.txt 2>&1; done
run() {
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --no-baseline > ${O}_$name.log 2>&1
  echo "$name: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_$name.log | head -1) $(grep -ao '"last_loss": [a-zA-Z0-9.+-]*' ${O}_$name.log)"
}
run unbatched_serial AGB_BATCH_WORKERS=0 AGB_PDL=0 AGB_WGRAD_STREAM=0
run unbatched_overlap AGB_BATCH_WORKERS=0 AGB_PDL=1 AGB_WGRAD_STREAM=1
run default AGB_X=0
run default_overlap AGB_PDL=1 AGB_WGRAD_STREAM=1

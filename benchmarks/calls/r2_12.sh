#!/bin/bash
# Round 2, GPU call 12 (1 GPU): ncu --set full captures of the shipped kernels (+ source pages) and the launch list of a batch-32 worker step.
mkdir -p gpurun_out
O=gpurun_out/r2c12
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:gemm_tcgen05_persistent -s 2 -c 1 -o ${O}_ncu_gemm_wide python benchmarks/ncu_targets.py gemm_wide > ${O}_ncu_gemm_wide.log 2>&1; tail -n 1 ${O}_ncu_gemm_wide.log
timeout 300 $NCU -k regex:gemm_tcgen05_pair -s 2 -c 1 -o ${O}_ncu_gemm_pair python benchmarks/ncu_targets.py gemm_pair > ${O}_ncu_gemm_pair.log 2>&1; tail -n 1 ${O}_ncu_gemm_pair.log
timeout 300 $NCU -k regex:gemm_tcgen05_persistent -s 2 -c 1 -o ${O}_ncu_gemm_tf32 python benchmarks/ncu_targets.py gemm_tf32 > ${O}_ncu_gemm_tf32.log 2>&1; tail -n 1 ${O}_ncu_gemm_tf32.log
timeout 300 $NCU -k regex:conv_tcgen05 -s 6 -c 3 -o ${O}_ncu_conv python benchmarks/ncu_targets.py conv > ${O}_ncu_conv.log 2>&1; tail -n 1 ${O}_ncu_conv.log
timeout 300 $NCU -k regex:gar_fused -s 2 -c 1 -o ${O}_ncu_gar python benchmarks/ncu_targets.py gar > ${O}_ncu_gar.log 2>&1; tail -n 1 ${O}_ncu_gar.log
timeout 300 $NCU -k regex:bn_fused -s 4 -c 2 -o ${O}_ncu_bn python benchmarks/ncu_targets.py bn > ${O}_ncu_bn.log 2>&1; tail -n 1 ${O}_ncu_bn.log
AGB_NATIVE_STRICT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${O}_launches_b32.csv python benchmarks/profile_step.py --batch-size 32 > ${O}_profile_step.log 2>&1
tail -n 2 ${O}_profile_step.log; python benchmarks/summarize_launches.py ${O}_launches_b32.csv 45 > ${O}_launches_b32.txt 2>&1; head -n 30 ${O}_launches_b32.txt
for rep in gpurun_out/r2c12_ncu_*.ncu-rep; do
  base=${rep%.ncu-rep}
  python benchmarks/ncu_summary.py $rep > ${base}_summary.txt 2>&1
  python benchmarks/ncu_source_top.py $rep 30 > ${base}_hotspots.txt 2>&1
done
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
du -sm gpurun_out | cut -f1
head -n 20 gpurun_out/r2c12_ncu_gemm_pair_hotspots.txt | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c12_bench_1gpu.log 2>&1; grep -a '^{"metric' gpurun_out/r2c12_bench_1gpu.log | cut -c1-900

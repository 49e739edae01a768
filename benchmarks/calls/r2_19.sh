#!/bin/bash
# Round 2, GPU call 19 (2 GPUs): compute-sanitizer memcheck + synccheck over the fused aggregation kernels running across two ranks
# (peer loads, multimem stores, system-scope flag barriers), every rule, small d.
mkdir -p gpurun_out
O=gpurun_out/r2c19
export AGB_FLAG_TIMEOUT_S=100
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --no-python"
for tool in memcheck synccheck; do
  timeout 400 $TR --master-port $((29730 + RANDOM % 20)) compute-sanitizer --tool $tool --log-file ${O}_${tool}_pid%p.log python benchmarks/gar_bench.py --gar-dim 200003 --gar-iters 1 --gar-out ${O}_gar_$tool > ${O}_${tool}_run.log 2>&1
  echo "$tool: exit $? | $(grep -ah 'ERROR SUMMARY' ${O}_${tool}_pid*.log | tr '\n' ' ') | $(tail -n 2 ${O}_${tool}_run.log | cut -c1-200 | tr '\n' ' ')"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c19_gar_*/gar_bench_2.json")):
    d = json.load(open(f))
    print(f, {r: (e["replicas_identical"], e["max_abs_diff_vs_baseline"]) for r, e in d["results"].items()})
PY

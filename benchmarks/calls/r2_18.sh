#!/bin/bash
# Round 2, GPU call 18 (1 GPU): batched residual loads in the single-launch batch norm, max-pool kernels with all taps in flight:
# layer tests, batch-norm micro-benchmark, launch list, step times.
mkdir -p gpurun_out
O=gpurun_out/r2c18
timeout 600 python -m pytest tests/test_layers_gpu.py -x -q > ${O}_pytest_layers.log 2>&1; echo "exit $?" >> ${O}_pytest_layers.log; tail -n 4 ${O}_pytest_layers.log | cut -c1-300
AGB_NATIVE_STRICT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${O}_launches_b32.csv python benchmarks/profile_step.py --batch-size 32 > ${O}_profile_step.log 2>&1
tail -n 1 ${O}_profile_step.log; python benchmarks/summarize_launches.py ${O}_launches_b32.csv 45 > ${O}_launches_b32.txt 2>&1; head -n 3 ${O}_launches_b32.txt; grep -a "maxpool\|add_relu" ${O}_launches_b32.txt
python - <<'PY'
import csv,collections
lines=[l for l in open("gpurun_out/r2c18_launches_b32.csv") if l.startswith('"')]
r=list(csv.reader(lines)); hdr=r[0]; data=r[1:]
idx={h:i for i,h in enumerate(hdr)}
names=collections.OrderedDict()
for x in data:
    n=x[idx["Kernel Name"]]
    if "bn_fused" in n:
        names.setdefault((n[22:40],x[idx["Grid Size"]]),[]).append(float(x[idx["Metric Value"]])/1000)
for k,v in names.items():
    print(k, len(v), "avg %.1f us"%(sum(v)/len(v)), " ".join("%.1f"%t for t in v[:14]))
PY
run() {
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --no-baseline > ${O}_$name.log 2>&1
  echo "$name: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_$name.log | head -1) $(grep -ao '"last_loss": [a-zA-Z0-9.+-]*' ${O}_$name.log)"
}
run default AGB_X=0
run unbatched_overlap AGB_BATCH_WORKERS=0 AGB_PDL=1 AGB_WGRAD_STREAM=1

#!/bin/bash
# Round 2, GPU call 13 (1 GPU): launch overlap (PDL + weight-gradient stream) after the batch-norm wait fix: regression test, loss and
# step time of 8 sequential batch-32 passes (what a single-worker rank runs) with each half of the switch.
mkdir -p gpurun_out
O=gpurun_out/r2c13
timeout 600 python -m pytest tests/test_layers_gpu.py -x -q -k "launch_overlap or deterministic" > ${O}_pytest_overlap.log 2>&1; echo "exit $?" >> ${O}_pytest_overlap.log; tail -n 6 ${O}_pytest_overlap.log | cut -c1-400
run() {
  name=$1; shift
  env "$@" AGB_BATCH_WORKERS=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-baseline > ${O}_$name.log 2>&1
  echo "$name: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_$name.log | head -1) $(grep -ao '"last_loss": [a-zA-Z0-9.+-]*' ${O}_$name.log)"
}
run serial AGB_PDL=0 AGB_WGRAD_STREAM=0
run pdl_wgrad AGB_PDL=1 AGB_WGRAD_STREAM=1
run pdl AGB_PDL=1 AGB_WGRAD_STREAM=0
run wgrad AGB_PDL=0 AGB_WGRAD_STREAM=1
timeout 400 python bench.py --steps 20 --warmup 5 --no-baseline --aggregator average --nb-workers 1 --nb-decl-byz-workers 0 > ${O}_single_worker.log 2>&1
echo "single worker, default: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_single_worker.log | head -1) $(grep -ao '"last_loss": [a-zA-Z0-9.+-]*' ${O}_single_worker.log) $(grep -ac 'launch overlap' ${O}_single_worker.log)"
AGB_LAUNCH_OVERLAP=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-baseline --aggregator average --nb-workers 1 --nb-decl-byz-workers 0 > ${O}_single_worker_serial.log 2>&1
echo "single worker, serial: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_single_worker_serial.log | head -1) $(grep -ao '"last_loss": [a-zA-Z0-9.+-]*' ${O}_single_worker_serial.log)"

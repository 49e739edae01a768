#!/bin/bash
# Round 2, GPU call 2 (1 GPU): input pipeline kernels, all 33 factory nets on the native provider, cnnet under graph replay,
# compute-sanitizer memcheck / synccheck over the aggregation, GEMM and convolution tests.
mkdir -p gpurun_out
O=gpurun_out/r2c2
timeout 300 python -m pytest tests/test_input_pipeline.py -m gpu -x -q > ${O}_pytest_input.log 2>&1; echo "exit $?" >> ${O}_pytest_input.log; tail -n 5 ${O}_pytest_input.log
timeout 900 python -m pytest tests/test_layers_gpu.py -q -s -k "every_factory_net" > ${O}_pytest_nets.log 2>&1; echo "exit $?" >> ${O}_pytest_nets.log
grep -a "aten fallbacks\|passed\|failed" ${O}_pytest_nets.log | tail -n 45
timeout 300 python bench.py --experiment cnnet --steps 20 --warmup 5 --no-baseline > ${O}_bench_cnnet.log 2>&1; grep -o '"ms_per_step": [0-9.]*' ${O}_bench_cnnet.log | head -2; grep -a "graph" ${O}_bench_cnnet.log | head -3
export AGB_NO_GRAPH=1
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --log-file ${O}_sanitize_${tool}.log \
    python -m pytest tests/test_gar_gpu.py tests/test_gemm_gpu.py -q -x -k "not 8192 and not 4096 and not 100352 and not 25088" > ${O}_sanitize_${tool}_pytest.log 2>&1
  echo "$tool exit $?: $(tail -n 1 ${O}_sanitize_${tool}_pytest.log)"; grep -a "ERROR SUMMARY" ${O}_sanitize_${tool}.log | tail -n 1
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file ${O}_sanitize_memcheck_layers.log \
  python -m pytest tests/test_layers_gpu.py -q -x -k "conv and not every_factory" > ${O}_sanitize_memcheck_layers_pytest.log 2>&1
echo "memcheck layers exit $?: $(tail -n 1 ${O}_sanitize_memcheck_layers_pytest.log)"; grep -a "ERROR SUMMARY" ${O}_sanitize_memcheck_layers.log | tail -n 1

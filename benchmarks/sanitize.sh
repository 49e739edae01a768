#!/bin/bash
# Race / memory / sync checking of every native kernel with compute-sanitizer (the reference has none).
# Run on a GPU box:  bash benchmarks/sanitize.sh [memcheck|racecheck|synccheck|initcheck]   -> gpurun_out/sanitize_<tool>.log
TOOL=${1:-memcheck}
mkdir -p gpurun_out
export AGB_NO_GRAPH=1
compute-sanitizer --tool ${TOOL} --error-exitcode 3 --log-file gpurun_out/sanitize_${TOOL}.log \
  python -m pytest tests/test_gemm_gpu.py tests/test_gar_gpu.py tests/test_layers_gpu.py -q -x \
  -k "not model_gradients and not 8192 and not 100352 and not 25088 and not 4096" > gpurun_out/sanitize_${TOOL}_pytest.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/sanitize_${TOOL}.log
tail -5 gpurun_out/sanitize_${TOOL}.log

#!/bin/bash
# Round 2, GPU call 21 (1 GPU): SHA-256 tree digest + authenticator on device rows, 2 x 2-block max-pool backward: tests, timings.
mkdir -p gpurun_out
O=gpurun_out/r2c21
timeout 600 python -m pytest tests/test_gar_gpu.py tests/test_layers_gpu.py -x -q -k "sha256 or authenticator or pools or drop_chunks" > ${O}_pytest.log 2>&1; echo "exit $?" >> ${O}_pytest.log; tail -n 4 ${O}_pytest.log | cut -c1-300
echo "max-pool: $(timeout 120 python benchmarks/maxpool_bench.py 32 2>&1 | tail -n 1) | $(timeout 120 python benchmarks/maxpool_bench.py 256 2>&1 | tail -n 1)"
timeout 120 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from aggregathor_b200.ops import gar as g
x = torch.randn(25558016, device="cuda")
g.sha256(x); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
  g.sha256(x)
b.record(); b.synchronize()
print("sha256 tree digest of 102 MB: %.3f ms (%.0f GB/s)" % (a.elapsed_time(b) / 10, x.numel() * 4 / (a.elapsed_time(b) / 10) / 1e6))
PY
timeout 400 python bench.py --steps 20 --warmup 5 --no-baseline --skip-e2e > ${O}_bench.log 2>&1; echo "bench: $(grep -ao '"ms_per_step": [0-9.]*' ${O}_bench.log | head -1)"

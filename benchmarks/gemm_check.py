#!/usr/bin/env python3
"""Correctness + throughput probe of the tcgen05 GEMM, one layout variant per process (a trap in one variant must not
take the others down). Usage: gemm_check.py {nt|nn|tn} ; appends JSON lines to gpurun_out/gemm_check.jsonl."""

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aggregathor_b200.ops import nn_native as nat  # noqa: E402

variant = sys.argv[1]
persistent = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nat.set_gemm_persistent(persistent)
shapes = [(128, 128, 64), (128, 64, 64), (256, 128, 256), (300, 200, 136), (32, 10, 104), (4096, 4096, 4096), (25088, 64, 256), (25088, 256, 64), (8192, 8192, 8192)]
if variant == "tn":
  shapes = [(128, 128, 64), (64, 64, 128), (256, 128, 256), (200, 300, 136), (64, 64, 100352), (512, 2048, 1568), (4096, 4096, 4096)]
os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/gemm_check.jsonl", "a")
for m, n, k in shapes:
  gen = torch.Generator(device="cuda").manual_seed(m + n + k)
  if variant == "nt":
    a, b = torch.randn((m, k), device="cuda", generator=gen).bfloat16(), torch.randn((n, k), device="cuda", generator=gen).bfloat16()
    ref = a.float() @ b.float().t()
    run = lambda: nat.mm_nt(a, b, out_dtype=torch.float32)
  elif variant == "nn":
    a, b = torch.randn((m, k), device="cuda", generator=gen).bfloat16(), torch.randn((k, n), device="cuda", generator=gen).bfloat16()
    ref = a.float() @ b.float()
    run = lambda: nat.mm_nn(a, b, out_dtype=torch.float32)
  else:
    a, b = torch.randn((k, m), device="cuda", generator=gen).bfloat16(), torch.randn((k, n), device="cuda", generator=gen).bfloat16()
    ref = a.float().t() @ b.float()
    run = lambda: nat.mm_tn(a, b)
  res = run()
  torch.cuda.synchronize()
  err = float((res - ref).abs().max())
  scale = float(ref.abs().max())
  for _ in range(3):
    run()
  torch.cuda.synchronize()
  begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  iters = 10
  begin.record()
  for _ in range(iters):
    run()
  end.record()
  torch.cuda.synchronize()
  ms = begin.elapsed_time(end) / iters
  a16, b16 = (a, b.t()) if variant == "nt" else (a, b) if variant == "nn" else (a.t(), b)
  for _ in range(3):
    torch.matmul(a16, b16)
  begin.record()
  for _ in range(iters):
    torch.matmul(a16, b16)
  end.record()
  torch.cuda.synchronize()
  lib_ms = begin.elapsed_time(end) / iters
  line = {"variant": variant, "persistent": persistent, "m": m, "n": n, "k": k, "max_err": err, "ref_scale": scale, "ok": err <= 3e-3 * max(1.0, scale), "ms": ms,
          "tflops": 2.0 * m * n * k / ms / 1e9, "cublas_ms": lib_ms, "cublas_tflops": 2.0 * m * n * k / lib_ms / 1e9}
  print(json.dumps(line), flush=True)
  out.write(json.dumps(line) + "\n")
  out.flush()

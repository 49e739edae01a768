#!/usr/bin/env python3
"""Throughput of the NT bf16 GEMM variants on large and layer-shaped problems: single-CTA persistent kernel (128 x 128 / 128 x 256 tiles),
CTA-pair kernel (`tcgen05.mma.cta_group::2`, 256 x 256 tiles) and cuBLAS (torch.matmul), bf16 output. JSON lines -> gpurun_out/gemm_pair_bench.jsonl."""

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aggregathor_b200.ops import nn_native as nat  # noqa: E402

shapes = [(8192, 8192, 8192), (4096, 4096, 4096), (16384, 4096, 1024), (100352, 256, 64), (100352, 64, 256), (25088, 512, 128), (25088, 128, 512), (25088, 1024, 256),
          (6272, 1024, 256), (6272, 256, 1024), (6272, 2048, 1024), (1568, 2048, 512), (256, 4096, 4096), (2048, 4096, 4096)]
os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/gemm_pair_bench.jsonl", "a")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > L2: operands come from HBM in every timed call


def timed(run, iters=10):
  for _ in range(3):
    run()
  total = 0.0
  for _ in range(iters):
    flush.zero_()
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    begin.record()
    run()
    end.record()
    torch.cuda.synchronize()
    total += begin.elapsed_time(end)
  return total / iters


for m, n, k in shapes:
  gen = torch.Generator(device="cuda").manual_seed(m + n + k)
  a, b = torch.randn((m, k), device="cuda", generator=gen).bfloat16(), torch.randn((n, k), device="cuda", generator=gen).bfloat16()
  ref = a.float() @ b.float().t() if m * n <= 64 << 20 else None
  line = {"m": m, "n": n, "k": k}
  for name, bn in (("cta1_bn128", 128), ("cta1_bn256", 256), ("pair_256x256", 512)):
    nat.set_gemm_pair("0" if bn != 512 else "1")
    res = nat.mm_nt(a, b, bn=bn)
    torch.cuda.synchronize()
    if ref is not None:
      line[name + "_max_err"] = float((res.float() - ref).abs().max()) / max(1.0, float(ref.abs().max()))
    ms = timed(lambda: nat.mm_nt(a, b, bn=bn))
    line[name + "_ms"], line[name + "_tflops"] = ms, 2.0 * m * n * k / ms / 1e9
  ms = timed(lambda: torch.matmul(a, b.t()))
  line["cublas_ms"], line["cublas_tflops"] = ms, 2.0 * m * n * k / ms / 1e9
  print(json.dumps(line), flush=True)
  out.write(json.dumps(line) + "\n")
  out.flush()

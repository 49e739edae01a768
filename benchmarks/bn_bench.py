#!/usr/bin/env python3
"""Batch-norm kernels at the ResNet-50 layer shapes: single-launch (resident tile) vs statistics + apply pair, forward and
backward, CUDA-event timed inside a CUDA graph of `reps` back-to-back launches (so launch gaps look like the training step)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aggregathor_b200.ops import nn as ops, nn_native  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--batch", type=int, default=32)
parser.add_argument("--reps", type=int, default=20)
parser.add_argument("--out", default="gpurun_out/bn_bench.json")
parser.add_argument("--only", default="", help="C,HW: a single shape")
parser.add_argument("--no-limits", action="store_true", help="lift the size envelope of the single-launch kernels")
parser.add_argument("--eager", action="store_true", help="no CUDA graph: plain launches (for ncu)")
args = parser.parse_args()
torch.cuda.set_device(0)
if args.no_limits:
  nn_native.set_bn_fused_limits(1 << 20, 1 << 20, 1)
shapes = [(64, 112), (64, 56), (256, 56), (128, 28), (512, 28), (256, 14), (1024, 14), (512, 7), (2048, 7)]
if args.only:
  shapes = [tuple(int(v) for v in args.only.split(","))]
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")  # 256 MB > L2
results = []
for c, hw in shapes:
  x = torch.randn((args.batch, c, hw, hw), device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
  dy = torch.randn_like(x).contiguous(memory_format=torch.channels_last)
  gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
  mm, mv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
  gg, gb = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
  entry = {"C": c, "HW": hw, "MB": x.numel() * 2 / 1e6}
  for fused in (True, False):
    nn_native.set_bn_fused(fused)
    y, mean, rstd = ops.batchnorm_forward("native", x, gamma, beta, mm, mv, 0.9, 1e-5, True)
    for name, fn in (("fwd", lambda: ops.batchnorm_forward("native", x, gamma, beta, mm, mv, 0.9, 1e-5, True)),
                     ("bwd", lambda: ops.batchnorm_backward("native", dy, x, y, gamma, mean, rstd, True, gg, gb))):
      for _ in range(3):
        fn()
      torch.cuda.synchronize()
      if args.eager:
        class graph:  # noqa: N801 - same interface as a captured graph
          @staticmethod
          def replay():
            for _ in range(args.reps):
              fn()
      else:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
          for _ in range(args.reps):
            fn()
      graph.replay()
      torch.cuda.synchronize()
      times = []
      for warm in (False, True):
        if not warm:
          flush.zero_()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        graph.replay()
        stop.record()
        torch.cuda.synchronize()
        times.append(start.elapsed_time(stop) * 1000.0 / args.reps)
      entry["%s_%s_us" % (name, "fused" if fused else "pair")] = round(times[1], 2)
  results.append(entry)
  print(entry)
nn_native.set_bn_fused(True)
os.makedirs(os.path.dirname(args.out), exist_ok=True)
with open(args.out, "w") as fd:
  json.dump({"batch": args.batch, "reps": args.reps, "note": "us per launch (pair = two kernels), back-to-back launches inside one CUDA graph", "results": results}, fd, indent=1)

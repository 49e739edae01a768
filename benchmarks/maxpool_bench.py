#!/usr/bin/env python3
"""Time the max-pool kernels on the ResNet stem shape (batch x 64 x 112 x 112, window 3, stride 2, TF "SAME"): us per call, CUDA events."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from aggregathor_b200.ops import nn_native

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = torch.randn((batch, 64, 112, 112), device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
pads = (0, 1, 0, 1)
y, arg = nn_native.maxpool_forward(x, 3, 2, pads)
dy = torch.randn_like(y).contiguous(memory_format=torch.channels_last)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=20):
  total = 0.0
  for _ in range(reps + 3):
    flush.zero_()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    fn()
    stop.record()
    stop.synchronize()
    if _ >= 3:
      total += start.elapsed_time(stop)
  return 1000.0 * total / reps


print("batch %d: forward %.1f us, backward %.1f us" % (batch, timed(lambda: nn_native.maxpool_forward(x, 3, 2, pads)), timed(lambda: nn_native.maxpool_backward(dy, x.shape, arg, 3, 2, pads))))

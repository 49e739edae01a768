#!/usr/bin/env python3
"""Per-output errors of the native batch-norm kernels (single-launch and pair) against an fp32 torch reference, with groups."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aggregathor_b200.ops import nn as ops, nn_native  # noqa: E402

torch.cuda.set_device(0)
CL = torch.channels_last
for rep in range(2):
  for c, hw, groups, batch in ((64, 28, 4, 8), (512, 7, 8, 8), (256, 14, 2, 8), (1024, 4, 1, 8), (64, 56, 1, 32), (64, 112, 1, 32)):
    gen = torch.Generator(device="cuda").manual_seed(c + hw)
    x = (torch.randn((batch * groups, c, hw, hw), device="cuda", generator=gen) + 0.25).to(torch.bfloat16).contiguous(memory_format=CL)
    dy = torch.randn((batch * groups, c, hw, hw), device="cuda", generator=gen).to(torch.bfloat16).contiguous(memory_format=CL)
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    outs = {}
    mask = None
    for tag in ("fused", "pair", "torch"):
      if tag != "torch":
        nn_native.set_bn_fused(tag == "fused")
      backend = "torch" if tag == "torch" else "native"
      xin, dyin = (x.float(), dy.float()) if tag == "torch" else (x, dy)
      mm, mv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
      y, mean, rstd = ops.batchnorm_forward(backend, xin, gamma, beta, mm, mv, 0.9, 1e-5, True, groups)
      if mask is None:
        mask = y
      grads = torch.zeros((groups, 2, c), device="cuda")
      dx = ops.batchnorm_backward(backend, dyin, xin, mask.to(xin.dtype), gamma, mean, rstd, True, grads[0, 0], grads[0, 1], groups, grads.stride(0))
      torch.cuda.synchronize()
      outs[tag] = (y.float(), mean, rstd, dx.float(), grads, mm, mv)
    names = ("y", "mean", "rstd", "dx", "grads", "mm", "mv")
    line = "rep %d C=%d hw=%d groups=%d:" % (rep, c, hw, groups)
    for tag in ("fused", "pair"):
      errs = ["%s %.2e" % (n, float((a - b).abs().max() / max(1e-3, float(b.abs().max())))) for n, a, b in zip(names, outs[tag], outs["torch"])]
      line += "\n   " + tag + ": " + ", ".join(errs)
    print(line)
nn_native.set_bn_fused(True)

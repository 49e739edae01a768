#!/usr/bin/env python3
"""Fused BN + residual add + ReLU (forward) and its backward with the masked-gradient output, with groups, vs composed ops."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aggregathor_b200.ops import nn as ops, nn_native  # noqa: E402

torch.cuda.set_device(0)
CL = torch.channels_last
for rep in range(2):
  for c, hw, groups, batch in ((256, 16, 4, 8), (512, 8, 4, 8), (1024, 4, 4, 8), (2048, 2, 4, 8), (256, 16, 1, 8), (2048, 2, 1, 8), (64, 32, 4, 8)):
    gen = torch.Generator(device="cuda").manual_seed(c + hw)
    mk = lambda: torch.randn((batch * groups, c, hw, hw), device="cuda", generator=gen).to(torch.bfloat16).contiguous(memory_format=CL)
    x, res, dy = mk() + 0.25, mk(), mk()
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    outs = {}
    for tag in ("fused", "pair"):
      nn_native.set_bn_fused(tag == "fused")
      mm, mv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
      y, mean, rstd = ops.batchnorm_add_relu_forward("native", x, gamma, beta, mm, mv, 0.997, 1e-5, res, groups)
      grads = torch.zeros((groups, 2, c), device="cuda")
      mask = y if tag == "fused" else mask
      dx, g = ops.batchnorm_add_relu_backward("native", dy, x, mask, gamma, mean, rstd, grads[0, 0], grads[0, 1], groups, grads.stride(0))
      torch.cuda.synchronize()
      outs[tag] = (y.float(), mean, rstd, dx.float(), g.float(), grads)
    names = ("y", "mean", "rstd", "dx", "g", "grads")
    errs = ["%s %.2e" % (n, float((a - b).abs().max() / max(1e-3, float(b.abs().max())))) for n, a, b in zip(names, outs["fused"], outs["pair"])]
    print("rep %d C=%d hw=%d groups=%d: " % (rep, c, hw, groups) + ", ".join(errs))
nn_native.set_bn_fused(True)

#!/usr/bin/env python3
"""Per-tensor error of the native and torch providers against an fp32 reference (conv fwd / dgrad / wgrad / bias grad)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aggregathor_b200.ops import nn as ops  # noqa: E402

CL = torch.channels_last


def rand(shape, seed, scale=1.0):
  gen = torch.Generator(device="cuda").manual_seed(seed)
  return (torch.randn(shape, device="cuda", generator=gen) * scale).to(torch.bfloat16)


def rel(a, b):
  return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp(min=1e-6))


for cin, cout, k, stride, hw, pads in [(64, 64, 3, 1, 56, (1, 1, 1, 1)), (128, 128, 3, 2, 28, (1, 1, 1, 1)), (256, 512, 1, 1, 14, (0, 0, 0, 0)), (64, 64, 5, 1, 16, (2, 2, 2, 2)), (3, 64, 7, 2, 64, (3, 3, 3, 3))]:
  n = 4
  x = rand((n, cin, hw, hw), 3).contiguous(memory_format=CL)
  w = rand((cout, k, k, cin), 4, (2.0 / (k * k * cin)) ** 0.5).contiguous()
  bias = torch.randn(cout, device="cuda") * 0.1
  xf, wf = x.float(), w.float().permute(0, 3, 1, 2)
  xp = F.pad(xf, (pads[2], pads[3], pads[0], pads[1]))
  xp.requires_grad_(True)
  wf = wf.clone().requires_grad_(True)
  bf = bias.clone().requires_grad_(True)
  yf = torch.relu(F.conv2d(xp, wf, bf, stride))
  dy = rand(tuple(yf.shape), 5).contiguous(memory_format=CL)
  yf.backward(dy.float())
  ref = {"y": yf.detach(), "gw": wf.grad.permute(0, 2, 3, 1), "gb": bf.grad, "dx": xp.grad[:, :, pads[0]:pads[0] + hw, pads[2]:pads[2] + hw]}
  line = {"case": [cin, cout, k, stride, hw]}
  for backend in ("torch", "native"):
    y = ops.conv2d_forward(backend, x, w, bias, stride, pads, True)
    gw, gb = torch.zeros((cout, k, k, cin), device="cuda"), torch.zeros(cout, device="cuda")
    dx, _, _ = ops.conv2d_backward(backend, dy, x, w, y, stride, pads, True, True, cin % 8 == 0, gw, gb)
    line[backend] = {"y": rel(y, ref["y"]), "gw": rel(gw, ref["gw"]), "gb": rel(gb, ref["gb"]), "dx": rel(dx, ref["dx"]) if dx is not None else None}
  print(json.dumps(line), flush=True)

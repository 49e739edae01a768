#!/usr/bin/env python3
"""Implicit vs im2col native wgrad/dgrad/forward on one configuration; prints where the two disagree."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aggregathor_b200.ops import nn_native as nat  # noqa: E402

CL = torch.channels_last
for (n, cin, cout, k, hw) in [(4, 256, 256, 3, 14), (2, 64, 64, 3, 14), (4, 64, 64, 3, 14), (4, 128, 128, 3, 14), (8, 256, 256, 3, 14), (32, 256, 256, 3, 14)]:
  gen = torch.Generator(device="cuda").manual_seed(1)
  x = torch.randn((n, cin, hw, hw), device="cuda", generator=gen).to(torch.bfloat16).contiguous(memory_format=CL)
  dy = torch.randn((n, cout, hw, hw), device="cuda", generator=gen).to(torch.bfloat16).contiguous(memory_format=CL)
  w = (torch.randn((cout, k, k, cin), device="cuda", generator=gen) * 0.05).to(torch.bfloat16)
  pads = ((k - 1) // 2,) * 4
  res = {}
  for tag, disabled in (("implicit", set()), ("im2col", {"implicit"})):
    nat._DISABLED = disabled
    gw = torch.zeros((cout, k, k, cin), device="cuda")
    y = nat.conv2d_forward(x, w, None, 1, pads, False)
    dx = nat.conv2d_backward(dy, x, w, None, 1, pads, False, False, True, gw, None)
    res[tag] = (y.float(), dx.float(), gw)
  ref = torch.nn.functional.conv2d(x.float(), w.float().permute(0, 3, 1, 2), None, 1, pads[0])
  out = []
  for idx, name in enumerate(("y", "dx", "gw")):
    a, b = res["implicit"][idx], res["im2col"][idx]
    err = (a - b).abs()
    out.append("%s max err %.4f / scale %.2f" % (name, float(err.max()), float(b.abs().max())))
    if name == "gw" and float(err.max()) > 1e-2 * float(b.abs().max()):
      per_tap = err.amax(dim=(0, 3)).flatten().tolist()
      per_co = err.amax(dim=(1, 2, 3))
      per_ci = err.amax(dim=(0, 1, 2))
      out.append("  per-tap max err: " + " ".join("%.2f" % v for v in per_tap))
      out.append("  bad co: %s" % (per_co > 0.5).nonzero().flatten().tolist()[:20])
      out.append("  bad ci: %s" % (per_ci > 0.5).nonzero().flatten().tolist()[:20])
  print((n, cin, cout, k, hw), "; ".join(out[:3]), "; y vs fp32 %.4f" % float((res["implicit"][0] - ref).abs().max()))
  for line in out[3:]:
    print(line)

#!/usr/bin/env python3
"""Representative launches of every hot kernel, for `ncu --set full --import-source on -k regex:<name> -c N` captures.

Usage: ncu_targets.py {gemm|conv|bn|gar|all}. Each target runs its kernel a few times after warm-up on ResNet-50-sized shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aggregathor_b200.aggregators import FusedSpec  # noqa: E402
from aggregathor_b200.ops import gar as gar_ops  # noqa: E402
from aggregathor_b200.ops import nn as ops  # noqa: E402
from aggregathor_b200.ops import nn_native as nat  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
CL = torch.channels_last
rand = lambda *shape: torch.randn(shape, device="cuda").to(torch.bfloat16)

if what in ("gemm", "all"):
  a, b = rand(8192, 8192), rand(8192, 8192)
  for _ in range(3):
    nat.mm_nt(a, b)                                    # large square forward product
  x, w = rand(100352, 64), rand(256, 64)
  for _ in range(3):
    nat.mm_nt(x, w)                                    # ResNet block1 1x1 conv (M = 32*56*56)
  dy = rand(100352, 256)
  for _ in range(3):
    nat.mm_tn(dy, x)                                   # its weight gradient (split-K)
if what == "gemm_wide":      # the shipped default for N >= 256: 128 x 256 tiles, persistent, double-buffered TMEM
  a, b = rand(8192, 8192), rand(8192, 8192)
  nat.set_gemm_pair("0")
  for _ in range(3):
    nat.mm_nt(a, b)
if what == "gemm_pair":      # CTA-pair kernel (cta_group::2)
  a, b = rand(8192, 8192), rand(8192, 8192)
  for _ in range(3):
    nat.mm_nt(a, b, bn=512)
if what == "gemm_tf32":
  a, b = torch.randn((8192, 8192), device="cuda"), torch.randn((8192, 8192), device="cuda")
  for _ in range(3):
    nat.mm_nt(a, b)
if what in ("conv", "all"):
  x = rand(32, 64, 56, 56).contiguous(memory_format=CL)
  w = rand(64, 3, 3, 64)
  gw = torch.zeros((64, 3, 3, 64), device="cuda")
  for _ in range(3):
    y = ops.conv2d_forward("native", x, w, None, 1, (1, 1, 1, 1), False)
    ops.conv2d_backward("native", y, x, w, None, 1, (1, 1, 1, 1), False, False, True, gw, None)
if what in ("bn", "all"):
  x = rand(32, 512, 28, 28).contiguous(memory_format=CL)     # 25.7 MB: inside the single-launch envelope in both directions
  gamma, beta = torch.ones(512, device="cuda"), torch.zeros(512, device="cuda")
  mm, mv = torch.zeros(512, device="cuda"), torch.ones(512, device="cuda")
  gg, gb = torch.zeros(512, device="cuda"), torch.zeros(512, device="cuda")
  for _ in range(3):
    y, mean, rstd = ops.batchnorm_forward("native", x, gamma, beta, mm, mv, 0.997, 1e-5, True)
    ops.batchnorm_backward("native", x, x, y, gamma, mean, rstd, True, gg, gb)
if what in ("gar", "all"):
  G = torch.randn((8, 25558016), device="cuda")
  for _ in range(3):
    gar_ops.aggregate(FusedSpec("krum", 8, f=2, m=4), G)
  for _ in range(2):
    gar_ops.aggregate(FusedSpec("median", 8), G)
torch.cuda.synchronize()
print("done", what)

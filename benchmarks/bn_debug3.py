#!/usr/bin/env python3
"""ResNet-50 (64 px, 4 workers x batch 8): sequential vs batched passes, single-launch vs pair batch-norm kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aggregathor_b200.engine.flat import FlatLayout  # noqa: E402
from aggregathor_b200.models import Context, get_network  # noqa: E402
from aggregathor_b200.ops import nn_native  # noqa: E402

torch.cuda.set_device(0)
workers, batch, image, classes = 4, 8, 64, 1000
model = get_network("resnet_v1_50", classes)
layout, shapes = FlatLayout(), {}
model.declare(layout, shapes)
layout.freeze()
init = torch.zeros(layout.padded_size)
init_states = {k: torch.zeros(v) for k, v in shapes.items()}
model.initialize(layout.views(init), init_states, torch.Generator().manual_seed(0))
params = init.cuda()
weights = params.to(torch.bfloat16)
gen = torch.Generator(device="cuda").manual_seed(7)
xs = [torch.randn((batch, 3, image, image), device="cuda", generator=gen).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(workers)]
ys = [torch.randint(0, classes, (batch,), device="cuda") for _ in range(workers)]


def context(rows):
  ctx = Context("native", True, torch.bfloat16, "cuda")
  ctx.master, ctx.weights = layout.views(params), layout.views(weights)
  ctx.state = {k: v.clone().cuda() for k, v in init_states.items()}
  ctx.grads = layout.views(rows)
  return ctx


runs = {}
for fused in (False, True, True):
  nn_native.set_bn_fused(fused)
  seq = torch.zeros((workers, layout.padded_size), device="cuda")
  seq_losses = [float(model.loss_and_backward(x, y, context(seq[i]))) for i, (x, y) in enumerate(zip(xs, ys))]
  bat = torch.zeros((workers, layout.padded_size), device="cuda")
  ctx = context(bat[0])
  ctx.groups, ctx.group_stride = workers, bat.stride(0)
  losses = model.loss_and_backward(torch.cat(xs, dim=0).contiguous(memory_format=torch.channels_last), torch.cat(ys), ctx).tolist()
  torch.cuda.synchronize()
  tag = ("fused" if fused else "pair") + str(len(runs))
  runs[tag] = (seq, seq_losses, bat, losses)
  print(tag, "seq losses", ["%.4f" % v for v in seq_losses], "bat losses", ["%.4f" % v for v in losses])
cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a, b, dim=0))
ref_seq, _, ref_bat, _ = runs["pair0"]
for tag, (seq, _, bat, _) in runs.items():
  print(tag, "cos(seq, pair seq)", ["%.4f" % cos(seq[i], ref_seq[i]) for i in range(workers)], "cos(bat, pair seq)", ["%.4f" % cos(bat[i], ref_seq[i]) for i in range(workers)])
# per-variable view of the worst worker
tag = [t for t in runs if t.startswith("fused")][0]
seq, _, bat, _ = runs[tag]
worst = []
for name in layout.names:
  a, b = layout.view(bat[0], name).flatten(), layout.view(ref_seq[0], name).flatten()
  worst.append((cos(a, b), name))
worst.sort()
print("lowest per-variable cosines (batched fused vs sequential pair, worker 0):", worst[:12])
worst = []
for name in layout.names:
  a, b = layout.view(seq[0], name).flatten(), layout.view(ref_seq[0], name).flatten()
  worst.append((cos(a, b), name))
worst.sort()
print("lowest per-variable cosines (sequential fused vs sequential pair, worker 0):", worst[:6])

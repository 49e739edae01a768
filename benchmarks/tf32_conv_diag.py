"""Layer-by-layer check of the TF32 native convolutions / dense layers inside a real ResNet step: every conv2d forward / backward of
slim resnet_v1_18 (64 x 64 input, batch 8) is run by the native provider AND by the aten provider in strict fp32 on the same inputs."""
import sys; sys.path.insert(0, "/root/repo")
import torch
from aggregathor_b200.engine.flat import FlatLayout
from aggregathor_b200.models import Context, nets_factory
from aggregathor_b200.ops import nn as ops
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False

def rel(a, b):
  if a is None or b is None: return -1.0
  return float((a.float() - b.float()).abs().max()) / max(float(b.float().abs().max()), 1e-9)

orig_fwd, orig_bwd = ops.conv2d_forward, ops.conv2d_backward
def fwd(backend, x, weight, bias, stride, pads, relu, aux=None):
  y = orig_fwd(backend, x, weight, bias, stride, pads, relu, aux)
  ref = orig_fwd("torch", x, weight, bias, stride, pads, relu, None)
  print("fwd  x%s w%s s%d pads%s relu%d: y %.2e" % (tuple(x.shape), tuple(weight.shape), stride, pads, relu, rel(y, ref)))
  return y
def bwd(backend, dy, x, weight, y, stride, pads, relu, has_bias, need_dx, grad_w, grad_b, groups=1, group_stride=0, aux=None):
  gw_ref = torch.zeros_like(grad_w); gb_ref = torch.zeros_like(grad_b) if grad_b is not None else None
  dx_ref, _, _ = orig_bwd("torch", dy, x, weight, y, stride, pads, relu, has_bias, need_dx, gw_ref, gb_ref)
  out = orig_bwd(backend, dy, x, weight, y, stride, pads, relu, has_bias, need_dx, grad_w, grad_b, groups, group_stride, aux)
  torch.cuda.synchronize()
  print("bwd  x%s w%s s%d pads%s: gw %.2e dx %.2e gb %.2e" % (tuple(x.shape), tuple(weight.shape), stride, pads, rel(grad_w, gw_ref), rel(out[0], dx_ref), rel(grad_b, gb_ref) if grad_b is not None else -1))
  return out
ops.conv2d_forward, ops.conv2d_backward = fwd, bwd
import aggregathor_b200.models.core as core
core.nn_ops.conv2d_forward, core.nn_ops.conv2d_backward = fwd, bwd

model = nets_factory.get_network("resnet_v1_18", 16)
layout, states = FlatLayout(), {}
model.declare(layout, states); layout.freeze()
host = torch.zeros(layout.padded_size); host_states = {k: torch.zeros(v) for k, v in states.items()}
model.initialize(layout.views(host), host_states, torch.Generator().manual_seed(1))
params = host.cuda()
ctx = Context("native", True, torch.float32, "cuda")
ctx.master = ctx.weights = layout.views(params)
ctx.state = {k: v.cuda() for k, v in host_states.items()}
grads = torch.zeros_like(params); ctx.grads = layout.views(grads)
gen = torch.Generator(device="cuda").manual_seed(3)
x = torch.randn((8, 3, 64, 64), device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
labels = torch.randint(0, 16, (8,), device="cuda", generator=gen)
print("loss", float(model.loss_and_backward(x, labels, ctx)))

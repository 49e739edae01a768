"""MobileNet v1 (slim `mobilenet_v1`): 3x3/2 stem conv + 13 depthwise-separable blocks (3x3 depthwise + BN +
ReLU6, 1x1 pointwise + BN + ReLU6), global average pool, dropout, 1x1 conv logits. Depthwise convolutions
run through the torch provider (grouped conv); pointwise 1x1 convolutions are plain GEMMs."""

import torch
import torch.nn.functional as F

from .core import BatchNorm, Conv2d, Dropout, GlobalAvgPool, Model, Module, Sequential, same_padding, _trunc_normal_


class DepthwiseConv2d(Module):
  def __init__(self, name, channels, k, stride):
    super().__init__(name)
    self.channels, self.k, self.stride = channels, k, stride

  def declare(self, layout, states):
    layout.add(self.name + "/depthwise_weights", (self.channels, self.k, self.k, 1))

  def initialize(self, master, states, generator):
    _trunc_normal_(master[self.name + "/depthwise_weights"], 0.09, generator)

  def forward(self, x, ctx):
    n, c, h, w = x.shape
    t, b = same_padding(h, self.k, self.stride)
    l, r = same_padding(w, self.k, self.stride)
    xp = F.pad(x, (l, r, t, b))
    self._saved = (xp, (t, b, l, r), (h, w))
    weight = ctx.weights[self.name + "/depthwise_weights"].permute(0, 3, 1, 2)
    return F.conv2d(xp, weight, None, self.stride, 0, 1, c).contiguous(memory_format=torch.channels_last)

  def backward(self, dy, ctx):
    xp, (t, b, l, r), (h, w) = self._saved
    self._saved = None
    weight = ctx.weights[self.name + "/depthwise_weights"].permute(0, 3, 1, 2)
    from ..ops.nn import group_view
    pieces = []
    for g, (dy_g, xp_g) in enumerate(zip(dy.chunk(ctx.groups, dim=0), xp.chunk(ctx.groups, dim=0))):
      dxp, dw, _ = torch.ops.aten.convolution_backward(dy_g, xp_g, weight, None, [self.stride] * 2, [0, 0], [1, 1], False, [0, 0], self.channels, [True, True, False])
      group_view(ctx.grads[self.name + "/depthwise_weights"], g, ctx.group_stride).copy_(dw.permute(0, 2, 3, 1))
      pieces.append(dxp)
    dxp = pieces[0] if len(pieces) == 1 else torch.cat(pieces, dim=0)
    return dxp[:, :, t:t + h, l:l + w].contiguous(memory_format=torch.channels_last)


class ReLU6(Module):
  def forward(self, x, ctx):
    self._saved_x = x
    return torch.clamp(x, 0.0, 6.0)

  def backward(self, dy, ctx):
    x, self._saved_x = self._saved_x, None
    return dy * ((x > 0) & (x < 6)).to(dy.dtype)


def mobilenet_v1(num_classes=1001, multiplier=1.0, name="mobilenet_v1"):
  s = "MobilenetV1"
  depth = lambda d: max(int(d * multiplier), 8)
  layers = [Conv2d(s + "/Conv2d_0", 3, depth(32), 3, stride=2, padding="SAME", init="truncated_normal", init_std=0.09),
            BatchNorm(s + "/Conv2d_0/BatchNorm", depth(32), decay=0.9997, epsilon=0.001), ReLU6(s + "/Conv2d_0/Relu6")]
  cin = depth(32)
  spec = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2)] + [(512, 1)] * 5 + [(1024, 2), (1024, 1)]
  for i, (cout, stride) in enumerate(spec):
    base = "%s/Conv2d_%d" % (s, i + 1)
    cout = depth(cout)
    layers += [DepthwiseConv2d(base + "_depthwise", cin, 3, stride), BatchNorm(base + "_depthwise/BatchNorm", cin, decay=0.9997, epsilon=0.001), ReLU6(base + "_depthwise/Relu6"),
               Conv2d(base + "_pointwise", cin, cout, 1, padding="SAME", init="truncated_normal", init_std=0.09),
               BatchNorm(base + "_pointwise/BatchNorm", cout, decay=0.9997, epsilon=0.001), ReLU6(base + "_pointwise/Relu6")]
    cin = cout
  layers += [GlobalAvgPool(s + "/AvgPool_1a"), Dropout(s + "/Dropout_1b", 0.999),
             Conv2d(s + "/Logits/Conv2d_1c_1x1", cin, num_classes, 1, padding="SAME", bias=True, init="truncated_normal", init_std=0.09)]
  return Model(name, Sequential(name, layers), (3, 224, 224), num_classes)

"""Static layer graph with explicit forward/backward — the model-side execution engine.

The reference builds a TF1 static graph per worker replica and lets TF differentiate it
(`graph.py:254-273`). Here models are trees of `Module`s with hand-written `forward` and
`backward`, executed over pre-planned buffers:

* parameters are *views* into the flat fp32 master buffer (and its bf16 compute copy); a module
  never owns storage, so the aggregation kernel and the layers share memory with zero copies;
* `backward` writes each variable's gradient straight into the current worker's row of the
  peer-mapped gradient matrix (`ctx.grads[name]`), in fp32;
* no autograd tape, no tracing compiler: a step is a fixed sequence of kernel launches, which is
  what CUDA-graph capture wants.

Activations are NHWC in memory (torch shape (N, C, H, W) with channels_last strides), bf16 on
GPU / fp32 on CPU. Conv weights are OHWI in memory (K-major rows for the implicit-GEMM kernels).
Every op has two providers selected by `ctx.backend`: `"native"` = the hand-written sm_100a
kernels of `native/op_nn` (ops/nn.py), `"torch"` = aten library calls (cuDNN/cuBLAS) used as
the numerical reference and as the baseline arm.
"""

import math

import torch
import torch.nn.functional as F

from .. import tools
from ..ops import nn as nn_ops


class Context:
  """Per-execution state handed to every module."""

  def __init__(self, backend="torch", training=True, dtype=torch.float32, device="cpu"):
    self.backend = backend      # "native" | "torch"
    self.training = training
    self.dtype = dtype          # activation/compute dtype
    self.device = torch.device(device)
    self.weights = None         # name -> compute-dtype view of the parameters
    self.master = None          # name -> fp32 master view
    self.grads = None           # name -> fp32 gradient view of the current worker
    self.state = None           # name -> non-trainable state tensor (BN moving statistics)
    self.generator = None       # torch.Generator for dropout
    self.need_input_grad = False
    self.groups = 1             # logical workers batched in one pass (their batches are consecutive along dim 0)
    self.group_stride = 0       # elements between two workers' gradient rows (`grads` are worker 0's views)


class Module:
  """Base layer: declares variables, runs forward, runs backward (storing what it needs in `self`)."""

  def __init__(self, name=None):
    self.name = name if name is not None else type(self).__name__.lower()

  def declare(self, layout, states):
    """Register trainable variables in `layout` (FlatLayout) and non-trainable ones in `states` (dict name -> shape)."""

  def initialize(self, master, states, generator):
    """Write initial values into the fp32 master views / state tensors."""

  def forward(self, x, ctx):
    raise NotImplementedError

  def backward(self, dy, ctx):
    raise NotImplementedError

  def children(self):
    return ()

  def release(self):
    """Drop tensors saved for backward."""
    for key in [k for k in self.__dict__ if k.startswith("_saved")]:
      self.__dict__[key] = None
    for child in self.children():
      child.release()


def _trunc_normal_(tensor, std, generator):
  torch.nn.init.trunc_normal_(tensor, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=generator)


def _conv_out(size, k, stride, pad_lo, pad_hi):
  return (size + pad_lo + pad_hi - k) // stride + 1


def same_padding(size, k, stride):
  """TensorFlow 'SAME' padding -> (lo, hi)."""
  out = -(-size // stride)
  total = max((out - 1) * stride + k - size, 0)
  return total // 2, total - total // 2


class Conv2d(Module):
  """2-D convolution, NHWC / OHWI. `padding`: "SAME" (TF semantics), "VALID", or an int (symmetric explicit,
  the `conv2d_same` convention of slim's resnet_utils: pad (k-1)//2, k-1-(k-1)//2 then VALID)."""

  def __init__(self, name, cin, cout, k, stride=1, padding="SAME", bias=False, relu=False, init="variance_scaling", init_std=None, bias_init=0.0):
    super().__init__(name)
    self.cin, self.cout, self.k, self.stride, self.padding = cin, cout, k, stride, padding
    self.bias, self.relu, self.init, self.init_std, self.bias_init = bias, relu, init, init_std, bias_init

  def declare(self, layout, states):
    layout.add(self.name + "/weights", (self.cout, self.k, self.k, self.cin))
    if self.bias:
      layout.add(self.name + "/biases", (self.cout,))

  def initialize(self, master, states, generator):
    w = master[self.name + "/weights"]
    fan_in = self.k * self.k * self.cin
    if self.init == "truncated_normal":
      _trunc_normal_(w, self.init_std, generator)
    elif self.init == "xavier":
      limit = math.sqrt(6.0 / (fan_in + self.k * self.k * self.cout))
      w.uniform_(-limit, limit, generator=generator)
    else:  # slim variance_scaling_initializer(): truncated normal, stddev = sqrt(2 / fan_in) / .8796...
      _trunc_normal_(w, math.sqrt(2.0 / fan_in) / 0.87962566103423978, generator)
    if self.bias:
      master[self.name + "/biases"].fill_(self.bias_init)

  def _pads(self, h, w):
    if self.padding == "SAME":
      return same_padding(h, self.k, self.stride) + same_padding(w, self.k, self.stride)
    if self.padding == "VALID":
      return (0, 0, 0, 0)
    total = self.k - 1
    return (total // 2, total - total // 2, total // 2, total - total // 2)

  def forward(self, x, ctx):
    n, c, h, w = x.shape
    pads = self._pads(h, w)
    weight = ctx.weights[self.name + "/weights"]
    bias = ctx.master[self.name + "/biases"] if self.bias else None
    y = nn_ops.conv2d_forward(ctx.backend, x, weight, bias, self.stride, pads, self.relu)
    if ctx.training:
      self._saved_x, self._saved_y, self._saved_pads = x, (y if self.relu else None), pads
    return y

  def backward(self, dy, ctx):
    x, pads = self._saved_x, self._saved_pads
    weight = ctx.weights[self.name + "/weights"]
    need_dx = ctx.need_input_grad or not getattr(self, "is_first", False)
    dx, dw, db = nn_ops.conv2d_backward(ctx.backend, dy, x, weight, self._saved_y, self.stride, pads, self.relu, self.bias, need_dx,
                                        ctx.grads[self.name + "/weights"], ctx.grads[self.name + "/biases"] if self.bias else None, ctx.groups, ctx.group_stride)
    self._saved_x = self._saved_y = None
    return dx


class Dense(Module):
  """Fully connected layer y = x W^T + b, W stored [out, in] (K-major rows)."""

  def __init__(self, name, fin, fout, relu=False, bias=True, init="xavier", init_std=None, bias_init=0.0):
    super().__init__(name)
    self.fin, self.fout, self.relu, self.bias, self.init, self.init_std, self.bias_init = fin, fout, relu, bias, init, init_std, bias_init

  def declare(self, layout, states):
    layout.add(self.name + "/weights", (self.fout, self.fin))
    if self.bias:
      layout.add(self.name + "/biases", (self.fout,))

  def initialize(self, master, states, generator):
    w = master[self.name + "/weights"]
    if self.init == "truncated_normal":
      _trunc_normal_(w, self.init_std, generator)
    else:  # glorot uniform: tf.get_variable's default initializer (reference: experiments/mnist.py:95-96)
      limit = math.sqrt(6.0 / (self.fin + self.fout))
      w.uniform_(-limit, limit, generator=generator)
    if self.bias:
      master[self.name + "/biases"].fill_(self.bias_init)

  def forward(self, x, ctx):
    x = x.reshape(x.shape[0], -1)
    weight = ctx.weights[self.name + "/weights"]
    bias = ctx.master[self.name + "/biases"] if self.bias else None
    y = nn_ops.linear_forward(ctx.backend, x, weight, bias, self.relu)
    if ctx.training:
      self._saved_x, self._saved_y = x, (y if self.relu else None)
    return y

  def backward(self, dy, ctx):
    need_dx = ctx.need_input_grad or not getattr(self, "is_first", False)
    dx = nn_ops.linear_backward(ctx.backend, dy, self._saved_x, ctx.weights[self.name + "/weights"], self._saved_y, self.relu, need_dx,
                                ctx.grads[self.name + "/weights"], ctx.grads[self.name + "/biases"] if self.bias else None, ctx.groups, ctx.group_stride)
    self._saved_x = self._saved_y = None
    return dx


class BatchNorm(Module):
  """Batch normalisation over (N, H, W) per channel with trainable gamma/beta and moving statistics
  (slim `batch_norm`: decay 0.997, epsilon 1e-5, scale=True in `resnet_arg_scope`), optionally fused with ReLU."""

  def __init__(self, name, channels, relu=False, decay=0.997, epsilon=1e-5, scale=True):
    super().__init__(name)
    self.channels, self.relu, self.decay, self.epsilon, self.scale = channels, relu, decay, epsilon, scale

  def declare(self, layout, states):
    if self.scale:
      layout.add(self.name + "/gamma", (self.channels,))
    layout.add(self.name + "/beta", (self.channels,))
    states[self.name + "/moving_mean"] = (self.channels,)
    states[self.name + "/moving_variance"] = (self.channels,)

  def initialize(self, master, states, generator):
    if self.scale:
      master[self.name + "/gamma"].fill_(1.0)
    master[self.name + "/beta"].zero_()
    states[self.name + "/moving_mean"].zero_()
    states[self.name + "/moving_variance"].fill_(1.0)

  def forward(self, x, ctx):
    gamma = ctx.master[self.name + "/gamma"] if self.scale else None
    beta = ctx.master[self.name + "/beta"]
    mean, var = ctx.state[self.name + "/moving_mean"], ctx.state[self.name + "/moving_variance"]
    if not ctx.training:
      return nn_ops.batchnorm_inference(ctx.backend, x, gamma, beta, mean, var, self.epsilon, self.relu)
    y, batch_mean, batch_rstd = nn_ops.batchnorm_forward(ctx.backend, x, gamma, beta, mean, var, self.decay, self.epsilon, self.relu, ctx.groups)
    self._saved = (x, y if self.relu else None, batch_mean, batch_rstd)
    return y

  def backward(self, dy, ctx):
    x, y, batch_mean, batch_rstd = self._saved
    gamma = ctx.master[self.name + "/gamma"] if self.scale else None
    dx = nn_ops.batchnorm_backward(ctx.backend, dy, x, y, gamma, batch_mean, batch_rstd, self.relu,
                                   ctx.grads[self.name + "/gamma"] if self.scale else None, ctx.grads[self.name + "/beta"], ctx.groups, ctx.group_stride)
    self._saved = None
    return dx


class LayerNorm(Module):
  """Layer normalisation over the feature dimension of a [rows, features] activation (trainable gamma / beta)."""

  def __init__(self, name, features, epsilon=1e-5):
    super().__init__(name)
    self.features, self.epsilon = features, epsilon

  def declare(self, layout, states):
    layout.add(self.name + "/gamma", (self.features,))
    layout.add(self.name + "/beta", (self.features,))

  def initialize(self, master, states, generator):
    master[self.name + "/gamma"].fill_(1.0)
    master[self.name + "/beta"].zero_()

  def forward(self, x, ctx):
    x = x.reshape(x.shape[0], -1)
    y, mean, rstd = nn_ops.layernorm_forward(ctx.backend, x, ctx.master[self.name + "/gamma"], ctx.master[self.name + "/beta"], self.epsilon)
    if ctx.training:
      self._saved = (x, mean, rstd)
    return y

  def backward(self, dy, ctx):
    x, mean, rstd = self._saved
    self._saved = None
    return nn_ops.layernorm_backward(ctx.backend, dy, x, ctx.master[self.name + "/gamma"], mean, rstd, ctx.grads[self.name + "/gamma"], ctx.grads[self.name + "/beta"],
                                     ctx.groups, ctx.group_stride)


class ReLU(Module):
  def forward(self, x, ctx):
    y = nn_ops.relu_forward(ctx.backend, x)
    if ctx.training:
      self._saved_y = y
    return y

  def backward(self, dy, ctx):
    dx = nn_ops.relu_backward(ctx.backend, dy, self._saved_y)
    self._saved_y = None
    return dx


class MaxPool(Module):
  """Max pooling; `padding` "SAME" / "VALID" (TF semantics; SAME pads with -inf)."""

  def __init__(self, name, k, stride, padding="VALID"):
    super().__init__(name)
    self.k, self.stride, self.padding = k, stride, padding

  def forward(self, x, ctx):
    n, c, h, w = x.shape
    pads = (same_padding(h, self.k, self.stride) + same_padding(w, self.k, self.stride)) if self.padding == "SAME" else (0, 0, 0, 0)
    y, index = nn_ops.maxpool_forward(ctx.backend, x, self.k, self.stride, pads)
    if ctx.training:
      self._saved = (x.shape, index, pads, x, y)
    return y

  def backward(self, dy, ctx):
    shape, index, pads, x, y = self._saved
    self._saved = None
    return nn_ops.maxpool_backward(ctx.backend, dy, shape, index, self.k, self.stride, pads, x, y)


class GlobalAvgPool(Module):
  """Mean over H, W keeping a 1x1 map (slim `pool5`)."""

  def forward(self, x, ctx):
    self._saved_shape = x.shape
    return nn_ops.global_avgpool_forward(ctx.backend, x)

  def backward(self, dy, ctx):
    return nn_ops.global_avgpool_backward(ctx.backend, dy, self._saved_shape)


class AvgPool(Module):
  """Average pooling (VALID), used by a few slim nets."""

  def __init__(self, name, k, stride):
    super().__init__(name)
    self.k, self.stride = k, stride

  def forward(self, x, ctx):
    self._saved_shape = x.shape
    return F.avg_pool2d(x, self.k, self.stride)

  def backward(self, dy, ctx):
    n, c, h, w = self._saved_shape
    ones = torch.ones((c, 1, self.k, self.k), dtype=dy.dtype, device=dy.device) / (self.k * self.k)
    return F.conv_transpose2d(dy, ones, stride=self.stride, groups=c, output_padding=((h - self.k) % self.stride, (w - self.k) % self.stride)).contiguous(memory_format=torch.channels_last)


class Subsample(Module):
  """slim `resnet_utils.subsample`: 1x1 max-pool with stride s == strided slicing."""

  def __init__(self, name, stride):
    super().__init__(name)
    self.stride = stride

  def forward(self, x, ctx):
    if self.stride == 1:
      return x
    self._saved_shape = x.shape
    return nn_ops.subsample_forward(ctx.backend, x, self.stride)

  def backward(self, dy, ctx):
    if self.stride == 1:
      return dy
    return nn_ops.subsample_backward(ctx.backend, dy, self._saved_shape, self.stride)


class Dropout(Module):
  def __init__(self, name, keep_prob):
    super().__init__(name)
    self.keep_prob = keep_prob

  def forward(self, x, ctx):
    if not ctx.training or self.keep_prob >= 1.0:
      self._saved_mask = None
      return x
    mask = (torch.rand(x.shape, device=x.device, generator=ctx.generator) < self.keep_prob).to(x.dtype) / self.keep_prob
    self._saved_mask = mask
    return x * mask

  def backward(self, dy, ctx):
    mask, self._saved_mask = self._saved_mask, None
    return dy if mask is None else dy * mask


class Flatten(Module):
  """NHWC flatten: (N, C, H, W) channels_last -> (N, H*W*C), matching TF's reshape of an NHWC tensor."""

  def forward(self, x, ctx):
    self._saved_shape = x.shape
    if x.dim() == 4:
      return x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)
    return x.reshape(x.shape[0], -1)

  def backward(self, dy, ctx):
    shape = self._saved_shape
    if len(shape) == 4:
      n, c, h, w = shape
      return dy.reshape(n, h, w, c).permute(0, 3, 1, 2)
    return dy.reshape(shape)


class Sequential(Module):
  def __init__(self, name, layers):
    super().__init__(name)
    self.layers = list(layers)

  def children(self):
    return self.layers

  def declare(self, layout, states):
    for layer in self.layers:
      layer.declare(layout, states)

  def initialize(self, master, states, generator):
    for layer in self.layers:
      layer.initialize(master, states, generator)

  def forward(self, x, ctx):
    for layer in self.layers:
      x = layer.forward(x, ctx)
    return x

  def backward(self, dy, ctx):
    for layer in reversed(self.layers):
      dy = layer.backward(dy, ctx)
    return dy


class Residual(Module):
  """y = relu(shortcut(x) + residual(x)) (ResNet v1 unit) or shortcut(x) + residual(x) (`relu=False`)."""

  def __init__(self, name, shortcut, residual, relu=True):
    super().__init__(name)
    self.shortcut, self.residual, self.relu = shortcut, residual, relu

  def children(self):
    return (self.shortcut, self.residual)

  def declare(self, layout, states):
    self.shortcut.declare(layout, states)
    self.residual.declare(layout, states)

  def initialize(self, master, states, generator):
    self.shortcut.initialize(master, states, generator)
    self.residual.initialize(master, states, generator)

  def forward(self, x, ctx):
    a = self.shortcut.forward(x, ctx)
    b = self.residual.forward(x, ctx)
    y = nn_ops.add_relu_forward(ctx.backend, a, b, self.relu)
    if ctx.training and self.relu:
      self._saved_y = y
    return y

  def backward(self, dy, ctx):
    if self.relu:
      dy = nn_ops.relu_backward(ctx.backend, dy, self._saved_y)
      self._saved_y = None
    da = self.shortcut.backward(dy, ctx)
    db = self.residual.backward(dy, ctx)
    return nn_ops.add_forward(ctx.backend, da, db)


class Identity(Module):
  def forward(self, x, ctx):
    return x

  def backward(self, dy, ctx):
    return dy


class Branches(Module):
  """Run several branches on the same input and concatenate along channels (Inception-style blocks)."""

  def __init__(self, name, branches):
    super().__init__(name)
    self.branches = list(branches)

  def children(self):
    return self.branches

  def declare(self, layout, states):
    for branch in self.branches:
      branch.declare(layout, states)

  def initialize(self, master, states, generator):
    for branch in self.branches:
      branch.initialize(master, states, generator)

  def forward(self, x, ctx):
    outs = [branch.forward(x, ctx) for branch in self.branches]
    self._saved_split = [o.shape[1] for o in outs]
    return torch.cat(outs, dim=1).contiguous(memory_format=torch.channels_last)

  def backward(self, dy, ctx):
    total = None
    for branch, piece in zip(self.branches, torch.split(dy, self._saved_split, dim=1)):
      dx = branch.backward(piece.contiguous(memory_format=torch.channels_last), ctx)
      total = dx if total is None else total + dx
    return total


class Model:
  """A root module + input description + loss head."""

  def __init__(self, name, root, input_shape, num_classes, label_smoothing=0.0, aux=None):
    self.name, self.root, self.input_shape, self.num_classes = name, root, tuple(input_shape), num_classes
    self.label_smoothing = label_smoothing
    first = self._first_layer(root)
    if first is not None:
      first.is_first = True

  @staticmethod
  def _first_layer(module):
    while True:
      kids = list(module.children())
      if not kids:
        return module if isinstance(module, (Conv2d, Dense)) else None
      if isinstance(module, (Residual, Branches)):
        return None
      module = kids[0]

  def declare(self, layout, states):
    self.root.declare(layout, states)

  def initialize(self, master, states, generator):
    self.root.initialize(master, states, generator)

  def logits(self, x, ctx):
    y = self.root.forward(x, ctx)
    self._raw_shape = y.shape
    return y.reshape(y.shape[0], -1)

  def loss_and_backward(self, x, labels, ctx):
    """Forward, mean softmax cross-entropy, backward. Returns the loss (0-d fp32 tensor; shape [ctx.groups] when several logical
    workers are batched: each worker's loss is the mean over its own slice of the batch)."""
    logits = self.logits(x, ctx)
    loss, dlogits = nn_ops.softmax_xent(ctx.backend, logits, labels, self.label_smoothing, ctx.groups)
    dlogits = dlogits.to(ctx.dtype).reshape(self._raw_shape)
    if dlogits.dim() == 4:
      dlogits = dlogits.contiguous(memory_format=torch.channels_last)
    self.root.backward(dlogits, ctx)
    return loss

  def accuracy(self, x, labels, ctx):
    logits = self.logits(x, ctx)
    return (logits.float().argmax(dim=1) == labels).float().mean()

"""Explicit backward passes of the static layer graph, checked against finite differences in float64 (CPU)."""

import pytest
import torch

from aggregathor_b200.engine.flat import FlatLayout
from aggregathor_b200.models import Context, get_network


def _setup(name, classes, dtype=torch.float64):
  model = get_network(name, classes)
  layout, shapes = FlatLayout(), {}
  model.declare(layout, shapes)
  layout.freeze()
  gen = torch.Generator().manual_seed(0)
  params = torch.zeros(layout.padded_size, dtype=torch.float32)
  states = {k: torch.zeros(v) for k, v in shapes.items()}
  model.initialize(layout.views(params), states, gen)
  return model, layout, params.to(dtype), {k: v.to(dtype) for k, v in states.items()}


def _loss(model, layout, params, states, x, labels, grads=None):
  ctx = Context("torch", True, params.dtype, "cpu")
  ctx.master = ctx.weights = layout.views(params)
  ctx.state = {k: v.clone() for k, v in states.items()}
  g = torch.zeros_like(params) if grads is None else grads
  ctx.grads = layout.views(g)
  return float(model.loss_and_backward(x, labels, ctx)), g


@pytest.mark.parametrize("name,classes,shape", [
  ("mlp", 10, (6, 784)), ("cnnet", 10, (3, 3, 32, 32)), ("resnet_v1_18", 7, (3, 3, 32, 32)), ("lenet", 10, (3, 1, 28, 28))])
def test_directional_derivative(name, classes, shape):
  torch.manual_seed(1)
  model, layout, params, states = _setup(name, classes)
  # dropout-free check: lenet has dropout -> evaluate with keep_prob 1 by patching
  from aggregathor_b200.models.core import Dropout
  def strip(module):
    if isinstance(module, Dropout):
      module.keep_prob = 1.0
    for child in module.children():
      strip(child)
  strip(model.root)
  x = torch.randn(shape, dtype=torch.float64)
  if len(shape) == 4:
    x = x.contiguous(memory_format=torch.channels_last)
  labels = torch.randint(0, classes, (shape[0],))
  loss, grad = _loss(model, layout, params, states, x, labels)
  mask = layout.mask()
  assert float(grad[~mask].abs().max()) == 0.0  # padding coordinates never receive gradient
  direction = torch.randn_like(params) * mask
  direction /= direction.norm()
  eps = 1e-5
  plus, _ = _loss(model, layout, params + eps * direction, states, x, labels)
  minus, _ = _loss(model, layout, params - eps * direction, states, x, labels)
  numeric = (plus - minus) / (2 * eps)
  analytic = float((grad * direction).sum())
  assert abs(numeric - analytic) <= 1e-5 + 1e-4 * abs(analytic), (numeric, analytic)


def test_parameter_counts_match_reference():
  expected = {"mlp": 79510, "cnnet": 1756426, "resnet_v1_50": 25557032}
  for name, count in expected.items():
    model = get_network(name, 1000 if name.startswith("resnet") else 10)
    layout = FlatLayout()
    model.declare(layout, {})
    assert layout.size == count, (name, layout.size)


def test_unbuilt_networks_raise_user_exception():
  from aggregathor_b200 import tools
  with pytest.raises(tools.UserException):
    get_network("inception_v3", 1000)
  with pytest.raises(tools.UserException):
    get_network("nope", 10)


def _check_stack(layers, shape, classes=5):
  """Per-variable directional derivatives of a small custom stack (float64)."""
  from aggregathor_b200.models.core import Model, Sequential
  model = Model("t", Sequential("t", layers), shape[1:], classes)
  layout, shapes = FlatLayout(), {}
  model.declare(layout, shapes)
  layout.freeze()
  params = torch.zeros(layout.padded_size)
  states = {k: torch.zeros(v) for k, v in shapes.items()}
  model.initialize(layout.views(params), states, torch.Generator().manual_seed(0))
  params, states = params.double(), {k: v.double() for k, v in states.items()}
  torch.manual_seed(2)
  x = torch.randn(shape, dtype=torch.float64).contiguous(memory_format=torch.channels_last)
  labels = torch.randint(0, classes, (shape[0],))
  _, grad = _loss(model, layout, params, states, x, labels)
  for name in layout.names:
    direction = torch.zeros_like(params)
    layout.view(direction, name).normal_()
    direction /= direction.norm()
    eps = 1e-6
    numeric = (_loss(model, layout, params + eps * direction, states, x, labels)[0] - _loss(model, layout, params - eps * direction, states, x, labels)[0]) / (2 * eps)
    analytic = float((grad * direction).sum())
    assert abs(numeric - analytic) <= 1e-7 + 1e-4 * abs(analytic), (name, numeric, analytic)


def test_resnet_v2_units_and_depthwise_blocks():
  from aggregathor_b200.models.core import BatchNorm, Conv2d, GlobalAvgPool, MaxPool
  from aggregathor_b200.models.mobilenet import DepthwiseConv2d, ReLU6
  from aggregathor_b200.models.resnet import _PreactUnit
  head = lambda c: [GlobalAvgPool("gap"), Conv2d("logits", c, 5, 1, padding="SAME", bias=True)]
  _check_stack([Conv2d("stem", 3, 16, 3, padding="SAME", bias=True), _PreactUnit("u1", 16, 32, 8, 1)] + head(32), (4, 3, 8, 8))
  _check_stack([Conv2d("stem", 3, 32, 7, stride=2, padding="explicit", bias=True), MaxPool("p", 3, 2, "SAME"), _PreactUnit("u1", 32, 32, 8, 2)] + head(32), (4, 3, 32, 32))
  from aggregathor_b200.models.core import Dense, Flatten, LayerNorm
  _check_stack([Flatten("f"), Dense("d1", 48, 24, relu=True), LayerNorm("ln", 24), Dense("d2", 24, 5)], (6, 3, 4, 4))
  _check_stack([Conv2d("stem", 3, 16, 3, stride=2, padding="SAME"), BatchNorm("bn0", 16), ReLU6("r0"), DepthwiseConv2d("dw", 16, 3, 2), BatchNorm("bn1", 16), ReLU6("r1"),
                Conv2d("pw", 16, 24, 1, padding="SAME"), BatchNorm("bn2", 24), ReLU6("r2")] + head(24), (4, 3, 16, 16))


@pytest.mark.parametrize("name,classes,shape", [("cnnet", 10, (2, 3, 32, 32)), ("resnet_v1_18", 7, (2, 3, 32, 32)), ("mlp", 10, (3, 784))])
def test_batched_workers_match_sequential_workers(name, classes, shape):
  """`ctx.groups` = W logical workers in one pass (per-worker BN statistics / loss means / gradient rows) == W separate passes."""
  workers = 3
  model, layout, params, states = _setup(name, classes, torch.float64)
  torch.manual_seed(3)
  xs = [torch.randn(shape, dtype=torch.float64) for _ in range(workers)]
  if len(shape) == 4:
    xs = [x.contiguous(memory_format=torch.channels_last) for x in xs]
  ys = [torch.randint(0, classes, (shape[0],)) for _ in range(workers)]
  rows = torch.zeros((workers, layout.padded_size), dtype=torch.float64)
  sequential = [_loss(model, layout, params, states, x, y, rows[i])[0] for i, (x, y) in enumerate(zip(xs, ys))]
  batched_rows = torch.zeros((workers, layout.padded_size), dtype=torch.float64)
  ctx = Context("torch", True, torch.float64, "cpu")
  ctx.master = ctx.weights = layout.views(params)
  ctx.state = {k: v.clone() for k, v in states.items()}
  ctx.grads = layout.views(batched_rows[0])
  ctx.groups, ctx.group_stride = workers, batched_rows.stride(0)
  x_all = torch.cat(xs, dim=0)
  if len(shape) == 4:
    x_all = x_all.contiguous(memory_format=torch.channels_last)
  losses = model.loss_and_backward(x_all, torch.cat(ys), ctx)
  assert losses.shape == (workers,)
  assert torch.allclose(losses, torch.tensor(sequential, dtype=losses.dtype), rtol=1e-9, atol=1e-12)
  assert torch.allclose(batched_rows, rows, rtol=1e-7, atol=1e-10)

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


def test_every_slim_name_builds():
  """All 33 names of the reference's slim factory (`external/slim/nets/nets_factory.py:39-72`) are constructible; unknown names raise."""
  from aggregathor_b200 import tools
  from aggregathor_b200.models import nets_factory
  assert len(nets_factory.networks_map) == 33
  sizes = {}
  for name in nets_factory.networks_map:
    model = get_network(name, 1001)
    layout, states = FlatLayout(), {}
    model.declare(layout, states)
    sizes[name] = layout.size
    assert model.input_shape[1] == nets_factory.default_image_size(name), name
  # published trainable-parameter counts (1001 classes)
  assert sizes["mobilenet_v2"] == 3506153
  assert 6.5e6 < sizes["inception_v1"] < 6.7e6 and 27.0e6 < sizes["inception_v3"] < 27.3e6
  assert 5.2e6 < sizes["nasnet_mobile"] - 2.5e6 < 5.4e6  # 5.3 M without the auxiliary head
  with pytest.raises(tools.UserException):
    get_network("nope", 10)


def _check_stack(layers, shape, classes=5, every=1):
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
  for name in layout.names[::every]:
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


def test_inception_pieces():
  """Rectangular kernels, SAME average pooling (stride 1 and 2, odd maps), scaled residual block, auxiliary head."""
  from aggregathor_b200.models.core import AuxHead, AvgPool, BatchNorm, Branches, Conv2d, Dense, Flatten, GlobalAvgPool, Identity, MaxPool, Residual, Scale, Sequential
  head = lambda c: [GlobalAvgPool("gap"), Conv2d("logits", c, 5, 1, padding="SAME", bias=True)]
  towers = Branches("mixed", [
    Sequential("b0", [Conv2d("b0/1x7", 8, 6, (1, 7)), BatchNorm("b0/bn", 6, relu=True, scale=False), Conv2d("b0/7x1", 6, 6, (7, 1))]),
    Sequential("b1", [AvgPool("b1/pool", 3, 1, "SAME"), Conv2d("b1/1x1", 8, 4, 1)]),
    Sequential("b2", [Conv2d("b2/3x1", 8, 4, (3, 1), padding="VALID"), Conv2d("b2/1x3", 4, 4, (1, 3), padding="VALID"), Conv2d("b2/fix", 4, 4, 3, padding="explicit")])])
  # b2 shrinks the map by 2 in each dimension: keep it out of the concat by using its own stack below
  towers.branches.pop()
  res = Residual("res", Identity("sc"), Sequential("r", [towers, Conv2d("up", 10, 8, 1, bias=True), Scale("scale", 0.17)]), relu=True)
  aux = AuxHead("aux", Sequential("auxh", [AvgPool("aux/pool", 5, 3, "VALID"), Conv2d("aux/conv", 8, 6, 1), Flatten("aux/flat"), Dense("aux/fc", 6, 5)]))
  _check_stack([Conv2d("stem", 3, 8, 3, stride=1, padding="SAME"), res, aux, AvgPool("red", 3, 2, "SAME"), MaxPool("mp", 3, 2, "SAME")] + head(8), (3, 3, 7, 7))
  _check_stack([Conv2d("stem", 3, 8, 3, padding="SAME"), Conv2d("b2/3x1", 8, 4, (3, 1), padding="VALID"), Conv2d("b2/1x3", 4, 4, (1, 3), stride=2, padding="SAME")] + head(4), (2, 3, 9, 8))


@pytest.mark.parametrize("family", ["nasnet", pytest.param("pnasnet", marks=pytest.mark.slow)])
def test_searched_cells(family):
  """A miniature NASNet-A / PNASNet-5 (cifar stem, 3 cells with both reductions, 4 filters): DAG backward vs finite differences."""
  from aggregathor_b200.models import nasnet
  model = nasnet._build("tiny", family, 5, 20, "cifar", num_cells=3, filters=4, stem_multiplier=1.0, drop_path_keep_prob=1.0, dense_keep_prob=1.0, skip_reduction_input=(family == "pnasnet"))
  _check_stack([model.root], (2, 3, 20, 20), every=3)


@pytest.mark.slow
def test_imagenet_stem_cells():
  from aggregathor_b200.models import nasnet
  model = nasnet._build("tiny", "nasnet", 5, 83, "imagenet", num_cells=3, filters=8, stem_multiplier=0.25, drop_path_keep_prob=1.0, dense_keep_prob=1.0, skip_reduction_input=True)
  _check_stack([model.root], (2, 3, 83, 83), every=9)


@pytest.mark.parametrize("name,shape", [("mobilenet_v2_035", (2, 3, 64, 64)), ("inception_v1", (2, 3, 64, 64)), ("nasnet_cifar", (2, 3, 32, 32))])
def test_new_families_step(name, shape):
  """fp32 smoke step at a reduced resolution: finite loss, gradient reaches the first and the last variable, drop-path / dropout active."""
  model, layout, params, states = _setup(name, 11, dtype=torch.float32)
  torch.manual_seed(3)
  x = torch.randn(shape).contiguous(memory_format=torch.channels_last)
  labels = torch.randint(0, 11, (shape[0],))
  ctx = Context("torch", True, torch.float32, "cpu")
  ctx.master = ctx.weights = layout.views(params)
  ctx.state = {k: v.clone() for k, v in states.items()}
  grads = torch.zeros_like(params)
  ctx.grads = layout.views(grads)
  ctx.generator = torch.Generator().manual_seed(5)
  loss = float(model.loss_and_backward(x, labels, ctx))
  assert loss == loss and loss < 1e3
  assert float(layout.view(grads, layout.names[0]).abs().sum()) > 0 and float(layout.view(grads, layout.names[-1]).abs().sum()) > 0
  ctx.training = False
  assert 0.0 <= float(model.accuracy(x, labels, ctx)) <= 1.0


def test_drop_path_and_dropout_masks():
  """Per-sample drop-path and element-wise dropout: kept entries are rescaled by 1 / keep_prob, backward applies the same mask,
  evaluation is the identity."""
  from aggregathor_b200.models.core import DropPath, Dropout
  ctx = Context("torch", True, torch.float32, "cpu")
  ctx.generator = torch.Generator().manual_seed(3)
  x = torch.ones((64, 4, 3, 3))
  for module, per_sample in ((DropPath("dp", 0.75), True), (Dropout("do", 0.75), False)):
    y = module.forward(x, ctx)
    kept = y != 0
    assert torch.allclose(y[kept], torch.full_like(y[kept], 1 / 0.75))
    assert 0.5 < float(kept.float().mean()) < 0.95
    if per_sample:
      flat = kept.reshape(64, -1)
      assert bool((flat.all(dim=1) | (~flat).all(dim=1)).all())     # a sample is kept or dropped as a whole
    dx = module.backward(torch.ones_like(x), ctx)
    assert torch.equal(dx, y)
    ctx.training = False
    assert module.forward(x, ctx) is x
    ctx.training = True


def test_graph_module_accumulates_fanout_gradients():
  """A value consumed by several nodes receives the sum of their gradients; multi-input graphs return one gradient per input."""
  from aggregathor_b200.models.core import Add, Concat, Graph, Scale
  ctx = Context("torch", True, torch.float64, "cpu")
  graph = Graph("g", [(Scale("a", 2.0), (0,)), (Scale("b", 3.0), (0,)), (Add("sum"), (2, 3)), (Concat("cat"), (4, 1)), (Scale("c", 0.5), (5,))], nb_inputs=2)
  x0, x1 = torch.randn(2, 3, 4, 4, dtype=torch.float64), torch.randn(2, 5, 4, 4, dtype=torch.float64)
  y = graph.forward([x0, x1], ctx)
  assert torch.allclose(y, 0.5 * torch.cat([5.0 * x0, x1], dim=1))
  g0, g1 = graph.backward(torch.ones_like(y), ctx)
  assert torch.allclose(g0, torch.full_like(x0, 2.5)) and torch.allclose(g1, torch.full_like(x1, 0.5))

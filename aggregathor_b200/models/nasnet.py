"""NASNet-A (`nasnet_cifar`, `nasnet_mobile`, `nasnet_large`) and PNASNet-5 (`pnasnet_large`, `pnasnet_mobile`) — the
searched-cell families of the slim factory (names: reference `external/slim/nets/nets_factory.py:39-72`).

A network is a chain of cells; cell c consumes the outputs of cells c-1 ("net") and c-2 ("prev"). Inside a cell both
inputs are projected to the cell's filter count (1x1 conv + BN, or a *factorised reduction* when the previous layer has a
larger map), then five combinations `h = op_left(h_i) + op_right(h_j)` are appended to the list of hidden states, and the
states nobody consumed are concatenated. Operations: stacked separable convolutions (ReLU -> depthwise k x k -> 1x1 -> BN,
twice, stride on the first), 3x3 average / max pooling, identity. Each cell is a `Graph` (static DAG with explicit
backward); the whole network is a `Graph` as well because of the skip from cell c-2.

Normalisation: BN decay 0.9997, epsilon 1e-3, with gamma; no biases. The auxiliary head sits before the second reduction
and is weighted 0.4 in the loss (reference `experiments/slims.py:122-125`). Drop-path is applied per sample on every
non-identity branch in training, with the cell-depth-scaled keep probability.
"""

from .core import (Add, AuxHead, AvgPool, BatchNorm, Concat, Conv2d, Dense, DepthwiseConv2d, DropPath, Dropout, Flatten, GlobalAvgPool, Graph, Identity, MaxPool, Model,
                   OffsetSubsample, ReLU, Sequential, Subsample)

_NASNET_NORMAL = dict(
  operations=["separable_5x5_2", "separable_3x3_2", "separable_5x5_2", "separable_3x3_2", "avg_pool_3x3", "none", "avg_pool_3x3", "avg_pool_3x3", "separable_3x3_2", "none"],
  used=[1, 0, 0, 0, 0, 0, 0], indices=[0, 1, 1, 1, 0, 1, 1, 1, 0, 0])
_NASNET_REDUCTION = dict(
  operations=["separable_5x5_2", "separable_7x7_2", "max_pool_3x3", "separable_7x7_2", "avg_pool_3x3", "separable_5x5_2", "none", "avg_pool_3x3", "separable_3x3_2", "max_pool_3x3"],
  used=[1, 1, 1, 0, 0, 0, 0], indices=[0, 1, 0, 1, 0, 1, 3, 2, 2, 0])
_PNASNET_CELL = dict(
  operations=["separable_5x5_2", "max_pool_3x3", "separable_7x7_2", "max_pool_3x3", "separable_5x5_2", "separable_3x3_2", "separable_3x3_2", "max_pool_3x3", "separable_3x3_2", "none"],
  used=[1, 1, 0, 0, 0, 0, 0], indices=[1, 1, 0, 0, 0, 0, 4, 0, 1, 0])


def _bn(name, channels):
  return BatchNorm(name, channels, decay=0.9997, epsilon=0.001, scale=True)


def _conv1x1(name, cin, cout, stride=1):
  layers = [Subsample(name + "/stride", stride)] if stride > 1 else []  # a strided 1x1 convolution only ever sees the kept pixels
  return layers + [Conv2d(name, cin, cout, 1, padding="SAME", bias=False)]


def _factorized_reduction(name, cin, cout, stride):
  """Halve the map without losing information at odd offsets: two 1x1 convolutions on the two pixel phases, concatenated."""
  if stride == 1:
    return Sequential(name, [Conv2d(name + "/path_conv", cin, cout, 1, padding="SAME", bias=False), _bn(name + "/path_bn", cout)])
  nodes = [
    (Sequential(name + "/path1", [Subsample(name + "/path1/pool", 2), Conv2d(name + "/path1_conv", cin, cout // 2, 1, padding="SAME", bias=False)]), (0,)),
    (Sequential(name + "/path2", [OffsetSubsample(name + "/path2/pool", 2, 1), Conv2d(name + "/path2_conv", cin, cout // 2 + cout % 2, 1, padding="SAME", bias=False)]), (0,)),
    (Concat(name + "/concat"), (1, 2)),
    (_bn(name + "/final_path_bn", cout), (3,))]
  return Graph(name, nodes)


def _operation(name, kind, cin, filters, stride, keep_prob):
  """One branch of a combination, applied to a hidden state of depth `cin`; returns a single-input module producing `filters` channels."""
  layers = []
  if kind.startswith("separable"):
    k = int(kind.split("_")[1].split("x")[0])
    repeats = int(kind.split("_")[2])
    depth = cin
    for layer in range(repeats):
      tag = "%s/separable_%dx%d_%d" % (name, k, k, layer + 1)
      layers += [ReLU(tag + "/relu"), DepthwiseConv2d(tag, depth, k, stride if layer == 0 else 1), Conv2d(tag + "/pointwise", depth, filters, 1, padding="SAME", bias=False),
                 _bn("%s/bn_sep_%dx%d_%d" % (name, k, k, layer + 1), filters)]
      depth = filters
  elif kind == "none":
    if stride > 1 or cin != filters:
      layers += [ReLU(name + "/relu")] + _conv1x1(name + "/1x1", cin, filters, stride) + [_bn(name + "/bn_1", filters)]
  else:
    pool = AvgPool if kind.startswith("avg") else MaxPool
    layers.append(pool(name + "/" + kind, 3, stride, "SAME"))
    if cin != filters:
      layers += _conv1x1(name + "/1x1", cin, filters) + [_bn(name + "/bn_1", filters)]
  if kind != "none" and keep_prob < 1.0:
    layers.append(DropPath(name + "/drop_path", keep_prob))
  if not layers:
    return Identity(name + "/identity")
  return Sequential(name, layers)


def _cell(name, spec, net_depth, prev_depth, prev_larger, filters, stride, keep_prob):
  """Build one cell as a two-input `Graph` (value 0 = net, value 1 = prev). Returns (graph, output depth)."""
  nodes = []

  def add(module, inputs):
    nodes.append((module, inputs))
    return 1 + len(nodes)  # id of the value just produced (2 inputs precede the nodes)

  # -- cell base: bring both inputs to `filters` channels at the resolution of `net`
  if prev_larger:
    prev = add(Sequential(name + "/prev", [ReLU(name + "/prev_relu"), _factorized_reduction(name + "/prev", prev_depth, filters, 2)]), (1,))
  elif prev_depth != filters:
    prev = add(Sequential(name + "/prev", [ReLU(name + "/prev_relu"), Conv2d(name + "/prev_1x1", prev_depth, filters, 1, padding="SAME", bias=False), _bn(name + "/prev_bn", filters)]), (1,))
  else:
    prev = 1
  net = add(Sequential(name + "/begin", [ReLU(name + "/relu"), Conv2d(name + "/1x1", net_depth, filters, 1, padding="SAME", bias=False), _bn(name + "/beginning_bn", filters)]), (0,))
  states = [net, prev]
  reduced = [False, False]  # whether a state already has the cell's output resolution (only matters for stride 2)
  # -- five combinations
  for step in range(5):
    sides = []
    for side, slot in (("left", 2 * step), ("right", 2 * step + 1)):
      source = spec["indices"][slot]
      original = source < 2
      op_stride = stride if original else 1
      module = _operation("%s/comb_iter_%d/%s" % (name, step, side), spec["operations"][slot], filters, filters, op_stride, keep_prob)
      sides.append(add(module, (states[source],)))
    states.append(add(Add("%s/comb_iter_%d/combine" % (name, step)), tuple(sides)))
    reduced.append(True)
  # -- concatenate the states nobody consumed (bringing a not-yet-reduced cell input to the output resolution first)
  unused = []
  for index, used in enumerate(spec["used"]):
    if used:
      continue
    if stride > 1 and not reduced[index]:
      unused.append(add(_factorized_reduction("%s/reduction_%d" % (name, index), filters, filters, 2), (states[index],)))
    else:
      unused.append(states[index])
  out = add(Concat(name + "/cell_output"), tuple(unused))
  return Graph(name, nodes, nb_inputs=2, output=out), filters * len(unused)


def _aux_head(name, cin, num_classes, spatial):
  pooled = (spatial - 5) // 3 + 1
  return Sequential(name, [
    ReLU(name + "/relu"), AvgPool(name + "/pool", 5, 3, "VALID"), Conv2d(name + "/proj", cin, 128, 1, padding="SAME", bias=False), _bn(name + "/aux_bn0", 128), ReLU(name + "/relu0"),
    Conv2d(name + "/conv768", 128, 768, pooled, padding="VALID", bias=False), _bn(name + "/aux_bn1", 768), ReLU(name + "/relu1"),
    Flatten(name + "/flatten"), Dense(name + "/FC", 768, num_classes)])


def _reduction_layers(num_cells, num_reduction_layers=2):
  return [int(float(pool) / (num_reduction_layers + 1) * num_cells) for pool in range(1, num_reduction_layers + 1)]


def _build(name, family, num_classes, image_size, stem, num_cells, filters, stem_multiplier, drop_path_keep_prob, dense_keep_prob, skip_reduction_input, rate=2.0):
  """Chain the cells. `family`: "nasnet" (separate reduction cells inserted before the normal cells at the reduction indices)
  or "pnasnet" (one cell type; the cells at the reduction indices run with stride 2)."""
  nodes = []  # network-level graph: value 0 = image

  def add(module, inputs):
    nodes.append((module, inputs))
    return len(nodes)

  reductions = _reduction_layers(num_cells)
  total_cells = num_cells + (2 if family == "nasnet" else 0) + (2 if stem == "imagenet" else 0)
  cell_index = [0]

  def keep(cell_num):
    if drop_path_keep_prob >= 1.0:
      return 1.0
    return 1.0 - (cell_num + 1) / float(total_cells) * (1.0 - drop_path_keep_prob)

  # every entry: (value id, depth, spatial size)
  def run_cell(tag, spec, net, prev, scaling, stride):
    cell_filters = int(filters * scaling)
    prev = prev if prev is not None else net
    cell, depth = _cell(tag, spec, net[1], prev[1], prev[2] != net[2], cell_filters, stride, keep(cell_index[0]))
    cell_index[0] += 1
    return add(cell, (net[0], prev[0])), depth, -(-net[2] // stride)

  normal = _NASNET_NORMAL if family == "nasnet" else _PNASNET_CELL
  reduction = _NASNET_REDUCTION if family == "nasnet" else _PNASNET_CELL
  if stem == "imagenet":
    stem_filters = int(32 * stem_multiplier)
    size = (image_size - 3) // 2 + 1
    value = add(Sequential(name + "/stem", [Conv2d(name + "/conv0", 3, stem_filters, 3, stride=2, padding="VALID", bias=False), _bn(name + "/conv0_bn", stem_filters)]), (0,))
    outputs = [None, (value, stem_filters, size)]
    scaling = 1.0 / (rate ** 2)
    for n in range(2):
      outputs.append(run_cell("%s/cell_stem_%d" % (name, n), reduction, outputs[-1], outputs[-2], scaling, 2))
      scaling *= rate
  else:
    stem_filters = int(filters * stem_multiplier)
    value = add(Sequential(name + "/stem", [Conv2d(name + "/l1_stem_3x3", 3, stem_filters, 3, padding="SAME", bias=False), _bn(name + "/l1_stem_bn", stem_filters)]), (0,))
    outputs = [None, (value, stem_filters, image_size)]
  scaling = 1.0
  aux_at = reductions[1] - 1
  for n in range(num_cells):
    if family == "nasnet":
      prev = outputs[-2]
      if n in reductions:
        scaling *= rate
        outputs.append(run_cell("%s/reduction_cell_%d" % (name, reductions.index(n)), reduction, outputs[-1], outputs[-2], scaling, 2))
      if not skip_reduction_input:
        prev = outputs[-2]
      outputs.append(run_cell("%s/cell_%d" % (name, n), normal, outputs[-1], prev, scaling, 1))
    else:
      is_reduction = n in reductions
      if is_reduction:
        scaling *= rate
      prev = outputs[-2] if (skip_reduction_input or not is_reduction) else None
      outputs.append(run_cell("%s/cell_%d" % (name, n), normal, outputs[-1], prev, scaling, 2 if is_reduction else 1))
    if n == aux_at:
      value, depth, size = outputs[-1]
      tapped = add(AuxHead(name + "/aux_%d" % n, _aux_head(name + "/aux_%d" % n, depth, num_classes, size)), (value,))
      outputs[-1] = (tapped, depth, size)
  value, depth, size = outputs[-1]
  layers = [ReLU(name + "/final_relu"), GlobalAvgPool(name + "/final_pool"), Flatten(name + "/flatten")]
  if dense_keep_prob < 1.0:
    layers.append(Dropout(name + "/dropout", dense_keep_prob))
  layers.append(Dense(name + "/FC", depth, num_classes))
  add(Sequential(name + "/head", layers), (value,))
  return Model(name, Graph(name, nodes), (3, image_size, image_size), num_classes)


def nasnet_cifar(num_classes=10, name="nasnet_cifar"):
  return _build(name, "nasnet", num_classes, 32, "cifar", num_cells=18, filters=32, stem_multiplier=3.0, drop_path_keep_prob=0.6, dense_keep_prob=1.0, skip_reduction_input=False)


def nasnet_mobile(num_classes=1001, name="nasnet_mobile"):
  return _build(name, "nasnet", num_classes, 224, "imagenet", num_cells=12, filters=44, stem_multiplier=1.0, drop_path_keep_prob=1.0, dense_keep_prob=0.5, skip_reduction_input=False)


def nasnet_large(num_classes=1001, name="nasnet_large"):
  return _build(name, "nasnet", num_classes, 331, "imagenet", num_cells=18, filters=168, stem_multiplier=3.0, drop_path_keep_prob=0.7, dense_keep_prob=0.5, skip_reduction_input=True)


def pnasnet_large(num_classes=1001, name="pnasnet_large"):
  return _build(name, "pnasnet", num_classes, 331, "imagenet", num_cells=12, filters=216, stem_multiplier=3.0, drop_path_keep_prob=0.6, dense_keep_prob=0.5, skip_reduction_input=True)


def pnasnet_mobile(num_classes=1001, name="pnasnet_mobile"):
  return _build(name, "pnasnet", num_classes, 224, "imagenet", num_cells=9, filters=54, stem_multiplier=1.0, drop_path_keep_prob=1.0, dense_keep_prob=0.5, skip_reduction_input=True)

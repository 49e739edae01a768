"""slim Inception families: `inception_v1`, `inception_v2`, `inception_v3`, `inception_v4`, `inception_resnet_v2`
(names from the reference's `external/slim/nets/nets_factory.py:39-72`; the architectures themselves live in the
tensorflow/models checkout the reference expects next to it).

Common conventions of slim's `inception_arg_scope`: every convolution is conv (no bias) + batch-norm (decay 0.9997,
epsilon 0.001, no gamma) + ReLU; weights from the variance-scaling initialiser. The mixed blocks are `Branches`
(parallel towers concatenated along channels); the 1x7 / 7x1 / 1x3 / 3x1 factorised kernels use rectangular `Conv2d`s.
Auxiliary classifiers (v3, v4, Inception-ResNet-v2) are `AuxHead` taps whose loss enters with weight 0.4, as the
reference's slim experiment does (`experiments/slims.py:122-125`).
"""

from .core import (AuxHead, AvgPool, BatchNorm, Branches, Conv2d, Dense, DepthwiseConv2d, Dropout, Flatten, GlobalAvgPool, Identity, MaxPool, Model, Residual, Scale,
                   Sequential)


def _conv(name, cin, cout, k, stride=1, padding="SAME", relu=True, **kwargs):
  return [Conv2d(name, cin, cout, k, stride=stride, padding=padding, bias=False, **kwargs), BatchNorm(name + "/BatchNorm", cout, relu=relu, decay=0.9997, epsilon=0.001, scale=False)]


def _tower(name, cin, spec):
  """`spec`: [(suffix, cout, k)] or [(suffix, cout, k, stride, padding)] chained convolutions; returns (Sequential, cout)."""
  layers = []
  for item in spec:
    suffix, cout, k = item[:3]
    stride, padding = (item[3], item[4]) if len(item) > 3 else (1, "SAME")
    layers += _conv(name + "/" + suffix, cin, cout, k, stride, padding)
    cin = cout
  return Sequential(name, layers), cin


def _pool_tower(name, cin, cout, kind="avg"):
  pool = AvgPool(name + "/AvgPool_0a_3x3", 3, 1, "SAME") if kind == "avg" else MaxPool(name + "/MaxPool_0a_3x3", 3, 1, "SAME")
  return Sequential(name, [pool] + _conv(name + "/Conv2d_0b_1x1", cin, cout, 1)), cout


def _mixed(name, towers):
  """`towers`: [(module, cout)] -> (Branches, total depth)."""
  return Branches(name, [module for module, _ in towers]), sum(cout for _, cout in towers)


def _valid(size, k, stride=1):
  """Output size of a VALID k-wide window."""
  return (size - k) // stride + 1


def _logits_conv(name, cin, num_classes):
  return Conv2d(name, cin, num_classes, 1, padding="SAME", bias=True, init="truncated_normal", init_std=0.09)


# ---------------------------------------------------------------------------- #
# Inception v1 (GoogLeNet) — 224 x 224

def inception_v1(num_classes=1001, name="inception_v1"):
  s = "InceptionV1"
  layers = _conv(s + "/Conv2d_1a_7x7", 3, 64, 7, 2) + [MaxPool(s + "/MaxPool_2a_3x3", 3, 2, "SAME")]
  layers += _conv(s + "/Conv2d_2b_1x1", 64, 64, 1) + _conv(s + "/Conv2d_2c_3x3", 64, 192, 3) + [MaxPool(s + "/MaxPool_3a_3x3", 3, 2, "SAME")]
  cin = 192

  def block(tag, b0, b1a, b1b, b2a, b2b, b3):
    nonlocal cin
    base = s + "/" + tag
    mixed, cout = _mixed(base, [
      _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", b0, 1)]),
      _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", b1a, 1), ("Conv2d_0b_3x3", b1b, 3)]),
      _tower(base + "/Branch_2", cin, [("Conv2d_0a_1x1", b2a, 1), ("Conv2d_0b_3x3", b2b, 3)]),
      _pool_tower(base + "/Branch_3", cin, b3, "max")])
    layers.append(mixed)
    cin = cout

  block("Mixed_3b", 64, 96, 128, 16, 32, 32)
  block("Mixed_3c", 128, 128, 192, 32, 96, 64)
  layers.append(MaxPool(s + "/MaxPool_4a_3x3", 3, 2, "SAME"))
  block("Mixed_4b", 192, 96, 208, 16, 48, 64)
  block("Mixed_4c", 160, 112, 224, 24, 64, 64)
  block("Mixed_4d", 128, 128, 256, 24, 64, 64)
  block("Mixed_4e", 112, 144, 288, 32, 64, 64)
  block("Mixed_4f", 256, 160, 320, 32, 128, 128)
  layers.append(MaxPool(s + "/MaxPool_5a_2x2", 2, 2, "SAME"))
  block("Mixed_5b", 256, 160, 320, 32, 128, 128)
  block("Mixed_5c", 384, 192, 384, 48, 128, 128)
  layers += [GlobalAvgPool(s + "/Logits/AvgPool_0a_7x7"), Dropout(s + "/Logits/Dropout_0b", 0.8), _logits_conv(s + "/Logits/Conv2d_0c_1x1", cin, num_classes)]
  return Model(name, Sequential(name, layers), (3, 224, 224), num_classes)


# ---------------------------------------------------------------------------- #
# Inception v2 (BN-Inception) — 224 x 224

def inception_v2(num_classes=1001, name="inception_v2"):
  s = "InceptionV2"
  # separable 7x7 stem: depthwise with 8 filters per input channel, then the 1x1 to 64 (+ BN + ReLU)
  layers = [DepthwiseConv2d(s + "/Conv2d_1a_7x7", 3, 7, 2, multiplier=8, init_std=1.0)] + _conv(s + "/Conv2d_1a_7x7/pointwise", 24, 64, 1)
  layers += [MaxPool(s + "/MaxPool_2a_3x3", 3, 2, "SAME")] + _conv(s + "/Conv2d_2b_1x1", 64, 64, 1) + _conv(s + "/Conv2d_2c_3x3", 64, 192, 3)
  layers.append(MaxPool(s + "/MaxPool_3a_3x3", 3, 2, "SAME"))
  cin = 192

  def block(tag, b0, b1a, b1b, b2a, b2b, b3, pool="avg"):
    nonlocal cin
    base = s + "/" + tag
    mixed, cout = _mixed(base, [
      _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", b0, 1)]),
      _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", b1a, 1), ("Conv2d_0b_3x3", b1b, 3)]),
      _tower(base + "/Branch_2", cin, [("Conv2d_0a_1x1", b2a, 1), ("Conv2d_0b_3x3", b2b, 3), ("Conv2d_0c_3x3", b2b, 3)]),
      _pool_tower(base + "/Branch_3", cin, b3, pool)])
    layers.append(mixed)
    cin = cout

  def reduction(tag, b0a, b0b, b1a, b1b):
    nonlocal cin
    base = s + "/" + tag
    mixed, cout = _mixed(base, [
      _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", b0a, 1), ("Conv2d_1a_3x3", b0b, 3, 2, "SAME")]),
      _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", b1a, 1), ("Conv2d_0b_3x3", b1b, 3), ("Conv2d_1a_3x3", b1b, 3, 2, "SAME")]),
      (MaxPool(base + "/Branch_2/MaxPool_1a_3x3", 3, 2, "SAME"), cin)])
    layers.append(mixed)
    cin = cout

  block("Mixed_3b", 64, 64, 64, 64, 96, 32)
  block("Mixed_3c", 64, 64, 96, 64, 96, 64)
  reduction("Mixed_4a", 128, 160, 64, 96)
  block("Mixed_4b", 224, 64, 96, 96, 128, 128)
  block("Mixed_4c", 192, 96, 128, 96, 128, 128)
  block("Mixed_4d", 160, 128, 160, 128, 160, 96)
  block("Mixed_4e", 96, 128, 192, 160, 192, 96)
  reduction("Mixed_5a", 128, 192, 192, 256)
  block("Mixed_5b", 352, 192, 320, 160, 224, 128)
  block("Mixed_5c", 352, 192, 320, 192, 224, 128, pool="max")
  layers += [GlobalAvgPool(s + "/Logits/AvgPool_1a_7x7"), Dropout(s + "/Logits/Dropout_1b", 0.8), _logits_conv(s + "/Logits/Conv2d_1c_1x1", cin, num_classes)]
  return Model(name, Sequential(name, layers), (3, 224, 224), num_classes)


# ---------------------------------------------------------------------------- #
# Inception v3 — 299 x 299

def _aux_conv_head(base, cin, num_classes, spatial):
  """avg-pool 5x5/3 -> 1x1 conv 128 -> `spatial` x `spatial` VALID conv 768 -> 1x1 logits (v3 flavour)."""
  layers = [AvgPool(base + "/AvgPool_1a_5x5", 5, 3, "VALID")] + _conv(base + "/Conv2d_1b_1x1", cin, 128, 1)
  layers += _conv(base + "/Conv2d_2a_%dx%d" % (spatial, spatial), 128, 768, spatial, 1, "VALID", init="truncated_normal", init_std=0.01)
  layers.append(Conv2d(base + "/Conv2d_2b_1x1", 768, num_classes, 1, padding="SAME", bias=True, init="truncated_normal", init_std=0.001))
  return Sequential(base, layers)


def inception_v3(num_classes=1001, name="inception_v3", image_size=299):
  s = "InceptionV3"
  layers = (_conv(s + "/Conv2d_1a_3x3", 3, 32, 3, 2, "VALID") + _conv(s + "/Conv2d_2a_3x3", 32, 32, 3, 1, "VALID") + _conv(s + "/Conv2d_2b_3x3", 32, 64, 3)
            + [MaxPool(s + "/MaxPool_3a_3x3", 3, 2, "VALID")] + _conv(s + "/Conv2d_3b_1x1", 64, 80, 1, 1, "VALID") + _conv(s + "/Conv2d_4a_3x3", 80, 192, 3, 1, "VALID")
            + [MaxPool(s + "/MaxPool_5a_3x3", 3, 2, "VALID")])
  cin = 192

  def push(mixed_and_depth):
    nonlocal cin
    layers.append(mixed_and_depth[0])
    cin = mixed_and_depth[1]

  for tag, pool in (("Mixed_5b", 32), ("Mixed_5c", 64), ("Mixed_5d", 64)):
    base = s + "/" + tag
    push(_mixed(base, [
      _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", 64, 1)]),
      _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 48, 1), ("Conv2d_0b_5x5", 64, 5)]),
      _tower(base + "/Branch_2", cin, [("Conv2d_0a_1x1", 64, 1), ("Conv2d_0b_3x3", 96, 3), ("Conv2d_0c_3x3", 96, 3)]),
      _pool_tower(base + "/Branch_3", cin, pool)]))
  base = s + "/Mixed_6a"
  push(_mixed(base, [
    _tower(base + "/Branch_0", cin, [("Conv2d_1a_1x1", 384, 3, 2, "VALID")]),
    _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 64, 1), ("Conv2d_0b_3x3", 96, 3), ("Conv2d_1a_1x1", 96, 3, 2, "VALID")]),
    (MaxPool(base + "/Branch_2/MaxPool_1a_3x3", 3, 2, "VALID"), cin)]))
  for tag, mid in (("Mixed_6b", 128), ("Mixed_6c", 160), ("Mixed_6d", 160), ("Mixed_6e", 192)):
    base = s + "/" + tag
    push(_mixed(base, [
      _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", 192, 1)]),
      _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", mid, 1), ("Conv2d_0b_1x7", mid, (1, 7)), ("Conv2d_0c_7x1", 192, (7, 1))]),
      _tower(base + "/Branch_2", cin, [("Conv2d_0a_1x1", mid, 1), ("Conv2d_0b_7x1", mid, (7, 1)), ("Conv2d_0c_1x7", mid, (1, 7)), ("Conv2d_0d_7x1", mid, (7, 1)),
                                       ("Conv2d_0e_1x7", 192, (1, 7))]),
      _pool_tower(base + "/Branch_3", cin, 192)]))
  size = _valid(_valid(image_size, 3, 2), 3)                       # 147 at 299
  size = _valid(_valid(_valid(size, 3, 2), 3), 3, 2)               # 73 -> 71 -> 35
  size = _valid(size, 3, 2)                                        # 17
  layers.append(AuxHead(s + "/AuxLogits", _aux_conv_head(s + "/AuxLogits", cin, num_classes, _valid(size, 5, 3))))
  base = s + "/Mixed_7a"
  push(_mixed(base, [
    _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", 192, 1), ("Conv2d_1a_3x3", 320, 3, 2, "VALID")]),
    _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 192, 1), ("Conv2d_0b_1x7", 192, (1, 7)), ("Conv2d_0c_7x1", 192, (7, 1)), ("Conv2d_1a_3x3", 192, 3, 2, "VALID")]),
    (MaxPool(base + "/Branch_2/MaxPool_1a_3x3", 3, 2, "VALID"), cin)]))
  for tag in ("Mixed_7b", "Mixed_7c"):
    base = s + "/" + tag
    head1, _ = _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 384, 1)])
    split1, _ = _mixed(base + "/Branch_1/split", [_tower(base + "/Branch_1/a", 384, [("Conv2d_0b_1x3", 384, (1, 3))]), _tower(base + "/Branch_1/b", 384, [("Conv2d_0b_3x1", 384, (3, 1))])])
    head2, _ = _tower(base + "/Branch_2", cin, [("Conv2d_0a_1x1", 448, 1), ("Conv2d_0b_3x3", 384, 3)])
    split2, _ = _mixed(base + "/Branch_2/split", [_tower(base + "/Branch_2/a", 384, [("Conv2d_0c_1x3", 384, (1, 3))]), _tower(base + "/Branch_2/b", 384, [("Conv2d_0d_3x1", 384, (3, 1))])])
    push(_mixed(base, [
      _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", 320, 1)]),
      (Sequential(base + "/Branch_1/seq", [head1, split1]), 768),
      (Sequential(base + "/Branch_2/seq", [head2, split2]), 768),
      _pool_tower(base + "/Branch_3", cin, 192)]))
  layers += [GlobalAvgPool(s + "/Logits/AvgPool_1a_8x8"), Dropout(s + "/Logits/Dropout_1b", 0.8), _logits_conv(s + "/Logits/Conv2d_1c_1x1", cin, num_classes)]
  return Model(name, Sequential(name, layers), (3, image_size, image_size), num_classes)


# ---------------------------------------------------------------------------- #
# Inception v4 — 299 x 299

def _aux_fc_head(base, cin, num_classes, spatial):
  """avg-pool 5x5/3 -> 1x1 conv 128 -> full-map VALID conv 768 -> flatten -> fully connected (v4 / Inception-ResNet flavour)."""
  layers = [AvgPool(base + "/AvgPool_1a_5x5", 5, 3, "VALID")] + _conv(base + "/Conv2d_1b_1x1", cin, 128, 1)
  layers += _conv(base + "/Conv2d_2a", 128, 768, spatial, 1, "VALID") + [Flatten(base + "/flatten"), Dense(base + "/Aux_logits", 768, num_classes)]
  return Sequential(base, layers)


def inception_v4(num_classes=1001, name="inception_v4", image_size=299):
  s = "InceptionV4"
  layers = _conv(s + "/Conv2d_1a_3x3", 3, 32, 3, 2, "VALID") + _conv(s + "/Conv2d_2a_3x3", 32, 32, 3, 1, "VALID") + _conv(s + "/Conv2d_2b_3x3", 32, 64, 3)
  cin = 64

  def push(mixed_and_depth):
    nonlocal cin
    layers.append(mixed_and_depth[0])
    cin = mixed_and_depth[1]

  base = s + "/Mixed_3a"
  push(_mixed(base, [(MaxPool(base + "/Branch_0/MaxPool_0a_3x3", 3, 2, "VALID"), cin), _tower(base + "/Branch_1", cin, [("Conv2d_0a_3x3", 96, 3, 2, "VALID")])]))
  base = s + "/Mixed_4a"
  push(_mixed(base, [
    _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", 64, 1), ("Conv2d_1a_3x3", 96, 3, 1, "VALID")]),
    _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 64, 1), ("Conv2d_0b_1x7", 64, (1, 7)), ("Conv2d_0c_7x1", 64, (7, 1)), ("Conv2d_1a_3x3", 96, 3, 1, "VALID")])]))
  base = s + "/Mixed_5a"
  push(_mixed(base, [_tower(base + "/Branch_0", cin, [("Conv2d_1a_3x3", 192, 3, 2, "VALID")]), (MaxPool(base + "/Branch_1/MaxPool_1a_3x3", 3, 2, "VALID"), cin)]))
  for tag in ("Mixed_5b", "Mixed_5c", "Mixed_5d", "Mixed_5e"):  # inception-A
    base = s + "/" + tag
    push(_mixed(base, [
      _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", 96, 1)]),
      _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 64, 1), ("Conv2d_0b_3x3", 96, 3)]),
      _tower(base + "/Branch_2", cin, [("Conv2d_0a_1x1", 64, 1), ("Conv2d_0b_3x3", 96, 3), ("Conv2d_0c_3x3", 96, 3)]),
      _pool_tower(base + "/Branch_3", cin, 96)]))
  base = s + "/Mixed_6a"  # reduction-A
  push(_mixed(base, [
    _tower(base + "/Branch_0", cin, [("Conv2d_1a_3x3", 384, 3, 2, "VALID")]),
    _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 192, 1), ("Conv2d_0b_3x3", 224, 3), ("Conv2d_1a_3x3", 256, 3, 2, "VALID")]),
    (MaxPool(base + "/Branch_2/MaxPool_1a_3x3", 3, 2, "VALID"), cin)]))
  for tag in ("Mixed_6b", "Mixed_6c", "Mixed_6d", "Mixed_6e", "Mixed_6f", "Mixed_6g", "Mixed_6h"):  # inception-B
    base = s + "/" + tag
    push(_mixed(base, [
      _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", 384, 1)]),
      _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 192, 1), ("Conv2d_0b_1x7", 224, (1, 7)), ("Conv2d_0c_7x1", 256, (7, 1))]),
      _tower(base + "/Branch_2", cin, [("Conv2d_0a_1x1", 192, 1), ("Conv2d_0b_7x1", 192, (7, 1)), ("Conv2d_0c_1x7", 224, (1, 7)), ("Conv2d_0d_7x1", 224, (7, 1)),
                                       ("Conv2d_0e_1x7", 256, (1, 7))]),
      _pool_tower(base + "/Branch_3", cin, 128)]))
  size = _valid(_valid(image_size, 3, 2), 3)                       # 147 at 299
  size = _valid(_valid(_valid(size, 3, 2), 3), 3, 2)               # 73 -> 71 -> 35
  size = _valid(size, 3, 2)                                        # 17
  layers.append(AuxHead(s + "/AuxLogits", _aux_fc_head(s + "/AuxLogits", cin, num_classes, _valid(size, 5, 3))))
  base = s + "/Mixed_7a"  # reduction-B
  push(_mixed(base, [
    _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", 192, 1), ("Conv2d_1a_3x3", 192, 3, 2, "VALID")]),
    _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 256, 1), ("Conv2d_0b_1x7", 256, (1, 7)), ("Conv2d_0c_7x1", 320, (7, 1)), ("Conv2d_1a_3x3", 320, 3, 2, "VALID")]),
    (MaxPool(base + "/Branch_2/MaxPool_1a_3x3", 3, 2, "VALID"), cin)]))
  for tag in ("Mixed_7b", "Mixed_7c", "Mixed_7d"):  # inception-C
    base = s + "/" + tag
    head1, _ = _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 384, 1)])
    split1, _ = _mixed(base + "/Branch_1/split", [_tower(base + "/Branch_1/a", 384, [("Conv2d_0b_1x3", 256, (1, 3))]), _tower(base + "/Branch_1/b", 384, [("Conv2d_0c_3x1", 256, (3, 1))])])
    head2, _ = _tower(base + "/Branch_2", cin, [("Conv2d_0a_1x1", 384, 1), ("Conv2d_0b_3x1", 448, (3, 1)), ("Conv2d_0c_1x3", 512, (1, 3))])
    split2, _ = _mixed(base + "/Branch_2/split", [_tower(base + "/Branch_2/a", 512, [("Conv2d_0d_1x3", 256, (1, 3))]), _tower(base + "/Branch_2/b", 512, [("Conv2d_0e_3x1", 256, (3, 1))])])
    push(_mixed(base, [
      _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", 256, 1)]),
      (Sequential(base + "/Branch_1/seq", [head1, split1]), 512),
      (Sequential(base + "/Branch_2/seq", [head2, split2]), 512),
      _pool_tower(base + "/Branch_3", cin, 256)]))
  layers += [GlobalAvgPool(s + "/Logits/AvgPool_1a"), Dropout(s + "/Logits/Dropout_1b", 0.8), Flatten(s + "/Logits/PreLogitsFlatten"), Dense(s + "/Logits/Logits", cin, num_classes)]
  return Model(name, Sequential(name, layers), (3, image_size, image_size), num_classes)


# ---------------------------------------------------------------------------- #
# Inception-ResNet-v2 — 299 x 299

def _res_block(base, cin, towers, scale, relu=True):
  """relu(x + scale * conv1x1(concat(towers(x)))) — the up-projection has a bias and neither batch-norm nor activation."""
  mixed, depth = _mixed(base + "/mixed", towers)
  up = Conv2d(base + "/Conv2d_1x1", depth, cin, 1, padding="SAME", bias=True)
  return Residual(base, Identity(base + "/shortcut"), Sequential(base + "/residual", [mixed, up, Scale(base + "/scale", scale)]), relu=relu)


def inception_resnet_v2(num_classes=1001, name="inception_resnet_v2", image_size=299):
  s = "InceptionResnetV2"
  layers = (_conv(s + "/Conv2d_1a_3x3", 3, 32, 3, 2, "VALID") + _conv(s + "/Conv2d_2a_3x3", 32, 32, 3, 1, "VALID") + _conv(s + "/Conv2d_2b_3x3", 32, 64, 3)
            + [MaxPool(s + "/MaxPool_3a_3x3", 3, 2, "VALID")] + _conv(s + "/Conv2d_3b_1x1", 64, 80, 1, 1, "VALID") + _conv(s + "/Conv2d_4a_3x3", 80, 192, 3, 1, "VALID")
            + [MaxPool(s + "/MaxPool_5a_3x3", 3, 2, "VALID")])
  cin = 192
  base = s + "/Mixed_5b"
  mixed, cin = _mixed(base, [
    _tower(base + "/Branch_0", cin, [("Conv2d_1x1", 96, 1)]),
    _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 48, 1), ("Conv2d_0b_5x5", 64, 5)]),
    _tower(base + "/Branch_2", cin, [("Conv2d_0a_1x1", 64, 1), ("Conv2d_0b_3x3", 96, 3), ("Conv2d_0c_3x3", 96, 3)]),
    _pool_tower(base + "/Branch_3", cin, 64)])
  layers.append(mixed)
  for i in range(1, 11):  # 10 x block35
    base = "%s/Repeat/block35_%d" % (s, i)
    layers.append(_res_block(base, cin, [
      _tower(base + "/Branch_0", cin, [("Conv2d_1x1", 32, 1)]),
      _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 32, 1), ("Conv2d_0b_3x3", 32, 3)]),
      _tower(base + "/Branch_2", cin, [("Conv2d_0a_1x1", 32, 1), ("Conv2d_0b_3x3", 48, 3), ("Conv2d_0c_3x3", 64, 3)])], 0.17))
  base = s + "/Mixed_6a"
  mixed, cin = _mixed(base, [
    _tower(base + "/Branch_0", cin, [("Conv2d_1a_3x3", 384, 3, 2, "VALID")]),
    _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 256, 1), ("Conv2d_0b_3x3", 256, 3), ("Conv2d_1a_3x3", 384, 3, 2, "VALID")]),
    (MaxPool(base + "/Branch_2/MaxPool_1a_3x3", 3, 2, "VALID"), cin)])
  layers.append(mixed)
  for i in range(1, 21):  # 20 x block17
    base = "%s/Repeat_1/block17_%d" % (s, i)
    layers.append(_res_block(base, cin, [
      _tower(base + "/Branch_0", cin, [("Conv2d_1x1", 192, 1)]),
      _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 128, 1), ("Conv2d_0b_1x7", 160, (1, 7)), ("Conv2d_0c_7x1", 192, (7, 1))])], 0.10))
  size = _valid(_valid(image_size, 3, 2), 3)                       # 147 at 299
  size = _valid(_valid(_valid(size, 3, 2), 3), 3, 2)               # 73 -> 71 -> 35
  size = _valid(size, 3, 2)                                        # 17
  layers.append(AuxHead(s + "/AuxLogits", _aux_fc_head(s + "/AuxLogits", cin, num_classes, _valid(size, 5, 3))))
  base = s + "/Mixed_7a"
  mixed, cin = _mixed(base, [
    _tower(base + "/Branch_0", cin, [("Conv2d_0a_1x1", 256, 1), ("Conv2d_1a_3x3", 384, 3, 2, "VALID")]),
    _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 256, 1), ("Conv2d_1a_3x3", 288, 3, 2, "VALID")]),
    _tower(base + "/Branch_2", cin, [("Conv2d_0a_1x1", 256, 1), ("Conv2d_0b_3x3", 288, 3), ("Conv2d_1a_3x3", 320, 3, 2, "VALID")]),
    (MaxPool(base + "/Branch_3/MaxPool_1a_3x3", 3, 2, "VALID"), cin)])
  layers.append(mixed)
  for i in range(1, 11):  # 9 x block8 + one without activation
    base = ("%s/Repeat_2/block8_%d" % (s, i)) if i < 10 else s + "/Block8"
    layers.append(_res_block(base, cin, [
      _tower(base + "/Branch_0", cin, [("Conv2d_1x1", 192, 1)]),
      _tower(base + "/Branch_1", cin, [("Conv2d_0a_1x1", 192, 1), ("Conv2d_0b_1x3", 224, (1, 3)), ("Conv2d_0c_3x1", 256, (3, 1))])], 0.20 if i < 10 else 1.0, relu=(i < 10)))
  layers += _conv(s + "/Conv2d_7b_1x1", cin, 1536, 1)
  layers += [GlobalAvgPool(s + "/Logits/AvgPool_1a_8x8"), Flatten(s + "/Logits/PreLogitsFlatten"), Dropout(s + "/Logits/Dropout", 0.8), Dense(s + "/Logits/Logits", 1536, num_classes)]
  return Model(name, Sequential(name, layers), (3, image_size, image_size), num_classes)

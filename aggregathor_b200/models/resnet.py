"""slim ResNet v1 / v2 families (reference: `external/slim/nets/resnet_v1.py`, `nets_factory.py:51-59`).

ResNet v1 as slim defines it: 7x7/2 stem (`conv2d_same`: explicit padding then VALID) + BN + ReLU,
3x3/2 max-pool (VALID... slim's `max_pool2d` default padding is VALID? No: `resnet_utils.resnet_arg_scope`
sets `padding='SAME'` for `slim.max_pool2d`), four blocks of *bottleneck* units where the stride sits on
the 3x3 conv of the **last** unit of each block (`resnet_v1.py:258-279`), projection shortcut (1x1 conv + BN,
no activation) whenever the depth changes and `subsample` (1x1 max-pool) otherwise, global mean pool,
1x1 conv logits with bias and no normaliser. BN after every conv (decay 0.997, eps 1e-5, scale).
`resnet_v1_18` is the reference's own bottleneck-based [2,2,2,2] variant (`resnet_v1.py:281-301`), not the
canonical basic-block ResNet-18. `resnet_v1_50` has 25 557 032 trainable parameters at 1000 classes.
"""

from .core import BatchNorm, Conv2d, GlobalAvgPool, MaxPool, Model, Module, Residual, Sequential, Subsample


def _conv_bn(name, cin, cout, k, stride=1, relu=True):
  padding = "SAME" if stride == 1 else "explicit"  # conv2d_same: stride > 1 => explicit pad + VALID
  return [Conv2d(name, cin, cout, k, stride=stride, padding=padding, bias=False),
          BatchNorm(name + "/BatchNorm", cout, relu=relu)]


def _bottleneck_v1(name, cin, depth, depth_bottleneck, stride):
  if depth == cin:
    shortcut = Subsample(name + "/shortcut", stride)
  else:
    shortcut = Sequential(name + "/shortcut", _conv_bn(name + "/shortcut", cin, depth, 1, stride=stride, relu=False))
  residual = Sequential(name + "/residual",
                        _conv_bn(name + "/conv1", cin, depth_bottleneck, 1)
                        + _conv_bn(name + "/conv2", depth_bottleneck, depth_bottleneck, 3, stride=stride)
                        + _conv_bn(name + "/conv3", depth_bottleneck, depth, 1, relu=False))
  return Residual(name, shortcut, residual, relu=True)


def resnet_v1(name, units_per_block, num_classes=1000, image_size=224):
  """`units_per_block`: e.g. [3, 4, 6, 3]; block strides are (2, 2, 2, 1) carried by each block's last unit."""
  scope = name
  layers = _conv_bn(scope + "/conv1", 3, 64, 7, stride=2) + [MaxPool(scope + "/pool1", 3, 2, "SAME")]
  cin = 64
  for b, (base, units, stride) in enumerate(zip((64, 128, 256, 512), units_per_block, (2, 2, 2, 1))):
    for u in range(units):
      unit_name = "%s/block%d/unit_%d/bottleneck_v1" % (scope, b + 1, u + 1)
      layers.append(_bottleneck_v1(unit_name, cin, base * 4, base, stride if u == units - 1 else 1))
      cin = base * 4
  layers.append(GlobalAvgPool(scope + "/pool5"))
  layers.append(Conv2d(scope + "/logits", cin, num_classes, 1, padding="SAME", bias=True, init="variance_scaling"))
  return Model(name, Sequential(scope, layers), (3, image_size, image_size), num_classes)


class _PreactUnit(Module):
  """ResNet v2 bottleneck: BN+ReLU pre-activation shared by the projection shortcut and the residual branch."""

  def __init__(self, name, cin, depth, depth_bottleneck, stride):
    super().__init__(name)
    self.preact = BatchNorm(name + "/preact", cin, relu=True)
    if depth == cin:
      self.shortcut, self.shortcut_on_preact = Subsample(name + "/shortcut", stride), False
    else:
      self.shortcut, self.shortcut_on_preact = Conv2d(name + "/shortcut", cin, depth, 1, stride=stride, padding="SAME" if stride == 1 else "explicit", bias=True), True
    self.residual = Sequential(name + "/residual",
                               _conv_bn(name + "/conv1", cin, depth_bottleneck, 1)
                               + _conv_bn(name + "/conv2", depth_bottleneck, depth_bottleneck, 3, stride=stride)
                               + [Conv2d(name + "/conv3", depth_bottleneck, depth, 1, padding="SAME", bias=True)])

  def children(self):
    return (self.preact, self.shortcut, self.residual)

  def declare(self, layout, states):
    for child in self.children():
      child.declare(layout, states)

  def initialize(self, master, states, generator):
    for child in self.children():
      child.initialize(master, states, generator)

  def forward(self, x, ctx):
    pre = self.preact.forward(x, ctx)
    a = self.shortcut.forward(pre if self.shortcut_on_preact else x, ctx)
    b = self.residual.forward(pre, ctx)
    return a + b

  def backward(self, dy, ctx):
    dpre = self.residual.backward(dy, ctx)
    da = self.shortcut.backward(dy, ctx)
    if self.shortcut_on_preact:
      return self.preact.backward(dpre + da, ctx)
    return self.preact.backward(dpre, ctx) + da


def resnet_v2(name, units_per_block, num_classes=1001, image_size=224):
  scope = name
  layers = [Conv2d(scope + "/conv1", 3, 64, 7, stride=2, padding="explicit", bias=True), MaxPool(scope + "/pool1", 3, 2, "SAME")]
  cin = 64
  for b, (base, units, stride) in enumerate(zip((64, 128, 256, 512), units_per_block, (2, 2, 2, 1))):
    for u in range(units):
      unit_name = "%s/block%d/unit_%d/bottleneck_v2" % (scope, b + 1, u + 1)
      layers.append(_PreactUnit(unit_name, cin, base * 4, base, stride if u == units - 1 else 1))
      cin = base * 4
  layers.append(BatchNorm(scope + "/postnorm", cin, relu=True))
  layers.append(GlobalAvgPool(scope + "/pool5"))
  layers.append(Conv2d(scope + "/logits", cin, num_classes, 1, padding="SAME", bias=True))
  return Model(name, Sequential(scope, layers), (3, image_size, image_size), num_classes)

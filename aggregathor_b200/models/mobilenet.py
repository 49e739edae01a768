"""MobileNet v1 / v2 — `mobilenet_v1[_075|_050|_025]`, `mobilenet_v2[_140|_035]` of the reference's slim factory
(`external/slim/nets/nets_factory.py:39-72`; the definitions live in the tensorflow/models checkout it expects).

MobileNet v1 (slim `mobilenet_v1`): 3x3/2 stem conv + 13 depthwise-separable blocks (3x3 depthwise + BN +
ReLU6, 1x1 pointwise + BN + ReLU6), global average pool, dropout, 1x1 conv logits. Depthwise convolutions
run through the torch provider (grouped conv); pointwise 1x1 convolutions are plain GEMMs."""

from .core import BatchNorm, Conv2d, DepthwiseConv2d, Dropout, GlobalAvgPool, Identity, Model, ReLU6, Residual, Sequential


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


def _make_divisible(value, divisor=8, min_value=8):
  new = max(min_value, int(value + divisor / 2) // divisor * divisor)
  return new + divisor if new < 0.9 * value else new


def mobilenet_v2(num_classes=1001, multiplier=1.0, name="mobilenet_v2"):
  """MobileNet v2 (slim `mobilenet_v2.V2_DEF`): 3x3/2 stem, 17 inverted-residual `expanded_conv` blocks (1x1 expansion x6 +
  BN + ReLU6, 3x3 depthwise + BN + ReLU6, linear 1x1 projection + BN, identity shortcut when stride 1 and the depth is kept),
  1x1 conv to 1280 (not shrunk for multipliers < 1), global average pool, dropout 0.8, 1x1 conv logits."""
  s = "MobilenetV2"
  bn = lambda base, c: BatchNorm(base + "/BatchNorm", c, decay=0.997, epsilon=0.001)
  depth = lambda d: _make_divisible(d * multiplier)
  cin = depth(32)
  layers = [Conv2d(s + "/Conv", 3, cin, 3, stride=2, padding="SAME", init="truncated_normal", init_std=0.09), bn(s + "/Conv", cin), ReLU6(s + "/Conv/Relu6")]
  spec = [(1, 16, 1), (6, 24, 2), (6, 24, 1), (6, 32, 2), (6, 32, 1), (6, 32, 1), (6, 64, 2), (6, 64, 1), (6, 64, 1), (6, 64, 1), (6, 96, 1), (6, 96, 1), (6, 96, 1),
          (6, 160, 2), (6, 160, 1), (6, 160, 1), (6, 320, 1)]
  for i, (expansion, cout, stride) in enumerate(spec):
    base = s + ("/expanded_conv" if i == 0 else "/expanded_conv_%d" % i)
    cout = depth(cout)
    inner = cin if expansion == 1 else _make_divisible(cin * expansion)
    block = []
    if inner > cin:
      block += [Conv2d(base + "/expand", cin, inner, 1, padding="SAME", init="truncated_normal", init_std=0.09), bn(base + "/expand", inner), ReLU6(base + "/expand/Relu6")]
    block += [DepthwiseConv2d(base + "/depthwise", inner, 3, stride), bn(base + "/depthwise", inner), ReLU6(base + "/depthwise/Relu6"),
              Conv2d(base + "/project", inner, cout, 1, padding="SAME", init="truncated_normal", init_std=0.09), bn(base + "/project", cout)]
    block = Sequential(base, block)
    layers.append(Residual(base + "/add", Identity(base + "/shortcut"), block, relu=False) if (stride == 1 and cin == cout) else block)
    cin = cout
  last = 1280 if multiplier < 1.0 else depth(1280)
  layers += [Conv2d(s + "/Conv_1", cin, last, 1, padding="SAME", init="truncated_normal", init_std=0.09), bn(s + "/Conv_1", last), ReLU6(s + "/Conv_1/Relu6"),
             GlobalAvgPool(s + "/Logits/AvgPool"), Dropout(s + "/Logits/Dropout", 0.8),
             Conv2d(s + "/Logits/Conv2d_1c_1x1", last, num_classes, 1, padding="SAME", bias=True, init="truncated_normal", init_std=0.09)]
  return Model(name, Sequential(name, layers), (3, 224, 224), num_classes)

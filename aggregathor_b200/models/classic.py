"""Classic slim networks without batch-norm: lenet, cifarnet, alexnet_v2, vgg_a / vgg_16 / vgg_19, overfeat
(names from the reference's `external/slim/nets/nets_factory.py:39-72`). Structures follow the slim
definitions (conv + bias + ReLU, max-pools, dropout 0.5, fully-connected layers expressed as convolutions
in slim are expressed as Dense layers on the flattened NHWC map here; parameter counts are identical)."""

from .core import Conv2d, Dense, Dropout, Flatten, MaxPool, Model, Sequential


def lenet(num_classes=10, name="lenet"):
  layers = [
    Conv2d("LeNet/conv1", 1, 32, 5, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=0.1), MaxPool("LeNet/pool1", 2, 2),
    Conv2d("LeNet/conv2", 32, 64, 5, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=0.1), MaxPool("LeNet/pool2", 2, 2),
    Flatten("LeNet/flatten"), Dense("LeNet/fc3", 7 * 7 * 64, 1024, relu=True, init="truncated_normal", init_std=0.1),
    Dropout("LeNet/dropout3", 0.5), Dense("LeNet/fc4", 1024, num_classes, init="truncated_normal", init_std=0.1)]
  return Model(name, Sequential(name, layers), (1, 28, 28), num_classes)


def cifarnet(num_classes=10, name="cifarnet"):
  layers = [
    Conv2d("CifarNet/conv1", 3, 64, 5, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=5e-2), MaxPool("CifarNet/pool1", 2, 2),
    Conv2d("CifarNet/conv2", 64, 64, 5, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=5e-2, bias_init=0.1), MaxPool("CifarNet/pool2", 2, 2),
    Flatten("CifarNet/flatten"), Dense("CifarNet/fc3", 8 * 8 * 64, 384, relu=True, init="truncated_normal", init_std=0.04, bias_init=0.1),
    Dropout("CifarNet/dropout3", 0.5), Dense("CifarNet/fc4", 384, 192, relu=True, init="truncated_normal", init_std=0.04, bias_init=0.1),
    Dense("CifarNet/logits", 192, num_classes, init="truncated_normal", init_std=1 / 192.0)]
  return Model(name, Sequential(name, layers), (3, 32, 32), num_classes)


def _vgg(name, scope, config, num_classes, image_size=224):
  layers, cin, size = [], 3, image_size
  for b, (reps, cout) in enumerate(config):
    for r in range(reps):
      layers.append(Conv2d("%s/conv%d/conv%d_%d" % (scope, b + 1, b + 1, r + 1), cin, cout, 3, padding="SAME", bias=True, relu=True, init="xavier"))
      cin = cout
    layers.append(MaxPool("%s/pool%d" % (scope, b + 1), 2, 2))
    size //= 2
  layers += [Flatten(scope + "/flatten"), Dense(scope + "/fc6", size * size * cin, 4096, relu=True), Dropout(scope + "/dropout6", 0.5),
             Dense(scope + "/fc7", 4096, 4096, relu=True), Dropout(scope + "/dropout7", 0.5), Dense(scope + "/fc8", 4096, num_classes)]
  return Model(name, Sequential(name, layers), (3, image_size, image_size), num_classes)


def vgg_a(num_classes=1000, name="vgg_a"):
  return _vgg(name, "vgg_a", [(1, 64), (1, 128), (2, 256), (2, 512), (2, 512)], num_classes)


def vgg_16(num_classes=1000, name="vgg_16"):
  return _vgg(name, "vgg_16", [(2, 64), (2, 128), (3, 256), (3, 512), (3, 512)], num_classes)


def vgg_19(num_classes=1000, name="vgg_19"):
  return _vgg(name, "vgg_19", [(2, 64), (2, 128), (4, 256), (4, 512), (4, 512)], num_classes)


def alexnet_v2(num_classes=1000, name="alexnet_v2"):
  s = "alexnet_v2"
  layers = [
    Conv2d(s + "/conv1", 3, 64, 11, stride=4, padding="VALID", bias=True, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1), MaxPool(s + "/pool1", 3, 2),
    Conv2d(s + "/conv2", 64, 192, 5, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1), MaxPool(s + "/pool2", 3, 2),
    Conv2d(s + "/conv3", 192, 384, 3, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1),
    Conv2d(s + "/conv4", 384, 384, 3, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1),
    Conv2d(s + "/conv5", 384, 256, 3, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1), MaxPool(s + "/pool5", 3, 2),
    Flatten(s + "/flatten"), Dense(s + "/fc6", 5 * 5 * 256, 4096, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1), Dropout(s + "/dropout6", 0.5),
    Dense(s + "/fc7", 4096, 4096, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1), Dropout(s + "/dropout7", 0.5),
    Dense(s + "/fc8", 4096, num_classes, init="truncated_normal", init_std=0.005)]
  return Model(name, Sequential(name, layers), (3, 224, 224), num_classes)


def overfeat(num_classes=1000, name="overfeat"):
  s = "overfeat"
  layers = [
    Conv2d(s + "/conv1", 3, 64, 11, stride=4, padding="VALID", bias=True, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1), MaxPool(s + "/pool1", 2, 2),
    Conv2d(s + "/conv2", 64, 256, 5, padding="VALID", bias=True, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1), MaxPool(s + "/pool2", 2, 2),
    Conv2d(s + "/conv3", 256, 512, 3, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1),
    Conv2d(s + "/conv4", 512, 1024, 3, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1),
    Conv2d(s + "/conv5", 1024, 1024, 3, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1), MaxPool(s + "/pool5", 2, 2),
    Flatten(s + "/flatten"), Dense(s + "/fc6", 6 * 6 * 1024, 3072, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1), Dropout(s + "/dropout6", 0.5),
    Dense(s + "/fc7", 3072, 4096, relu=True, init="truncated_normal", init_std=0.005, bias_init=0.1), Dropout(s + "/dropout7", 0.5),
    Dense(s + "/fc8", 4096, num_classes, init="truncated_normal", init_std=0.005)]
  return Model(name, Sequential(name, layers), (3, 231, 231), num_classes)

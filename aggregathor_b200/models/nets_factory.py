"""Name -> network constructor registry (reference: `external/slim/nets/nets_factory.py:39-72,115-153`).

The reference exposes whatever a user-supplied checkout of tensorflow/models `research/slim` contains (33
names) and only ships the ResNet-v1 override adding `resnet_v1_18`. All 33 are built here on the static layer
graph of `models/core.py`: ResNet v1/v2, VGG, AlexNet, OverFeat, LeNet, CifarNet, MobileNet v1/v2,
Inception v1-v4, Inception-ResNet-v2, NASNet-A and PNASNet-5.
"""

from .. import tools
from . import classic, inception, mobilenet, nasnet, resnet, simple

_RESNET_UNITS = {"18": [2, 2, 2, 2], "50": [3, 4, 6, 3], "101": [3, 4, 23, 3], "152": [3, 8, 36, 3], "200": [3, 24, 36, 3]}


def _resnet_v1(depth):
  return lambda num_classes: resnet.resnet_v1("resnet_v1_" + depth, _RESNET_UNITS[depth], num_classes)


def _resnet_v2(depth):
  return lambda num_classes: resnet.resnet_v2("resnet_v2_" + depth, _RESNET_UNITS[depth], num_classes)


networks_map = {
  "alexnet_v2": classic.alexnet_v2, "cifarnet": classic.cifarnet, "overfeat": classic.overfeat,
  "vgg_a": classic.vgg_a, "vgg_16": classic.vgg_16, "vgg_19": classic.vgg_19, "lenet": classic.lenet,
  "resnet_v1_18": _resnet_v1("18"), "resnet_v1_50": _resnet_v1("50"), "resnet_v1_101": _resnet_v1("101"),
  "resnet_v1_152": _resnet_v1("152"), "resnet_v1_200": _resnet_v1("200"),
  "resnet_v2_50": _resnet_v2("50"), "resnet_v2_101": _resnet_v2("101"), "resnet_v2_152": _resnet_v2("152"), "resnet_v2_200": _resnet_v2("200"),
  "mobilenet_v1": lambda num_classes: mobilenet.mobilenet_v1(num_classes, 1.0, "mobilenet_v1"),
  "mobilenet_v1_075": lambda num_classes: mobilenet.mobilenet_v1(num_classes, 0.75, "mobilenet_v1_075"),
  "mobilenet_v1_050": lambda num_classes: mobilenet.mobilenet_v1(num_classes, 0.50, "mobilenet_v1_050"),
  "mobilenet_v1_025": lambda num_classes: mobilenet.mobilenet_v1(num_classes, 0.25, "mobilenet_v1_025")}
networks_map.update({
  "inception_v1": inception.inception_v1, "inception_v2": inception.inception_v2, "inception_v3": inception.inception_v3, "inception_v4": inception.inception_v4,
  "inception_resnet_v2": inception.inception_resnet_v2,
  "mobilenet_v2": lambda num_classes: mobilenet.mobilenet_v2(num_classes, 1.0, "mobilenet_v2"),
  "mobilenet_v2_140": lambda num_classes: mobilenet.mobilenet_v2(num_classes, 1.4, "mobilenet_v2_140"),
  "mobilenet_v2_035": lambda num_classes: mobilenet.mobilenet_v2(num_classes, 0.35, "mobilenet_v2_035"),
  "nasnet_cifar": nasnet.nasnet_cifar, "nasnet_mobile": nasnet.nasnet_mobile, "nasnet_large": nasnet.nasnet_large,
  "pnasnet_large": nasnet.pnasnet_large, "pnasnet_mobile": nasnet.pnasnet_mobile})

# extra names of this framework (not slim): the two hand-written reference experiments' models
extra_networks = {"mlp": lambda num_classes: simple.mlp((784, 100, num_classes)), "cnnet": simple.cnnet}

_image_sizes = {"cifarnet": 32, "lenet": 28, "overfeat": 231, "nasnet_cifar": 32, "inception_v3": 299, "inception_v4": 299,
                "inception_resnet_v2": 299, "nasnet_large": 331, "pnasnet_large": 331}


def default_image_size(name):
  return _image_sizes.get(name, 224)


def get_network(name, num_classes):
  """Build the `Model` called `name` with `num_classes` outputs."""
  table = networks_map if name in networks_map else extra_networks
  if name not in table:
    raise tools.UserException("Name of network unknown " + repr(name))
  return table[name](num_classes)

"""`slim-<model>-<dataset>`: the cross product of the network factory and the known datasets
(reference: `experiments/slims.py:47-196`), e.g. `slim-resnet_v1_50-imagenet`.

Args: `batch-size:32`, `eval-batch-size:1024`, `weight-decay:4e-5` (accepted; as in the reference the
regularisation losses are *not* added to the training loss, `slims.py:122-125`), `label-smoothing:0`,
`labels-offset:0`, `preprocessing:<model default>`, `image-size:<model default>`, `augment:none|flip|crop-flip`. Loss = softmax cross-entropy on
one-hot labels (+ label smoothing). Images are uint8 NHWC on the host; `vgg` preprocessing (mean subtraction) or
`inception` preprocessing (scale to [-1, 1]) happens on the device, fused with the bf16 cast.
"""


from .. import tools
from ..models import nets_factory
from . import _Experiment, register
from ._data import BatchStream, Dataset, known_datasets
from .cnnet import cifarnet_preprocess

_VGG_MEANS = (123.68, 116.78, 103.94)
_PREPROCESSING = {"cifarnet": "cifarnet", "lenet": "lenet"}
for _name in nets_factory.networks_map:
  if _name.startswith(("resnet", "vgg")):
    _PREPROCESSING[_name] = "vgg"
  elif _name not in _PREPROCESSING:
    _PREPROCESSING[_name] = "inception"


class SlimExperiment(_Experiment):
  @staticmethod
  def _make(dataset, model):
    return lambda args: SlimExperiment(dataset, model, args)

  def __init__(self, dataset, model, args):
    self.args = tools.parse_keyval(args if args is not None else [], defaults={
      "batch-size": 32, "eval-batch-size": 1024, "weight-decay": 0.00004, "label-smoothing": 0., "labels-offset": 0,
      "nb-fetcher-threads": 1, "nb-batcher-threads": 1, "image-size": nets_factory.default_image_size(model), "seed": 0, "synthetic-samples": 512,
      "augment": "none"})
    if self.args["augment"] not in ("none", "flip", "crop-flip"):
      raise tools.UserException("augment must be one of none, flip, crop-flip")
    # random augmentation draws from the device generator every step: such a step cannot be replayed from a CUDA graph
    self.stochastic_preprocess = self.args["augment"] != "none"
    if self.args["batch-size"] <= 0:
      raise tools.UserException("Cannot make batches of non-positive size")
    self.dataset_name, self.model_name = dataset, model
    self.preprocessing = self.args.get("preprocessing", _PREPROCESSING.get(model, "inception"))
    with tools.Context("slim", None):
      print("Dataset name in use:   " + repr(dataset))
      print("Dataset preprocessing: " + repr(self.preprocessing))
      print("Model name in use:     " + repr(model))
    self.data = Dataset(dataset, image_size=self.args["image-size"], synthetic_limit=self.args["synthetic-samples"])
    self.num_classes = self.data.classes - self.args["labels-offset"]
    self._streams = {}

  def model(self):
    net = nets_factory.get_network(self.model_name, self.num_classes)
    net.label_smoothing = self.args["label-smoothing"]
    return net

  def _offset(self, x, y):
    return (x, y - self.args["labels-offset"]) if self.args["labels-offset"] else (x, y)

  def train_stream(self, worker, nbworkers, device):
    if worker not in self._streams:
      self._streams[worker] = BatchStream(self.data.x_train, self.data.y_train, self.args["batch-size"], device, seed=self.args["seed"] + worker, transform=self._offset)
    return self._streams[worker]

  def eval_batch(self, device):
    if not hasattr(self, "_eval_stream"):
      self._eval_stream = BatchStream(self.data.x_test, self.data.y_test, min(self.args["eval-batch-size"], len(self.data.y_test)), device, shuffle=False, transform=self._offset)
    return next(self._eval_stream)

  def _augment(self, images, generator):
    """Training-time augmentation on the device, uint8 NHWC in and out: random horizontal flip, optionally preceded by a random
    crop of the image padded by 1/8 of its size on every side (a light stand-in for slim's random-resize-and-crop)."""
    import torch
    batch, height, width, _ = images.shape
    if self.args["augment"] == "crop-flip":
      pad_h, pad_w = max(1, height // 8), max(1, width // 8)
      padded = torch.nn.functional.pad(images.permute(0, 3, 1, 2), (pad_w, pad_w, pad_h, pad_h), mode="replicate").permute(0, 2, 3, 1)
      top = torch.randint(0, 2 * pad_h + 1, (batch,), device=images.device, generator=generator)
      left = torch.randint(0, 2 * pad_w + 1, (batch,), device=images.device, generator=generator)
      rows = top[:, None] + torch.arange(height, device=images.device)[None, :]
      cols = left[:, None] + torch.arange(width, device=images.device)[None, :]
      images = padded[torch.arange(batch, device=images.device)[:, None, None], rows[:, :, None], cols[:, None, :]]
    flip = torch.rand(batch, device=images.device, generator=generator) < 0.5
    return torch.where(flip.view(-1, 1, 1, 1), images.flip(2), images)

  def preprocess(self, inputs, ctx, training):
    if training and self.stochastic_preprocess:
      inputs = self._augment(inputs, ctx.generator)
    if self.preprocessing == "cifarnet":
      return cifarnet_preprocess(inputs, ctx.dtype, training, ctx.generator)
    from ..ops import nn as nn_ops
    return nn_ops.image_normalize(ctx.backend, inputs, self.preprocessing, ctx.dtype)


for _model in nets_factory.networks_map:
  for _dataset in known_datasets():
    register("slim-" + _model + "-" + _dataset, SlimExperiment._make(_dataset, _model))

"""`cnnet`: small CNN on CIFAR-10 (reference: `experiments/cnnet.py:54-196`).

Args: `batch-size:32`, `eval-batch-size:1024`, `nb-fetcher-threads`, `nb-batcher-threads` (accepted; the input
pipeline here is one prefetch thread per stream), `preprocessing:cifarnet`. Each worker dequeues its own batches.
`cifarnet` preprocessing = per-image standardisation (train: + random horizontal flip), done on the device.
"""

import torch

from .. import tools
from ..models import simple
from . import _Experiment, register
from ._data import BatchStream, Dataset


def cifarnet_preprocess(images, dtype, training, generator=None):
  """uint8 NHWC batch -> standardised channels_last activations of `dtype`."""
  x = images.permute(0, 3, 1, 2).to(torch.float32)
  if training:
    flip = torch.rand(x.shape[0], device=x.device, generator=generator) < 0.5
    x = torch.where(flip.view(-1, 1, 1, 1), x.flip(3), x)
  mean = x.mean(dim=(1, 2, 3), keepdim=True)
  std = x.std(dim=(1, 2, 3), keepdim=True, unbiased=False)
  x = (x - mean) / torch.clamp(std, min=1.0 / (x[0].numel() ** 0.5))
  return x.to(dtype).contiguous(memory_format=torch.channels_last)


class CNNetExperiment(_Experiment):
  def __init__(self, args):
    self.args = tools.parse_keyval(args if args is not None else [], defaults={
      "batch-size": 32, "eval-batch-size": 1024, "nb-fetcher-threads": 1, "nb-batcher-threads": 1, "preprocessing": "cifarnet", "seed": 0})
    if self.args["batch-size"] <= 0:
      raise tools.UserException("Cannot make batches of non-positive size")
    self.data = Dataset("cifar10", synthetic_limit=8192)
    if self.data.synthetic:
      tools.warning("CIFAR-10 files not found: using the synthetic CIFAR-10-shaped dataset", context="cnnet")
    self._streams = {}
    self.stochastic_preprocess = True   # the random flip of `cifarnet` preprocessing: no CUDA-graph replay of the step

  def model(self):
    return simple.cnnet(10)

  def train_stream(self, worker, nbworkers, device):
    if worker not in self._streams:
      self._streams[worker] = BatchStream(self.data.x_train, self.data.y_train, self.args["batch-size"], device, seed=self.args["seed"] + worker)
    return self._streams[worker]

  def eval_batch(self, device):
    if not hasattr(self, "_eval_stream"):
      self._eval_stream = BatchStream(self.data.x_test, self.data.y_test, min(self.args["eval-batch-size"], len(self.data.y_test)), device, shuffle=False)
    return next(self._eval_stream)

  def preprocess(self, inputs, ctx, training):
    return cifarnet_preprocess(inputs, ctx.dtype, training, ctx.generator)


register("cnnet", CNNetExperiment)

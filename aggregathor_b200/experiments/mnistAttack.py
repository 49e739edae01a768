"""`mnistAttack`: the MNIST MLP trained on malformed data (reference: `experiments/mnistAttack.py:48-174`).

The reference multiplies every training pixel by -1e12 and permutes images and labels independently
(`malformed_severity == 2`, `:83-92`), and — because the dataset is cached — feeds that poisoned stream to *every*
worker. Here the poisoning is per worker: the last `nb-byz` logical workers (default: all, the reference behaviour)
draw from the poisoned stream, the others from clean MNIST; combined with a robust GAR this is the data-poisoning
experiment the reference intended. `severity:1` = scale by -100, `severity:2` = scale by -1e12 + permutation.
"""

import numpy as np

from .. import tools
from . import register
from .mnist import MNIST


class MNISTAttack(MNIST):
  def __init__(self, args):
    args = list(args if args is not None else [])
    own = tools.parse_keyval([a for a in args if a.split(":")[0] in ("severity", "nb-byz")], defaults={"severity": 2, "nb-byz": -1})
    super().__init__([a for a in args if a.split(":")[0] not in ("severity", "nb-byz")])
    self.severity, self.nb_byz = own["severity"], own["nb-byz"]
    rng = np.random.default_rng(4321)
    self._bad_labels = rng.permutation(self.data.y_train) if self.severity == 2 else self.data.y_train
    self._bad_images = rng.permutation(self.data.x_train) if self.severity == 2 else self.data.x_train
    self._scale = {0: 1.0, 1: -100.0, 2: -1e12}[self.severity]

  def _is_poisoned(self, worker, nbworkers):
    return self.nb_byz < 0 or worker >= nbworkers - self.nb_byz

  def train_stream(self, worker, nbworkers, device):
    stream = super().train_stream(worker, nbworkers, device)
    stream.poisoned = self._is_poisoned(worker, nbworkers)
    return stream

  def _train_arrays(self, worker, nbworkers):
    if self._is_poisoned(worker, nbworkers):
      return self._bad_images, self._bad_labels
    return self.data.x_train, self.data.y_train

  def losses(self, model, batches, contexts, trace=None):
    out = []
    for (inputs, labels), ctx in zip(batches, contexts):
      x = self.preprocess(inputs, ctx, True)
      worker, nbworkers = getattr(ctx, "worker_id", 0), getattr(ctx, "nbworkers", 1)
      if self._is_poisoned(worker, nbworkers):
        x = (x.float() * self._scale).to(ctx.dtype)
      out.append(model.loss_and_backward(x, labels, ctx))
    return out


register("mnistAttack", MNISTAttack)

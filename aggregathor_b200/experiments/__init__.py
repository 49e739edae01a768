"""Experiments (model + dataset + loss/metric) — plug-in registry and base class.

Registry contract kept from the reference (`experiments/__init__.py:40-81`): `register(name, ctor)`,
`instantiate(name, args)`, `itemize()`; an experiment is built as `ctor(args)` with the raw
`--experiment-args key:value` list, and dropping a `.py` file here auto-registers it.

What changes is what an experiment *returns*. The reference's `losses(device_dataset, device_models, trace)`
built TF graph nodes; here an experiment provides
  * `model()`                      -> `models.Model` (static layer graph, flat parameters),
  * `train_stream(worker, nbworkers, device)` -> iterator of `(inputs, labels)` device batches (one independent
                                      stream per logical worker; `shared-batch:1` reproduces the reference's
                                      mnist quirk where every worker sees the same batch),
  * `eval_batch(device)`           -> one evaluation batch,
and the base class implements `losses(...)` / `accuracy(...)` on top of them: `losses` runs forward + backward for
each of this rank's logical workers (writing gradients into their rows) and returns the per-worker loss tensors,
`accuracy` returns `{"top1-X-acc": tensor}` — the same metric name as the reference.
"""

import pathlib

import torch

from .. import tools

__all__ = ["_Experiment", "register", "instantiate", "itemize"]


class _Experiment:
  """Base experiment."""

  def __init__(self, args):
    raise NotImplementedError

  def model(self):
    raise NotImplementedError

  def train_stream(self, worker, nbworkers, device):
    raise NotImplementedError

  def eval_batch(self, device):
    raise NotImplementedError

  def preprocess(self, inputs, ctx, training):
    """Device-side conversion of a raw input batch into the model's activation format."""
    return inputs.to(ctx.dtype)

  # -- reference-named entry points ------------------------------------------- #
  def losses(self, model, batches, contexts, trace=None):
    """Forward + backward of every local worker. `batches[j]`/`contexts[j]` belong to local worker j;
    gradients land in `contexts[j].grads`. Returns the list of per-worker losses (0-d tensors)."""
    out = []
    for j, ((inputs, labels), ctx) in enumerate(zip(batches, contexts)):
      if trace is not None:
        with trace.span("Worker " + str(getattr(ctx, "worker_id", j)) + ": loss + gradient computation"):
          out.append(model.loss_and_backward(self.preprocess(inputs, ctx, True), labels, ctx))
      else:
        out.append(model.loss_and_backward(self.preprocess(inputs, ctx, True), labels, ctx))
    return out

  def losses_batched(self, model, batches, ctx):
    """All local workers in ONE pass: their batches are concatenated along dim 0 and `ctx.groups` = number of workers (per-worker
    batch-norm statistics, per-worker loss means, per-worker parameter gradients `ctx.group_stride` elements apart).
    Returns a tensor of `len(batches)` losses."""
    inputs = torch.cat([b[0] for b in batches], dim=0)
    labels = torch.cat([b[1] for b in batches], dim=0)
    return model.loss_and_backward(self.preprocess(inputs, ctx, True), labels, ctx).reshape(-1)

  def accuracy(self, model, batch, ctx, trace=None):
    inputs, labels = batch
    with torch.no_grad():
      return {"top1-X-acc": model.accuracy(self.preprocess(inputs, ctx, False), labels, ctx)}


_register = tools.ClassRegister("experiment")
itemize = _register.itemize
register = _register.register
instantiate = _register.instantiate
del _register

with tools.Context("experiments", None):
  tools.import_directory(pathlib.Path(__file__).parent, globals())

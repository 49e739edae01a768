"""Gradient attacks that need no knowledge of the honest workers.

* `flip`      g <- factor * g (default factor -1; `factor:-1e12` is the gradient-space analogue of the reference's
              `mnistAttack` inputs scaled by -1e12, `experiments/mnistAttack.py:83-92`)
* `random`    g <- N(0, deviation^2) noise
* `nan`       g <- NaN everywhere (a fully lost gradient); `inf` likewise with +inf
* `zero`      g <- 0
* `constant`  g <- value
* `replay`    g <- the gradient this worker sent at the previous step (stale update)
* `drop-chunks`  lossy-transport emulation (reference UDP path, `tf_patches/patches/mpi_rendezvous_mgr.patch:814-843`):
              each 65 000-byte chunk of the serialized gradient is lost with probability `rate` and replaced by
              NaN (`fill:nan`, the intent of the reference), zeros (`fill:zero`, what its byte-fill actually did) or the
              bytes of the previous gradient (`fill:clever`, the reference's `CLEVER=1`).
* `forge`     in-flight tampering: scales the first `fraction` of the row by `factor` *after* the row was signed, so that gradient
              authentication (`runner.py --authenticate`, `parallel/signing.py`) must reject exactly the touched slices; without
              authentication it behaves like a partial `flip`.
"""

import torch

from .. import config, tools
from . import _Attack, register


class _Simple(_Attack):
  defaults = {}

  def __init__(self, nbworkers, nbbyzwrks, args):
    self.args = tools.parse_keyval(args if args is not None else [], defaults=self.defaults)


class FlipAttack(_Simple):
  defaults = {"factor": -1.0}

  def apply(self, row, worker, step, state):
    row.mul_(self.args["factor"])


class RandomAttack(_Simple):
  defaults = {"deviation": 1.0, "seed": 0}

  def apply(self, row, worker, step, state):
    gen = state.get("generator")
    if gen is None:
      gen = state["generator"] = torch.Generator(device=row.device)
      gen.manual_seed(self.args["seed"] * 7919 + worker)
    row.normal_(0.0, self.args["deviation"], generator=gen)


class NaNAttack(_Simple):
  def apply(self, row, worker, step, state):
    row.fill_(float("nan"))


class InfAttack(_Simple):
  def apply(self, row, worker, step, state):
    row.fill_(float("inf"))


class ZeroAttack(_Simple):
  def apply(self, row, worker, step, state):
    row.zero_()


class ConstantAttack(_Simple):
  defaults = {"value": 1.0}

  def apply(self, row, worker, step, state):
    row.fill_(self.args["value"])


class ReplayAttack(_Simple):
  def apply(self, row, worker, step, state):
    previous = state.get("previous")
    current = row.clone()
    if previous is not None:
      row.copy_(previous)
    state["previous"] = current


class DropChunksAttack(_Simple):
  defaults = {"rate": 0.1, "fill": "nan", "chunk-bytes": config.udp_chunk_bytes, "seed": 0}

  def __init__(self, nbworkers, nbbyzwrks, args):
    super().__init__(nbworkers, nbbyzwrks, args)
    if self.args["fill"] not in ("nan", "zero", "clever"):
      raise tools.UserException("drop-chunks: fill must be one of nan, zero, clever")

  def apply(self, row, worker, step, state):
    fill, rate, chunk = self.args["fill"], self.args["rate"], self.args["chunk-bytes"]
    previous = state.get("previous") if fill == "clever" else None
    if fill == "clever":
      current = row.clone()
      if previous is None:
        previous = torch.zeros_like(row)
    seed = (self.args["seed"] * 1000003 + worker) * 1000003 + step
    if row.is_cuda:
      from ..ops import gar as gar_ops
      gar_ops.drop_chunks_(row, rate, fill, previous, chunk, seed)
    else:
      elems = max(1, chunk // 4)
      nchunks = (row.numel() + elems - 1) // elems
      gen = torch.Generator().manual_seed(seed % (2 ** 63))
      lost = (torch.rand(nchunks, generator=gen) < rate).repeat_interleave(elems)[:row.numel()]
      if fill == "nan":
        row[lost] = float("nan")
      elif fill == "zero":
        row[lost] = 0.0
      else:
        row[lost] = previous[lost]
    if fill == "clever":
      state["previous"] = current


class ForgeAttack(_Simple):
  defaults = {"factor": -1.0, "fraction": 1.0}
  forges = True

  def apply(self, row, worker, step, state):
    count = max(1, min(row.numel(), int(row.numel() * self.args["fraction"])))
    row[:count].mul_(self.args["factor"])


register("flip", FlipAttack)
register("forge", ForgeAttack)
register("random", RandomAttack)
register("nan", NaNAttack)
register("inf", InfAttack)
register("zero", ZeroAttack)
register("constant", ConstantAttack)
register("replay", ReplayAttack)
register("drop-chunks", DropChunksAttack)

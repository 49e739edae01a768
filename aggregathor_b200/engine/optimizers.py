"""Optimizers applied to the aggregated gradient (reference: `graph.py:58-66`).

Same five choices and CLI keys as the reference (TF 1.x semantics and defaults):
`sgd`, `adam` (`adam-beta1`, `adam-beta2`), `rmsprop`, `adagrad` (`initial-accumulator-value`),
`adadelta` (`adadelta-rho`, `opt-epsilon`). An `OptimizerSpec` is a *description*: the fused
sm_100a kernel applies it on the coordinate slice owned by each rank (slots live in flat fp32
buffers indexed like the parameters); `apply_torch` is the same math in torch ops for the
host/gloo mode, the baseline path and the tests.
"""

import math

import torch


class OptimizerSpec:
  """Name, slot layout, hyper-parameters and reference (torch) update of an optimizer."""

  def __init__(self, name, nbslots, slot_init=(0.0, 0.0), **hyper):
    self.name, self.nbslots, self.slot_init, self.hyper = name, nbslots, slot_init, hyper

  def __repr__(self):
    return self.name + "(" + ", ".join(k + "=" + repr(v) for k, v in self.hyper.items()) + ")"

  def kernel_args(self, rate, step):
    """(effective lr, h0, h1, h2) as consumed by `native/op_gar` for update number `step` (1-based)."""
    h = self.hyper
    if self.name == "adam":
      correction = math.sqrt(1.0 - h["beta2"] ** step) / (1.0 - h["beta1"] ** step)
      return rate * correction, (h["beta1"], h["beta2"], h["epsilon"])
    if self.name == "rmsprop":
      return rate, (h["decay"], h["momentum"], h["epsilon"])
    if self.name == "adadelta":
      return rate, (h["rho"], h["epsilon"], 0.0)
    return rate, (0.0, 0.0, 0.0)

  def make_slots(self, like):
    return [torch.full_like(like, self.slot_init[i]) for i in range(self.nbslots)]

  def apply_torch(self, param, grad, slots, rate, step):
    """In-place reference update of `param` (any device) with the aggregated `grad`."""
    lr, (h0, h1, h2) = self.kernel_args(rate, step)
    if self.name == "sgd":
      param.sub_(grad, alpha=lr)
    elif self.name == "adam":
      m, v = slots
      m.mul_(h0).add_(grad, alpha=1.0 - h0)
      v.mul_(h1).addcmul_(grad, grad, value=1.0 - h1)
      param.sub_(lr * m / (v.sqrt() + h2))
    elif self.name == "rmsprop":
      ms, mom = slots
      ms.mul_(h0).addcmul_(grad, grad, value=1.0 - h0)
      mom.mul_(h1).add_(lr * grad * torch.rsqrt(ms + h2))
      param.sub_(mom)
    elif self.name == "adagrad":
      (acc,) = slots
      acc.addcmul_(grad, grad)
      param.sub_(lr * grad * torch.rsqrt(acc))
    elif self.name == "adadelta":
      acc, acc_update = slots
      acc.mul_(h0).addcmul_(grad, grad, value=1.0 - h0)
      update = torch.sqrt(acc_update + h1) * torch.rsqrt(acc + h1) * grad
      acc_update.mul_(h0).addcmul_(update, update, value=1.0 - h0)
      param.sub_(update, alpha=lr)
    else:
      raise AssertionError(self.name)


optimizers = {
  "adadelta": (lambda rho=None, epsilon=None: OptimizerSpec("adadelta", 2, rho=rho, epsilon=epsilon),
               {"adadelta-rho": (0.95, "rho"), "opt-epsilon": (1., "epsilon")}),
  "adagrad": (lambda initial_accumulator_value=None: OptimizerSpec("adagrad", 1, slot_init=(initial_accumulator_value, 0.0), initial_accumulator_value=initial_accumulator_value),
              {"initial-accumulator-value": (0.1, "initial_accumulator_value")}),
  "adam": (lambda beta1=None, beta2=None: OptimizerSpec("adam", 2, beta1=beta1, beta2=beta2, epsilon=1e-8),
           {"adam-beta1": (0.9, "beta1"), "adam-beta2": (0.999, "beta2")}),
  # TF 1.x `RMSPropOptimizer._create_slots`: the mean-square slot "rms" starts at ONE, "momentum" at zero (first update ~ lr * g)
  "rmsprop": (lambda: OptimizerSpec("rmsprop", 2, slot_init=(1.0, 0.0), decay=0.9, momentum=0.0, epsilon=1e-10), {}),
  "sgd": (lambda: OptimizerSpec("sgd", 0), {})}

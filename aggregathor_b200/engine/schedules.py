"""Learning-rate schedules and the generic `build` helper (reference: `graph.py:51-91`).

`learning_rates` maps a schedule name to `(constructor, {cli-key: (default, kwarg)})`, exactly the
table shape of the reference, so `build(learning_rates, "learning rate decay", name, args)` parses
`--learning-rate-args key:value ...` the same way (unknown keys are ignored, values are coerced to
the default's type). Schedules are host-side callables `rate(step) -> float`: the rate is a kernel
argument of the fused update, not a graph node.
"""

from .. import config, tools


class LearningRate:
  """Callable `step -> rate` with a printable description."""

  def __init__(self, name, fn, **params):
    self.name, self._fn, self.params = name, fn, params

  def __call__(self, step):
    return float(self._fn(int(step)))

  def __repr__(self):
    return self.name + "(" + ", ".join(k + "=" + repr(v) for k, v in self.params.items()) + ")"


def _fixed(initial_rate=None):
  return LearningRate("fixed", lambda step: initial_rate, initial_rate=initial_rate)


def _polynomial(initial_rate=None, decay_step=None, end_rate=None, power=None):
  # tf.train.polynomial_decay(cycle=False): (lr0 - lr_end) * (1 - min(step, T)/T)^power + lr_end
  def rate(step):
    frac = min(step, decay_step) / float(decay_step)
    return (initial_rate - end_rate) * (1.0 - frac) ** power + end_rate
  return LearningRate("polynomial", rate, initial_rate=initial_rate, decay_step=decay_step, end_rate=end_rate, power=power)


def _exponential(initial_rate=None, decay_step=None, decay_rate=None):
  # tf.train.exponential_decay(staircase=False): lr0 * decay_rate^(step / T)
  return LearningRate("exponential", lambda step: initial_rate * decay_rate ** (step / float(decay_step)),
                      initial_rate=initial_rate, decay_step=decay_step, decay_rate=decay_rate)


learning_rates = {
  "fixed": (_fixed, {"initial-rate": (config.default_learning_rate, "initial_rate")}),
  "polynomial": (_polynomial, {
    "initial-rate": (config.default_learning_rate, "initial_rate"), "end-rate": (config.default_end_learning_rate, "end_rate"),
    "decay-step": (config.default_decay_step, "decay_step"), "power": (1., "power")}),
  "exponential": (_exponential, {
    "initial-rate": (config.default_learning_rate, "initial_rate"), "decay-step": (config.default_decay_step, "decay_step"),
    "decay-rate": (config.default_decay_rate, "decay_rate")})}


def build(struct, name, select, args, **kwargs):
  """Instantiate `struct[select]` with its `key:value` CLI arguments (+ forwarded kwargs)."""
  if select not in struct:
    raise tools.UserException("Unknown " + name + " " + repr(select) + ", " + ("no " + name + " available" if len(struct) == 0 else "expected one of: '" + "', '".join(struct.keys()) + "'"))
  construct, arg_table = struct[select]
  parsed = tools.parse_keyval(args if args is not None else [], defaults={key: entry[0] for key, entry in arg_table.items()})
  call_kwargs = {entry[1]: parsed[key] for key, entry in arg_table.items()}  # supplementary keys are ignored
  call_kwargs.update(kwargs)
  return construct(**call_kwargs)

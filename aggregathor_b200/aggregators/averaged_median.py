"""`averaged-median`: per coordinate, mean of the beta = n - f values closest to the median
(reference: `aggregators/averaged-median.py:53-61` -> `deprecated_native/native.cpp:714-740`).

`--aggregator-args beta:<int>` overrides beta."""

from .. import tools
from . import _GAR, FusedSpec, register
from . import _ops


class AveragedMedianGAR(_GAR):
  def __init__(self, nbworkers, nbbyzwrks, args):
    parsed = tools.parse_keyval(args if args is not None else [], defaults={"beta": nbworkers - nbbyzwrks})
    self._n = nbworkers
    self._beta = parsed["beta"]
    if not 1 <= self._beta <= nbworkers:
      raise tools.UserException("averaged-median needs 1 <= beta <= n (got beta = %d, n = %d)" % (self._beta, nbworkers))

  def aggregate(self, gradients):
    G = _ops.stack(gradients)
    beta = min(self._beta, G.shape[0])
    return _ops.dispatch(G, lambda M: _ops.host_averaged_median(M, beta), FusedSpec("averaged-median", G.shape[0], beta=beta))

  def fused_spec(self):
    return FusedSpec("averaged-median", self._n, beta=self._beta)


register("averaged-median", AveragedMedianGAR)

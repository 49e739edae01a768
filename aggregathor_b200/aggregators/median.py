"""`median`: coordinate-wise median, element of rank floor(n/2) with non-finite values ordered
last — i.e. the *upper* median for even n, no averaging of the two middles (reference:
`aggregators/median.py:57-62` -> `deprecated_native/native.cpp:678-697`).

sm_100a path: per-coordinate selection network in registers over the n values streamed
from the peers' gradient buffers."""

from . import _GAR, FusedSpec, register
from . import _ops


class MedianGAR(_GAR):
  def __init__(self, nbworkers, nbbyzwrks, args):
    self._n = nbworkers

  def aggregate(self, gradients):
    G = _ops.stack(gradients)
    return _ops.dispatch(G, _ops.host_median, FusedSpec("median", G.shape[0]))

  def fused_spec(self):
    return FusedSpec("median", self._n)


register("median", MedianGAR)

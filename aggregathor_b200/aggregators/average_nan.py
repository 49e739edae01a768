"""`average-nan`: coordinate-wise mean over the *finite* values only (reference:
`aggregators/average-nan.py:57-62` -> `deprecated_native/native.cpp:756-775`).

Exists because the lossy (UDP) transport turns lost chunks into NaN coordinates. A
coordinate that is non-finite for every worker yields 0/0 = NaN, which trips the
NaN-loss guard of the runner, exactly as in the reference."""

from . import _GAR, FusedSpec, register
from . import _ops


class AverageNaNGAR(_GAR):
  def __init__(self, nbworkers, nbbyzwrks, args):
    self._n = nbworkers

  def aggregate(self, gradients):
    G = _ops.stack(gradients)
    return _ops.dispatch(G, _ops.host_average_nan, FusedSpec("average-nan", G.shape[0]))

  def fused_spec(self):
    return FusedSpec("average-nan", self._n)


register("average-nan", AverageNaNGAR)

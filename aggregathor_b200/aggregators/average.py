"""`average`: arithmetic mean of the n gradients, NaN propagates (reference: `aggregators/average.py:47-54`).

Fused path: in-switch NVLS reduction (`multimem.ld_reduce.add`) or P2P loads of the
peers' slices, fused with the optimizer update and the parameter multicast."""

from . import _GAR, FusedSpec, register
from . import _ops


class AverageGAR(_GAR):
  def __init__(self, nbworkers, nbbyzwrks, args):
    self._n = nbworkers

  def aggregate(self, gradients):
    G = _ops.stack(gradients)
    return _ops.dispatch(G, _ops.host_average if G.dtype.is_floating_point and G.dtype.itemsize >= 4 else _ops.torch_average, self.fused_spec_for(G.shape[0]))

  def fused_spec_for(self, n):
    return FusedSpec("average", n)

  def fused_spec(self):
    return self.fused_spec_for(self._n)


register("average", AverageGAR)

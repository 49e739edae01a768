"""Bulyan over Multi-Krum (reference: `aggregators/bulyan.py:43-94`, `native/op_bulyan/cpu.cpp:52-195`).

theta = n - 2f - 2 selection rounds; round k outputs the mean of the (m - k) best-scoring
remaining gradients (m = n - f - 2), removes the best one and updates the scores with the
pruned distances; then, per coordinate, the beta = theta - 2f intermediates closest to their
median are averaged. Needs n >= 4f + 3 (the reference underflows a size_t otherwise,
`op_bulyan/cpu.cpp:57-58`; here it is a `UserException`).

Because every intermediate is a fixed linear combination of the inputs, the sm_100a kernel
never materialises the [theta, d] intermediates: the selection stage yields a [theta, n]
weight matrix and the coordinate stage rebuilds the theta values in registers.

Flavours: `bulyan-py` (host C++), `bulyan-co` / `bulyan` (sm_100a kernel, host for CPU tensors).
"""

from .. import tools
from . import _GAR, FusedSpec, register
from . import _ops


class _BulyanBase(_GAR):
  def __init__(self, nbworkers, nbbyzwrks, args):
    parsed = tools.parse_keyval(args if args is not None else [], defaults={"m": nbworkers - nbbyzwrks - 2})
    _ops.check_bulyan(nbworkers, nbbyzwrks, parsed["m"])
    self._n, self._f, self._m = nbworkers, nbbyzwrks, parsed["m"]

  def _params(self, G):
    n = G.shape[0]
    _ops.check_bulyan(n, self._f, min(self._m, n))
    return self._f, min(self._m, n)

  def fused_spec(self):
    return FusedSpec("bulyan", self._n, f=self._f, m=self._m, beta=self._n - 4 * self._f - 2)


class PYBulyanGAR(_BulyanBase):
  def aggregate(self, gradients):
    G = _ops.stack(gradients)
    f, m = self._params(G)
    return _ops.host_bulyan(G, f, m)


class COBulyanGAR(_BulyanBase):
  def aggregate(self, gradients):
    G = _ops.stack(gradients)
    f, m = self._params(G)
    n = G.shape[0]
    return _ops.dispatch(G, lambda M: _ops.host_bulyan(M, f, m), FusedSpec("bulyan", n, f=f, m=m, beta=n - 4 * f - 2))


register("bulyan-py", PYBulyanGAR)
register("bulyan-co", COBulyanGAR)
register("bulyan", COBulyanGAR)

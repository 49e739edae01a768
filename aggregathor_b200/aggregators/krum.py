"""Multi-Krum (reference: `aggregators/krum.py:45-169`, `native/op_krum/cpu.cpp:51-127`).

score_i = sum of the n - f - 2 smallest squared distances from g_i to the other gradients
(non-finite distance -> +inf); output = mean of the m = n - f - 2 smallest-scoring gradients
(`--aggregator-args m:<int>` overrides m; m:1 is the original Krum). Ties -> lower index.

Registered flavours, all computing the same function:
  `krum-py` host C++ library through ctypes (GPU tensors round-trip through the host),
  `krum-tf` plain torch ops on whatever device holds the gradients,
  `krum-co` stand-alone sm_100a kernel (host library for CPU tensors),
  `krum`    alias of `krum-co`; this is the one the fused multi-GPU kernel implements.
"""

from .. import tools
from . import _GAR, FusedSpec, register
from . import _ops


class _KrumBase(_GAR):
  def __init__(self, nbworkers, nbbyzwrks, args):
    parsed = tools.parse_keyval(args if args is not None else [], defaults={"m": nbworkers - nbbyzwrks - 2})
    _ops.check_krum(nbworkers, nbbyzwrks, parsed["m"])
    self._n, self._f, self._m = nbworkers, nbbyzwrks, parsed["m"]

  def _params(self, G):
    n = G.shape[0]  # like the custom op, n comes from the stacked tensor (reference: op_krum/op.cpp:77)
    _ops.check_krum(n, self._f, min(self._m, n))
    return self._f, min(self._m, n)

  def fused_spec(self):
    return FusedSpec("krum", self._n, f=self._f, m=self._m)


class PYKrumGAR(_KrumBase):
  def aggregate(self, gradients):
    G = _ops.stack(gradients)
    f, m = self._params(G)
    if m == G.shape[0]:
      return _ops.host_average(G)  # fast path of the reference (krum.py:56-62)
    return _ops.host_krum(G, f, m)


class TorchKrumGAR(_KrumBase):
  def aggregate(self, gradients):
    G = _ops.stack(gradients)
    f, m = self._params(G)
    return _ops.torch_krum(G, f, m)


class COKrumGAR(_KrumBase):
  def aggregate(self, gradients):
    G = _ops.stack(gradients)
    f, m = self._params(G)
    return _ops.dispatch(G, lambda M: _ops.host_krum(M, f, m), FusedSpec("krum", G.shape[0], f=f, m=m))


register("krum-py", PYKrumGAR)
register("krum-tf", TorchKrumGAR)
register("krum-co", COKrumGAR)
register("krum", COKrumGAR)

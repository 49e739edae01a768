"""Back-ends shared by the GAR plug-ins.

Three interchangeable implementations of every rule, mirroring the reference's
`-py` / `-tf` / `-co` triplets (`aggregators/krum.py:45-169`, `bulyan.py:43-94`):

* `host_*`   — the C++ host library `native/py_gars` through ctypes (the `-py` flavour;
               GPU tensors take a device->host->device round trip, as `tf.py_func` did);
* `torch_*`  — plain torch ops, device agnostic (the `-tf` flavour; also the fp32 test oracle);
* `cuda_*`   — the stand-alone sm_100a kernels of `native/op_gar` (the `-co` flavour).

Ordering convention everywhere: finite ascending, then non-finite; ties -> lower worker index.
"""

import ctypes

import torch

from .. import tools

# ---------------------------------------------------------------------------- #
# Helpers

def stack(gradients):
  """List of n flat tensors (or an [n, d] tensor) -> contiguous [n, d] tensor."""
  if isinstance(gradients, torch.Tensor):
    if gradients.dim() != 2:
      raise tools.UserException("Expected an [n, d] tensor of gradients, got shape " + repr(tuple(gradients.shape)))
    return gradients.contiguous()
  if len(gradients) == 0:
    raise tools.UserException("Empty list of gradient to aggregate")
  return torch.stack([g.reshape(-1) for g in gradients], dim=0)


def check_krum(n, f, m=None):
  if n - f - 2 < 1:
    raise tools.UserException("Multi-Krum needs n - f - 2 >= 1 (got n = %d, f = %d)" % (n, f))
  if m is not None and not 1 <= m <= n:
    raise tools.UserException("Multi-Krum needs 1 <= m <= n (got m = %d, n = %d)" % (m, n))


def check_bulyan(n, f, m=None):
  if n < 4 * f + 3:
    raise tools.UserException("Bulyan needs n >= 4 f + 3 (got n = %d, f = %d): beta = n - 4 f - 2 must be >= 1" % (n, f))
  theta = n - 2 * f - 2
  if m is not None and not theta <= m <= n:
    raise tools.UserException("Bulyan needs n - 2 f - 2 <= m <= n (got m = %d)" % m)


def _rank_key(values):
  """Sort key implementing (finite ascending, non-finite last); argsort(stable) then breaks ties by index."""
  return torch.where(torch.isfinite(values), values, torch.full_like(values, float("inf")))


# ---------------------------------------------------------------------------- #
# Host C++ back-end

_host_cache = {}


def _host(symbol, dtype):
  from .. import native
  suffix = {torch.float32: "float", torch.float64: "double"}.get(dtype)
  if suffix is None:
    raise tools.UserException("Unsupported floating point type " + repr(dtype) + " for the host GARs")
  key = symbol + "_" + suffix
  if key not in _host_cache:
    lib = native.library("py_gars")
    _host_cache[key] = getattr(lib, "agb_cpu_" + key)
  return _host_cache[key]


def _ptr(tensor):
  return ctypes.c_void_p(tensor.data_ptr())


def _host_call(symbol, G, *extra, outputs=()):
  """Run `agb_cpu_<symbol>_<type>(G, n, d, *extra, out, *outputs)` on a CPU copy of G; result on G's device."""
  Gc = G.detach().to("cpu").contiguous()
  n, d = Gc.shape
  out = torch.empty(d, dtype=Gc.dtype)
  func = _host(symbol, Gc.dtype)
  status = func(_ptr(Gc), ctypes.c_size_t(n), ctypes.c_size_t(d), *[ctypes.c_size_t(e) for e in extra], _ptr(out), *[(_ptr(o) if o is not None else None) for o in outputs])
  if status != 0:
    raise tools.UserException("Host GAR " + repr(symbol) + " rejected its arguments (n = %d, d = %d, extra = %r)" % (n, d, extra))
  return out.to(G.device)


def host_average(G):
  return _host_call("average", G)


def host_average_nan(G):
  return _host_call("average_nan", G)


def host_median(G):
  return _host_call("median", G)


def host_averaged_median(G, beta):
  return _host_call("averaged_median", G, beta)


def host_krum(G, f, m, return_selected=False):
  selected = torch.empty(m, dtype=torch.int64)
  out = _host_call("krum", G, f, m, outputs=(selected, None))
  return (out, selected) if return_selected else out


def host_bulyan(G, f, m, return_weights=False):
  n = G.shape[0]
  weights = torch.empty((n - 2 * f - 2, n), dtype=G.dtype if G.dtype in (torch.float32, torch.float64) else torch.float32)
  out = _host_call("bulyan", G, f, m, outputs=(weights,))
  return (out, weights) if return_weights else out


def host_pairwise_distances(G):
  Gc = G.detach().to("cpu").contiguous()
  n, d = Gc.shape
  dist = torch.empty((n, n), dtype=Gc.dtype)
  _host("pairwise_distances", Gc.dtype)(_ptr(Gc), ctypes.c_size_t(n), ctypes.c_size_t(d), _ptr(dist))
  return dist


def host_bulyan_weights(dist, f, m):
  """Selection stage of Bulyan on an [n, n] distance matrix -> [theta, n] weight matrix."""
  dist = dist.detach().to("cpu").contiguous()
  n = dist.shape[0]
  weights = torch.empty((n - 2 * f - 2, n), dtype=dist.dtype)
  status = _host("bulyan_weights", dist.dtype)(_ptr(dist), ctypes.c_size_t(n), ctypes.c_size_t(f), ctypes.c_size_t(m), _ptr(weights))
  if status != 0:
    raise tools.UserException("Invalid Bulyan parameters (n = %d, f = %d, m = %d)" % (n, f, m))
  return weights


def host_squared_distance(a, b):
  a = a.detach().to("cpu").contiguous().reshape(-1)
  b = b.detach().to("cpu").contiguous().reshape(-1)
  func = _host("squared_distance", a.dtype)
  func.restype = ctypes.c_float if a.dtype == torch.float32 else ctypes.c_double
  return float(func(_ptr(a), _ptr(b), ctypes.c_size_t(a.numel())))


# ---------------------------------------------------------------------------- #
# Pure torch back-end (device agnostic; the fp32 oracle of the CUDA kernels)

def torch_average(G):
  return G.sum(dim=0) / G.shape[0]


def torch_average_nan(G):
  finite = torch.isfinite(G)
  total = torch.where(finite, G, torch.zeros_like(G)).sum(dim=0)
  return total / finite.sum(dim=0).to(G.dtype)


def torch_median(G):
  n = G.shape[0]
  order = torch.argsort(_rank_key(G), dim=0, stable=True)
  return torch.gather(G, 0, order[n // 2:n // 2 + 1]).squeeze(0)


def torch_averaged_median(G, beta):
  n = G.shape[0]
  zero = torch_median(G)
  dev = (G - zero.unsqueeze(0)).abs()
  order = torch.argsort(_rank_key(dev), dim=0, stable=True)[:beta]
  keep = torch.zeros_like(G, dtype=torch.bool).scatter_(0, order, True)
  return torch.where(keep, G, torch.zeros_like(G)).sum(dim=0) / beta


def torch_pairwise_distances(G):
  n = G.shape[0]
  dist = torch.zeros((n, n), dtype=G.dtype, device=G.device)
  for i in range(n - 1):
    delta = G[i + 1:] - G[i].unsqueeze(0)
    row = (delta * delta).sum(dim=1)
    row = torch.where(torch.isfinite(row), row, torch.full_like(row, float("inf")))
    dist[i, i + 1:] = row
    dist[i + 1:, i] = row
  return dist


def krum_select(dist, f, m):
  """[n, n] distances -> (scores [n], ids of the m best, ascending score / index-stable)."""
  n = dist.shape[0]
  off = dist.clone()
  off.fill_diagonal_(float("inf"))  # a worker is not its own neighbour
  ranked, _ = torch.sort(_rank_key(off), dim=1, stable=True)
  scores = ranked[:, :n - f - 2].sum(dim=1)
  order = torch.argsort(_rank_key(scores), stable=True)
  return scores, order[:m]


def torch_krum(G, f, m, return_selected=False):
  dist = torch_pairwise_distances(G)
  _, selected = krum_select(dist, f, m)
  selected, _ = torch.sort(selected)
  out = G[selected].sum(dim=0) / m
  return (out, selected) if return_selected else out


def torch_bulyan_weights(dist, f, m):
  """Pure-python transcription of the selection loop (small n): returns the [theta, n] weight matrix."""
  n = dist.shape[0]
  theta = n - 2 * f - 2
  inscore = n - f - 2
  D = dist.detach().to("cpu", torch.float64)
  D = torch.where(torch.isfinite(D), D, torch.full_like(D, float("inf")))
  key = lambda value, index: (0 if value != float("inf") and value == value else 1, value if value == value else 0.0, index)
  pruned = D.clone()
  scores = []
  for i in range(n):
    others = sorted((j for j in range(n) if j != i), key=lambda j: key(float(D[i, j]), j))
    scores.append(sum(float(D[i, j]) for j in others[:inscore]))
    for j in others[inscore:]:
      pruned[i, j] = 0.0
  removed = [False] * n
  weights = torch.zeros((theta, n), dtype=dist.dtype)
  for k in range(theta):
    order = sorted(range(n), key=lambda i: (removed[i],) + key(scores[i], i))
    count = m - k
    for i in order[:count]:
      weights[k, i] = 1.0 / count
    best = order[0]
    removed[best] = True
    for i in range(n):
      if not removed[i]:
        scores[i] -= float(pruned[i, best])
  return weights


def torch_bulyan(G, f, m, return_weights=False):
  n = G.shape[0]
  theta = n - 2 * f - 2
  beta = theta - 2 * f
  dist = torch_pairwise_distances(G)
  weights = torch_bulyan_weights(dist, f, m).to(G.device)
  inter = torch.stack([G[weights[k] != 0].sum(dim=0) / int((weights[k] != 0).sum()) for k in range(theta)], dim=0)
  out = torch_averaged_median(inter, beta)
  return (out, weights) if return_weights else out


# ---------------------------------------------------------------------------- #
# Stand-alone CUDA back-end (native/op_gar)

def cuda_aggregate(spec, G):
  """Run the stand-alone sm_100a kernel for `spec` on the [n, d] CUDA matrix G."""
  from ..ops import gar as gar_ops
  return gar_ops.aggregate(spec, G)


def dispatch(G, host_fn, cuda_spec):
  """Default placement policy: CUDA tensors -> sm_100a kernel, CPU tensors -> host C++ library."""
  if G.is_cuda:
    return cuda_aggregate(cuda_spec, G)
  return host_fn(G)

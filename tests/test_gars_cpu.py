"""Host C++ GARs vs the torch oracle, vs the reference's own `native.cpp` (compiled at test time), plus properties."""

import ctypes
import pathlib
import subprocess

import numpy as np
import pytest
import torch

from aggregathor_b200 import aggregators, tools
from aggregathor_b200.aggregators import _ops

ROOT = pathlib.Path(__file__).resolve().parent.parent
REF_SOURCES = [pathlib.Path("/root/reference/aggregators/deprecated_native/native.cpp"), ROOT / "baseline/_ref/aggregathor/aggregators/deprecated_native/native.cpp"]


@pytest.fixture(scope="module")
def reference(tmp_path_factory):
  """The reference's deprecated native library, built with its documented command line (deprecated_native/__init__.py:59)."""
  source = next((p for p in REF_SOURCES if p.is_file()), None)
  if source is None:
    pytest.skip("reference sources unavailable")
  lib = tmp_path_factory.mktemp("ref") / "lib.so"
  cmd = ["c++", "-Wall", "-Wextra", "-Wfatal-errors", "-O2", "-std=c++14", "-fPIC", "-shared", "-o", str(lib), str(source)]
  if subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE).returncode != 0:
    pytest.skip("reference native.cpp does not build here")
  return ctypes.CDLL(str(lib))


def _data(n, d, seed, outliers=0, nan=False):
  gen = torch.Generator().manual_seed(seed)
  G = torch.randn(n, d, generator=gen)
  for k in range(outliers):
    G[n - 1 - k] = G[n - 1 - k] * 25 - 4
  if nan:
    G[0, ::7] = float("nan")
    G[1, 3] = float("inf")
  return G


def _ref_call(lib, name, G, *extra):
  arr = np.ascontiguousarray(G.numpy().copy())
  out = np.empty(arr.shape[1], dtype=np.float32)
  getattr(lib, name + "_float")(ctypes.c_size_t(arr.shape[1]), ctypes.c_size_t(arr.shape[0]), *[ctypes.c_size_t(e) for e in extra],
                                ctypes.c_void_p(arr.ctypes.data), ctypes.c_void_p(out.ctypes.data))
  return torch.from_numpy(out)


def _close(a, b, tol=1e-5):
  assert bool((torch.isnan(a) == torch.isnan(b)).all())
  assert float((torch.nan_to_num(a) - torch.nan_to_num(b)).abs().max()) <= tol * max(1.0, float(torch.nan_to_num(b).abs().max()))


@pytest.mark.parametrize("n,d", [(4, 50), (7, 1001), (8, 4096), (11, 333)])
def test_against_reference_native(reference, n, d):
  G = _data(n, d, n + d, nan=True)
  _close(_ops.host_median(G), _ref_call(reference, "median", G))
  _close(_ops.host_average_nan(G), _ref_call(reference, "average_nan", G))
  clean = _data(n, d, n * d)
  _close(_ops.host_averaged_median(clean, n - 2), _ref_call(reference, "averaged_median", clean, n - 2))
  a, b = clean[0].contiguous(), clean[1].contiguous()
  reference.squared_distance_float.restype = ctypes.c_float
  expected = reference.squared_distance_float(ctypes.c_size_t(d), ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()))
  assert abs(_ops.host_squared_distance(a, b) - expected) <= 1e-4 * expected


@pytest.mark.parametrize("n,f", [(11, 2), (15, 3), (7, 1)])
def test_bulyan_against_reference_native(reference, n, f):
  """On well separated data the reference's deprecated bulyan (flagged possibly buggy by its authors) and ours agree."""
  G = _data(n, 500, n, outliers=f)
  arr = np.ascontiguousarray(G.numpy().copy())
  s = n - 2 * f - 2
  sel = np.empty((s, arr.shape[1]), dtype=np.float32)
  out = np.empty(arr.shape[1], dtype=np.float32)
  reference.bulyan_float(ctypes.c_size_t(arr.shape[1]), ctypes.c_size_t(n), ctypes.c_size_t(f), ctypes.c_size_t(s), ctypes.c_void_p(arr.ctypes.data),
                         ctypes.c_void_p(sel.ctypes.data), ctypes.c_void_p(out.ctypes.data))
  ours = _ops.host_bulyan(G, f, n - f - 2)
  # both must stay inside the honest cloud: far from the outliers, close to the honest mean
  honest = G[:n - f].mean(dim=0)
  assert float((ours - honest).norm()) < 0.6 * float(honest.norm() + G[:n - f].std(dim=0).norm())
  assert float((torch.from_numpy(out) - ours).norm()) < float((G[n - 1] - honest).norm()) * 0.05


@pytest.mark.parametrize("n,f,d", [(8, 2, 1000), (5, 1, 77), (16, 3, 4000), (11, 2, 513)])
def test_host_matches_torch_oracle(n, f, d):
  G = _data(n, d, d, outliers=f)
  m = n - f - 2
  _close(_ops.host_krum(G, f, m), _ops.torch_krum(G, f, m))
  _close(_ops.host_average(G), _ops.torch_average(G))
  _close(_ops.host_median(G), _ops.torch_median(G))
  _close(_ops.host_averaged_median(G, n - f), _ops.torch_averaged_median(G, n - f))
  _close(_ops.host_pairwise_distances(G), _ops.torch_pairwise_distances(G), 1e-4)
  if n >= 4 * f + 3:
    _close(_ops.host_bulyan(G, f, m), _ops.torch_bulyan(G, f, m))
    dist = _ops.host_pairwise_distances(G)
    assert torch.equal(_ops.host_bulyan_weights(dist, f, m) != 0, _ops.torch_bulyan_weights(dist, f, m) != 0)


def test_double_precision_and_determinism():
  G = _data(9, 3001, 5, outliers=2).double()
  a, b = _ops.host_krum(G, 2, 5), _ops.host_krum(G, 2, 5)
  assert a.dtype == torch.float64 and torch.equal(a, b)
  assert torch.equal(_ops.host_pairwise_distances(G), _ops.host_pairwise_distances(G))


def test_properties():
  G = _data(8, 600, 9, outliers=2)
  perm = torch.tensor([3, 0, 7, 1, 6, 2, 5, 4])
  for name, f in (("krum", 2), ("median", 0), ("averaged-median", 2), ("average", 0)):
    gar = aggregators.instantiate(name, 8, f, [])
    _close(gar.aggregate(list(G)), gar.aggregate(list(G[perm])), 1e-5)            # permutation invariance
    same = G[0:1].repeat(8, 1)
    _close(gar.aggregate(list(same)), same[0], 1e-6)                              # identical gradients => identity
  _close(aggregators.instantiate("krum", 8, 0, ["m:8"]).aggregate(list(G)), G.mean(dim=0), 1e-5)   # f = 0, m = n => average
  poisoned = G.clone()
  poisoned[2] = float("nan")
  out, selected = _ops.host_krum(poisoned, 2, 4, return_selected=True)
  assert 2 not in selected.tolist() and bool(torch.isfinite(out).all())                          # NaN rows are never selected
  out, selected = _ops.host_krum(G, 2, 4, return_selected=True)
  assert 6 not in selected.tolist() and 7 not in selected.tolist()                                # outliers rejected
  assert float((aggregators.instantiate("bulyan", 11, 2, []).aggregate(list(_data(11, 300, 2, outliers=2))) ).abs().max()) < 3.0


def test_invalid_configurations_raise():
  with pytest.raises(tools.UserException):
    aggregators.instantiate("bulyan", 8, 2, [])   # beta = n - 4f - 2 < 1: the reference underflows a size_t here
  with pytest.raises(tools.UserException):
    aggregators.instantiate("krum", 4, 2, [])
  with pytest.raises(tools.UserException):
    aggregators.instantiate("unknown-gar", 4, 0, [])
  with pytest.raises(tools.UserException):
    aggregators.instantiate("averaged-median", 4, 0, ["beta:9"])


def test_registry_names_match_reference():
  names = set(aggregators.itemize())
  assert {"average", "average-nan", "median", "averaged-median", "krum-py", "krum-tf", "krum-co", "bulyan-py", "bulyan-co"} <= names

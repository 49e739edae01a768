#!/usr/bin/env python3
"""Host (CPU) aggregation rules: this framework's C++ library (`native/py_gars`) vs the reference's own `deprecated_native/native.cpp`
compiled with its documented command line — the one part of the reference that builds offline, so the one place where the two can
be timed against each other on the same machine. Prints ms per call (median of `--reps` after warm-up) and the ratio; writes JSON."""
import argparse
import ctypes
import json
import os
import pathlib
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from aggregathor_b200.aggregators import _ops  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--d", type=int, default=1756426, help="gradient dimension (default: cnnet)")
parser.add_argument("--n", type=int, default=8)
parser.add_argument("--f", type=int, default=1)
parser.add_argument("--reps", type=int, default=15)
parser.add_argument("--out", default="profiles/host_gar_bench.json")
args = parser.parse_args()
source = next((p for p in (pathlib.Path("/root/reference/aggregators/deprecated_native/native.cpp"),
                           ROOT / "baseline/_ref/aggregathor/aggregators/deprecated_native/native.cpp") if p.is_file()), None)
lib = None
if source is not None:
  target = pathlib.Path(tempfile.mkdtemp()) / "ref.so"
  if subprocess.run(["c++", "-Wall", "-Wextra", "-Wfatal-errors", "-O2", "-std=c++14", "-fPIC", "-shared", "-o", str(target), str(source)], stdout=subprocess.PIPE, stderr=subprocess.PIPE).returncode == 0:
    lib = ctypes.CDLL(str(target))
torch.set_num_threads(1)   # no OpenMP team spinning next to the two libraries' own thread pools
n, f, d = args.n, args.f, args.d
G = torch.randn(n, d, generator=torch.Generator().manual_seed(0))
arr = np.ascontiguousarray(G.numpy())


def measure(ours, theirs):
  """Median ms of `reps` interleaved calls (ours, reference, ours, ...) after warm-up, so both see the same machine state."""
  for _ in range(5):
    ours()
    if theirs is not None:
      theirs()
  mine, other = [], []
  for _ in range(args.reps):
    begin = time.perf_counter()
    ours()
    mine.append((time.perf_counter() - begin) * 1e3)
    if theirs is not None:
      begin = time.perf_counter()
      theirs()
      other.append((time.perf_counter() - begin) * 1e3)
  middle = lambda values: sorted(values)[len(values) // 2]
  return middle(mine), (middle(other) if other else None)


def ref(name, *extra, scratch_rows=0):
  def call():
    work = arr.copy()  # the reference mutates its input (nth_element in place); the copy is part of what its py_func wrapper pays too
    out = np.empty(d, dtype=np.float32)
    selected = np.empty((max(scratch_rows, 1), d), dtype=np.float32)   # bulyan: the s selected gradients
    scratch = [ctypes.c_void_p(selected.ctypes.data)] if scratch_rows else []
    getattr(lib, name + "_float")(ctypes.c_size_t(d), ctypes.c_size_t(n), *[ctypes.c_size_t(e) for e in extra], ctypes.c_void_p(work.ctypes.data), *scratch, ctypes.c_void_p(out.ctypes.data))
  return call


m = n - f - 2
rows = {
  "median": (lambda: _ops.host_median(G), ref("median") if lib else None),
  "averaged-median": (lambda: _ops.host_averaged_median(G, n - f), ref("averaged_median", n - f) if lib else None),
  "average-nan": (lambda: _ops.host_average_nan(G), ref("average_nan") if lib else None),
  "bulyan": (lambda: _ops.host_bulyan(G, f, m), ref("bulyan", f, n - 2 * f - 2, scratch_rows=n - 2 * f - 2) if lib else None),
  "krum": (lambda: _ops.host_krum(G, f, m), None),
}
results = {"n": n, "f": f, "d": d, "threads": os.cpu_count(), "rules": {}}
deadline = time.perf_counter() + 3.0   # let the (virtual) cores spin up before anything is timed
while time.perf_counter() < deadline:
  _ops.host_average(G)
for name, (ours, theirs) in rows.items():
  mine, other = measure(ours, theirs)
  entry = {"ours_ms": round(mine, 3)}
  if other is not None:
    entry["reference_ms"] = round(other, 3)
    entry["speedup"] = round(other / mine, 2)
  results["rules"][name] = entry
  print(name, entry)
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
with open(args.out, "w") as fd:
  json.dump(results, fd, indent=1)

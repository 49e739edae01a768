"""Tracing and phase timing (reference: `tools/tf.py:41-58`, `graph.py:264-283`).

The reference wraps key graph ops with `tf.Print("[TRACE] (begin|end) what")`.
Here `Tracer.span(what)` prints the same markers around a host-side phase, opens
an NVTX range, and records CUDA events so that per-phase *device* time can be
reported (max over ranks) without host synchronisation inside the step.
"""

import contextlib
import sys
import time

__all__ = ["Tracer", "trace_graph", "device_from_tuple"]


def trace_graph(optn, what, stream=None):
  """Wrap the callable `optn` so that every call prints `[TRACE] (begin) what` / `[TRACE] (end)   what` around it
  (the eager counterpart of the reference's `tools.trace_graph`, which wrapped a graph op with `tf.Print` nodes)."""
  import functools

  @functools.wraps(optn)
  def traced(*args, **kwargs):
    out = stream if stream is not None else sys.stderr
    print("[TRACE] (begin) " + what, file=out)
    try:
      return optn(*args, **kwargs)
    finally:
      print("[TRACE] (end)   " + what, file=out)
  return traced


def device_from_tuple(job, taskid, devtype, devid):
  """`(job, task, type, id)` -> `/job:J/replica:0/task:T/device:TYPE:ID` (reference naming)."""
  return "/job:" + str(job) + "/replica:0/task:" + str(taskid) + "/device:" + str(devtype) + ":" + str(devid)


class Tracer:
  """Begin/end markers + NVTX + CUDA-event timing of named phases."""

  def __init__(self, enabled=False, cuda=False, stream=sys.stderr):
    self.enabled = enabled
    self.cuda = cuda
    self.stream = stream
    self._pending = []   # (what, start_event, stop_event)
    self._host = {}      # what -> [count, seconds]
    self._device = {}    # what -> [count, milliseconds]

  @contextlib.contextmanager
  def span(self, what, timed=True):
    if self.enabled:
      print("[TRACE] (begin) " + what, file=self.stream)
    nvtx = None
    start = stop = None
    if self.cuda:
      import torch
      nvtx = torch.cuda.nvtx
      nvtx.range_push(what)
      if timed:
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record()
    t0 = time.perf_counter()
    try:
      yield
    finally:
      entry = self._host.setdefault(what, [0, 0.0])
      entry[0] += 1
      entry[1] += time.perf_counter() - t0
      if start is not None:
        stop.record()
        self._pending.append((what, start, stop))
      if nvtx is not None:
        nvtx.range_pop()
      if self.enabled:
        print("[TRACE] (end)   " + what, file=self.stream)

  def collect(self):
    """Fold finished CUDA-event pairs into the per-phase device totals (call after a sync)."""
    for what, start, stop in self._pending:
      try:
        ms = start.elapsed_time(stop)
      except Exception:
        continue
      entry = self._device.setdefault(what, [0, 0.0])
      entry[0] += 1
      entry[1] += ms
    self._pending.clear()

  def report(self):
    """`{phase: {"count", "host_s", "device_ms"}}`."""
    self.collect()
    out = {}
    for what, (count, seconds) in self._host.items():
      out[what] = {"count": count, "host_s": seconds, "device_ms": self._device.get(what, [0, None])[1]}
    return out

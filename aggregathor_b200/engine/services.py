"""Evaluation / checkpoint / summary services (reference: the three threads of `runner.py:356-494`).

Each service fires when a step delta and/or a wall-clock period has elapsed, once more when training stops, and —
after a restore — starts its timers "now" instead of firing immediately (`runner.py:433-438,472-477`).

Two drivers share the `Cadence` logic:
* single-process runs use real threads named `test`, `checkpoint`, `summary` polling every `config.thread_idle_delay`
  seconds like the reference; they take the manager's lock, i.e. they run *between* steps (the static layer graph keeps
  per-module saved tensors, so a concurrent forward would corrupt a training step);
* multi-rank runs poll the same cadences on rank 0 inside the step loop and ship the decision to every rank in the
  flag lane of the per-step loss all-reduce, because checkpointing sharded optimizer slots is a collective.
"""

import os
import pathlib
import threading
import time

from .. import config, tools

FLAG_EVAL, FLAG_CHECKPOINT, FLAG_SUMMARY, FLAG_STOP = 1, 2, 4, 8


class Cadence:
  """Step-delta / period trigger with the reference's initial conditions."""

  def __init__(self, delta, period, restored=False, step=0):
    self.delta, self.period = delta, period
    if restored:
      self.last_step, self.last_time = step, time.time()
    else:
      self.last_step, self.last_time = -delta, -period

  @property
  def disabled(self):
    return self.delta < 0 and self.period < 0

  def due(self, step, now):
    return (self.delta >= 0 and step - self.last_step >= self.delta) or (self.period >= 0. and now - self.last_time >= self.period)

  def mark(self, step):
    self.last_step, self.last_time = step, time.time()


class Evaluator:
  """Runs `manager.evaluate()`, appends `walltime<TAB>step<TAB>name:value...` to the evaluation file, logs the metrics."""

  def __init__(self, manager, path):
    self.manager, self.fd = manager, None
    self.latest = {}
    if path:
      try:
        path = pathlib.Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        self.fd = path.open("a")
      except Exception:
        self.fd = None

  def run(self, step):
    begin = time.time()
    accuracies = self.manager.evaluate()
    self.latest = accuracies
    if self.fd is not None:
      self.fd.write(str(begin) + "\t" + str(step) + "".join("\t" + key + ":" + str(val) for key, val in accuracies.items()) + os.linesep)
      self.fd.flush()
    tools.info(" Step " + str(step) + ": " + ", ".join(key + " = " + str(val) for key, val in accuracies.items()) + " (took " + repr(time.time() - begin) + " s)")

  def close(self):
    if self.fd is not None:
      self.fd.close()


class Checkpointer:
  def __init__(self, manager, checkpoints, meta=None, write=True):
    self.manager, self.checkpoints, self.meta, self.write = manager, checkpoints, meta, write

  def run(self, step):
    begin = time.time()
    state = self.manager.state_dict()  # collective when optimizer slots are sharded
    if self.write:
      self.checkpoints.save(state, step, self.meta)
      tools.info("Checkpoint saved (took " + repr(time.time() - begin) + " s)")

  def close(self):
    pass


class Summarizer:
  """Scalars `learning_rate` and `eval-<metric>` in a TensorBoard event file + START/STOP session markers."""

  def __init__(self, manager, evaluator, path):
    self.manager, self.evaluator = manager, evaluator
    self.writer = tools.SummaryWriter(path)
    self.writer.add_session_log(tools.SummaryWriter.SESSION_START, manager.step)

  def run(self, step):
    begin = time.time()
    scalars = {"learning_rate": self.manager.rate(step)}
    if self.manager.total_loss is not None:
      scalars["total_loss"] = float(self.manager.total_loss)
    for key, val in (self.evaluator.latest if self.evaluator is not None else {}).items():
      scalars["eval-" + key] = val
    self.writer.add_scalars(scalars, step)
    tools.info("Summaries saved (took " + repr(time.time() - begin) + " s)")

  def close(self):
    self.writer.add_session_log(tools.SummaryWriter.SESSION_STOP, self.manager.step)
    self.writer.close()


class ServiceThread(threading.Thread):
  """Reference-style polling thread around one service."""

  def __init__(self, name, service, cadence, manager, lock, stop_event, first_event=None):
    super().__init__(name=name, daemon=True)
    self.service, self.cadence, self.manager, self.lock, self.stop_event, self.first_event = service, cadence, manager, lock, stop_event, first_event
    self.error = None

  def run(self):
    try:
      if self.cadence.disabled:
        tools.info({"test": "Evaluation is", "checkpoint": "Checkpoint saving is", "summary": "Summary saving is"}.get(self.name, self.name + " is") + " effectively disabled")
        return
      while True:
        stop = self.stop_event.wait(config.thread_idle_delay)
        step = self.manager.step
        if stop or self.cadence.due(step, time.time()):
          with self.lock:
            self.service.run(self.manager.step)
          if self.first_event is not None:
            self.first_event.set()
          self.cadence.mark(self.manager.step)
          if stop:
            break
    except BaseException as err:  # surfaced by the main loop
      self.error = err
    finally:
      if self.first_event is not None:
        self.first_event.set()
      self.service.close()

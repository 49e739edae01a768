import io
import os
import threading

import pytest
import torch

from aggregathor_b200 import cluster, tools


def test_parse_keyval_types_and_errors():
  parsed = tools.parse_keyval(["a:3", "b:x:y", "flag:true"], {"a": 1, "c": 2.5, "flag": False})
  assert parsed == {"a": 3, "b": "x:y", "c": 2.5, "flag": True}
  with pytest.raises(tools.UserException):
    tools.parse_keyval(["novalue"])
  with pytest.raises(tools.UserException):
    tools.parse_keyval(["a:1", "a:2"])
  with pytest.raises(tools.UserException):
    tools.parse_keyval(["a:notanint"], {"a": 1})


def test_class_register():
  reg = tools.ClassRegister("thing")
  reg.register("x", dict)
  assert list(reg.itemize()) == ["x"] and reg.instantiate("x", a=1) == {"a": 1}
  with pytest.raises(tools.UserException) as err:
    reg.instantiate("y")
  assert "available thing(s): 'x'" in str(err.value)
  with pytest.raises(AssertionError):
    reg.register("x", list)


def test_context_headers_and_threads():
  buf = io.StringIO()
  out = tools.ContextIOWrapper(buf, nocolor=True)
  with tools.Context("outer", "info"):
    with tools.Context("inner", None):
      out.write("a\nb\n")
    out.write("c")
    out.write("d\n")
  def worker():
    with tools.Context("w", None):
      out.write("t\n")
  thread = threading.Thread(target=worker, name="summary")
  thread.start()
  thread.join()
  assert buf.getvalue() == "[outer] [inner] a\n[outer] [inner] b\n[outer] cd\n[summary] [w] t\n"


def test_method_call_replicator():
  a, b = io.StringIO(), io.StringIO()
  tee = tools.MethodCallReplicator(a, b)
  tee.write("x")
  assert a.getvalue() == b.getvalue() == "x"


def test_make_interface():
  log = []
  cls = tools.make_interface(lambda v: {"v": v}, lambda h: log.append("destroyed"), get=lambda h: h["v"], add=lambda h, x: h["v"] + x)
  inst = cls(3)
  assert inst.get() == 3 and inst.add(4) == 7 and inst()["v"] == 3
  del inst
  assert log == ["destroyed"]


def test_checkpoints_layout_and_retention(tmp_path):
  ckpt = tools.Checkpoints(tmp_path, keep=2)
  assert not ckpt.can_restore()
  for step in (5, 10, 200):
    ckpt.save({"global_step": step, "params": torch.arange(4.0) + step}, step, {"note": "x"})
  names = sorted(os.listdir(tmp_path))
  assert names == ["model-10.data-00000-of-00001", "model-10.index", "model-10.meta", "model-200.data-00000-of-00001", "model-200.index", "model-200.meta"]
  assert ckpt.can_restore() and ckpt.latest().endswith("model-200")
  assert ckpt.restore()["global_step"] == 200
  assert [os.path.basename(p) for p in ckpt.get()] == ["model-10", "model-200"] and ckpt.get() == []


def test_summary_roundtrip(tmp_path):
  with tools.SummaryWriter(tmp_path) as writer:
    writer.add_session_log(tools.SummaryWriter.SESSION_START, 0)
    writer.add_scalars({"learning_rate": 0.05, "eval-top1-X-acc": 0.75}, 12)
  events = tools.read_events(writer.path)
  assert events[0]["file_version"] == "brain.Event:2" and events[1]["session_status"] == 1
  assert events[2]["step"] == 12 and abs(events[2]["scalars"]["eval-top1-X-acc"] - 0.75) < 1e-6


def test_cluster_parse_and_access(tmp_path, monkeypatch):
  assert tools.cluster_parse('{"ps": ["a:1"], "workers": ["b:2", "c:3"]}')["workers"] == ["b:2", "c:3"]
  nodes = tmp_path / "nodes"
  nodes.write_text("n1\nn1\nn2\nn3\n")
  monkeypatch.setenv("OAR_FILE_NODES", str(nodes))
  import aggregathor_b200.tools.cluster_spec as spec
  spec._g5k_cache = None
  assert tools.cluster_parse("G5k") == {"ps": ["n1:7000"], "workers": ["n2:7000", "n3:7000"]}
  with pytest.raises(tools.UserException):
    tools.cluster_parse("{not json")
  assert tools.can_access(tmp_path, read=True) and not tools.can_access(tmp_path / "missing")


def test_cluster_allocation_spread_and_reuse():
  mgr = cluster.Manager.from_world(4, [0, 1, 2, 3], None, "ps", "workers", "eval", devs=("GPU", "CPU"), reuse=("GPU", "CPU"))
  workers = mgr.allocate("worker", 8, jobs={"workers"})
  assert [int(t) for _, t, _, _ in workers] == [0, 1, 2, 3, 0, 1, 2, 3] and all(d == "GPU" for _, _, d, _ in workers)
  assert mgr.allocate("ps", 1, jobs={"ps"})[0][0] == "ps"
  strict = cluster.Manager.from_world(2, [0, 1], None, devs=("GPU", "CPU"), reuse=("CPU",))
  got = strict.allocate("worker", 3, jobs={"workers"})
  assert [d for _, _, d, _ in got] == ["GPU", "GPU", "CPU"]  # GPUs are exclusive without --reuse-gpu
  none = cluster.Manager.from_world(1, [0], None, devs=("GPU",), reuse=())
  assert none.allocate("worker", 2, jobs={"workers"}) is None
  assert len(none.allocate("worker", 2, jobs={"workers"}, partial=True)) == 1


def test_learning_rates_and_optimizer_table():
  from aggregathor_b200.engine import build, learning_rates, optimizers
  fixed = build(learning_rates, "learning rate decay", "fixed", ["initial-rate:0.05"])
  assert fixed(0) == fixed(1000) == 0.05
  poly = build(learning_rates, "learning rate decay", "polynomial", ["initial-rate:0.1", "end-rate:0.01", "decay-step:100", "power:1"])
  assert abs(poly(50) - 0.055) < 1e-9 and abs(poly(1000) - 0.01) < 1e-12
  expo = build(learning_rates, "learning rate decay", "exponential", ["initial-rate:0.1", "decay-step:10", "decay-rate:0.5"])
  assert abs(expo(20) - 0.025) < 1e-12
  with pytest.raises(tools.UserException):
    build(learning_rates, "learning rate decay", "cosine", [])
  adam = build(optimizers, "optimizer", "adam", ["adam-beta1:0.8", "unknown:1"])
  assert adam.hyper["beta1"] == 0.8 and adam.nbslots == 2


def test_clock_sampler_without_nvidia_smi():
  """On a box without nvidia-smi the sampler reports that instead of failing (bench.py records it in its JSON line)."""
  from aggregathor_b200.utils.clocks import ClockSampler
  sampler = ClockSampler(0)
  sampler.start()
  report = sampler.stop()
  assert set(report) >= {"sm_mhz", "sm_max_mhz", "reasons"}
  assert report["sm_mhz"] is None or report["sm_mhz"] > 0


def test_cadence_delta_and_period_semantics():
  """Step-delta and wall-clock triggers of the evaluation / checkpoint / summary services (reference: `runner.py:356-494`): fire at
  start, then every `delta` steps or `period` seconds, whichever comes first; both negative = never; a restored run counts from its step."""
  from aggregathor_b200.engine.services import Cadence
  cadence = Cadence(5, -1)
  assert not cadence.disabled and cadence.due(0, 100.0)                   # first trigger right away
  cadence.mark(0)
  assert not cadence.due(4, 1e9) and cadence.due(5, 0.0)
  timed = Cadence(-1, 10.0)
  assert timed.due(0, 0.0)
  timed.mark(0)
  mark_time = timed.last_time
  assert not timed.due(10 ** 6, mark_time + 9.9) and timed.due(1, mark_time + 10.0)
  assert Cadence(-1, -1).disabled and not Cadence(-1, -1).due(10 ** 6, 1e12)
  restored = Cadence(5, -1, restored=True, step=300)
  assert not restored.due(303, 0.0) and restored.due(305, 0.0)            # no immediate trigger after a restore

"""End-to-end CPU runs: runner CLI golden files, resume, attacks, the 2-process gloo plumbing config."""

import os
import pathlib
import re
import subprocess
import sys

import pytest
import torch

from aggregathor_b200 import aggregators, attacks, experiments, tools
from aggregathor_b200.engine.trainer import Manager

ROOT = pathlib.Path(__file__).resolve().parent.parent
LOCAL = ["--server", '{"local": ["127.0.0.1:7000"]}', "--ps-job-name", "local", "--wk-job-name", "local", "--ev-job-name", "local", "--no-wait"]


def _run(args, timeout=300, launcher=None):
  env = dict(os.environ, AGB_NUM_THREADS="2", OMP_NUM_THREADS="2")
  cmd = (launcher or [sys.executable]) + [str(ROOT / "runner.py")] + args
  proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, env=env, cwd=str(ROOT))
  return proc.returncode, proc.stdout.decode(errors="replace")


def test_runner_mnist_average_golden_outputs(tmp_path):
  ckpt = tmp_path / "ckpt"
  args = LOCAL + ["--experiment", "mnist", "--aggregator", "average", "--nb-workers", "4", "--max-step", "12", "--learning-rate-args", "initial-rate:0.05",
                  "--evaluation-delta", "1000", "--evaluation-period", "-1", "--checkpoint-dir", str(ckpt), "--checkpoint-delta", "1000", "--checkpoint-period", "-1",
                  "--summary-delta", "1000", "--summary-period", "-1", "--stdout-to", str(tmp_path / "out.txt")]
  code, out = _run(args)
  assert code == 0, out
  losses = [float(x) for x in re.findall(r"Step \d+: total loss = ([0-9.eE+-]+)", out)]
  assert len(losses) == 12 and losses[-1] < losses[0]
  assert "step(s)/s (all steps)" in out and "Cluster structure and allocation report" in out
  lines = (ckpt / "eval").read_text().strip().splitlines()
  assert len(lines) == 2  # first evaluation before training + final evaluation
  wall, step, metric = lines[-1].split("\t")
  assert int(step) == 12 and metric.startswith("top1-X-acc:") and float(wall) > 0
  names = sorted(p.name for p in ckpt.iterdir())
  assert "model-12.index" in names and "model-12.data-00000-of-00001" in names and "model-12.meta" in names and any(n.startswith("events.out.tfevents") for n in names)
  events = tools.read_events(next(p for p in ckpt.iterdir() if p.name.startswith("events")))
  assert events[1]["session_status"] == 1 and events[-1]["session_status"] == 2 and any("learning_rate" in e.get("scalars", {}) for e in events)
  plain = (tmp_path / "out.txt").read_text()
  assert "\033[" not in plain and "[train] Step 0: total loss" in plain
  # resume: --max-step counts additional steps from the restored global step
  code, out = _run(args[:args.index("--max-step") + 1] + ["3"] + args[args.index("--max-step") + 2:])
  assert code == 0, out
  assert "Loading latest checkpoint" in out and "Step 12: total loss" in out and "Step 14: total loss" in out and "Step 15:" not in out.replace("Step 15: top1", "")
  assert (ckpt / "model-15.index").exists()


def test_runner_argument_errors():
  code, out = _run(["--experiment", "mnist", "--aggregator", "average", "--nb-workers", "2"])
  assert code != 0 and "One and only one of '--client' and '--server'" in out
  code, out = _run(LOCAL + ["--experiment", "mnist", "--aggregator", "nope", "--nb-workers", "2", "--max-step", "1"])
  assert code != 0 and "Unknown name 'nope'" in out and "Traceback" not in out
  code, out = _run(LOCAL + ["--experiment", "mnist", "--aggregator", "bulyan", "--nb-workers", "8", "--nb-decl-byz-workers", "2", "--max-step", "1"])
  assert code != 0 and "Bulyan needs n >= 4 f + 3" in out


def test_two_process_gloo_plumbing(tmp_path):
  """BASELINE.json config 1: mnist + average, nb-workers = 2, CPU / gloo, one worker per rank."""
  port = 29500 + os.getpid() % 400
  launcher = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port)]
  args = ["--server", '{"ps": ["127.0.0.1:7000"], "workers": ["127.0.0.1:7001", "127.0.0.1:7002"], "eval": ["127.0.0.1:7000"]}', "--no-wait",
          "--experiment", "mnist", "--aggregator", "average", "--nb-workers", "2", "--max-step", "8", "--learning-rate-args", "initial-rate:0.05",
          "--evaluation-delta", "4", "--evaluation-period", "-1", "--checkpoint-dir", str(tmp_path / "c"), "--checkpoint-delta", "8", "--checkpoint-period", "-1",
          "--summary-dir", "-", "--debug-checksum"]
  code, out = _run(args, timeout=600, launcher=launcher)
  assert code == 0, out
  assert "Step 7: total loss" in out and "Replica divergence" not in out
  assert (tmp_path / "c" / "model-8.index").exists()
  assert len((tmp_path / "c" / "eval").read_text().strip().splitlines()) >= 2


_RATES = {"sgd": 0.05, "adam": 0.002, "rmsprop": 0.002, "adagrad": 0.02, "adadelta": 1.0}


def _manager(gar_name, n, f, attack=None, real=0, opt="sgd", exp="mnist", exp_args=("batch-size:16",), **kwargs):
  experiment = experiments.instantiate(exp, list(exp_args))
  gar = aggregators.instantiate(gar_name, n, f, [])
  return Manager(experiment, gar, n, opt, [], "fixed", ["initial-rate:" + str(_RATES[opt])], device="cpu", attack=attack, nb_real_byz=real, **kwargs)


@pytest.mark.parametrize("attack_name,attack_args", [("flip", ["factor:-50"]), ("nan", []), ("random", ["deviation:100"]), ("drop-chunks", ["rate:0.5", "fill:nan"]), ("replay", [])])
def test_krum_survives_attacks_average_does_not(attack_name, attack_args):
  attack = attacks.instantiate(attack_name, 7, 2, attack_args)
  robust = _manager("krum", 7, 2, attack, 2)
  first = float(robust.train())
  for _ in range(25):
    last = float(robust.train())
  assert last == last and last < first
  assert robust.evaluate()["top1-X-acc"] > 0.5
  if attack_name in ("flip", "nan", "random"):
    naive = _manager("average", 7, 2, attacks.instantiate(attack_name, 7, 2, attack_args), 2)
    for _ in range(25):
      loss = float(naive.train())
    assert not (loss == loss and naive.evaluate()["top1-X-acc"] > 0.5 and loss < first)


def test_average_nan_handles_lossy_transport():
  attack = attacks.instantiate("drop-chunks", 4, 2, ["rate:0.3", "fill:nan", "chunk-bytes:4000"])
  mgr = _manager("average-nan", 4, 2, attack, 2)
  losses = [float(mgr.train()) for _ in range(20)]
  assert all(l == l for l in losses) and losses[-1] < losses[0]


@pytest.mark.parametrize("opt", ["sgd", "adam", "rmsprop", "adagrad", "adadelta"])
def test_optimizers_decrease_loss_and_checkpoint_roundtrip(opt):
  mgr = _manager("median", 3, 0, opt=opt)
  losses = [float(mgr.train()) for _ in range(15)]
  assert losses[-1] < losses[0]
  state = mgr.state_dict()
  clone = _manager("median", 3, 0, opt=opt, seed=99)
  clone.load_state_dict(state)
  assert clone.step == 15 and torch.equal(clone.params, mgr.params)
  for a, b in zip(clone.aggregation.slots, mgr.aggregation.slots):
    assert torch.equal(a, b)


def test_mnist_shared_batch_gives_every_worker_the_same_batch():
  shared = experiments.instantiate("mnist", ["batch-size:8", "shared-batch:1"])
  a, b = shared.train_stream(0, 2, "cpu"), shared.train_stream(1, 2, "cpu")
  assert a is not b
  for _ in range(3):
    (xa, ya), (xb, yb) = next(a), next(b)
    assert torch.equal(xa, xb) and torch.equal(ya, yb)
  own = experiments.instantiate("mnist", ["batch-size:8"])
  (xa, _), (xb, _) = next(own.train_stream(0, 2, "cpu")), next(own.train_stream(1, 2, "cpu"))
  assert not torch.equal(xa, xb)


def test_rmsprop_first_step_matches_tensorflow():
  """TF 1.x RMSPropOptimizer: rms slot starts at 1, momentum at 0 => first update = lr * g / sqrt(0.9 + 0.1 g^2 + eps)."""
  from aggregathor_b200.engine.optimizers import optimizers
  from aggregathor_b200.engine.schedules import build
  spec = build(optimizers, "optimizer", "rmsprop", [])
  param, grad = torch.tensor([1.0, -2.0, 0.5]), torch.tensor([0.5, -4.0, 0.0])
  slots = spec.make_slots(param)
  assert torch.equal(slots[0], torch.ones(3)) and torch.equal(slots[1], torch.zeros(3))
  expected = param - 0.01 * grad / torch.sqrt(0.9 + 0.1 * grad * grad + 1e-10)
  spec.apply_torch(param, grad, slots, 0.01, 1)
  assert torch.allclose(param, expected, atol=1e-7)


def test_regularization_and_mnist_attack_experiment():
  mgr = _manager("average", 2, 0, regularizations=(1e-4, 1e-3))
  assert float(mgr.train()) > 0
  poisoned = _manager("krum", 7, 2, exp="mnistAttack", exp_args=("batch-size:16", "nb-byz:2"))
  for _ in range(25):
    loss = float(poisoned.train())
  assert loss == loss and poisoned.evaluate()["top1-X-acc"] > 0.5


def test_experiment_registry_covers_reference_names():
  names = set(experiments.itemize())
  assert {"mnist", "mnistAttack", "cnnet", "slim-resnet_v1_50-imagenet", "slim-resnet_v1_18-cifar10", "slim-vgg_16-imagenet", "slim-inception_v3-imagenet"} <= names
  assert len([name for name in names if name.startswith("slim-") and name.endswith("-imagenet")]) == 33
  assert experiments.instantiate("slim-inception_v3-imagenet", []).model().input_shape == (3, 299, 299)
  with pytest.raises(tools.UserException):
    experiments.instantiate("slim-nope-imagenet", [])


def test_cnnet_step_on_cpu():
  mgr = _manager("median", 3, 0, exp="cnnet", exp_args=("batch-size:4", "eval-batch-size:16"))
  a = float(mgr.train())
  b = float(mgr.train())
  assert a == a and b == b and 0.0 <= mgr.evaluate()["top1-X-acc"] <= 1.0


def test_authentication_rejects_forged_rows():
  """ed25519-signed gradient digests: honest rows pass, a row tampered with after signing is dropped (NaN) and Krum trains on."""
  honest = _manager("krum", 7, 2, authenticate=True)
  for _ in range(3):
    loss = float(honest.train())
  assert loss == loss and honest.authenticator.rejected_total == 0
  forged = _manager("krum", 7, 2, attacks.instantiate("forge", 7, 2, ["factor:-1e6", "fraction:0.5"]), 2, authenticate=True)
  first = float(forged.train())
  for _ in range(20):
    last = float(forged.train())
  assert forged.authenticator.rejected_total == 21 * 2  # two forging workers, one (single-rank) slice each, every step
  assert last == last and last < first and forged.evaluate()["top1-X-acc"] > 0.5
  # without authentication the same tampering reaches the aggregator as a plain attack
  naive = _manager("average", 7, 2, attacks.instantiate("forge", 7, 2, ["factor:-1e6", "fraction:0.5"]), 2)
  for _ in range(10):
    loss = float(naive.train())
  assert not (loss == loss and loss < first)


def test_authenticator_signature_checks():
  from aggregathor_b200.engine.flat import FlatLayout
  from aggregathor_b200.parallel.signing import Authenticator
  layout = FlatLayout()
  layout.add("theta", (1000,))
  layout.freeze()
  auth = Authenticator(layout, 3)
  rows = {i: torch.randn(layout.padded_size) for i in range(3)}
  records = auth.publish(5, list(rows.items()))
  assert auth.verify(5, rows, records, [0]) == []
  assert auth.verify(6, {0: rows[0].clone()}, records, [0]) == [(0, 0)]                      # replayed record of another step
  bad = dict(records)
  bad[1] = (records[1][0], bytes(64))
  probe = {1: rows[1].clone()}
  assert auth.verify(5, probe, bad, [0]) == [(1, 0)] and bool(torch.isnan(probe[1]).all())   # forged signature -> whole slice dropped
  missing = {2: rows[2].clone()}
  assert auth.verify(5, missing, {}, [0]) == [(2, 0)]                                        # no record at all


def test_two_process_authenticated_run(tmp_path):
  """2 ranks x 2 workers (gloo): one forging worker on rank 1; every rank must drop its slices and replicas stay identical."""
  port = 29100 + os.getpid() % 300
  launcher = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port)]
  args = ["--server", '{"ps": ["127.0.0.1:7000"], "workers": ["127.0.0.1:7001", "127.0.0.1:7002"], "eval": ["127.0.0.1:7000"]}', "--no-wait",
          "--experiment", "mnist", "--aggregator", "average-nan", "--nb-workers", "4", "--nb-decl-byz-workers", "1", "--nb-real-byz-workers", "1", "--attack", "forge",
          "--attack-args", "factor:-1e9", "fraction:0.2", "--authenticate", "--max-step", "6", "--learning-rate-args", "initial-rate:0.05", "--evaluation-file", "-",
          "--checkpoint-dir", str(tmp_path / "c"), "--checkpoint-delta", "-1", "--checkpoint-period", "-1", "--summary-dir", "-", "--debug-checksum"]
  code, out = _run(args, timeout=600, launcher=launcher)
  assert code == 0, out
  assert "failing authentication" in out and "Replica divergence" not in out and "Step 5: total loss" in out
  # only the tampered part is dropped: the forger touches the first 20 % of its row = slice 0 of 2, on every rank, every step
  dropped = re.findall(r"dropped (\d+) gradient slice\(s\) failing authentication: \[\((\d+), (\d+)\)\]", out)
  assert len(dropped) == 12 and all(count == "1" and piece == "0" for count, _, piece in dropped), dropped
  assert len({slot for _, slot, _ in dropped}) == 1
  losses = [float(x) for x in re.findall(r"Step \d+: total loss = ([0-9.eE+-]+)", out)]
  assert all(l == l and l < 100 for l in losses), losses


def test_deploy_local_cluster_with_lossy_workers(tmp_path):
  """`deploy.py --deploy` on a local cluster specification: two CPU ranks as child processes, `--UDP 1` turns one worker into a
  lossy-transport (drop-chunks) worker, the NaN-aware rule trains through it, the deployer exits 0 when its ranks are done."""
  port = 7200 + os.getpid() % 500
  cluster = '{"ps": ["127.0.0.1:%d"], "workers": ["127.0.0.1:%d", "127.0.0.1:%d"]}' % (port, port + 1, port + 2)
  runner = ("--experiment mnist --aggregator average-nan --nb-workers 4 --nb-decl-byz-workers 1 --max-step 5 --learning-rate-args initial-rate:0.05 "
            "--evaluation-file - --checkpoint-dir %s --checkpoint-delta -1 --checkpoint-period -1 --summary-dir -" % (tmp_path / "c"))
  env = dict(os.environ, AGB_NUM_THREADS="2", OMP_NUM_THREADS="2")
  proc = subprocess.run([sys.executable, str(ROOT / "deploy.py"), "--cluster", cluster, "--deploy", "--UDP", "1", "--MPI", "--runner", runner],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, env=env, cwd=str(ROOT))
  out = proc.stdout.decode(errors="replace")
  assert proc.returncode == 0, out[-3000:]
  assert "rank 0/2" in out and "rank 1/2" in out and "'--MPI' is accepted for compatibility" in out
  losses = [float(x) for x in re.findall(r"Step \d+: total loss = ([0-9.eE+-]+)", out)]
  assert len(losses) == 5 and all(l == l for l in losses) and losses[-1] < losses[0]
  assert "drop-chunks" in out


def test_deploy_ships_its_source_through_the_pipe(tmp_path):
  """NFS-free deployment (`--ship always`): every rank unpacks the tarball it receives on stdin into a private directory and runs from
  there (what a remote `ssh host` does; here through `sh -c`), no file of this checkout is read by the ranks."""
  port = 7700 + os.getpid() % 250
  cluster = '{"ps": ["127.0.0.1:%d"], "workers": ["127.0.0.1:%d", "127.0.0.1:%d"]}' % (port, port + 1, port + 2)
  runner = ("--experiment mnist --aggregator median --nb-workers 2 --max-step 3 --learning-rate-args initial-rate:0.05 "
            "--evaluation-file - --checkpoint-dir %s --checkpoint-delta -1 --checkpoint-period -1 --summary-dir -" % (tmp_path / "c"))
  env = dict(os.environ, AGB_NUM_THREADS="2", OMP_NUM_THREADS="2", AGB_PRINT_ROOT="1")
  proc = subprocess.run([sys.executable, str(ROOT / "deploy.py"), "--cluster", cluster, "--deploy", "--ship", "always", "--runner", runner],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, env=env, cwd=str(tmp_path))
  out = proc.stdout.decode(errors="replace")
  assert proc.returncode == 0, out[-3000:]
  assert len(re.findall(r"Step \d+: total loss = ([0-9.eE+-]+)", out)) == 3
  roots = set(re.findall(r"package root: (\S+)", out))
  assert roots and all("agb-rank-" in root for root in roots), roots


def test_slim_augmentation_options():
  """`augment:flip` / `augment:crop-flip`: every output image is a (possibly mirrored) window of the replicate-padded input, evaluation is untouched."""
  exp = experiments.instantiate("slim-resnet_v1_18-cifar10", ["batch-size:4", "synthetic-samples:64", "augment:crop-flip", "image-size:16"])
  assert exp.stochastic_preprocess
  generator = torch.Generator().manual_seed(1)
  images = torch.randint(0, 255, (6, 16, 16, 3), dtype=torch.uint8)
  out = exp._augment(images, generator)
  assert out.shape == images.shape and out.dtype == torch.uint8
  pad = 2
  padded = torch.nn.functional.pad(images.permute(0, 3, 1, 2), (pad,) * 4, mode="replicate").permute(0, 2, 3, 1)
  for b in range(6):
    windows = [padded[b, t:t + 16, l:l + 16] for t in range(2 * pad + 1) for l in range(2 * pad + 1)]
    assert any(torch.equal(out[b], w) or torch.equal(out[b], w.flip(1)) for w in windows)
  plain = experiments.instantiate("slim-resnet_v1_18-cifar10", ["batch-size:4", "synthetic-samples:64", "image-size:16"])
  assert not plain.stochastic_preprocess
  with pytest.raises(tools.UserException):
    experiments.instantiate("slim-resnet_v1_18-cifar10", ["augment:rotate"])


def test_engine_selection(monkeypatch):
  """`auto`: host engine on CPU; on CUDA the fused engine when the rule has a kernel and every rank shares one machine, the NCCL
  baseline when the ranks span hosts or the rule has no kernel (engines replaced by stand-ins: no GPU needed to check the choice)."""
  from aggregathor_b200.engine.flat import FlatLayout
  from aggregathor_b200.engine.optimizers import optimizers
  from aggregathor_b200.engine.schedules import build
  from aggregathor_b200.parallel import aggregation
  layout = FlatLayout()
  layout.add("theta", (10,))
  layout.freeze()
  sgd = build(optimizers, "optimizer", "sgd", [])
  made = []
  for name in ("FusedAggregation", "BaselineAggregation"):
    monkeypatch.setattr(aggregation, name, lambda *a, _name=name, **k: made.append(_name) or _name)
  krum = aggregators.instantiate("krum", 7, 2, [])
  assert aggregation.make_aggregation("auto", krum, layout, 7, sgd, device="cpu").name == "host"
  monkeypatch.setattr(aggregation, "_single_host", lambda group=None: True)
  assert aggregation.make_aggregation("auto", krum, layout, 7, sgd, device="cuda:0") == "FusedAggregation"
  monkeypatch.setattr(aggregation, "_single_host", lambda group=None: False)
  assert aggregation.make_aggregation("auto", krum, layout, 7, sgd, device="cuda:0") == "BaselineAggregation"
  assert aggregation.make_aggregation("fused", krum, layout, 7, sgd, device="cuda:0") == "FusedAggregation"      # an explicit choice is honoured

  class Custom(aggregators._GAR):
    def __init__(self):
      pass
    def aggregate(self, gradients):
      return gradients[0]
  monkeypatch.setattr(aggregation, "_single_host", lambda group=None: True)
  assert aggregation.make_aggregation("auto", Custom(), layout, 7, sgd, device="cuda:0") == "BaselineAggregation"
  assert aggregation._single_host.__name__ == "<lambda>"


def test_single_host_detection_across_processes(tmp_path):
  script = tmp_path / "probe.py"
  script.write_text("import sys, torch.distributed as dist\nsys.path.insert(0, %r)\nfrom aggregathor_b200.parallel.aggregation import _single_host\n"
                    "dist.init_process_group('gloo')\nsys.stdout.write('single host: ' + str(_single_host()) + chr(10))\nsys.stdout.flush()\ndist.destroy_process_group()\n" % str(ROOT))
  port = 29050 + os.getpid() % 40
  proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300, cwd=str(ROOT))
  out = proc.stdout.decode(errors="replace")
  assert proc.returncode == 0 and out.count("True") == 2 and "False" not in out, out[-2000:]   # the two ranks' lines may interleave


def test_sha256_tree_digest_specification():
  """The digest signed on GPUs, on the host: a SHA-256 tree over 1024-byte leaves whose nodes carry (level, index) headers."""
  import hashlib
  import struct
  from aggregathor_b200.ops.gar import sha256_tree_host
  leaf = bytes(range(256)) * 4
  assert sha256_tree_host(b"") == hashlib.sha256(struct.pack("<IIQ", 0, 0, 0)).digest()
  assert sha256_tree_host(leaf[:1000]) == hashlib.sha256(struct.pack("<IIQ", 0, 0, 0) + leaf[:1000]).digest()
  two = leaf + leaf[:4]
  nodes = hashlib.sha256(struct.pack("<IIQ", 0, 0, 0) + leaf).digest() + hashlib.sha256(struct.pack("<IIQ", 0, 0, 1) + leaf[:4]).digest()
  assert sha256_tree_host(two) == hashlib.sha256(struct.pack("<IIQ", 1, 0, 0) + nodes).digest()
  assert sha256_tree_host(leaf + leaf) != sha256_tree_host(leaf)                     # length extension changes the root
  swapped = leaf[4:8] + leaf[:4] + leaf[8:]
  assert sha256_tree_host(swapped) != sha256_tree_host(leaf)                         # not a commutative sum
  big = bytes(40000)
  assert len(sha256_tree_host(big)) == 32 and sha256_tree_host(big) != sha256_tree_host(bytes(40004))

#!/usr/bin/env python3
"""Headline benchmark: steps/s of slim ResNet-50 (v1) + Multi-Krum f=2 with n = 8 logical workers
(per-worker batch 32, 224x224, synthetic ImageNet-shaped data, random-init weights) on N B200s of one box.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 ...

n stays 8 whatever N is (Krum with f=2 needs n >= 5): with fewer GPUs several logical workers share a GPU, the
reference's `--reuse-gpu`. Total work per step is therefore fixed => "strong" scaling.

Arms
  --impl ours       fused sm_100a aggregation + native nn kernels. Unless `--no-baseline`, the SAME invocation then measures the
                    baseline arm on the same GPUs and fills `vs_baseline` = ours / baseline (device-timed and end-to-end):
                    BASELINE.md publishes no number, so the same-box, same-run, same-dtype baseline is the anchor.
  --impl baseline   reference-style path made of library kernels only (`baseline/reference_style.py`): torch.nn ResNet-50 in
                    channels_last + cuDNN/cuBLAS under CUDA-graph replay, NCCL all-gather of the flat gradients, ONE stand-alone
                    GAR kernel, separate SGD kernel.
  --impl reference  the unmodified TF1 reference: cannot be installed offline, prints `{"impl": "reference", "unavailable": ...}`.
`--dtype bf16|tf32` selects the compute precision of BOTH arms (tf32 = fp32 storage, TF32 tensor-core products: the parity
precision of the fp32 reference, `graph.py:267-273`).

Timing: W warm-up steps, then exactly K steps bracketed by barrier + cuda synchronize, CUDA events on the launching
stream, max over ranks. `value` = device-timed steps/s with inputs resident on the device; `e2e.value` = the same loop
through the public `Manager.train()` API with the per-step pinned-host -> device input copy and the device -> host
read of the loss inside the timed region.
"""

import argparse
import gc
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))


def parse():
  parser = argparse.ArgumentParser()
  parser.add_argument("--gpus", type=int, default=1)
  parser.add_argument("--steps", type=int, default=10)
  parser.add_argument("--warmup", type=int, default=3)
  parser.add_argument("--impl", type=str, default="ours", choices=("ours", "baseline", "baseline-eager", "reference"))
  parser.add_argument("--dtype", type=str, default="bf16", choices=("bf16", "tf32", "fp32"))
  parser.add_argument("--no-baseline", action="store_true", help="--impl ours: skip the in-run baseline arm (vs_baseline stays null)")
  parser.add_argument("--experiment", type=str, default="", help="experiment name overriding slim-<model>-<dataset> (e.g. cnnet: BASELINE.json configuration 2)")
  parser.add_argument("--model", type=str, default="resnet_v1_50")
  parser.add_argument("--dataset", type=str, default="imagenet")
  parser.add_argument("--aggregator", type=str, default="krum")
  parser.add_argument("--nb-workers", type=int, default=8)
  parser.add_argument("--nb-decl-byz-workers", type=int, default=2)
  parser.add_argument("--nb-real-byz-workers", type=int, default=0, help="the last k logical workers run --attack (BASELINE.json config 5: 2 + flip)")
  parser.add_argument("--attack", type=str, default="")
  parser.add_argument("--attack-args", nargs="*", default=[])
  parser.add_argument("--batch-size", type=int, default=32, help="per logical worker")
  parser.add_argument("--image-size", type=int, default=0)
  parser.add_argument("--nn-backend", type=str, default="auto")
  parser.add_argument("--engine", type=str, default="")
  parser.add_argument("--skip-e2e", action="store_true")
  return parser.parse_args()


class Ring:
  """Device-resident batches served round-robin (the device-timed arm: no host traffic inside the timed region)."""

  def __init__(self, items):
    self.items, self.i = items, 0

  def __next__(self):
    self.i += 1
    return self.items[self.i % len(self.items)]


class Harness:
  """Process-group set-up, timing and reporting shared by the arms."""

  def __init__(self, args):
    import torch
    import torch.distributed as dist
    self.torch, self.dist, self.args = torch, dist, args
    self.world = int(os.environ.get("WORLD_SIZE", "1"))
    self.rank = int(os.environ.get("RANK", "0"))
    self.local = int(os.environ.get("LOCAL_RANK", "0"))
    if self.world != args.gpus and self.world == 1 and args.gpus > 1:
      raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
      raise SystemExit("bench.py needs a CUDA device")
    self.device = torch.device("cuda", self.local)
    torch.cuda.set_device(self.device)
    if self.world > 1:
      os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
      dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=self.device)

  def sync(self):
    self.torch.cuda.synchronize(self.device)
    if self.world > 1:
      self.dist.barrier()
      self.torch.cuda.synchronize(self.device)

  def timed(self, run_step, steps):
    """(device ms, wall ms) of `steps` calls, max over ranks."""
    torch = self.torch
    self.sync()
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall = time.perf_counter()
    begin.record()
    for _ in range(steps):
      run_step()
    end.record()
    self.sync()
    wall = time.perf_counter() - wall
    ms = torch.tensor([begin.elapsed_time(end), wall * 1000.0], dtype=torch.float64, device=self.device)
    if self.world > 1:
      self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
    return float(ms[0]), float(ms[1])

  def measure(self, step_resident, step_e2e, h2d_bytes, counter=None):
    """Warm-up + timed device-resident loop (+ clocks), then warm-up + timed end-to-end loop."""
    from aggregathor_b200.utils.clocks import ClockSampler
    args, torch = self.args, self.torch
    for _ in range(args.warmup):
      step_resident()
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else self.local)
    before = counter() if counter else 0
    if self.rank == 0:
      sampler.start()
    dev_ms, _ = self.timed(step_resident, args.steps)
    clocks = sampler.stop() if self.rank == 0 else None
    launches = (counter() - before) if counter else 0
    result = {"value": args.steps / (dev_ms / 1000.0), "ms_per_step": dev_ms / args.steps, "clocks": clocks, "gpu_launches": launches, "e2e": None}
    if not args.skip_e2e and step_e2e is not None:
      losses = []
      for _ in range(args.warmup):
        losses.append(step_e2e())
      e2e_dev_ms, e2e_wall_ms = self.timed(lambda: losses.append(step_e2e()), args.steps)
      e2e_ms = max(e2e_dev_ms, e2e_wall_ms)
      h2d = torch.tensor([float(h2d_bytes)], dtype=torch.float64, device=self.device)
      if self.world > 1:
        self.dist.all_reduce(h2d)
      result["e2e"] = {"value": args.steps / (e2e_ms / 1000.0), "unit": "steps/s", "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(h2d.item()),
                       "d2h_bytes_per_step": 4 * self.world, "last_loss": losses[-1] if losses else None}
    return result


def make_experiment(args):
  from aggregathor_b200 import experiments
  exp_args = ["batch-size:" + str(args.batch_size), "synthetic-samples:" + str(max(256, 4 * args.batch_size))]
  if args.image_size and not args.experiment:
    exp_args.append("image-size:" + str(args.image_size))
  name = args.experiment or ("slim-" + args.model + "-" + args.dataset)
  return name, experiments.instantiate(name, exp_args)


def run_ours(harness, args, impl):
  """The product (impl "ours") or the round-1 eager baseline engine ("baseline-eager": our layer graph on aten ops + NCCL all-gather)."""
  import torch
  from aggregathor_b200 import aggregators, attacks
  from aggregathor_b200.engine.trainer import Manager
  from aggregathor_b200.ops import counters
  n, f = args.nb_workers, args.nb_decl_byz_workers
  name, experiment = make_experiment(args)
  gar = aggregators.instantiate(args.aggregator, n, f, [])
  engine = args.engine or ("fused" if impl == "ours" else "baseline")
  backend = args.nn_backend if impl == "ours" else "torch"
  attack = attacks.instantiate(args.attack, n, f, args.attack_args) if (args.attack and args.nb_real_byz_workers > 0) else None
  dtype = {"bf16": torch.bfloat16, "tf32": torch.float32, "fp32": torch.float32}[args.dtype]
  manager = Manager(experiment, gar, n, "sgd", [], "fixed", ["initial-rate:0.01"], device=harness.device, engine=engine, backend=backend, seed=0,
                    attack=attack, nb_real_byz=args.nb_real_byz_workers if attack is not None else 0, dtype=dtype)
  e2e_streams = manager.streams
  resident = []
  for stream in e2e_streams:
    ring = [next(stream) for _ in range(2)]
    resident.append(Ring([(x.clone(), y.clone()) for x, y in ring]))

  def step_resident():
    manager.streams = resident
    manager.train()

  def step_e2e():
    manager.streams = e2e_streams
    return float(manager.train())   # .item(): device -> host read of the step's result

  result = harness.measure(step_resident, step_e2e, manager.h2d_bytes_per_step, lambda: counters.launches)
  result.update({"experiment": name, "engine": manager.aggregation.name, "nn_backend": manager.backend, "d": manager.layout.size,
                 "image_size": manager.model.input_shape[-1], "attack": attack is not None,
                 "dtype": args.dtype if dtype == torch.float32 else "bf16"})
  manager.streams = e2e_streams
  manager.close()
  del manager, resident, e2e_streams
  return result


def run_baseline(harness, args):
  """Library-kernel reference-style arm (`baseline/reference_style.py`)."""
  import torch
  sys.path.insert(0, os.path.join(ROOT, "baseline"))
  from reference_style import ReferenceStyleTrainer
  from aggregathor_b200 import aggregators
  from aggregathor_b200.ops import counters
  n, f = args.nb_workers, args.nb_decl_byz_workers
  if args.experiment or args.model != "resnet_v1_50":
    return {"unavailable": "the library baseline arm implements slim-resnet_v1_50 only"}
  name, experiment = make_experiment(args)
  gar = aggregators.instantiate(args.aggregator, n, f, [])
  spec = gar.fused_spec()
  image_size = args.image_size or 224
  trainer = ReferenceStyleTrainer(n, spec, args.batch_size, image_size, harness.device, precision=args.dtype, lr=0.01, num_classes=experiment.num_classes)
  first = harness.rank * trainer.w
  streams = [experiment.train_stream(first + j, n, harness.device) for j in range(trainer.w)]
  resident = [Ring([tuple(t.clone() for t in next(stream)) for _ in range(2)]) for stream in streams]

  def step_resident():
    trainer.train([next(r) for r in resident])

  def step_e2e():
    return float(trainer.train([next(s) for s in streams]))

  h2d = sum(getattr(s, "h2d_bytes", 0) for s in streams)
  result = harness.measure(step_resident, step_e2e, h2d, lambda: counters.launches)
  result.update({"experiment": name, "engine": "NCCL all_gather_into_tensor + stand-alone GAR kernel + SGD kernel", "nn_backend": "torch.nn + cuDNN/cuBLAS, channels_last, CUDA-graph replay",
                 "d": trainer.d, "image_size": image_size, "attack": False, "dtype": args.dtype})
  for stream in streams:
    close = getattr(stream, "close", None)
    if close is not None:
      close()
  del trainer, resident
  return result


def main():
  args = parse()
  if args.impl == "reference":
    print(json.dumps({"impl": "reference", "unavailable": "LPD-EPFL/AggregaThor needs TensorFlow 1.10 / Python 3.5 and has no setup.py/pyproject: "
                      "`pip install --no-index --target baseline/_ref /root/reference` fails ('not installable') and tensorflow is absent from /opt/wheelhouse"}))
    return 0
  sys.path.insert(0, ROOT)
  harness = Harness(args)
  import torch
  from aggregathor_b200 import tools
  if harness.rank != 0:
    tools.set_rank_tag("r" + str(harness.rank))
    sys.stdout = open(os.devnull, "w")  # rank 0 reports

  n, f = args.nb_workers, args.nb_decl_byz_workers
  main_arm = run_baseline(harness, args) if args.impl == "baseline" else run_ours(harness, args, args.impl)
  if "unavailable" in main_arm:
    raise SystemExit(main_arm["unavailable"])
  baseline_arm = None
  if args.impl == "ours" and not args.no_baseline:
    gc.collect()
    torch.cuda.empty_cache()
    try:
      baseline_arm = run_baseline(harness, args)
    except Exception as err:   # the anchor is optional: never lose the product's own number
      baseline_arm = {"unavailable": type(err).__name__ + ": " + str(err).splitlines()[0][:200]}

  experiment_name = main_arm["experiment"]
  world = harness.world
  headline = (experiment_name, args.aggregator, n, f, main_arm["attack"]) == ("slim-resnet_v1_50-imagenet", "krum", 8, 2, False)
  metric = "steps/sec (whole box, device-timed, max over ranks) ResNet-50 slim + Krum f=2" if headline else (
    "steps/sec (whole box, device-timed, max over ranks) %s + %s n=%d f=%d%s" % (experiment_name, args.aggregator, n, f, (" attack=" + args.attack) if main_arm["attack"] else ""))
  if harness.rank == 0:
    sys.stdout = sys.__stdout__
    value, e2e = main_arm["value"], main_arm["e2e"]
    vs_baseline, baseline_report = None, None
    if baseline_arm is not None:
      if "unavailable" in baseline_arm:
        baseline_report = baseline_arm
      else:
        vs_baseline = value / baseline_arm["value"]
        baseline_report = {"what": "same run, same GPUs, same dtype: " + baseline_arm["nn_backend"] + " + " + baseline_arm["engine"],
                           "value": baseline_arm["value"], "ms_per_step": baseline_arm["ms_per_step"], "dtype": baseline_arm["dtype"],
                           "e2e_value": baseline_arm["e2e"]["value"] if baseline_arm["e2e"] else None, "clocks": baseline_arm["clocks"]}
        if e2e and baseline_arm["e2e"]:
          e2e["vs_baseline"] = e2e["value"] / baseline_arm["e2e"]["value"]
    line = {
      "metric": metric, "value": value, "unit": "steps/s", "n_gpus": world,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_arm["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": vs_baseline,
      "dtype": main_arm["dtype"], "data": "synthetic (ImageNet-shaped uint8 images, random-init weights)", "impl": args.impl,
      "config": {"model": experiment_name, "aggregator": args.aggregator, "nb_workers": n, "nb_decl_byz_workers": f,
                 "global_batch": n * args.batch_size, "per_worker_batch": args.batch_size, "image_size": main_arm["image_size"], "seq_len": None,
                 "parallelism": "dp%d (x%d logical workers per GPU)" % (world, n // world), "engine": main_arm["engine"], "nn_backend": main_arm["nn_backend"],
                 "d": main_arm["d"], "l2": "per-step working set (activations + 8 x 102 MB gradients) exceeds the 126 MB L2; no explicit flush",
                 "vs_baseline_is": "value / the in-run baseline arm's value (BASELINE.md publishes no number)"},
      "clocks": main_arm["clocks"], "e2e": e2e, "gpu_launches": main_arm["gpu_launches"],
      "images_per_s": value * n * args.batch_size, "baseline": baseline_report}
    print(json.dumps(line))
  if world > 1:
    harness.dist.barrier()
    harness.dist.destroy_process_group()
  return 0


if __name__ == "__main__":
  sys.exit(main())

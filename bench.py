#!/usr/bin/env python3
"""Headline benchmark: steps/s of slim ResNet-50 (v1) + Multi-Krum f=2 with n = 8 logical workers
(per-worker batch 32, 224x224, synthetic ImageNet-shaped data, random-init weights) on N B200s of one box.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 ...

n stays 8 whatever N is (Krum with f=2 needs n >= 5): with fewer GPUs several logical workers share a GPU, the
reference's `--reuse-gpu`. Total work per step is therefore fixed => "strong" scaling.

Arms: `--impl ours` (fused sm_100a aggregation + native nn kernels), `--impl baseline` (our reference-style path: NCCL
all-gather + stand-alone GAR kernel + separate update kernel + torch/cuDNN model ops), `--impl reference` (the
unmodified TF1 reference: cannot be installed offline, prints `{"impl": "reference", "unavailable": ...}`).

Timing: W warm-up steps, then exactly K steps bracketed by barrier + cuda synchronize, CUDA events on the launching
stream, max over ranks. `value` = device-timed steps/s with inputs resident on the device; `e2e.value` = the same loop
through the public `Manager.train()` API with the per-step pinned-host -> device input copy and the device -> host
read of the loss inside the timed region.
"""

import argparse
import json
import os
import sys
import time


def parse():
  parser = argparse.ArgumentParser()
  parser.add_argument("--gpus", type=int, default=1)
  parser.add_argument("--steps", type=int, default=10)
  parser.add_argument("--warmup", type=int, default=3)
  parser.add_argument("--impl", type=str, default="ours", choices=("ours", "baseline", "reference"))
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


def main():
  args = parse()
  if args.impl == "reference":
    print(json.dumps({"impl": "reference", "unavailable": "LPD-EPFL/AggregaThor needs TensorFlow 1.10 / Python 3.5 and has no setup.py/pyproject: "
                      "`pip install --no-index --target baseline/_ref /root/reference` fails ('not installable') and tensorflow is absent from /opt/wheelhouse"}))
    return 0
  import torch
  import torch.distributed as dist
  sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
  from aggregathor_b200 import aggregators, attacks, experiments, tools
  from aggregathor_b200.utils.clocks import ClockSampler
  from aggregathor_b200.engine.trainer import Manager
  from aggregathor_b200.ops import counters

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if world != args.gpus:
    if world == 1 and args.gpus > 1:
      raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device")
  device = torch.device("cuda", local)
  torch.cuda.set_device(device)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
  if rank != 0:
    tools.set_rank_tag("r" + str(rank))
    sys.stdout = open(os.devnull, "w")  # rank 0 reports

  n, f = args.nb_workers, args.nb_decl_byz_workers
  exp_args = ["batch-size:" + str(args.batch_size), "synthetic-samples:" + str(max(256, 4 * args.batch_size))]
  if args.image_size:
    exp_args.append("image-size:" + str(args.image_size))
  experiment_name = args.experiment or ("slim-" + args.model + "-" + args.dataset)
  if args.experiment:
    exp_args = [a for a in exp_args if not a.startswith("image-size:")]
  experiment = experiments.instantiate(experiment_name, exp_args)
  gar = aggregators.instantiate(args.aggregator, n, f, [])
  engine = args.engine or ("fused" if args.impl == "ours" else "baseline")
  backend = args.nn_backend if args.impl == "ours" else "torch"
  attack = attacks.instantiate(args.attack, n, f, args.attack_args) if (args.attack and args.nb_real_byz_workers > 0) else None
  manager = Manager(experiment, gar, n, "sgd", [], "fixed", ["initial-rate:0.01"], device=device, engine=engine, backend=backend, seed=0,
                    attack=attack, nb_real_byz=args.nb_real_byz_workers if attack is not None else 0)

  def sync():
    torch.cuda.synchronize(device)
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize(device)

  def timed(run_step, steps):
    sync()
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall = time.perf_counter()
    begin.record()
    for _ in range(steps):
      run_step()
    end.record()
    sync()
    wall = time.perf_counter() - wall
    ms = torch.tensor([begin.elapsed_time(end), wall * 1000.0], dtype=torch.float64, device=device)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms[0]), float(ms[1])

  # ---- device-resident arm: inputs already on the device, loss stays on the device ---- #
  e2e_streams = manager.streams
  resident = []
  for stream in e2e_streams:
    ring = [next(stream) for _ in range(2)]
    ring = [(x.clone(), y.clone()) for x, y in ring]
    resident.append(ring)

  class Ring:
    def __init__(self, items):
      self.items, self.i = items, 0

    def __next__(self):
      self.i += 1
      return self.items[self.i % len(self.items)]

  manager.streams = [Ring(items) for items in resident]
  for _ in range(args.warmup):
    manager.train()
  sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
  launches_before = counters.launches
  if rank == 0:
    sampler.start()
  dev_ms, _ = timed(manager.train, args.steps)
  clocks = sampler.stop() if rank == 0 else None
  launches = counters.launches - launches_before
  value = args.steps / (dev_ms / 1000.0)

  # ---- end-to-end arm: public API, pinned-host input copy + loss read-back every step ---- #
  e2e = None
  if not args.skip_e2e:
    manager.streams = e2e_streams
    losses = []

    def e2e_step():
      losses.append(float(manager.train()))  # .item(): device -> host read of the step's result
    for _ in range(args.warmup):
      e2e_step()
    e2e_dev_ms, e2e_wall_ms = timed(e2e_step, args.steps)
    e2e_ms = max(e2e_dev_ms, e2e_wall_ms)
    h2d = torch.tensor([float(manager.h2d_bytes_per_step)], dtype=torch.float64, device=device)
    if world > 1:
      dist.all_reduce(h2d)
    e2e = {"value": args.steps / (e2e_ms / 1000.0), "unit": "steps/s", "ms_per_step": e2e_ms / args.steps,
           "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": 4 * world, "last_loss": losses[-1] if losses else None}

  headline = (experiment_name, args.aggregator, n, f, attack) == ("slim-resnet_v1_50-imagenet", "krum", 8, 2, None)
  metric = "steps/sec (whole box, device-timed, max over ranks) ResNet-50 slim + Krum f=2" if headline else (
    "steps/sec (whole box, device-timed, max over ranks) %s + %s n=%d f=%d%s" % (experiment_name, args.aggregator, n, f, (" attack=" + args.attack) if attack is not None else ""))
  if rank == 0:
    sys.stdout = sys.__stdout__
    line = {
      "metric": metric, "value": value, "unit": "steps/s", "n_gpus": world,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
      "dtype": {"bfloat16": "bf16", "float16": "fp16", "float32": "fp32"}.get(str(manager.dtype).replace("torch.", ""), str(manager.dtype)), "data": "synthetic (ImageNet-shaped uint8 images, random-init weights)", "impl": args.impl,
      "config": {"model": experiment_name, "aggregator": args.aggregator, "nb_workers": n, "nb_decl_byz_workers": f,
                 "global_batch": n * args.batch_size, "per_worker_batch": args.batch_size, "image_size": manager.model.input_shape[-1], "seq_len": None,
                 "parallelism": "dp%d (x%d logical workers per GPU)" % (world, n // world), "engine": manager.aggregation.name, "nn_backend": manager.backend,
                 "d": manager.layout.size, "l2": "per-step working set (activations + 8 x 102 MB gradients) exceeds the 126 MB L2; no explicit flush"},
      "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
      "images_per_s": value * n * args.batch_size}
    print(json.dumps(line))
  manager.close()
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  return 0


if __name__ == "__main__":
  sys.exit(main())

#!/usr/bin/env python3
"""One eager training step of a single worker between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum` launch lists (per-kernel device time)."""
import argparse
import os
import sys

import torch

os.environ.setdefault("AGB_NO_GRAPH", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aggregathor_b200 import aggregators, experiments  # noqa: E402
from aggregathor_b200.engine.trainer import Manager  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--model", default="resnet_v1_50")
parser.add_argument("--batch-size", type=int, default=32)
parser.add_argument("--nn-backend", default="native")
parser.add_argument("--workers", type=int, default=1)
args = parser.parse_args()
experiment = experiments.instantiate("slim-" + args.model + "-imagenet", ["batch-size:" + str(args.batch_size), "synthetic-samples:128"])
gar = aggregators.instantiate("average", args.workers, 0, [])
manager = Manager(experiment, gar, args.workers, "sgd", [], "fixed", ["initial-rate:0.01"], device="cuda:0", engine="fused", backend=args.nn_backend)
for _ in range(3):
  manager.train()
torch.cuda.synchronize()
torch.cuda.profiler.start()
manager.train()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
begin.record()
for _ in range(5):
  manager.train()
end.record()
torch.cuda.synchronize()
print("eager ms/step (%d worker(s) x batch %d): %.3f" % (args.workers, args.batch_size, begin.elapsed_time(end) / 5))
manager.close()

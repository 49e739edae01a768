"""Device-side timing helpers shared by the benchmarks: CUDA events on the launching stream, maximum over the ranks."""

import torch
import torch.distributed as dist


def synchronize(device):
  """Barrier across the ranks (when there are several) with a device synchronisation on both sides."""
  torch.cuda.synchronize(device)
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    dist.barrier()
    torch.cuda.synchronize(device)


def device_ms(fn, iterations, device, warmup=3):
  """Milliseconds per call of `fn`, timed with CUDA events after `warmup` untimed calls; the slowest rank's figure on every rank."""
  for _ in range(warmup):
    fn()
  synchronize(device)
  start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  start.record()
  for _ in range(iterations):
    fn()
  stop.record()
  synchronize(device)
  elapsed = torch.tensor([start.elapsed_time(stop) / iterations], dtype=torch.float64, device=device)
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
  return float(elapsed)

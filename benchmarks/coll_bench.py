#!/usr/bin/env python3
"""Own collectives (`aggregathor_b200.parallel.collectives`) vs `torch.distributed` (NCCL on GPUs, gloo on CPU).

Launch with torchrun for N > 1. Checks all-reduce (sum / mean; fp32, int32, int64) and all-gather (equal and different first
dimensions, rows that are not multiples of 16 bytes) against the library collectives, then — on GPUs — times both with CUDA
events (max over ranks) and reports algorithm bandwidth. Writes `<out>/coll_bench_<N>.json` from rank 0; exits non-zero on
any mismatch.
"""

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from aggregathor_b200.parallel import collectives  # noqa: E402


def _time(fn, iters, device):
  for _ in range(3):
    fn()
  torch.cuda.synchronize(device)
  dist.barrier()
  start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  start.record()
  for _ in range(iters):
    fn()
  stop.record()
  torch.cuda.synchronize(device)
  ms = torch.tensor([start.elapsed_time(stop) / iters], device=device)
  dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  return float(ms)


def main():
  parser = argparse.ArgumentParser()
  parser.add_argument("--coll-device", dest="device", default="cuda")
  parser.add_argument("--coll-numel", dest="numel", type=int, default=25557032)
  parser.add_argument("--coll-iters", dest="iters", type=int, default=20)
  parser.add_argument("--coll-out", dest="out", default="gpurun_out")
  args = parser.parse_args()
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  on_gpu = args.device == "cuda"
  device = torch.device("cuda", local) if on_gpu else torch.device("cpu")
  if on_gpu:
    torch.cuda.set_device(device)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if on_gpu:
      dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    else:
      dist.init_process_group("gloo", rank=rank, world_size=world)
  comm = collectives.init(capacity=max(args.numel * 4, 1 << 20) + 4096, device=device)
  assert collectives.size() == world and collectives.rank() == rank and collectives.local_rank() == local
  failures = []
  generator = torch.Generator().manual_seed(1234 + rank)

  def reference_allreduce(t):
    out = t.clone()
    if world > 1:
      dist.all_reduce(out)
    return out

  # ---- all-reduce ------------------------------------------------------------- #
  for dtype, numel in ((torch.float32, 1000003), (torch.float32, 7), (torch.int32, 4099), (torch.int64, 513)):
    if dtype == torch.float32:
      t = torch.randn(numel, generator=generator).to(device)
    else:
      t = torch.randint(-1000, 1000, (numel,), generator=generator, dtype=dtype).to(device)
    got, want = collectives.allreduce(t), reference_allreduce(t)
    err = float((got.double() - want.double()).abs().max())
    tol = 1e-4 if dtype == torch.float32 else 0
    if got.shape != t.shape or got.dtype != dtype or err > tol:
      failures.append("allreduce %s[%d]: error %g" % (dtype, numel, err))
    if dtype == torch.float32:
      mean = collectives.allreduce(t.view(-1, 1), mean=True)
      if mean.shape != (numel, 1) or float((mean.view(-1) - want / world).abs().max()) > 1e-4:
        failures.append("allreduce mean %s[%d]" % (dtype, numel))
  # every rank must hold the same bits
  t = torch.randn(65537, generator=generator).to(device)
  got = collectives.allreduce(t)
  digest = got.view(torch.int32).to(torch.int64).sum().reshape(1)
  digests = [torch.zeros_like(digest) for _ in range(world)]
  if world > 1:
    dist.all_gather(digests, digest)
    if any(int(x) != int(digest) for x in digests):
      failures.append("allreduce: replicas differ")

  # ---- all-gather --------------------------------------------------------------- #
  cases = ((lambda r: 5, 3, torch.float32), (lambda r: r + 1, 7, torch.float32), (lambda r: 2 * r + 3, 1, torch.int64), (lambda r: 4, 33, torch.bfloat16 if on_gpu else torch.float32))
  for rows_of, cols, dtype in cases:
    rows, counts = rows_of(rank), [rows_of(r) for r in range(world)]
    t = (torch.arange(rows * cols, dtype=torch.float32).reshape(rows, cols) + 1000 * rank).to(dtype).to(device)
    got = collectives.allgather(t)
    if not torch.equal(got, collectives.allgather(t, counts)):
      failures.append("allgather with known counts differs, rows=%s" % counts)
    want = torch.cat([(torch.arange(c * cols, dtype=torch.float32).reshape(c, cols) + 1000 * r).to(dtype) for r, c in enumerate(counts)]).to(device)
    if got.shape != want.shape or got.dtype != dtype or not torch.equal(got, want):
      failures.append("allgather rows=%s cols=%d %s" % (counts, cols, dtype))

  # ---- timing (GPU, world > 1) ---------------------------------------------------- #
  report = {"world": world, "device": args.device, "multicast": bool(getattr(comm, "multicast", False)), "failures": failures}
  if on_gpu and world > 1 and not failures:
    numel = args.numel
    staged = comm.buffer(numel)
    staged.normal_()
    plain = torch.randn(numel, device=device)
    ours = _time(lambda: comm.allreduce_(numel), args.iters, device)
    nccl = _time(lambda: dist.all_reduce(plain), args.iters, device)
    nbytes = numel * 4
    block = torch.randn(numel // world, device=device)
    outs = torch.empty(numel // world * world, device=device)
    ours_ag = _time(lambda: comm.allgather(block), args.iters, device)
    known = [numel // world] * world
    ours_ag_known = _time(lambda: comm.allgather(block, known), args.iters, device)
    nccl_ag = _time(lambda: dist.all_gather_into_tensor(outs, block), args.iters, device)
    bus = 2.0 * (world - 1) / world * nbytes
    report.update({
      "allreduce": {"bytes": nbytes, "ours_ms": ours, "nccl_ms": nccl, "ours_busbw_gbs": bus / ours / 1e6, "nccl_busbw_gbs": bus / nccl / 1e6, "speedup": nccl / ours},
      "allgather": {"bytes_out": numel // world * world * 4, "ours_ms_including_staging_copy": ours_ag, "ours_ms_known_sizes": ours_ag_known, "nccl_ms": nccl_ag}})
  if failures:
    print("[rank %d] FAILED: %s" % (rank, "; ".join(failures)), file=sys.stderr)
  if rank == 0:
    os.makedirs(args.out, exist_ok=True)
    with open(os.path.join(args.out, "coll_bench_%d.json" % world), "w") as fd:
      json.dump(report, fd, indent=1)
    print(json.dumps(report))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  sys.exit(1 if failures else 0)


if __name__ == "__main__":
  main()

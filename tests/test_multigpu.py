"""Multi-GPU checks (need >= 2 CUDA devices; skipped otherwise): the fused P2P aggregation kernel must give every rank
bit-identical parameters that match the NCCL all-gather baseline, for every rule, and a full training step must keep
replicas identical."""

import json
import os
import pathlib
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


def _gpus():
  return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(nproc, script_args, timeout=600):
  port = 29700 + os.getpid() % 200
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
  proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, cwd=str(ROOT))
  return proc.returncode, proc.stdout.decode(errors="replace")


@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
def test_fused_matches_baseline_on_all_ranks(tmp_path):
  nproc = _ranks_for(8)
  code, out = _torchrun(nproc, [str(ROOT / "benchmarks" / "gar_bench.py"), "--gar-dim", "1000003", "--gar-iters", "3", "--gar-out", str(tmp_path)])
  assert code == 0, out[-4000:]
  results = json.loads((tmp_path / ("gar_bench_%d.json" % nproc)).read_text())["results"]
  assert set(results) == {"average", "average-nan", "median", "averaged-median", "krum", "bulyan"}
  for rule, entry in results.items():
    assert entry["replicas_identical"], (rule, entry)
    assert entry["max_abs_diff_vs_baseline"] < 1e-4, (rule, entry)


def _ranks_for(nb_workers):
  """Largest rank count <= visible GPUs that divides the number of logical workers (every visible GPU when possible)."""
  return max(r for r in range(1, min(_gpus(), nb_workers) + 1) if nb_workers % r == 0)


@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
def test_overlapped_whole_step_graph_keeps_replicas_identical(tmp_path):
  """No attack, Krum on cnnet: the step is ONE CUDA graph (forward/backward + bucketed distance pass on a side stream + finish kernel,
  loss through the kernel's mailbox, no NCCL on the step path) on every visible GPU; replicas stay bit-identical and the loss falls."""
  nproc = _ranks_for(8)
  workers = ", ".join('"127.0.0.1:%d"' % (7001 + i) for i in range(nproc))
  args = [str(ROOT / "runner.py"), "--server", '{"ps": ["127.0.0.1:7000"], "workers": [%s], "eval": ["127.0.0.1:7000"]}' % workers, "--no-wait",
          "--experiment", "cnnet", "--experiment-args", "batch-size:16", "--aggregator", "krum", "--nb-workers", "8", "--nb-decl-byz-workers", "2",
          "--max-step", "14", "--use-gpu", "--reuse-gpu", "--debug-checksum", "--learning-rate-args", "initial-rate:0.02",
          "--evaluation-delta", "7", "--evaluation-period", "-1", "--checkpoint-dir", str(tmp_path / "c"), "--checkpoint-delta", "14", "--checkpoint-period", "-1", "--summary-dir", "-"]
  os.environ["AGB_OVERLAP"] = "2"   # the overlapped distance pass also when a rank hosts several workers
  try:
    code, out = _torchrun(nproc, args)
  finally:
    del os.environ["AGB_OVERLAP"]
  assert code == 0, out[-4000:]
  assert "Replica divergence" not in out and "Step 13: total loss" in out
  assert "forward/backward + aggregation into a CUDA graph" in out and " 3 bucket(s)" in out
  import re
  losses = [float(x) for x in re.findall(r"Step \d+: total loss = ([0-9.eE+-]+)", out)]
  assert len(losses) == 14 and all(l == l for l in losses) and min(losses[-4:]) < losses[0]


@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
def test_training_keeps_replicas_identical(tmp_path):
  nproc = _ranks_for(8)
  workers = ", ".join('"127.0.0.1:%d"' % (7001 + i) for i in range(nproc))
  args = [str(ROOT / "runner.py"), "--server", '{"ps": ["127.0.0.1:7000"], "workers": [%s], "eval": ["127.0.0.1:7000"]}' % workers, "--no-wait",
          "--experiment", "cnnet", "--experiment-args", "batch-size:16", "--aggregator", "krum", "--nb-workers", "8", "--nb-decl-byz-workers", "2",
          "--nb-real-byz-workers", "2", "--attack", "flip", "--attack-args", "factor:-20", "--max-step", "12", "--use-gpu", "--reuse-gpu", "--debug-checksum",
          "--evaluation-delta", "6", "--evaluation-period", "-1", "--checkpoint-dir", str(tmp_path / "c"), "--checkpoint-delta", "12", "--checkpoint-period", "-1", "--summary-dir", "-"]
  code, out = _torchrun(nproc, args)
  assert code == 0, out[-4000:]
  assert "Replica divergence" not in out and "Step 11: total loss" in out
  assert (tmp_path / "c" / "model-12.index").exists()


@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
def test_own_collectives_match_nccl(tmp_path):
  """All-reduce / all-gather kernels over peer-mapped memory (`parallel/collectives.py`) vs the NCCL results, every dtype and the
  uneven all-gather; replicas must end with identical bits."""
  nproc = 2 if _gpus() < 4 else 4
  code, out = _torchrun(nproc, [str(ROOT / "benchmarks" / "coll_bench.py"), "--coll-numel", "2000003", "--coll-iters", "3", "--coll-out", str(tmp_path)])
  assert code == 0, out[-4000:]
  report = json.loads((tmp_path / ("coll_bench_%d.json" % nproc)).read_text())
  assert report["failures"] == [] and report["allreduce"]["ours_ms"] > 0


@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
def test_authenticated_training_drops_only_forged_slices(tmp_path):
  """`--authenticate` on 2 ranks x 4 workers with the fused engine: digests are recomputed through the peer mapping; the two forging
  workers lose the slice they tampered with (the first 40 % of the row = slice 0), honest rows pass, Krum keeps training."""
  args = [str(ROOT / "runner.py"), "--server", '{"ps": ["127.0.0.1:7000"], "workers": ["127.0.0.1:7001", "127.0.0.1:7002"], "eval": ["127.0.0.1:7000"]}', "--no-wait",
          "--experiment", "cnnet", "--experiment-args", "batch-size:16", "--aggregator", "krum", "--nb-workers", "8", "--nb-decl-byz-workers", "2",
          "--nb-real-byz-workers", "2", "--attack", "forge", "--attack-args", "factor:-50", "fraction:0.4", "--authenticate", "--max-step", "8", "--use-gpu", "--reuse-gpu",
          "--debug-checksum", "--evaluation-file", "-", "--checkpoint-dir", str(tmp_path / "c"), "--checkpoint-delta", "-1", "--checkpoint-period", "-1", "--summary-dir", "-"]
  code, out = _torchrun(2, args)
  assert code == 0, out[-4000:]
  assert "Replica divergence" not in out and "Step 7: total loss" in out
  import re
  dropped = re.findall(r"dropped (\d+) gradient slice\(s\) failing authentication: (\[[^\]]*\])", out)
  assert dropped and all(count == "2" for count, _ in dropped), dropped[:4]          # rank 0 (owner of slice 0) drops the two forgers' slice, every step
  assert all(piece.count(", 0)") == 2 for _, piece in dropped), dropped[:4]

"""Collectives API on the CPU / gloo plumbing path (1 and 3 processes; the GPU kernels are covered by tests/test_multigpu.py)."""

import json
import os
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_single_process(tmp_path):
  proc = subprocess.run([sys.executable, str(ROOT / "benchmarks" / "coll_bench.py"), "--coll-device", "cpu", "--coll-out", str(tmp_path)], stdout=subprocess.PIPE,
                        stderr=subprocess.STDOUT, timeout=300, cwd=str(ROOT))
  assert proc.returncode == 0, proc.stdout.decode(errors="replace")[-3000:]
  assert json.loads((tmp_path / "coll_bench_1.json").read_text())["failures"] == []


def test_three_processes_gloo(tmp_path):
  port = 29300 + os.getpid() % 300
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "3", "--master-addr", "127.0.0.1", "--master-port", str(port),
         str(ROOT / "benchmarks" / "coll_bench.py"), "--coll-device", "cpu", "--coll-out", str(tmp_path)]
  proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, cwd=str(ROOT))
  assert proc.returncode == 0, proc.stdout.decode(errors="replace")[-3000:]
  report = json.loads((tmp_path / "coll_bench_3.json").read_text())
  assert report["world"] == 3 and report["failures"] == []

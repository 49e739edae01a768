"""Cluster bring-up (reference: `deploy.py:1-329`).

The reference starts one `tf.train.Server` per cluster node (over SSH, piping its own source to a remote
`python`), optionally through `mpirun` with the lossy-UDP environment, then sleeps until signalled. Here a cluster node
is one *rank* of the SPMD job (one process per GPU): `deploy.py` starts the ranks — local ones as child processes,
remote ones over SSH — with the torch.distributed rendezvous environment (`RANK`, `WORLD_SIZE`, `LOCAL_RANK`,
`MASTER_ADDR`, `MASTER_PORT`) and forwards `--runner "<runner.py arguments>"` to each of them.

Same flags as the reference:
  --cluster JSON|G5k|local   cluster specification; one rank per entry of the worker job, rendezvous at the `ps` entry
  --deploy                   start every rank of the cluster (except this node's own when `--id` is given with `--omit`)
  --id job:idx               this node's identity: start (only) that rank here
  --nice [jobs...]           maximise the niceness of this process / of the ranks of the listed jobs
  --omit                     with `--id` and `--deploy`: do not start this node's own rank
  --MPI                      accepted for compatibility (ranks talk through NCCL/NVLink or gloo, never MPI)
  --runner "..."             arguments passed to `runner.py` on every rank
  --UDP n                    the first n workers use the lossy transport: mapped to `--nb-real-byz-workers n --attack drop-chunks`
                             (the emulation of the reference's UDP path) unless `--runner` already selects an attack
  --ship auto|always|never   NFS-free deployment (reference: `deploy.py:201-229` pipes its own source into `ssh host python`): the
                             framework's source tree (package, entry points, native sources and built libraries) is packed into
                             a tarball and piped through the SSH connection; the remote side unpacks it into a temporary directory,
                             runs its rank from there and removes it afterwards — no shared filesystem, no pre-installed checkout.
                             `auto` (default) ships to remote hosts only, `never` expects the same checkout path on every host.
Stays in the foreground until SIGINT/SIGTERM (or until orphaned), then terminates what it started.
"""

import argparse
import io
import os
import pathlib
import shlex
import signal
import socket
import subprocess
import sys
import time

try:
  from .. import tools
  cluster_parse, cluster_parsers = tools.cluster_parse, tools.cluster_parsers
except Exception:  # stand-alone use (piped through ssh): JSON only
  import json
  cluster_parse, cluster_parsers = (lambda text: json.loads(text)), ""

REPO = pathlib.Path(__file__).resolve().parents[2]
exit_pending = False
children = []


def mark_exit(*args, **kwargs):
  global exit_pending
  exit_pending = True


def clean_exit(code):
  for child in children:
    if child.poll() is None:
      child.terminate()
  deadline = time.time() + 10
  for child in children:
    try:
      child.wait(timeout=max(0.1, deadline - time.time()))
    except Exception:
      child.kill()
  sys.exit(code)


def parse_id(text):
  job, sep, index = text.rpartition(":")
  if not sep or not job:
    raise ValueError("Invalid ID format, expected <job>:<id>")
  index = int(index)
  if index < 0:
    raise ValueError("Expected non-negative node ID")
  return job, index


def _is_local(host):
  if host in ("localhost", "127.0.0.1", "::1", socket.gethostname()):
    return True
  try:
    return socket.gethostbyname(host) in ("127.0.0.1", socket.gethostbyname(socket.gethostname()))
  except OSError:
    return False


def plan(cluster, wk_job="workers", ps_job="ps"):
  """Cluster spec -> (master address, master port, [(job, index, host, rank, local rank)])."""
  if wk_job not in cluster:
    wk_job = next((job for job in cluster if job != ps_job), ps_job)
  entries = cluster[wk_job]
  master = cluster.get(ps_job, entries)[0]
  master_host, _, master_port = master.rpartition(":")
  ranks, per_host = [], {}
  for rank, entry in enumerate(entries):
    host = entry.rpartition(":")[0]
    if not host:
      raise ValueError("Invalid hostname:port format " + repr(entry))
    local = per_host.get(host, 0)
    per_host[host] = local + 1
    ranks.append((wk_job, rank, host, rank, local))
  return master_host, int(master_port), ranks


_BUNDLE_SUFFIXES = {".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".so", ".sh", ".md", ".toml", ""}
_bundle_cache = None


def source_bundle():
  """gzip-compressed tarball (bytes) of everything a rank needs: the package (sources + built native libraries), the entry points."""
  global _bundle_cache
  if _bundle_cache is not None:
    return _bundle_cache
  import tarfile
  buffer = io.BytesIO()
  with tarfile.open(fileobj=buffer, mode="w:gz", compresslevel=3) as tar:
    for entry in ("runner.py", "deploy.py", "pyproject.toml"):
      if (REPO / entry).is_file():
        tar.add(str(REPO / entry), arcname=entry)
    package = REPO / "aggregathor_b200"
    for path in sorted(package.rglob("*")):
      relative = path.relative_to(REPO)
      if "__pycache__" in relative.parts or not path.is_file() or path.name.endswith(".o") or ".tmp" in path.name:
        continue
      if path.suffix in _BUNDLE_SUFFIXES or path.name in ("DEPS",):
        tar.add(str(path), arcname=str(relative))
      elif path.is_symlink():
        tar.add(str(path), arcname=str(relative))
  _bundle_cache = buffer.getvalue()
  return _bundle_cache


def remote_script(env, runner_args, nice):
  """Shell run by `ssh host`: unpack the tarball arriving on stdin into a private directory, run the rank there, clean up."""
  assignments = " ".join(k + "=" + shlex.quote(v) for k, v in env.items())
  command = ("nice -n 19 " if nice else "") + "python3 runner.py " + " ".join(shlex.quote(c) for c in runner_args)
  return ("set -e; d=$(mktemp -d -t agb-rank-XXXXXX); trap 'rm -rf \"$d\"' EXIT; tar xzf - -C \"$d\"; cd \"$d\"; "
          "env " + assignments + " AGB_SHIPPED=1 " + command + " < /dev/null")


def rank_command(runner_args, nice):
  cmd = [sys.executable, str(REPO / "runner.py")] + runner_args
  return (["nice", "-n", "19"] if nice else []) + cmd


def start_rank(job, index, host, rank, local, world, master_host, master_port, runner_args, nice, ship="auto", ssh=("ssh", "-o", "BatchMode=yes")):
  env = {"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(local), "LOCAL_WORLD_SIZE": str(world), "MASTER_ADDR": master_host, "MASTER_PORT": str(master_port)}
  command = rank_command(runner_args, nice)
  local_host = _is_local(host)
  if local_host and ship != "always":
    child = subprocess.Popen(command, env=dict(os.environ, **env), cwd=str(REPO))
  elif ship != "never":
    # NFS-free: the source tree travels through the connection itself
    launcher = ["sh", "-c"] if local_host else list(ssh) + [host]
    child = subprocess.Popen(launcher + [remote_script(env, runner_args, nice)], stdin=subprocess.PIPE)
    try:
      child.stdin.write(source_bundle())
      child.stdin.close()
    except BrokenPipeError:
      pass
  else:
    remote = "cd " + shlex.quote(str(REPO)) + " && env " + " ".join(k + "=" + shlex.quote(v) for k, v in env.items()) + " " + " ".join(shlex.quote(c) for c in command)
    child = subprocess.Popen(["ssh", "-o", "BatchMode=yes", host, remote])
  children.append(child)
  print("\033[1;30m[" + job + ":" + str(index) + "]\033[1;32m rank " + str(rank) + "/" + str(world) + " on " + host + " (local rank " + str(local) + ")" + (" (nice)" if nice else "") + "\033[0m")
  return child


def main(argv=None):
  signal.signal(signal.SIGINT, mark_exit)
  signal.signal(signal.SIGTERM, mark_exit)
  parser = argparse.ArgumentParser(formatter_class=argparse.RawTextHelpFormatter)
  parser.add_argument("--cluster", type=str, required=True, help="Full cluster specification, JSON format: {\"<jobname>\": [\"hostname:port\", ...], ...}" + ("" if not cluster_parsers else ", or special value(s): " + cluster_parsers))
  parser.add_argument("--deploy", action="store_true", default=False, help="Whether this instance must deploy the whole cluster (local child processes, SSH for remote hosts)")
  parser.add_argument("--id", type=str, help="This node's role, format: <job>:<id>")
  parser.add_argument("--nice", nargs="*", help="Make this process nice, or list of job(s) which tasks must maximize their respective niceness level")
  parser.add_argument("--omit", action="store_true", default=False, help="Do not start the node's own rank, can be used only with '--id' and '--deploy'")
  parser.add_argument("--MPI", action="store_true", default=False, help="Accepted for compatibility with the reference (grpc+mpi transport); ignored")
  parser.add_argument("--runner", type=str, default="", help="Arguments passed to runner.py on every rank")
  parser.add_argument("--UDP", type=int, default=0, help="Number of workers running over the lossy transport (could be seen as Byzantine)")
  parser.add_argument("--ship", type=str, default="auto", choices=("auto", "always", "never"), help="Ship the source tree through the SSH pipe (NFS-free deployment): remote hosts only (auto), every rank, or never")
  args = parser.parse_args(sys.argv[1:] if argv is None else argv)
  cluster = cluster_parse(args.cluster)
  nices = args.nice if args.nice is not None else []
  for job in nices:
    if job not in cluster:
      print("\033[1;33mJob " + repr(job) + " does not appear in the cluster specification, hence cannot be nice\033[0m")
  if args.id is None:
    if not args.deploy:
      raise RuntimeError("Nothing to do (no deployment and no node ID)")
    if args.omit:
      raise RuntimeError("Cannot omit starting the node's server instance if its identity is unknown")
    this_job, this_id = None, None
  else:
    if args.omit and not args.deploy:
      raise RuntimeError("Nothing to do (no server start and no deployment)")
    this_job, this_id = parse_id(args.id)
    if this_job not in cluster or this_id >= len(cluster[this_job]):
      raise ValueError("Role is not in the specified cluster")
  if args.nice is not None and (args.nice == [] or (this_job is not None and this_job in args.nice)):
    os.nice(19)
  if args.MPI:
    print("\033[1;33m'--MPI' is accepted for compatibility: ranks communicate through NCCL/NVLink (GPU) or gloo (CPU)\033[0m")
  runner_args = shlex.split(args.runner)
  if runner_args and "--server" not in runner_args and "--client" not in runner_args:
    runner_args = ["--server", repr(cluster).replace("'", "\"")] + runner_args
  if runner_args and "--ev-job-name" not in runner_args and "eval" not in cluster and "ps" in cluster:
    runner_args += ["--ev-job-name", "ps"]  # the reference's experiments.sh does the same
  if args.UDP > 0 and "--attack" not in runner_args:
    runner_args += ["--nb-real-byz-workers", str(args.UDP), "--attack", "drop-chunks"]
  if "--no-wait" not in runner_args:
    runner_args.append("--no-wait")
  master_host, master_port, ranks = plan(cluster)
  world = len(ranks)
  if not runner_args or "--experiment" not in runner_args:
    print("\033[1;33mNo '--runner' arguments: nothing will be trained; the cluster of " + str(world) + " rank(s) is only described\033[0m")
    for job, index, host, rank, local in ranks:
      print("\033[1;30m[" + job + ":" + str(index) + "]\033[0m rank " + str(rank) + " on " + host + " (local rank " + str(local) + "), rendezvous " + master_host + ":" + str(master_port))
  else:
    for job, index, host, rank, local in ranks:
      if exit_pending:
        break
      own = this_job is not None and job == this_job and index == this_id
      if own and args.omit:
        print("\033[1;30m[" + job + ":" + str(index) + "]\033[1;34m No server running\033[0m")
        continue
      if args.deploy or own:
        start_rank(job, index, host, rank, local, world, master_host, master_port, runner_args, job in nices, ship=args.ship)
  sys.stdout.flush()
  while not exit_pending:
    time.sleep(1)
    if os.getppid() <= 1:
      break
    if children and all(child.poll() is not None for child in children):
      codes = [child.returncode for child in children]
      clean_exit(0 if all(code == 0 for code in codes) else 1)
  clean_exit(0)


if __name__ == "__main__":
  main()

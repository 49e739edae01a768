"""Cluster specification parsing (reference: `tools/cluster.py:45-91`).

A cluster is `{"<job>": ["host:port", ...], ...}` given as JSON text, or the
name of a special parser. `G5k` reads the Grid5000 `$OAR_FILE_NODES` file
(first distinct host -> `ps`, the others -> `workers`, port 7000). `local`
(new) describes the B200 box itself: one `ps` and one worker task per visible GPU.
"""

import json
import os
import pathlib

__all__ = ["cluster_parse", "cluster_parsers", "cluster_hosts"]

_G5K_ENV = "OAR_FILE_NODES"
_g5k_cache = None


def _g5k_parser():
  global _g5k_cache
  if _g5k_cache is None:
    from . import UserException
    if _G5K_ENV not in os.environ:
      raise UserException("Key " + repr(_G5K_ENV) + " not found in environment; are you running on Grid5000?")
    hosts = []
    for line in pathlib.Path(os.environ[_G5K_ENV]).read_text().split():
      entry = line.strip() + ":7000"
      if entry not in hosts:
        hosts.append(entry)
    _g5k_cache = {"ps": hosts[:1], "workers": hosts[1:]}
  return _g5k_cache


def _local_parser():
  try:
    import torch
    count = max(1, torch.cuda.device_count())
  except Exception:
    count = 1
  return {"ps": ["127.0.0.1:7000"], "workers": ["127.0.0.1:" + str(7001 + i) for i in range(count)], "eval": ["127.0.0.1:7000"]}


_parsers = {"G5k": _g5k_parser, "local": _local_parser}
cluster_parsers = ", ".join("'" + name + "'" for name in _parsers)


def cluster_parse(text):
  """JSON text (or special parser name) -> `{job: [host:port, ...]}`."""
  if text in _parsers:
    return _parsers[text]()
  from . import UserException
  try:
    spec = json.loads(text)
  except ValueError as err:
    raise UserException("Invalid cluster specification (expected JSON or one of " + cluster_parsers + "): " + str(err))
  if not isinstance(spec, dict) or not all(isinstance(v, list) for v in spec.values()):
    raise UserException("Invalid cluster specification: expected {\"<job>\": [\"host:port\", ...], ...}")
  return spec


def cluster_hosts(spec):
  """Ordered, de-duplicated list of (job, index, host, port) of a parsed spec."""
  nodes = []
  for job, entries in spec.items():
    for index, entry in enumerate(entries):
      host, _, port = entry.rpartition(":")
      if not host:
        raise ValueError("Invalid hostname:port format " + repr(entry))
      nodes.append((job, index, host, int(port)))
  return nodes

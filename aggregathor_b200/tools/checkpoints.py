"""Checkpoint directory manager (reference: `tools/tf.py:78-173`).

Layout kept from the reference so that existing tooling keeps working:
`<dir>/<base>-<step>.index` (JSON manifest), `<base>-<step>.data-00000-of-00001`
(the `torch.save`d state) and `<base>-<step>.meta` (JSON run description).
"Latest" is the highest numeric suffix among `*.index` files. All ranks of an
SPMD run hold identical state, so only rank 0 writes.
"""

import json
import os
import pathlib
import time

__all__ = ["Checkpoints"]

_DATA_SUFFIX = ".data-00000-of-00001"


class Checkpoints:
  """Save/restore/list `model-<step>` checkpoints of one directory."""

  def __init__(self, path, base=None, keep=None):
    from .. import config
    self._path = pathlib.Path(path)
    self._base = base if base is not None else config.default_checkpoint_base_name
    self._keep = config.checkpoint_keep if keep is None else keep
    self._available = []
    self._processed = set()

  def _stem(self, step):
    return str(self._path / (self._base + "-" + str(int(step))))

  def _update(self):
    found = []
    if self._path.exists():
      prefix = self._base + "-"
      for item in self._path.iterdir():
        if item.is_file() and item.suffix == ".index" and item.stem.startswith(prefix):
          tail = item.stem[len(prefix):]
          if tail.isdigit():
            found.append((int(tail), str(item)[:-len(".index")]))
    found.sort()
    self._available = [stem for _, stem in found]

  def get(self, no_filter=False):
    """Available checkpoint stems (oldest first), minus those already returned."""
    self._update()
    got = list(self._available) if no_filter else [s for s in self._available if s not in self._processed]
    self._processed.update(got)
    return got

  def can_restore(self):
    self._update()
    return len(self._available) > 0

  def latest(self):
    self._update()
    return self._available[-1] if self._available else None

  def restore(self, path=None, map_location="cpu"):
    """Load and return the state dict of `path` (a stem), default the latest one."""
    import torch
    from . import UserException
    if path is None:
      path = self.latest()
      if path is None:
        raise UserException("No storage file to restore")
    from . import tf_checkpoint
    if tf_checkpoint.is_tf_bundle(path):  # written by the reference (tf.train.Saver): the trainer maps the variables onto its layout
      return {"tf_variables": tf_checkpoint.read_bundle(path), "source": str(path)}
    return torch.load(str(path) + _DATA_SUFFIX, map_location=map_location, weights_only=False)

  def save(self, state, step, meta=None):
    """Atomically write `state` (any picklable/torch-saveable dict) for `step`."""
    import torch
    self._path.mkdir(parents=True, exist_ok=True)
    stem = self._stem(step)
    tmp = stem + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, stem + _DATA_SUFFIX)
    with open(stem + ".meta", "w") as fd:
      json.dump(meta if meta is not None else {}, fd, indent=1, default=str)
    with open(stem + ".index.tmp", "w") as fd:
      json.dump({"step": int(step), "time": time.time(), "data": os.path.basename(stem) + _DATA_SUFFIX,
                 "keys": sorted(state.keys()) if isinstance(state, dict) else None}, fd)
    os.replace(stem + ".index.tmp", stem + ".index")  # the index appears last: readers never see partial data
    self._update()
    if self._keep and self._keep > 0:
      for old in self._available[:-self._keep]:
        for suffix in (".index", _DATA_SUFFIX, ".meta"):
          try:
            os.remove(old + suffix)
          except OSError:
            pass
    return stem

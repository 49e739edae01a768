"""Filesystem permission probe (reference: `tools/access.py:42-79`)."""

import os
import pathlib
import stat

__all__ = ["can_access"]


def _mode_allows(info, read, write):
  """Apply owner/group/other permission bits the way the kernel picks them."""
  if os.geteuid() == info.st_uid:
    rbit, wbit = stat.S_IRUSR, stat.S_IWUSR
  elif os.getegid() == info.st_gid:
    rbit, wbit = stat.S_IRGRP, stat.S_IWGRP
  else:
    rbit, wbit = stat.S_IROTH, stat.S_IWOTH
  return (not read or bool(info.st_mode & rbit)) and (not write or bool(info.st_mode & wbit))


def can_access(path, read=False, write=False, recurse=False):
  """Whether `path` exists and every file below it (one level, or the whole
  tree with `recurse`) grants the requested access to the effective user."""
  try:
    path = pathlib.Path(path)
    if not path.exists():
      return False
    info = path.stat()
    if not stat.S_ISDIR(info.st_mode):
      return _mode_allows(info, read, write)
    for child in path.iterdir():
      if child.is_dir() and not recurse:
        continue
      if not can_access(child, read, write, recurse):
        return False
    return True
  except OSError:
    return False

"""Logging contexts, user-facing errors and plug-in discovery.

Functional equivalent of the reference's `tools/__init__.py:44-318`:

* `UserException`: an error the user can fix; reported as `[fatal] message`, exit 1.
* `Context(name, color)`: per-thread stack of `[name] ` headers + ANSI colours;
  non-main threads get their thread name prepended.
* `ContextIOWrapper`: text stream wrapper that prefixes every line with the
  current header (used for stdout/stderr, and colour-stripped tee files).
* `trace/info/success/warning/error/fatal`: coloured prints.
* `import_directory`: auto-import every `*.py` of a package directory (the
  plug-in discovery mechanism for `aggregators/`, `attacks/`, `experiments/`).

Unlike the reference, importing this module has no side effect on
`sys.stdout`/`sys.excepthook`; the CLIs call `install()` explicitly so that the
package stays usable as a library (and under pytest).
"""

import importlib
import os
import pathlib
import sys
import threading
import traceback

__all__ = [
  "UserException", "Context", "ContextIOWrapper", "trace", "info", "success",
  "warning", "error", "fatal", "install", "uninstall", "import_directory"]


class UserException(Exception):
  """Error caused by (and fixable by) the user: printed without traceback."""
  pass


# ---------------------------------------------------------------------------- #
# Contexts

_COLORS = {
  "header": "\033[1;30m",
  "red": "\033[1;31m", "error": "\033[1;31m",
  "green": "\033[1;32m", "success": "\033[1;32m",
  "yellow": "\033[1;33m", "warning": "\033[1;33m",
  "blue": "\033[1;34m", "info": "\033[1;34m",
  "gray": "\033[1;30m", "trace": "\033[1;30m"}
_COLOR_END = "\033[0m"

_tls = threading.local()
_rank_tag = None  # Optional "r3"-style tag inserted in every header (SPMD runs)


def set_rank_tag(tag):
  """Prefix every header with `[tag]` (used by non-zero ranks of an SPMD run)."""
  global _rank_tag
  _rank_tag = tag


def _stack():
  stack = getattr(_tls, "stack", None)
  if stack is None:
    stack = _tls.stack = []
  return stack


class Context:
  """Scoped `[name]` header and/or colour for everything printed by this thread."""

  def __init__(self, cntxtname, colorname):
    if colorname is not None and colorname not in _COLORS:
      raise AssertionError("Unknown color name " + repr(colorname))
    self._entry = (cntxtname, None if colorname is None else _COLORS[colorname])

  def __enter__(self):
    _stack().append(self._entry)
    return self

  def __exit__(self, *args):
    _stack().pop()
    return False

  @staticmethod
  def current():
    """Returns (header text, header colour, body colour, colour reset)."""
    header = ""
    color = None
    for name, code in _stack():
      if name is not None:
        header += "[" + name + "] "
      if code is not None:
        color = code  # innermost colour wins
    thread = threading.current_thread()
    if thread is not threading.main_thread():
      header = "[" + thread.name + "] " + header
    if _rank_tag is not None:
      header = "[" + _rank_tag + "] " + header
    return header, _COLORS["header"], (color if color is not None else _COLOR_END), _COLOR_END


class ContextIOWrapper:
  """Text stream decorator: every new line starts with the current context header."""

  def __init__(self, output, nocolor=False):
    self._output = output
    self._nocolor = nocolor
    self._at_line_start = True
    self._lock = threading.Lock()

  def __getattr__(self, name):
    return getattr(self._output, name)

  def write(self, text):
    if not text:
      return 0
    header, c_head, c_body, c_end = Context.current()
    if self._nocolor:
      c_head = c_body = c_end = ""
    out = []
    with self._lock:
      for line in text.splitlines(True):
        if self._at_line_start:
          out.append(c_head + header)
        out.append(c_body + line)
        self._at_line_start = line.endswith("\n")
      out.append(c_end)
      return self._output.write("".join(out))


def _colored(color):
  def color_print(*args, context=None, **kwargs):
    with Context(context, color):
      stream = kwargs.get("file", sys.stdout)
      if isinstance(stream, (ContextIOWrapper,)) or getattr(stream, "_is_context_tee", False):
        return print(*args, **kwargs)
      # Stream is not context-aware (library use): render the header ourselves
      kwargs["file"] = ContextIOWrapper(stream, nocolor=not getattr(stream, "isatty", lambda: False)())
      return print(*args, **kwargs)
  color_print.__name__ = color
  color_print.__doc__ = "print() inside a " + repr(color) + " coloured context; `context=` adds a header."
  return color_print


trace = _colored("trace")
info = _colored("info")
success = _colored("success")
warning = _colored("warning")
error = _colored("error")


def fatal(*args, **kwargs):
  """`error(...)` then `exit(1)`."""
  error(*args, **kwargs)
  sys.exit(1)


# ---------------------------------------------------------------------------- #
# Process-wide installation (CLI entry points only)

_installed = None


def _excepthook(etype, evalue, tb):
  if issubclass(etype, UserException):
    with Context("fatal", "error"):
      print(evalue)
    sys.stdout.flush()
    return  # the interpreter exits with status 1 after an uncaught exception
  with Context("uncaught", "error"):
    _installed[2](etype, evalue, tb)


def install():
  """Wrap stdout/stderr with context-aware writers and report `UserException`s cleanly."""
  global _installed
  if _installed is not None:
    return
  _installed = (sys.stdout, sys.stderr, sys.excepthook)
  sys.stdout = ContextIOWrapper(sys.stdout)
  sys.stderr = ContextIOWrapper(sys.stderr)
  sys.excepthook = _excepthook


def uninstall():
  """Undo `install()` (tests)."""
  global _installed
  if _installed is None:
    return
  sys.stdout, sys.stderr, sys.excepthook = _installed
  _installed = None


# ---------------------------------------------------------------------------- #
# Plug-in discovery

def import_directory(dirpath, scope, ignore=("__init__.py",)):
  """Import every `*.py` module of the package directory `dirpath`.

  Modules register themselves (e.g. `aggregators.register(...)`) at import; a
  module that fails to import is reported as a warning and skipped, so that a
  broken/unsupported plug-in never takes the framework down (reference
  behaviour, `tools/__init__.py:292-318`). Symbols listed in a module's
  `__all__` are re-exported into `scope`.
  """
  package = scope["__package__"]
  for path in sorted(pathlib.Path(dirpath).iterdir()):
    if not (path.is_file() and path.suffix == ".py") or path.name in ignore or path.name.startswith("_"):
      continue
    name = path.stem
    with Context(name, None):
      try:
        module = importlib.import_module("." + name, package)
        for symbol in getattr(module, "__all__", ()):
          if not hasattr(module, symbol):
            warning("Symbol " + repr(symbol) + " exported but not defined")
          elif symbol in scope and scope[symbol] is not getattr(module, symbol):
            warning("Symbol " + repr(symbol) + " already defined in " + repr(package))
          else:
            scope[symbol] = getattr(module, symbol)
      except Exception as err:
        warning("Loading failed for module " + repr(path.name) + ": " + str(err))
        with Context("traceback", "trace"):
          traceback.print_exc()


from .misc import *          # noqa: E402,F401,F403
from .access import *        # noqa: E402,F401,F403
from .cluster_spec import *  # noqa: E402,F401,F403
from .checkpoints import *   # noqa: E402,F401,F403
from .tracing import *       # noqa: E402,F401,F403
from .summary import *       # noqa: E402,F401,F403

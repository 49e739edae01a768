"""Small generic helpers (reference: `tools/misc.py:45-283`).

`MethodCallReplicator` (tee), `ClassRegister` (name -> class plug-in registry),
`parse_keyval` ("key:value" lists -> typed dict), `print_args`, `ExpandPath`,
`make_interface` (pimpl wrapper over native create/destroy/method functions).
"""

import sys

__all__ = ["MethodCallReplicator", "ClassRegister", "parse_keyval", "print_args", "ExpandPath", "make_interface"]


def _user_exception(text):
  from . import UserException
  return UserException(text)


class MethodCallReplicator:
  """Forward every method call to several instances, returning the list of results."""

  def __init__(self, *instances):
    if len(instances) == 0:
      raise AssertionError("Expected at least one instance on which to forward method calls")
    self._instances = instances
    self._is_context_tee = any(getattr(i, "_is_context_tee", False) or type(i).__name__ == "ContextIOWrapper" for i in instances)

  def __getattr__(self, name):
    targets = [getattr(instance, name) for instance in self._instances]
    if not all(callable(t) for t in targets):
      return targets[0]
    def replicated(*args, **kwargs):
      return [target(*args, **kwargs) for target in targets]
    return replicated


class ClassRegister:
  """Name -> constructor registry with a helpful unknown-name error."""

  def __init__(self, singular, optplural=None):
    self._singular = singular
    self._optplural = singular + "(s)" if optplural is None else optplural
    self._entries = {}

  def itemize(self):
    return self._entries.keys()

  def register(self, name, cls):
    if name in self._entries:
      raise AssertionError("Name " + repr(name) + " already in use while registering " + repr(getattr(cls, "__name__", "<unknown " + self._singular + " class name>")))
    self._entries[name] = cls

  def get(self, name):
    if name not in self._entries:
      if len(self._entries) == 0:
        cause = "no registered " + self._singular
      else:
        cause = "available " + self._optplural + ": '" + "', '".join(self._entries.keys()) + "'"
      raise _user_exception("Unknown name " + repr(name) + ", " + cause)
    return self._entries[name]

  def instantiate(self, name, *args, **kwargs):
    return self.get(name)(*args, **kwargs)


def _coerce(kind, text):
  if kind is bool:
    lowered = text.strip().lower()
    if lowered in ("1", "true", "yes", "on"):
      return True
    if lowered in ("0", "false", "no", "off", ""):
      return False
    raise ValueError(text)
  return kind(text)


def parse_keyval(list_keyval, defaults={}):
  """Parse `["key:value", ...]` into a dict.

  Values of keys present in `defaults` are converted to the type of their
  default; other keys stay `str`. Missing keys take their default. Duplicate
  keys and entries without `:` raise `UserException`.
  """
  parsed = {}
  for entry in list_keyval:
    key, sep, val = entry.partition(":")
    if not sep:
      raise _user_exception("Expected list of " + repr("<key>:<value>") + ", got " + repr(entry) + " as one entry")
    if key in parsed:
      raise _user_exception("Key " + repr(key) + " had already been specified with value " + repr(parsed[key]))
    if key in defaults:
      kind = type(defaults[key])
      try:
        val = _coerce(kind, val)
      except Exception:
        raise _user_exception("Required key " + repr(key) + " expected a value of type " + repr(getattr(kind, "__name__", "<unknown>")))
    parsed[key] = val
  for key, val in defaults.items():
    parsed.setdefault(key, val)
  return parsed


def print_args(name, selected, list_keyval, head="[ARGS] "):
  """Print the selected plug-in and its key:value arguments."""
  print(head + "Selected " + name + ": " + (selected if selected else "<none>"))
  for key, val in parse_keyval(list_keyval).items():
    print(head + "· " + key + ": " + str(val))


class ExpandPath:
  """Context manager temporarily appending directories to `sys.path`."""

  def __init__(self, *paths):
    self._extra = [str(path) for path in paths]
    self._saved = None

  def __enter__(self):
    self._saved = sys.path
    sys.path = sys.path + self._extra
    return self

  def __exit__(self, *args):
    sys.path = self._saved
    return False


def make_interface(_create, _destroy, **methods):
  """Build a class wrapping a native handle: `_create(...) -> handle`,
  `_destroy(handle)`, and `methods[name](handle, ...)` exposed as bound methods.
  Calling the instance returns the raw handle."""
  if "_native" in methods:
    raise _user_exception("Method name '_native' is reserved")

  class Interface:
    def __init__(self, *args):
      self._native = _create(*args)

    def __del__(self):
      handle = self.__dict__.get("_native", None)
      if handle is not None:
        _destroy(handle)

    def __getattr__(self, name):
      if "_native" not in self.__dict__:
        raise _user_exception("Unable to access instance as its creation failed")
      if name not in methods:
        raise AttributeError(name)
      method, handle = methods[name], self.__dict__["_native"]
      return lambda *args: method(handle, *args)

    def __call__(self):
      if "_native" not in self.__dict__:
        raise _user_exception("Unable to access instance as its creation failed")
      return self._native

  return Interface

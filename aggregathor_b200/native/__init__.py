"""In-tree native build + load system for sm_100a (reference: `native/__init__.py:99-439`).

Every sub-directory `<prefix>_<name>/` of this package is one shared library:

* `so_<name>`: plain shared object other libraries link against (e.g. the host thread pool);
* `py_<name>`: host C++ library with a C ABI, loaded with `ctypes` and registered
  in the *py* register (`instantiate_py(name)` returns the `CDLL`);
* `op_<name>`: CUDA (sm_100a) kernel library with a C ABI; every function listed
  by its exported `agb_op_list()` is registered in the *op* register
  (`instantiate_op("krum", ...)` calls it).

Dependencies between libraries are expressed as in the reference: a symlink
(or a line in a `DEPS` text file) inside the dependent directory naming the
dependee directory. Builds are incremental (mtime of sources, of the shared
headers in `include/`, and of this very file) and happen *in-tree*
(`native/<dir>.so`), so the built objects travel with a snapshot of the repo.
CUDA sources are compiled with `-gencode arch=compute_100a,code=sm_100a
-lineinfo`; when `nvcc` is missing they are skipped and the library is reported
unavailable (ops then fail loudly at call time on a GPU box).

The C ABI + ctypes design (instead of torch C++ extensions) keeps a full rebuild
at a few seconds per file and keeps the kernels independent from the torch ABI:
wrappers pass `tensor.data_ptr()` and the current `cudaStream_t`.
"""

import contextlib
import ctypes
import fcntl
import os
import pathlib
import shlex
import shutil
import subprocess
import sys
import threading

from .. import tools, config

__all__ = [
  "build_all", "library", "available", "itemize_op", "register_op", "instantiate_op", "get_op",
  "itemize_py", "register_py", "instantiate_py", "import_py", "dump_sass", "build_log"]

_HERE = pathlib.Path(__file__).resolve().parent
_INCLUDE = _HERE / "include"
_SELF_MTIME = pathlib.Path(__file__).stat().st_mtime

_EXT_HDR = {".h", ".hh", ".hpp", ".hxx", ".cuh"}
_EXT_CPP = {".cc", ".cpp", ".cxx"}
_EXT_CUDA = {".cu"}
_PREFIXES = ("so_", "py_", "op_")

_CXX = os.environ.get("CXX", "c++")
_NVCC = os.environ.get("NVCC", shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc")
_CXX_FLAGS = ["-std=c++17", "-O3", "-DNDEBUG", "-fPIC", "-Wall", "-Wextra", "-Wfatal-errors", "-pthread", "-fno-math-errno"]
_NVCC_FLAGS = ["-std=c++17", "-O3", "-DNDEBUG", "-lineinfo", "--expt-relaxed-constexpr", "--expt-extended-lambda",
               "-Xcompiler", "-fPIC,-Wall,-Wextra,-Wno-unused-parameter", "-Xptxas", "-v", "-diag-suppress", "177"] + list(config.cuda_arch_flags)
_LINK_FLAGS = ["-shared", "-pthread", "-Wl,--no-as-needed", "-Wl,-rpath,$ORIGIN"]

build_log = []       # (library name, command, returncode, output) for every executed command
_lock = threading.RLock()
_libs = {}           # name ("op_gar") -> ctypes.CDLL or None (failed)
_failed = {}         # name -> reason

_reg_op = tools.ClassRegister("custom operation")
itemize_op = _reg_op.itemize
register_op = _reg_op.register
get_op = _reg_op.get
_reg_py = tools.ClassRegister("foreign import")
itemize_py = _reg_py.itemize
register_py = _reg_py.register
instantiate_py = _reg_py.instantiate


def instantiate_op(name, *args, **kwargs):
  """Call the registered native op `name` (a C function; integer status 0 = success)."""
  return _reg_op.get(name)(*args, **kwargs)


def _has_nvcc():
  return _NVCC is not None and pathlib.Path(_NVCC).exists()


def _run(libname, command, quiet):
  proc = subprocess.run(command, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
  output = proc.stdout.decode("utf8", errors="replace")
  build_log.append((libname, " ".join(shlex.quote(c) for c in command), proc.returncode, output))
  if proc.returncode != 0:
    with tools.Context(libname, "warning"):
      print("Command failed (" + str(proc.returncode) + "): " + " ".join(command))
      print(output)
  elif not quiet:
    tools.trace("built " + command[-1], context=libname)
  return proc.returncode == 0


@contextlib.contextmanager
def _process_lock():
  """Serialise builds across PROCESSES (every rank of a torchrun / deploy.py job calls `library()` at the same moment on a
  fresh checkout): an exclusive `flock` on a lock file beside the libraries. Products are additionally written under a
  temporary name and renamed into place, so a concurrent `CDLL` never maps a half-written file."""
  try:
    handle = open(_HERE / ".build.lock", "a+")
  except OSError:   # read-only install: nothing can be built anyway
    yield
    return
  try:
    fcntl.flock(handle, fcntl.LOCK_EX)
    yield
  finally:
    try:
      fcntl.flock(handle, fcntl.LOCK_UN)
    finally:
      handle.close()


def _run_into(libname, command, product, quiet):
  """Run `command` (whose output file argument is the literal "@OUT@") writing to a temporary sibling of `product`, then rename."""
  tmp = product.with_name(product.name + ".tmp" + str(os.getpid()))
  ok = _run(libname, [str(tmp) if c == "@OUT@" else c for c in command], quiet)
  if ok:
    os.replace(tmp, product)
  else:
    with contextlib.suppress(OSError):
      tmp.unlink()
  return ok


def _stale(product, sources):
  if not product.exists():
    return True
  stamp = product.stat().st_mtime
  if _SELF_MTIME >= stamp:
    return True
  return any(src.stat().st_mtime >= stamp for src in sources)


def _scan(libdir):
  """Returns (dependency dirs, headers, c++ sources, cuda sources) of a library directory."""
  deps, headers, cpps, cudas = [], [], [], []
  for path in sorted(libdir.iterdir()):
    if path.is_symlink() or path.is_dir():
      try:
        target = path.resolve(strict=True)
      except OSError:
        raise RuntimeError("missing dependency " + repr(os.readlink(str(path))))
      if target.is_dir() and target.parent == _HERE and target.name[:3] in _PREFIXES and target != libdir:
        deps.append(target)
      continue
    if path.name == "DEPS":
      for line in path.read_text().split():
        target = (_HERE / line.strip()).resolve()
        if not target.is_dir():
          raise RuntimeError("missing dependency " + repr(line.strip()))
        deps.append(target)
      continue
    suffixes = path.suffixes
    if not suffixes:
      continue
    if suffixes[-1] in _EXT_HDR:
      headers.append(path)
    elif suffixes[-1] in _EXT_CUDA or (suffixes[-1] in _EXT_CPP and len(suffixes) > 1 and suffixes[-2] in _EXT_CUDA):
      cudas.append(path)
    elif suffixes[-1] in _EXT_CPP:
      cpps.append(path)
  return deps, headers, cpps, cudas


def _so_path(libdir):
  return libdir.parent / (libdir.name + ".so")


def _build(libdir, stack, quiet):
  """(Re)build `libdir` and its dependencies; returns the .so path or raises."""
  name = libdir.name
  if libdir in stack:
    raise RuntimeError("dependency cycle through " + repr(name))
  deps, headers, cpps, cudas = _scan(libdir)
  shared_headers = [p for p in sorted(_INCLUDE.iterdir()) if p.suffix in _EXT_HDR] if _INCLUDE.is_dir() else []
  dep_sos = []
  for dep in deps:
    dep_sos.append(_build(dep, stack + [libdir], quiet))
    _, dep_headers, _, _ = _scan(dep)
    headers = headers + dep_headers
  if not cpps and not cudas:
    raise RuntimeError("no source file")
  if cudas and not _has_nvcc():
    raise RuntimeError("nvcc not found, cannot compile " + ", ".join(p.name for p in cudas))
  incl = ["-I" + str(_INCLUDE)] + ["-I" + str(dep) for dep in deps] + ["-I" + str(libdir)]
  objects = []
  for src in cpps:
    obj = pathlib.Path(str(src) + ".o")
    if _stale(obj, headers + shared_headers + [src]):
      if not _run_into(name, [_CXX] + _CXX_FLAGS + incl + ["-c", "-o", "@OUT@", str(src)], obj, quiet):
        raise RuntimeError("C++ source " + repr(src.name) + " did not compile")
    objects.append(obj)
  for src in cudas:
    obj = pathlib.Path(str(src) + ".o")
    if _stale(obj, headers + shared_headers + [src]):
      if not _run_into(name, [_NVCC] + _NVCC_FLAGS + incl + ["-c", "-o", "@OUT@", str(src)], obj, quiet):
        raise RuntimeError("CUDA source " + repr(src.name) + " did not compile")
    objects.append(obj)
  so_path = _so_path(libdir)
  if _stale(so_path, objects + dep_sos):
    linker = [_NVCC, "-shared", "-Xcompiler", "-fPIC", "-Xlinker", "-rpath=$ORIGIN", "-Xlinker", "--no-as-needed"] if cudas else [_CXX] + _LINK_FLAGS
    command = linker + ["-o", "@OUT@"] + [str(o) for o in objects] + ["-L" + str(_HERE)] + ["-l:" + p.name for p in dep_sos]
    if not _run_into(name, command, so_path, quiet):
      raise RuntimeError("final shared object " + repr(so_path.name) + " could not be linked")
  return so_path


def _load(libdir, so_path):
  name = libdir.name
  for dep in _scan(libdir)[0]:
    library(dep.name)
  lib = ctypes.CDLL(str(so_path), mode=ctypes.RTLD_GLOBAL)
  _libs[name] = lib
  kind, short = name[:3], name[3:]
  if kind == "py_":
    register_py(short, lambda lib=lib: lib)
  elif kind == "op_":
    lister = getattr(lib, "agb_op_list", None)
    if lister is not None:
      lister.restype = ctypes.c_char_p
      for opname in lister().decode().split(","):
        opname = opname.strip()
        if opname and opname not in itemize_op():
          func = getattr(lib, "agb_" + opname)
          func.restype = ctypes.c_int
          register_op(opname, func)
  return lib


def library(name, quiet=True):
  """Build (if stale) and load the native library `name` (e.g. "op_gar"); raises `UserException` if unavailable."""
  with _lock:
    if name in _libs and _libs[name] is not None:
      return _libs[name]
    if name in _failed:
      raise tools.UserException("Native library " + repr(name) + " is unavailable: " + _failed[name])
    libdir = _HERE / name
    try:
      if not libdir.is_dir():
        raise RuntimeError("no such library directory")
      if os.environ.get("AGB_NATIVE_NO_BUILD") and _so_path(libdir).exists():
        so_path = _so_path(libdir)
      else:
        with _process_lock():
          so_path = _build(libdir, [], quiet)
      return _load(libdir, so_path)
    except Exception as err:
      _failed[name] = str(err)
      _libs[name] = None
      raise tools.UserException("Native library " + repr(name) + " is unavailable: " + str(err))


def available(name):
  """Whether `library(name)` works (never raises)."""
  try:
    library(name)
    return True
  except tools.UserException:
    return False


def build_all(quiet=True, load=True, strict=False):
  """Build every library directory; returns {name: so path or None}. CUDA libraries are
  cross-compiled even without a GPU; loading them only needs the CUDA runtime."""
  results = {}
  with _lock:
    for libdir in sorted(_HERE.iterdir()):
      if not (libdir.is_dir() and libdir.name[:3] in _PREFIXES):
        continue
      if not any(_scan(libdir)[2:]):
        continue  # nothing to compile (yet)
      try:
        with _process_lock():
          so_path = _build(libdir, [], quiet)
        results[libdir.name] = so_path
        if load and libdir.name not in _libs:
          _load(libdir, so_path)
      except Exception as err:
        results[libdir.name] = None
        _failed[libdir.name] = str(err)
        tools.warning("Build failed: " + str(err), context=libdir.name)
        if strict:
          raise
  return results


def import_py(lib, name, args, ret, defs=None, echk=None):
  """Declare the prototype of a C function of a ctypes library (reference: `native/__init__.py:309-357`).
  `defs` are default values for the trailing arguments."""
  func = getattr(lib, name, None)
  if func is None:
    raise RuntimeError("Function " + repr(name) + " is not exported in " + repr(getattr(lib, "_name", "?")))
  func.argtypes = args
  func.restype = ret
  if echk is not None:
    func.errcheck = echk
  if defs is None:
    return func
  defs = tuple(defs)
  if len(defs) > len(args):
    raise AssertionError("More defaults provided than possible arguments")
  def call(*given):
    missing = len(args) - len(given)
    if missing > len(defs):
      raise AssertionError("Not enough parameters provided in native function call")
    return func(*(given + (defs[len(defs) - missing:] if missing > 0 else ())))
  return call


def dump_sass(outdir, names=None):
  """Write `cuobjdump -sass` and a resource-usage summary of every built CUDA library into `outdir`
  (text files, committed under `profiles/sass/` as kernel evidence)."""
  outdir = pathlib.Path(outdir)
  outdir.mkdir(parents=True, exist_ok=True)
  cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
  written = []
  for libdir in sorted(_HERE.iterdir()):
    if not (libdir.is_dir() and libdir.name.startswith("op_")) or (names and libdir.name not in names):
      continue
    so_path = _so_path(libdir)
    if not so_path.exists():
      continue
    proc = subprocess.run([cuobjdump, "-sass", str(so_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    target = outdir / (libdir.name + ".sass")
    target.write_bytes(proc.stdout)
    written.append(target)
  return written

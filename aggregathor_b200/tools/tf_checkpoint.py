"""TensorFlow checkpoints ("tensor bundles") without TensorFlow: read the `model-<step>.index` + `.data-00000-of-00001` pairs
written by the reference (`tools/tf.py:141-173`, a `tf.train.Saver` around every global variable) and map their variables onto
a `FlatLayout`, so that a run can resume from — or be initialised with — what the reference trained; and write the same format
back, so that parameters trained here can be handed to TensorFlow tooling.

Formats (decoded by hand):
* `.index`: a LevelDB-style sorted string table. Footer (last 48 bytes) = block handles of the meta-index and index blocks
  (varint64 offset, size) + magic 0xdb4775248b80fb57. A block = entries (shared-prefix length, unshared length, value length,
  key suffix, value) + restart array + restart count, followed by a 1-byte compression tag and a 4-byte masked CRC32C.
  Key "" -> `BundleHeaderProto` (num_shards, endianness, version); key <variable name> -> `BundleEntryProto`
  (dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5, crc32c = 6).
* `.data-00000-of-00001`: the raw little-endian tensors at the recorded offsets.

Layout conversion (TensorFlow -> here): convolution kernels HWIO -> OHWI, depthwise kernels [kh, kw, C, m] -> [C * m, kh, kw, 1],
dense kernels [in, out] -> [out, in] (a 4-D "fully connected as convolution" kernel of slim is flattened in NHWC order first);
vectors and batch-norm statistics are copied. Names: identical for mnist and the slim networks; the reference's cnnet variables
`shared/<layer>_<kind>` map to `<layer>/<kind>`.
"""

import pathlib
import struct

import numpy as np

from . import UserException, info, warning
from .summary import _field_bytes, _field_varint, _masked_crc, _parse, _read_varint, _varint

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_, 19: np.float16}
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}
_DATA_SUFFIX = ".data-00000-of-00001"


# ---------------------------------------------------------------------------- #
# Sorted string table

def _block(data, offset, size):
  if data[offset + size] != 0:
    raise UserException("Compressed checkpoint index blocks (tag %d) are not supported" % data[offset + size])
  block = data[offset:offset + size]
  (restarts,) = struct.unpack_from("<I", block, size - 4)
  end = size - 4 - 4 * restarts
  pos, key, entries = 0, b"", []
  while pos < end:
    shared, pos = _read_varint(block, pos)
    unshared, pos = _read_varint(block, pos)
    length, pos = _read_varint(block, pos)
    key = key[:shared] + bytes(block[pos:pos + unshared])
    pos += unshared
    entries.append((key, bytes(block[pos:pos + length])))
    pos += length
  return entries


def _handle(buf, pos=0):
  offset, pos = _read_varint(buf, pos)
  size, pos = _read_varint(buf, pos)
  return offset, size, pos


def is_tf_bundle(stem):
  """Whether `<stem>.index` is a TensorFlow tensor-bundle index (as opposed to this framework's JSON manifest)."""
  path = pathlib.Path(str(stem) + ".index")
  if not path.is_file() or path.stat().st_size < 48:
    return False
  with open(path, "rb") as fd:
    fd.seek(-8, 2)
    return struct.unpack("<Q", fd.read(8))[0] == _MAGIC


def read_bundle(stem):
  """`<stem>.index` + `<stem>.data-00000-of-00001` -> {variable name: numpy array}."""
  index = pathlib.Path(str(stem) + ".index").read_bytes()
  if len(index) < 48 or struct.unpack_from("<Q", index, len(index) - 8)[0] != _MAGIC:
    raise UserException(repr(str(stem) + ".index") + " is not a TensorFlow checkpoint index")
  footer = index[-48:]
  _, _, pos = _handle(footer)                       # meta-index block: unused
  index_offset, index_size, _ = _handle(footer, pos)
  records = {}
  for _, value in _block(index, index_offset, index_size):
    offset, size, _ = _handle(value)
    for key, entry in _block(index, offset, size):
      records[key.decode()] = entry
  header = dict(_parse(records.pop("", b"")))
  if header.get(1, 1) != 1:
    raise UserException("Sharded checkpoints (%d shards) are not supported" % header.get(1))
  if header.get(2, 0) != 0:
    raise UserException("Big-endian checkpoints are not supported")
  data_path = pathlib.Path(str(stem) + _DATA_SUFFIX)
  data = np.memmap(data_path, dtype=np.uint8, mode="r") if data_path.stat().st_size else np.zeros(0, dtype=np.uint8)
  variables = {}
  for name, entry in records.items():
    fields = _parse(entry)
    dtype = next((v for k, v in fields if k == 1), 1)
    shape_blob = next((v for k, v in fields if k == 2), b"")
    shape = tuple(dict(_parse(dim)).get(1, 0) for k, dim in _parse(shape_blob) if k == 2)
    offset = next((v for k, v in fields if k == 4), 0)
    size = next((v for k, v in fields if k == 5), 0)
    if any(k == 7 for k, _ in fields):
      raise UserException("Partitioned variable " + repr(name) + " is not supported")
    if dtype not in _DTYPES:
      continue  # strings and other non-numeric entries
    raw = np.asarray(data[offset:offset + size])
    variables[name] = raw.view(_DTYPES[dtype]).reshape(shape).copy()
  return variables


def _build_block(entries):
  """Entries without prefix compression (shared = 0), a restart point at every entry: valid, if not minimal."""
  body, restarts = bytearray(), []
  for key, value in entries:
    restarts.append(len(body))
    body += _varint(0) + _varint(len(key)) + _varint(len(value)) + key + value
  for offset in restarts or [0]:
    body += struct.pack("<I", offset)
  body += struct.pack("<I", max(1, len(restarts)))
  return bytes(body)


def write_bundle(stem, variables):
  """Write {name: numpy array} as `<stem>.index` / `<stem>.data-00000-of-00001` (single shard, uncompressed index)."""
  stem = str(stem)
  pathlib.Path(stem).parent.mkdir(parents=True, exist_ok=True)
  entries, offset = [(b"", _field_varint(1, 1) + _field_bytes(3, _field_varint(1, 1)))], 0   # header: one shard, little endian, version { producer: 1 }
  with open(stem + _DATA_SUFFIX, "wb") as fd:
    for name in sorted(variables):
      array = np.asarray(variables[name], order="C")   # (ascontiguousarray would turn a scalar into a vector)
      if array.dtype not in _DTYPE_CODES:
        raise UserException("Cannot store dtype " + str(array.dtype) + " of " + repr(name))
      raw = array.tobytes()
      shape = b"".join(_field_bytes(2, _field_varint(1, int(extent))) for extent in array.shape)
      entry = _field_varint(1, _DTYPE_CODES[array.dtype]) + _field_bytes(2, shape) + _field_varint(4, offset) + _field_varint(5, len(raw))
      entry += _varint((6 << 3) | 5) + struct.pack("<I", _masked_crc(raw))
      entries.append((name.encode(), entry))
      fd.write(raw)
      offset += len(raw)
  blob = bytearray()

  def emit(block):
    handle = _varint(len(blob)) + _varint(len(block))
    blob.extend(block + b"\x00" + struct.pack("<I", _masked_crc(block + b"\x00")))
    return handle
  data_handle = emit(_build_block(entries))
  meta_handle = emit(_build_block([]))
  index_handle = emit(_build_block([(entries[-1][0] + b"\xff", data_handle)]))
  footer = meta_handle + index_handle
  blob.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))
  pathlib.Path(stem + ".index").write_bytes(bytes(blob))


# ---------------------------------------------------------------------------- #
# Mapping onto a layout

def _candidates(name):
  yield name
  head, _, kind = name.rpartition("/")
  if head and "/" not in head:
    yield "shared/" + head + "_" + kind          # the reference's cnnet: shared/conv1_weights


def _convert(array, shape, name):
  """TensorFlow array -> array of this framework's `shape` (see the module docstring)."""
  shape = tuple(shape)
  if array.ndim == 4 and len(shape) == 4:
    if name.endswith("depthwise_weights"):
      kh, kw, c, m = array.shape
      out = array.transpose(2, 3, 0, 1).reshape(c * m, kh, kw, 1)
    else:
      out = array.transpose(3, 0, 1, 2)
  elif array.ndim == 4 and len(shape) == 2:
    out = array.transpose(3, 0, 1, 2).reshape(array.shape[3], -1)
  elif array.ndim == 2 and len(shape) == 2:
    out = array.T
  elif array.ndim == 2 and len(shape) == 4 and shape[1] == shape[2] == 1:
    out = array.T.reshape(shape)
  else:
    out = array
  if tuple(out.shape) != shape:
    raise UserException("Variable %r: checkpoint shape %r cannot be mapped to %r" % (name, tuple(array.shape), shape))
  return np.ascontiguousarray(out, dtype=np.float32)


def to_layout(variables, layout, states):
  """(flat fp32 parameter tensor of `layout`, {state name: tensor}, global step or None) from the variables of a bundle.
  `states`: {name: tensor or shape} of the non-trainable state (batch-norm moving statistics). Missing variables raise."""
  import torch
  flat = torch.zeros(layout.padded_size, dtype=torch.float32)
  views = layout.views(flat)
  used, missing = set(), []
  for name in layout.names:
    key = next((c for c in _candidates(name) if c in variables), None)
    if key is None:
      missing.append(name)
      continue
    views[name].copy_(torch.from_numpy(_convert(variables[key], views[name].shape, name)))
    used.add(key)
  if missing:
    raise UserException("Checkpoint lacks %d variable(s) of the model, e.g. %r (it holds e.g. %r)" % (len(missing), missing[:3], sorted(variables)[:3]))
  state_out = {}
  for name, like in states.items():
    key = next((c for c in _candidates(name) if c in variables), None)
    if key is None:
      warning("Checkpoint has no %r: keeping the initial value" % name)
      continue
    shape = tuple(like.shape) if hasattr(like, "shape") else tuple(like)
    state_out[name] = torch.from_numpy(_convert(variables[key], shape, name))
    used.add(key)
  step = variables.get("global_step")
  ignored = sorted(k for k in variables if k not in used and k != "global_step")
  if ignored:
    info("Ignored %d checkpoint entr%s without counterpart (optimizer slots, ...), e.g. %r" % (len(ignored), "y" if len(ignored) == 1 else "ies", ignored[:3]))
  return flat, state_out, (int(step) if step is not None else None)


def from_layout(layout, params, states, step):
  """The inverse: {TensorFlow variable name: array in TensorFlow's layout} for `write_bundle`."""
  variables = {"global_step": np.asarray(int(step), dtype=np.int64)}
  views = layout.views(params)
  for name in layout.names:
    array = views[name].detach().cpu().numpy()
    if array.ndim == 4 and name.endswith("depthwise_weights"):
      variables[name] = np.ascontiguousarray(array[..., 0].transpose(1, 2, 0)[..., None])      # [C, kh, kw, 1] -> [kh, kw, C, 1]
    elif array.ndim == 4:
      variables[name] = np.ascontiguousarray(array.transpose(1, 2, 3, 0))
    elif array.ndim == 2:
      variables[name] = np.ascontiguousarray(array.T)
    else:
      variables[name] = np.ascontiguousarray(array)
  for name, value in states.items():
    variables[name] = np.ascontiguousarray(value.detach().cpu().numpy())
  return variables


def main(argv=None):
  """`python -m aggregathor_b200.tools.tf_checkpoint export <checkpoint dir> <experiment> [experiment args ...] --output <prefix>`:
  rewrite the latest checkpoint of this framework as a TensorFlow tensor bundle; `list <prefix>` prints the variables of a bundle."""
  import argparse
  import sys
  parser = argparse.ArgumentParser(description=main.__doc__)
  parser.add_argument("command", choices=("export", "list"))
  parser.add_argument("source", help="export: checkpoint directory of this framework; list: prefix of a TensorFlow checkpoint (without .index)")
  parser.add_argument("experiment", nargs="?", help="export: experiment name (defines the variable layout)")
  parser.add_argument("experiment_args", nargs="*", default=[])
  parser.add_argument("--output", type=str, default=None, help="export: prefix of the TensorFlow checkpoint to write")
  args = parser.parse_args(sys.argv[1:] if argv is None else argv)
  if args.command == "list":
    for name, array in sorted(read_bundle(args.source).items()):
      print("%-80s %-10s %r" % (name, array.dtype, tuple(array.shape)))
    return 0
  if args.experiment is None or args.output is None:
    raise UserException("export needs an experiment name and --output")
  from .. import experiments
  from ..engine.flat import FlatLayout
  from .checkpoints import Checkpoints
  model = experiments.instantiate(args.experiment, args.experiment_args).model()
  layout, shapes = FlatLayout(), {}
  model.declare(layout, shapes)
  layout.freeze()
  state = Checkpoints(args.source).restore()
  if "tf_variables" in state:
    raise UserException("The latest checkpoint of " + repr(args.source) + " already is a TensorFlow checkpoint")
  write_bundle(args.output, from_layout(layout, state["params"], state.get("states", {}), state.get("global_step", 0)))
  info("Wrote %s.index / %s%s (%d variables, global step %d)" % (args.output, args.output, _DATA_SUFFIX, len(layout.names) + len(state.get("states", {})) + 1, state.get("global_step", 0)))
  return 0


if __name__ == "__main__":
  import sys
  sys.exit(main())

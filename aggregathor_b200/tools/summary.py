"""TensorBoard event-file writer with no TensorFlow dependency.

The reference writes `tf.summary` scalars (`learning_rate`, `eval-<metric>`) and
`SessionLog` START/STOP markers through `tf.summary.FileWriter`
(`graph.py:243,291-292`, `runner.py:479-494`). This module emits the same
`events.out.tfevents.*` files: TFRecord framing (length, masked CRC32C of the
length, payload, masked CRC32C of the payload) around hand-encoded `Event`
protobufs, so TensorBoard reads them unchanged.
"""

import os
import pathlib
import socket
import struct
import threading
import time

__all__ = ["SummaryWriter", "read_events"]

# ---- CRC32C (Castagnoli), table driven ------------------------------------- #
_CRC_TABLE = []
for _i in range(256):
  _c = _i
  for _ in range(8):
    _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
  _CRC_TABLE.append(_c)


_native_crc = None


def _crc32c(data):
  """CRC32C of a bytes-like object: the C++ routine of `native/py_gars` for anything sizeable, the table loop otherwise / as fallback."""
  global _native_crc
  if len(data) >= 4096 and _native_crc is not False:
    if _native_crc is None:
      try:
        import ctypes
        from .. import native
        _native_crc = native.library("py_gars").agb_crc32c
        _native_crc.restype = ctypes.c_uint32
        _native_crc.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint32]
      except Exception:
        _native_crc = False
    if _native_crc:
      return int(_native_crc(bytes(data), len(data), 0))
  crc = 0xFFFFFFFF
  for byte in data:
    crc = _CRC_TABLE[(crc ^ byte) & 0xFF] ^ (crc >> 8)
  return crc ^ 0xFFFFFFFF


def _masked_crc(data):
  crc = _crc32c(data)
  return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- minimal protobuf encoding ---------------------------------------------- #
def _varint(value):
  value &= (1 << 64) - 1
  out = bytearray()
  while True:
    byte = value & 0x7F
    value >>= 7
    if value:
      out.append(byte | 0x80)
    else:
      out.append(byte)
      return bytes(out)


def _field_bytes(number, payload):
  return _varint((number << 3) | 2) + _varint(len(payload)) + payload


def _field_varint(number, value):
  return _varint((number << 3) | 0) + _varint(value)


def _field_double(number, value):
  return _varint((number << 3) | 1) + struct.pack("<d", value)


def _field_float(number, value):
  return _varint((number << 3) | 5) + struct.pack("<f", value)


def _event(wall_time, step=None, file_version=None, scalars=None, session_status=None):
  msg = _field_double(1, wall_time)
  if step is not None:
    msg += _field_varint(2, int(step))
  if file_version is not None:
    msg += _field_bytes(3, file_version.encode())
  if scalars is not None:
    summary = b""
    for tag, value in scalars.items():
      summary += _field_bytes(1, _field_bytes(1, tag.encode()) + _field_float(2, float(value)))
    msg += _field_bytes(5, summary)
  if session_status is not None:
    msg += _field_bytes(7, _field_varint(1, session_status))
  return msg


class SummaryWriter:
  """Append scalar summaries to an `events.out.tfevents.<time>.<host>` file."""

  SESSION_START = 1
  SESSION_STOP = 2

  def __init__(self, logdir):
    self._dir = pathlib.Path(logdir)
    self._dir.mkdir(parents=True, exist_ok=True)
    name = "events.out.tfevents.%010d.%s.%d" % (int(time.time()), socket.gethostname(), os.getpid())
    self.path = self._dir / name
    self._fd = open(self.path, "ab")
    self._lock = threading.Lock()
    self._write(_event(time.time(), file_version="brain.Event:2"))

  def _write(self, payload):
    header = struct.pack("<Q", len(payload))
    with self._lock:
      self._fd.write(header + struct.pack("<I", _masked_crc(header)) + payload + struct.pack("<I", _masked_crc(payload)))
      self._fd.flush()

  def add_scalars(self, scalars, step):
    self._write(_event(time.time(), step=step, scalars=scalars))

  def add_session_log(self, status, step):
    self._write(_event(time.time(), step=step, session_status=status))

  def close(self):
    with self._lock:
      if not self._fd.closed:
        self._fd.close()

  def __enter__(self):
    return self

  def __exit__(self, *args):
    self.close()
    return False


# ---- reader (tests / tooling) ----------------------------------------------- #
def _read_varint(buf, pos):
  shift = value = 0
  while True:
    byte = buf[pos]
    pos += 1
    value |= (byte & 0x7F) << shift
    if not byte & 0x80:
      return value, pos
    shift += 7


def _parse(buf):
  pos, out = 0, []
  while pos < len(buf):
    key, pos = _read_varint(buf, pos)
    number, kind = key >> 3, key & 7
    if kind == 0:
      val, pos = _read_varint(buf, pos)
    elif kind == 1:
      val, pos = struct.unpack_from("<d", buf, pos)[0], pos + 8
    elif kind == 5:
      val, pos = struct.unpack_from("<f", buf, pos)[0], pos + 4
    elif kind == 2:
      size, pos = _read_varint(buf, pos)
      val, pos = bytes(buf[pos:pos + size]), pos + size
    else:
      raise ValueError("unsupported wire type " + str(kind))
    out.append((number, val))
  return out


def read_events(path):
  """Decode an event file into a list of dicts (checks both CRCs of every record)."""
  data = pathlib.Path(path).read_bytes()
  pos, events = 0, []
  while pos < len(data):
    header = data[pos:pos + 8]
    (size,) = struct.unpack("<Q", header)
    if struct.unpack_from("<I", data, pos + 8)[0] != _masked_crc(header):
      raise ValueError("corrupted record length")
    payload = data[pos + 12:pos + 12 + size]
    if struct.unpack_from("<I", data, pos + 12 + size)[0] != _masked_crc(payload):
      raise ValueError("corrupted record payload")
    pos += 16 + size
    event = {}
    for number, val in _parse(payload):
      if number == 1:
        event["wall_time"] = val
      elif number == 2:
        event["step"] = val
      elif number == 3:
        event["file_version"] = val.decode()
      elif number == 5:
        scalars = {}
        for _, value_msg in _parse(val):
          fields = dict(_parse(value_msg))
          scalars[fields[1].decode()] = fields.get(2)
        event["scalars"] = scalars
      elif number == 7:
        event["session_status"] = dict(_parse(val)).get(1)
    events.append(event)
  return events

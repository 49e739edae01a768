"""TensorFlow tensor-bundle reader / writer and the mapping onto the flat layout (the reference saves its variables with
`tf.train.Saver`, `tools/tf.py:141-173`): round trips, layout conversions, resuming a run from a TensorFlow-format checkpoint."""

import re
import struct
import subprocess
import sys
import pathlib

import numpy as np
import pytest
import torch

from aggregathor_b200 import tools
from aggregathor_b200.engine.flat import FlatLayout
from aggregathor_b200.models import get_network
from aggregathor_b200.tools import tf_checkpoint

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_bundle_roundtrip_and_format(tmp_path):
  rng = np.random.default_rng(0)
  variables = {"dense_1/weights": rng.standard_normal((784, 100)).astype(np.float32), "dense_1/biases": rng.standard_normal(100).astype(np.float32),
               "global_step": np.asarray(1234, dtype=np.int64), "scope/a/very/long/variable/name/to/cross/varint/boundaries" * 3: rng.integers(0, 9, (3, 1, 2)).astype(np.int32),
               "half": rng.standard_normal(5).astype(np.float16)}
  stem = tmp_path / "model-1234"
  tf_checkpoint.write_bundle(stem, variables)
  assert tf_checkpoint.is_tf_bundle(stem) and not tf_checkpoint.is_tf_bundle(tmp_path / "nope")
  index = (tmp_path / "model-1234.index").read_bytes()
  assert struct.unpack_from("<Q", index, len(index) - 8)[0] == 0xdb4775248b80fb57      # the table magic of the format
  assert (tmp_path / "model-1234.data-00000-of-00001").stat().st_size == sum(v.nbytes for v in variables.values())
  back = tf_checkpoint.read_bundle(stem)
  assert set(back) == set(variables)
  for name, array in variables.items():
    assert back[name].dtype == array.dtype and back[name].shape == array.shape and np.array_equal(back[name], array)


def test_prefix_compressed_blocks_are_decoded(tmp_path):
  """Real writers share key prefixes between consecutive entries and restart every 16 keys: build such a block by hand."""
  from aggregathor_b200.tools.summary import _field_bytes, _field_varint, _masked_crc, _varint
  names = ["layer/a/weights", "layer/a/weights_momentum", "layer/b/biases"]
  arrays = [np.arange(6, dtype=np.float32).reshape(2, 3), np.ones((2, 3), dtype=np.float32), np.asarray([7.0], dtype=np.float32)]
  entries, offset, data = [(b"", _field_varint(1, 1))], 0, b""
  for name, array in zip(names, arrays):
    shape = b"".join(_field_bytes(2, _field_varint(1, e)) for e in array.shape)
    entries.append((name.encode(), _field_varint(1, 1) + _field_bytes(2, shape) + _field_varint(4, offset) + _field_varint(5, array.nbytes)))
    data += array.tobytes()
    offset += array.nbytes
  body, previous = bytearray(), b""
  for key, value in entries:
    shared = 0
    while shared < min(len(key), len(previous)) and key[shared] == previous[shared]:
      shared += 1
    body += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
    previous = key
  body += struct.pack("<II", 0, 1)                                                     # one restart point (offset 0)
  blob = bytearray()
  def emit(block):
    handle = _varint(len(blob)) + _varint(len(block))
    blob.extend(bytes(block) + b"\x00" + struct.pack("<I", _masked_crc(bytes(block) + b"\x00")))
    return handle
  data_handle = emit(body)
  empty = struct.pack("<II", 0, 1)
  meta_handle = emit(empty)
  index_block = _varint(0) + _varint(1) + _varint(len(data_handle)) + b"~" + data_handle + struct.pack("<II", 0, 1)
  index_handle = emit(index_block)
  footer = meta_handle + index_handle
  blob.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57))
  (tmp_path / "m.index").write_bytes(bytes(blob))
  (tmp_path / "m.data-00000-of-00001").write_bytes(data)
  back = tf_checkpoint.read_bundle(tmp_path / "m")
  assert sorted(back) == sorted(names) and all(np.array_equal(back[n], a) for n, a in zip(names, arrays))


@pytest.mark.parametrize("name,classes", [("cnnet", 10), ("resnet_v1_18", 7), ("mobilenet_v1_025", 5), pytest.param("vgg_a", 3, marks=pytest.mark.slow)])
def test_layout_conversion_roundtrip(name, classes):
  """from_layout -> TensorFlow shapes (HWIO kernels, [in, out] dense, [kh, kw, C, 1] depthwise) -> to_layout gives the parameters back."""
  model = get_network(name, classes)
  layout, shapes = FlatLayout(), {}
  model.declare(layout, shapes)
  layout.freeze()
  params = torch.randn(layout.padded_size) * layout.mask()
  states = {k: torch.randn(v) for k, v in shapes.items()}
  variables = tf_checkpoint.from_layout(layout, params, states, 42)
  for key, array in variables.items():
    if key.endswith("depthwise_weights"):
      assert array.shape[0] == array.shape[1] and array.shape[3] == 1
    elif array.ndim == 4:
      assert array.shape[0] == array.shape[1] or 1 in array.shape[:2]        # spatial dimensions first (HWIO)
  if name == "cnnet":  # the reference's own naming for this model
    variables = {("shared/" + k.replace("/", "_")) if k != "global_step" else k: v for k, v in variables.items()}
    assert "shared/conv1_weights" in variables and variables["shared/dense3_weights"].shape == (4096, 384)
  flat, got_states, step = tf_checkpoint.to_layout(variables, layout, states)
  assert step == 42 and torch.equal(flat, params) and all(torch.equal(got_states[k], states[k]) for k in states)
  with pytest.raises(tools.UserException):
    tf_checkpoint.to_layout({k: v for k, v in variables.items() if "biases" not in k}, layout, states)


def test_runner_resumes_from_a_tensorflow_checkpoint(tmp_path):
  """A checkpoint directory holding a TensorFlow bundle (as the reference leaves it): the runner restores it, trains on and saves its own."""
  local = ["--server", '{"local": ["127.0.0.1:7000"]}', "--ps-job-name", "local", "--wk-job-name", "local", "--ev-job-name", "local", "--no-wait"]
  rng = np.random.default_rng(3)
  variables = {"dense_1/weights": (rng.standard_normal((784, 100)) * 0.05).astype(np.float32), "dense_1/biases": np.zeros(100, dtype=np.float32),
               "dense_2/weights": (rng.standard_normal((100, 10)) * 0.05).astype(np.float32), "dense_2/biases": np.zeros(10, dtype=np.float32),
               "global_step": np.asarray(300, dtype=np.int64), "dense_1/weights/Adam": np.zeros((784, 100), dtype=np.float32)}
  ckpt = tmp_path / "ckpt"
  tf_checkpoint.write_bundle(ckpt / "model-300", variables)
  cmd = [sys.executable, str(ROOT / "runner.py")] + local + ["--experiment", "mnist", "--aggregator", "average", "--nb-workers", "2", "--max-step", "3",
         "--learning-rate-args", "initial-rate:0.05", "--evaluation-file", "-", "--checkpoint-dir", str(ckpt), "--checkpoint-delta", "1000", "--checkpoint-period", "-1", "--summary-dir", "-"]
  proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300, cwd=str(ROOT))
  out = proc.stdout.decode(errors="replace")
  assert proc.returncode == 0, out[-3000:]
  assert "Imported the TensorFlow checkpoint" in out and "Step 300: total loss" in out and "Step 302: total loss" in out
  assert (ckpt / "model-303.index").exists() and not tf_checkpoint.is_tf_bundle(ckpt / "model-303")
  # and back: export the new checkpoint as a bundle, the weights come out in TensorFlow's [in, out] layout
  assert tf_checkpoint.main(["export", str(ckpt), "mnist", "--output", str(tmp_path / "exported" / "model-303")]) == 0
  exported = tf_checkpoint.read_bundle(tmp_path / "exported" / "model-303")
  assert exported["dense_1/weights"].shape == (784, 100) and int(exported["global_step"]) == 303
  assert not np.array_equal(exported["dense_1/weights"], variables["dense_1/weights"])     # it was trained on

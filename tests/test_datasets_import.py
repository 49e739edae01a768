"""Dataset import tool: slim TFRecords (hand-decoded `tf.train.Example`), MNIST IDX, CIFAR binary / pickles, image folders -> the
`.npz` layout consumed by the experiments. Source files are synthesised here in their exact on-disk formats."""

import gzip
import io
import pickle
import struct

import numpy as np
import pytest

from aggregathor_b200 import tools
from aggregathor_b200.tools import datasets
from aggregathor_b200.tools.summary import _field_bytes, _field_varint, _masked_crc, _varint


def _example(encoded, label, fmt=b"png"):
  def feature_bytes(value):
    return _field_bytes(1, _field_bytes(1, value))                      # Feature.bytes_list.value
  def feature_int(value):
    return _field_bytes(3, _field_bytes(1, _varint(value)))             # Feature.int64_list.value (packed)
  entries = b""
  for key, feature in ((b"image/encoded", feature_bytes(encoded)), (b"image/format", feature_bytes(fmt)), (b"image/class/label", feature_int(label)),
                       (b"image/height", feature_int(12)), (b"image/width", feature_int(12))):
    entries += _field_bytes(1, _field_bytes(1, key) + _field_bytes(2, feature))
  return _field_bytes(1, entries)                                       # Example.features


def _write_tfrecord(path, payloads):
  with open(path, "wb") as fd:
    for payload in payloads:
      header = struct.pack("<Q", len(payload))
      fd.write(header + struct.pack("<I", _masked_crc(header)) + payload + struct.pack("<I", _masked_crc(payload)))


def _png(array):
  from PIL import Image
  buffer = io.BytesIO()
  Image.fromarray(array).save(buffer, format="PNG")
  return buffer.getvalue()


def test_slim_tfrecords_roundtrip(tmp_path):
  rng = np.random.default_rng(0)
  images = rng.integers(0, 255, size=(7, 12, 12, 3), dtype=np.uint8)
  labels = [3, 1, 4, 1, 5, 9, 2]
  source = tmp_path / "src"
  source.mkdir()
  _write_tfrecord(source / "toy_train.tfrecord", [_example(_png(images[i]), labels[i]) for i in range(5)])
  _write_tfrecord(source / "toy_test.tfrecord", [_example(_png(images[i]), labels[i]) for i in range(5, 7)])
  target = tmp_path / "toy.npz"
  assert datasets.main(["slim", str(source), "toy", "--output", str(target)]) == 0
  with np.load(target) as blob:
    assert np.array_equal(blob["x_train"], images[:5]) and blob["y_train"].tolist() == labels[:5]
    assert np.array_equal(blob["x_test"], images[5:]) and blob["y_test"].tolist() == labels[5:]
  # resizing + label offset (slim's ImageNet convention), and the held-out split when there is no test file
  (source / "toy_test.tfrecord").unlink()
  arrays = datasets.from_slim_tfrecords(source, "toy", image_size=8, labels_offset=1)
  assert arrays[0].shape[1:] == (8, 8, 3) and len(arrays[1]) + len(arrays[3]) == 5 and int(arrays[1][0]) == labels[0] - 1
  # a corrupted length checksum is reported, not swallowed
  raw = bytearray((source / "toy_train.tfrecord").read_bytes())
  raw[9] ^= 0xFF
  (source / "toy_train.tfrecord").write_bytes(bytes(raw))
  with pytest.raises(tools.UserException):
    datasets.from_slim_tfrecords(source, "toy")


def test_mnist_idx_and_cifar_formats(tmp_path):
  rng = np.random.default_rng(1)
  mnist = tmp_path / "mnist"
  mnist.mkdir()
  x, y = rng.integers(0, 255, size=(6, 28, 28), dtype=np.uint8), rng.integers(0, 10, size=6, dtype=np.uint8)
  for stem, array, gz in (("train-images-idx3-ubyte", x[:4], True), ("train-labels-idx1-ubyte", y[:4], False), ("t10k-images-idx3-ubyte", x[4:], False), ("t10k-labels-idx1-ubyte", y[4:], True)):
    blob = struct.pack(">HBB", 0, 8, array.ndim) + struct.pack(">" + "I" * array.ndim, *array.shape) + array.tobytes()
    if gz:
      with gzip.open(str(mnist / stem) + ".gz", "wb") as fd:
        fd.write(blob)
    else:
      (mnist / stem).write_bytes(blob)
  x_train, y_train, x_test, y_test = datasets.from_mnist_idx(mnist)
  assert x_train.shape == (4, 28, 28, 1) and np.array_equal(x_train[..., 0], x[:4]) and y_test.tolist() == y[4:].tolist()
  # CIFAR-10 binary: label byte + 3 x 1024 planar bytes
  images = rng.integers(0, 255, size=(5, 32, 32, 3), dtype=np.uint8)
  labels = rng.integers(0, 10, size=5, dtype=np.uint8)
  planar = images.transpose(0, 3, 1, 2).reshape(5, -1)
  binary = tmp_path / "bin"
  binary.mkdir()
  (binary / "data_batch_1.bin").write_bytes(np.concatenate([labels[:3, None], planar[:3]], axis=1).tobytes())
  (binary / "test_batch.bin").write_bytes(np.concatenate([labels[3:, None], planar[3:]], axis=1).tobytes())
  got = datasets.from_cifar(binary)
  assert np.array_equal(got[0], images[:3]) and got[1].tolist() == labels[:3].tolist() and np.array_equal(got[2], images[3:])
  # CIFAR-10 python pickles
  py = tmp_path / "py"
  py.mkdir()
  for name, lo, hi in (("data_batch_1", 0, 3), ("test_batch", 3, 5)):
    with open(py / name, "wb") as fd:
      pickle.dump({b"data": planar[lo:hi], b"labels": labels[lo:hi].tolist()}, fd)
  got = datasets.from_cifar(py)
  assert np.array_equal(got[0], images[:3]) and got[3].tolist() == labels[3:].tolist()


def test_image_folder_and_experiment_pickup(tmp_path, monkeypatch):
  from PIL import Image
  rng = np.random.default_rng(2)
  for split, count in (("train", 4), ("val", 2)):
    for cls in ("cat", "dog"):
      (tmp_path / "src" / split / cls).mkdir(parents=True)
      for i in range(count):
        Image.fromarray(rng.integers(0, 255, size=(20, 30, 3), dtype=np.uint8)).save(tmp_path / "src" / split / cls / ("%d.png" % i))
  root = tmp_path / "datasets"
  target = root / "pets" / "pets.npz"
  assert datasets.main(["folder", str(tmp_path / "src"), "pets", "--image-size", "16", "--output", str(target)]) == 0
  with np.load(target) as blob:
    assert blob["x_train"].shape == (8, 16, 16, 3) and sorted(set(blob["y_train"].tolist())) == [0, 1] and blob["x_test"].shape[0] == 4
  # the experiments' loader finds it through $AGB_DATASETS
  monkeypatch.setenv("AGB_DATASETS", str(root))
  from aggregathor_b200.experiments._data import Dataset
  data = Dataset("pets")
  assert not data.synthetic and data.classes == 2 and data.shape == (16, 16, 3)

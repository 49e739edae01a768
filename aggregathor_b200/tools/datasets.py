"""Import datasets prepared for the reference (or downloaded in their canonical formats) into the `.npz` layout of
`experiments/_data.py` (`x_train, y_train, x_test, y_test`; uint8 NHWC images, integer labels).

The reference feeds its experiments from `experiments/datasets/<name>/` directories holding the TFRecords written by
tensorflow/models `research/slim` (`download_and_convert_data.py`; `experiments/cnnet.py:123-132`, `experiments/slims.py:100-120`)
and MNIST from `tf.keras.datasets.mnist.load_data()` (`experiments/mnist.py:114`). A user switching over has exactly those files,
so they are read here without TensorFlow: the TFRecord framing and the `tf.train.Example` protobuf are decoded by hand, images
by Pillow.

    python -m aggregathor_b200.tools.datasets slim   <dir with cifar10_train.tfrecord ...> cifar10
    python -m aggregathor_b200.tools.datasets slim   <dir with train-00000-of-01024 ...>   imagenet --image-size 224 --limit 50000
    python -m aggregathor_b200.tools.datasets mnist  <dir with train-images-idx3-ubyte(.gz) ...> mnist
    python -m aggregathor_b200.tools.datasets cifar  <cifar-10-batches-py | cifar-10-batches-bin> cifar10
    python -m aggregathor_b200.tools.datasets folder <dir with train/<class>/*.jpg and val|test/<class>/*.jpg> flowers --image-size 224

The result goes to `experiments/datasets/<name>/<name>.npz` (or `--output`); keras' own `mnist.npz` is picked up as it is.

Datasets larger than host memory are written as *shards* for the streaming reader (`native/py_loader`, `experiments/_data.py`):

    python -m aggregathor_b200.tools.datasets slim <dir with train-00000-of-01024 ...> imagenet --shards 0 --store-size 256 --labels-offset 1

`--shards 0` = one shard per source TFRecord file, converted one file at a time (nothing but the current file is held in memory);
`--shards N` (> 0) splits an in-memory import into N training shards. Images are stored at `--store-size` (short side resized, centre
crop to a square): slim's aspect-preserving resize / distorted-box crops are then taken from the stored image on the device.
"""

import argparse
import gzip
import io
import pathlib
import pickle
import struct
import sys

import numpy as np

from . import UserException, info, warning
from .summary import _masked_crc, _parse


# ---------------------------------------------------------------------------- #
# TFRecord + tf.train.Example

def read_tfrecords(path, verify=False):
  """Yield the payload of every record of a TFRecord file. The 4-byte length checksum is always checked; the payload checksum
  only with `verify` (pure-Python CRC32C: slow on image-sized records)."""
  data = pathlib.Path(path).read_bytes()
  pos = 0
  while pos < len(data):
    if pos + 12 > len(data):
      raise UserException("Truncated TFRecord file " + repr(str(path)))
    header = data[pos:pos + 8]
    (size,) = struct.unpack("<Q", header)
    if struct.unpack_from("<I", data, pos + 8)[0] != _masked_crc(header):
      raise UserException("Corrupted record length in " + repr(str(path)))
    payload = data[pos + 12:pos + 12 + size]
    if len(payload) != size:
      raise UserException("Truncated TFRecord file " + repr(str(path)))
    if verify and struct.unpack_from("<I", data, pos + 12 + size)[0] != _masked_crc(payload):
      raise UserException("Corrupted record payload in " + repr(str(path)))
    pos += 16 + size
    yield payload


def _packed_varints(buf):
  out, pos = [], 0
  while pos < len(buf):
    shift = value = 0
    while True:
      byte = buf[pos]
      pos += 1
      value |= (byte & 0x7F) << shift
      if not byte & 0x80:
        break
      shift += 7
    out.append(value - (1 << 64) if value >= 1 << 63 else value)
  return out


def parse_example(payload):
  """`tf.train.Example` bytes -> {feature name: list of bytes | ints | floats}."""
  features = {}
  for number, blob in _parse(payload):                    # Example { Features features = 1; }
    if number != 1:
      continue
    for number2, entry in _parse(blob):                   # Features { map<string, Feature> feature = 1; }
      if number2 != 1:
        continue
      key, value = None, None
      for number3, field in _parse(entry):                # map entry { string key = 1; Feature value = 2; }
        if number3 == 1:
          key = field.decode()
        elif number3 == 2:
          value = field
      if key is None or value is None:
        continue
      values = []
      for kind, body in _parse(value):                    # Feature { BytesList = 1 | FloatList = 2 | Int64List = 3 }
        if kind == 1:
          values += [item for tag, item in _parse(body) if tag == 1]
        elif kind == 2:
          for tag, item in _parse(body):
            values += list(struct.unpack("<%df" % (len(item) // 4), item)) if isinstance(item, bytes) else [item]
        elif kind == 3:
          for tag, item in _parse(body):
            values += _packed_varints(item) if isinstance(item, bytes) else [item - (1 << 64) if item >= 1 << 63 else item]
      features[key] = values
  return features


def _decode_image(encoded, image_size=None, keep_aspect=False):
  from PIL import Image
  image = Image.open(io.BytesIO(encoded))
  image = image.convert("L" if image.mode in ("L", "1") else "RGB")
  if image_size is not None and image.size != (image_size, image_size):
    if keep_aspect:   # short side -> image_size, then the central square
      w, h = image.size
      scale = image_size / min(w, h)
      nw, nh = max(image_size, round(w * scale)), max(image_size, round(h * scale))
      image = image.resize((nw, nh), Image.BILINEAR)
      left, top = (nw - image_size) // 2, (nh - image_size) // 2
      image = image.crop((left, top, left + image_size, top + image_size))
    else:
      image = image.resize((image_size, image_size), Image.BILINEAR)
  array = np.asarray(image, dtype=np.uint8)
  return array[..., None] if array.ndim == 2 else array


def from_slim_tfrecords(directory, name, image_size=None, limit=None, labels_offset=0):
  """The splits written by slim's converters: files whose name contains `train` and files containing `validation` or `test`."""
  directory = pathlib.Path(directory)
  files = sorted(p for p in directory.iterdir() if p.is_file() and ("tfrecord" in p.name or "-of-" in p.name))
  splits = {"train": [p for p in files if "train" in p.name], "test": [p for p in files if "validation" in p.name or "test" in p.name]}
  if not splits["train"]:
    raise UserException("No '*train*' TFRecord file found in " + repr(str(directory)))
  out = {}
  for split, paths in splits.items():
    images, labels = [], []
    for path in paths:
      for payload in read_tfrecords(path):
        example = parse_example(payload)
        if "image/encoded" not in example or "image/class/label" not in example:
          raise UserException("Record without 'image/encoded' / 'image/class/label' in " + repr(str(path)))
        images.append(_decode_image(example["image/encoded"][0], image_size))
        labels.append(int(example["image/class/label"][0]) - labels_offset)
        if limit is not None and len(images) >= limit:
          break
      if limit is not None and len(images) >= limit:
        break
    shapes = {image.shape for image in images}
    if len(shapes) > 1:
      raise UserException("Images of different sizes in split %r of %r: pass --image-size" % (split, name))
    out[split] = (np.stack(images) if images else np.zeros((0,), dtype=np.uint8), np.asarray(labels, dtype=np.int64))
  if not len(out["test"][1]):
    warning("No validation / test split found for %r: the last 10%% of the training split is held out" % name)
    x, y = out["train"]
    cut = max(1, len(y) // 10)
    out = {"train": (x[:-cut], y[:-cut]), "test": (x[-cut:], y[-cut:])}
  return out["train"] + out["test"]


# ---------------------------------------------------------------------------- #
# Canonical download formats

def _open_maybe_gz(path):
  path = pathlib.Path(path)
  if path.is_file():
    return open(path, "rb")
  if pathlib.Path(str(path) + ".gz").is_file():
    return gzip.open(str(path) + ".gz", "rb")
  raise UserException("File " + repr(str(path)) + "(.gz) not found")


def _read_idx(path):
  with _open_maybe_gz(path) as fd:
    data = fd.read()
  zero, dtype, dims = struct.unpack_from(">HBB", data, 0)
  if zero != 0 or dtype != 8:
    raise UserException("Not an unsigned-byte IDX file: " + repr(str(path)))
  shape = struct.unpack_from(">" + "I" * dims, data, 4)
  return np.frombuffer(data, dtype=np.uint8, offset=4 + 4 * dims).reshape(shape)


def from_mnist_idx(directory):
  directory = pathlib.Path(directory)
  x_train, y_train = _read_idx(directory / "train-images-idx3-ubyte"), _read_idx(directory / "train-labels-idx1-ubyte")
  x_test, y_test = _read_idx(directory / "t10k-images-idx3-ubyte"), _read_idx(directory / "t10k-labels-idx1-ubyte")
  return x_train[..., None], y_train.astype(np.int64), x_test[..., None], y_test.astype(np.int64)


def from_cifar(directory):
  """`cifar-10-batches-py` (pickles), `cifar-10-batches-bin` (1 label byte + 3072 pixel bytes per image) or the CIFAR-100 pickles."""
  directory = pathlib.Path(directory)

  def planar(rows):
    return np.ascontiguousarray(rows.reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1))

  pickles = sorted(directory.glob("data_batch_*")) if not list(directory.glob("*.bin")) else []
  if pickles or (directory / "train").is_file():
    def load(path):
      with open(path, "rb") as fd:
        blob = pickle.load(fd, encoding="bytes")
      labels = blob.get(b"labels", blob.get(b"fine_labels"))
      return planar(np.asarray(blob[b"data"], dtype=np.uint8)), np.asarray(labels, dtype=np.int64)
    train = [load(p) for p in (pickles or [directory / "train"])]
    test = load(directory / ("test_batch" if pickles else "test"))
    return np.concatenate([t[0] for t in train]), np.concatenate([t[1] for t in train]), test[0], test[1]
  bins = sorted(directory.glob("data_batch_*.bin"))
  if not bins:
    raise UserException("No CIFAR batch files in " + repr(str(directory)))

  def load_bin(path):
    raw = np.frombuffer(pathlib.Path(path).read_bytes(), dtype=np.uint8).reshape(-1, 3073)
    return planar(raw[:, 1:]), raw[:, 0].astype(np.int64)
  train = [load_bin(p) for p in bins]
  test = load_bin(directory / "test_batch.bin")
  return np.concatenate([t[0] for t in train]), np.concatenate([t[1] for t in train]), test[0], test[1]


def from_image_folder(directory, image_size, limit=None):
  """`train/<class>/*` and `val|validation|test/<class>/*` image files; classes are numbered in sorted order of the training split."""
  directory = pathlib.Path(directory)
  test_dir = next((directory / n for n in ("val", "validation", "test") if (directory / n).is_dir()), None)
  if not (directory / "train").is_dir():
    raise UserException("Expected a 'train' directory in " + repr(str(directory)))
  classes = sorted(p.name for p in (directory / "train").iterdir() if p.is_dir())

  def load(root):
    images, labels = [], []
    for label, cls in enumerate(classes):
      for path in sorted((root / cls).glob("*")) if (root / cls).is_dir() else []:
        try:
          images.append(_decode_image(path.read_bytes(), image_size))
        except Exception:
          continue
        labels.append(label)
        if limit is not None and len(images) >= limit:
          return np.stack(images), np.asarray(labels, dtype=np.int64)
    return (np.stack(images) if images else np.zeros((0, image_size, image_size, 3), dtype=np.uint8)), np.asarray(labels, dtype=np.int64)
  x_train, y_train = load(directory / "train")
  if test_dir is not None:
    x_test, y_test = load(test_dir)
  else:
    order = np.random.default_rng(0).permutation(len(y_train))
    cut = max(1, len(order) // 10)
    x_test, y_test, x_train, y_train = x_train[order[:cut]], y_train[order[:cut]], x_train[order[cut:]], y_train[order[cut:]]
  return x_train, y_train, x_test, y_test


# ---------------------------------------------------------------------------- #
# Shards for the streaming reader

def _shard_dir(name, output=None):
  from ..experiments._data import DATASETS_DIR
  target = pathlib.Path(output) if output else DATASETS_DIR / name
  target.mkdir(parents=True, exist_ok=True)
  return target


def write_shards(name, x_train, y_train, x_test, y_test, shards, output=None):
  """Split an in-memory import into `shards` training shards (+ 1 test shard per 8) under `experiments/datasets/<name>/`."""
  import json
  from ..experiments._data import SHARD_SUFFIX, write_shard
  target = _shard_dir(name, output)
  written = []
  for split, x, y, count in (("train", x_train, y_train, max(1, shards)), ("test", x_test, y_test, max(1, shards // 8))):
    bounds = np.linspace(0, len(y), count + 1).astype(np.int64)
    for i in range(count):
      if bounds[i + 1] > bounds[i]:
        written.append(write_shard(target / ("%s-%05d-of-%05d%s" % (split, i, count, SHARD_SUFFIX)), x[bounds[i]:bounds[i + 1]], y[bounds[i]:bounds[i + 1]]))
  classes = int(max(np.max(y_train), np.max(y_test) if len(y_test) else 0)) + 1
  (target / "meta.json").write_text(json.dumps({"name": name, "classes": classes, "train": int(len(y_train)), "test": int(len(y_test)), "shape": list(np.asarray(x_train).shape[1:])}))
  info("Dataset %r: %d training and %d test images in %d shard(s) -> %s" % (name, len(y_train), len(y_test), len(written), target))
  return written


def slim_to_shards(directory, name, store_size, labels_offset=0, limit=None, output=None):
  """One shard per slim TFRecord file, one file in memory at a time (ImageNet: 1024 training + 128 validation files)."""
  import json
  from ..experiments._data import SHARD_SUFFIX, write_shard
  directory = pathlib.Path(directory)
  files = sorted(p for p in directory.iterdir() if p.is_file() and ("tfrecord" in p.name or "-of-" in p.name))
  splits = {"train": [p for p in files if "train" in p.name], "test": [p for p in files if "validation" in p.name or "test" in p.name]}
  if not splits["train"]:
    raise UserException("No '*train*' TFRecord file found in " + repr(str(directory)))
  target = _shard_dir(name, output)
  counts, top_label, shape = {}, 0, None
  for split, paths in splits.items():
    counts[split] = 0
    for index, path in enumerate(paths):
      images, labels = [], []
      for payload in read_tfrecords(path):
        example = parse_example(payload)
        if "image/encoded" not in example or "image/class/label" not in example:
          raise UserException("Record without 'image/encoded' / 'image/class/label' in " + repr(str(path)))
        images.append(_decode_image(example["image/encoded"][0], store_size, keep_aspect=True))
        labels.append(int(example["image/class/label"][0]) - labels_offset)
        if limit is not None and counts[split] + len(images) >= limit:
          break
      if not images:
        continue
      if len({image.shape for image in images}) > 1:
        raise UserException("Images of different sizes in %r: pass --store-size" % str(path))
      write_shard(target / ("%s-%05d-of-%05d%s" % (split, index, len(paths), SHARD_SUFFIX)), np.stack(images), np.asarray(labels, dtype=np.int64))
      counts[split] += len(images)
      top_label = max(top_label, max(labels))
      shape = images[0].shape
      if limit is not None and counts[split] >= limit:
        break
  (target / "meta.json").write_text(json.dumps({"name": name, "classes": top_label + 1, "train": counts["train"], "test": counts.get("test", 0), "shape": list(shape)}))
  info("Dataset %r: %d training and %d test images streamed into shards -> %s" % (name, counts["train"], counts.get("test", 0), target))
  return target


def write_npz(name, x_train, y_train, x_test, y_test, output=None):
  from ..experiments._data import DATASETS_DIR
  target = pathlib.Path(output) if output else DATASETS_DIR / name / (name + ".npz")
  target.parent.mkdir(parents=True, exist_ok=True)
  np.savez(target, x_train=np.asarray(x_train, dtype=np.uint8), y_train=np.asarray(y_train, dtype=np.int64),
           x_test=np.asarray(x_test, dtype=np.uint8), y_test=np.asarray(y_test, dtype=np.int64))
  info("Dataset %r: %d training and %d test images of shape %r, %d classes -> %s" % (
    name, len(y_train), len(y_test), tuple(np.asarray(x_train).shape[1:]), int(max(np.max(y_train), np.max(y_test))) + 1 if len(y_train) else 0, target))
  return target


def main(argv=None):
  parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
  parser.add_argument("format", choices=("slim", "mnist", "cifar", "folder"))
  parser.add_argument("source", help="directory holding the files of the chosen format")
  parser.add_argument("name", help="dataset name (the <dataset> of slim-<model>-<dataset>)")
  parser.add_argument("--image-size", type=int, default=None, help="resize every image to this square size (needed for ImageNet-like sources)")
  parser.add_argument("--limit", type=int, default=None, help="keep at most this many images per split (everything is held in memory)")
  parser.add_argument("--labels-offset", type=int, default=0, help="subtracted from the stored labels (slim's ImageNet labels start at 1)")
  parser.add_argument("--output", type=str, default=None, help="target .npz (default: experiments/datasets/<name>/<name>.npz) or shard directory")
  parser.add_argument("--shards", type=int, default=None, help="write shards for the streaming reader instead of one .npz: N training shards, 0 = one per source file (slim)")
  parser.add_argument("--store-size", type=int, default=None, help="storage resolution of sharded images (short side resized + central square crop); default --image-size")
  args = parser.parse_args(sys.argv[1:] if argv is None else argv)
  if args.shards is not None and args.format == "slim" and args.shards == 0:
    slim_to_shards(args.source, args.name, args.store_size or args.image_size, args.labels_offset, args.limit, args.output)
    return 0
  if args.format == "slim":
    arrays = from_slim_tfrecords(args.source, args.name, args.image_size, args.limit, args.labels_offset)
  elif args.format == "mnist":
    arrays = from_mnist_idx(args.source)
  elif args.format == "cifar":
    arrays = from_cifar(args.source)
  else:
    if args.image_size is None:
      raise UserException("'folder' sources need --image-size")
    arrays = from_image_folder(args.source, args.image_size, args.limit)
  if args.shards is not None:
    write_shards(args.name, *arrays, shards=args.shards, output=args.output)
  else:
    write_npz(args.name, *arrays, output=args.output)
  return 0


if __name__ == "__main__":
  sys.exit(main())

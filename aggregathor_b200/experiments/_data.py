"""Datasets: on-disk loaders when the files exist, deterministic synthetic stand-ins otherwise.

The reference reads MNIST through `tf.keras.datasets.mnist.load_data()` and CIFAR-10 / ImageNet TFRecords through
slim from user-provided `experiments/datasets/<name>` directories (`README.md:190-195`). This sandbox has no network
and no datasets, so every dataset has a *synthetic* twin of the same shape and dtype: class-conditional prototypes
plus noise (learnable, so the training loop still shows a falling loss / rising accuracy). Real data is picked up from
`experiments/datasets/<name>/` (`.npz` with `x_train,y_train,x_test,y_test`, uint8 NHWC) or `$AGB_DATASETS/<name>/`.

`BatchStream` is the input pipeline: a background thread assembles batches into pinned host buffers and issues the
host->device copies on a side stream, double-buffered, so the step never waits on the input unless it outruns it.
"""

import os
import pathlib
import queue
import threading

import numpy as np
import torch

from .. import tools

DATASETS_DIR = pathlib.Path(__file__).parent / "datasets"

_SHAPES = {  # name -> (train size, test size, (H, W, C), classes)
  "mnist": (60000, 10000, (28, 28, 1), 10),
  "cifar10": (50000, 10000, (32, 32, 3), 10),
  "cifar100": (50000, 10000, (32, 32, 3), 100),
  "flowers": (3320, 350, (224, 224, 3), 5),
  "imagenet": (1281167, 50000, (224, 224, 3), 1000)}


def known_datasets():
  names = dict(_SHAPES)
  for root in (DATASETS_DIR, pathlib.Path(os.environ.get("AGB_DATASETS", "/nonexistent"))):
    if root.is_dir():
      for path in root.iterdir():
        if path.is_dir() and path.name not in names and tools.can_access(path, read=True):
          names[path.name] = None
  return names


def _find_npz(name):
  for root in (DATASETS_DIR, pathlib.Path(os.environ.get("AGB_DATASETS", "/nonexistent")), pathlib.Path.home() / ".keras" / "datasets"):
    for candidate in (root / name / (name + ".npz"), root / (name + ".npz")):
      if candidate.is_file():
        return candidate
  return None


class Dataset:
  """In-memory (or lazily synthesised) classification dataset: uint8 NHWC images + int64 labels."""

  def __init__(self, name, image_size=None, synthetic_limit=4096, seed=1234):
    self.name = name
    path = _find_npz(name)
    if path is not None:
      with np.load(path) as blob:
        self.x_train, self.y_train = blob["x_train"], blob["y_train"].astype(np.int64).reshape(-1)
        self.x_test, self.y_test = blob["x_test"], blob["y_test"].astype(np.int64).reshape(-1)
      if self.x_train.ndim == 3:
        self.x_train, self.x_test = self.x_train[..., None], self.x_test[..., None]
      self.synthetic = False
      self.classes = int(self.y_train.max()) + 1
      self.shape = tuple(self.x_train.shape[1:])
      return
    if name not in _SHAPES or _SHAPES[name] is None:
      raise tools.UserException("Dataset " + repr(name) + " not found (expected " + repr(str(DATASETS_DIR / name / (name + ".npz"))) + ")")
    ntrain, ntest, shape, classes = _SHAPES[name]
    if image_size is not None:
      shape = (image_size, image_size, shape[2])
    self.synthetic, self.classes, self.shape = True, classes, shape
    rng = np.random.default_rng(seed)
    # low-resolution class prototypes, upsampled: cheap to store, still linearly separable
    low = max(4, min(8, shape[0] // 4))
    self._protos = rng.integers(32, 224, size=(classes, low, low, shape[2]), dtype=np.uint8)
    ntrain, ntest = min(ntrain, synthetic_limit), min(ntest, max(256, synthetic_limit // 4))
    self.y_train = rng.integers(0, classes, size=ntrain, dtype=np.int64)
    self.y_test = rng.integers(0, classes, size=ntest, dtype=np.int64)
    self.x_train = self._render(self.y_train, rng)
    self.x_test = self._render(self.y_test, rng)

  def _render(self, labels, rng):
    h, w, c = self.shape
    low = self._protos.shape[1]
    reps_h, reps_w = -(-h // low), -(-w // low)
    out = np.empty((len(labels), h, w, c), dtype=np.uint8)
    for start in range(0, len(labels), 512):
      chunk = labels[start:start + 512]
      imgs = np.repeat(np.repeat(self._protos[chunk], reps_h, axis=1), reps_w, axis=2)[:, :h, :w, :].astype(np.int16)
      imgs += rng.integers(-48, 48, size=imgs.shape, dtype=np.int16)
      out[start:start + 512] = np.clip(imgs, 0, 255).astype(np.uint8)
    return out


class BatchStream:
  """Infinite shuffled batch iterator with pinned double buffering and asynchronous H2D copies."""

  def __init__(self, images, labels, batch_size, device, seed=0, shuffle=True, depth=2, transform=None):
    self.images, self.labels, self.batch = images, labels, int(batch_size)
    self.device = torch.device(device)
    self.shuffle, self.transform = shuffle, transform
    self._rng = np.random.default_rng(seed)
    self._order = np.arange(len(labels))
    self._cursor = len(labels)  # forces a shuffle at first use
    self._cuda = self.device.type == "cuda"
    self._depth = depth
    self._queue = None
    self._thread = None
    self._stop = False
    self.h2d_bytes = self.batch * int(np.prod(images.shape[1:])) * images.dtype.itemsize + self.batch * 8

  def _next_indices(self):
    if self._cursor + self.batch > len(self._order):
      if self.shuffle:
        self._rng.shuffle(self._order)
      self._cursor = 0
    idx = self._order[self._cursor:self._cursor + self.batch]
    self._cursor += self.batch
    return np.sort(idx) if not self.shuffle else idx

  def _host_batch(self, slot=None):
    idx = self._next_indices()
    x = torch.from_numpy(self.images[idx])
    y = torch.from_numpy(self.labels[idx])
    if slot is not None:
      slot[0].copy_(x)
      slot[1].copy_(y)
      return slot
    return x, y

  def _producer(self):
    torch.cuda.set_device(self.device)
    stream = torch.cuda.Stream(self.device)
    shape = (self.batch,) + tuple(self.images.shape[1:])
    slots = [(torch.empty(shape, dtype=torch.from_numpy(self.images[:1]).dtype).pin_memory(), torch.empty(self.batch, dtype=torch.int64).pin_memory()) for _ in range(self._depth + 2)]
    i = 0
    events = [None] * len(slots)
    while not self._stop:
      if events[i % len(slots)] is not None:
        events[i % len(slots)].synchronize()  # the previous copy out of this pinned slot has completed
      host = self._host_batch(slots[i % len(slots)])
      with torch.cuda.stream(stream):
        x = host[0].to(self.device, non_blocking=True)
        y = host[1].to(self.device, non_blocking=True)
        done = torch.cuda.Event()
        done.record(stream)
      events[i % len(slots)] = done
      while not self._stop:
        try:
          self._queue.put((x, y, done), timeout=0.1)
          break
        except queue.Full:
          continue
      i += 1

  def __iter__(self):
    return self

  def __next__(self):
    if not self._cuda:
      x, y = self._host_batch()
      x, y = x.clone(), y.clone()
    else:
      if self._thread is None:
        self._queue = queue.Queue(maxsize=self._depth)
        self._thread = threading.Thread(target=self._producer, name="input", daemon=True)
        self._thread.start()
      x, y, done = self._queue.get()
      consumer = torch.cuda.current_stream(self.device)
      consumer.wait_event(done)
      # allocated on the producer's side stream, consumed here: tell the caching allocator, or the block could be handed back to the
      # producer's next H2D copy while kernels of this stream still read it
      x.record_stream(consumer)
      y.record_stream(consumer)
    return (x, y) if self.transform is None else self.transform(x, y)

  def close(self):
    self._stop = True

"""Stand-alone collectives over peer-mapped memory: all-reduce and (variable first dimension) all-gather.

The reference ships these as MPI-backed TensorFlow ops — `MPIInit / MPISize / MPIRank / MPILocalRank / MPIAllreduce /
MPIAllgather` (`tf_patches/kernels/mpi_ops.cc:861-1126`) over ring algorithms (`tf_patches/kernels/ring.h:155-318`) with an
element-wise accumulate CUDA kernel per ring step (`ring.cu.cc:88-105`) and a rank-0 coordinator thread that matches tensor
names across ranks (`mpi_ops.cc:573-795`). Here:

* SPMD lock-step replaces the coordinator: every rank issues the same collectives in the same order on its stream;
* the data plane is ONE kernel per rank (`native/op_comm`): `multimem.ld_reduce` + `multimem.st` through the NVSwitch for
  fp32 all-reduce (P2P loads/stores for int32 / int64 or without a multicast mapping), P2P pulls for all-gather, device-side
  flag barriers in the symmetric signal pad — no NCCL call, no ring, no host round trip;
* on CPU tensors (the gloo plumbing configuration) the calls go through `torch.distributed`.

Module-level functions mirror the reference's op set: `init()`, `size()`, `rank()`, `local_rank()`, `allreduce(t)`, `allgather(t)`.
"""

import ctypes
import os

import torch
import torch.distributed as dist

from .. import tools
from .symm import SymmetricHeap

_MAX_RANKS = 16
_DTYPES = {torch.float32: 0, torch.int32: 1, torch.int64: 2}


class Communicator:
  """A symmetric staging buffer of `capacity` bytes per rank + the kernels operating on it."""

  def __init__(self, capacity=256 << 20, device=None, group=None, max_blocks=0):
    self.group = group
    self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    self.rank = dist.get_rank(group) if self.world > 1 else 0
    if self.world > _MAX_RANKS:
      raise tools.UserException("At most %d ranks per communicator" % _MAX_RANKS)
    if device is None:
      device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    self.device = torch.device(device)
    self.capacity = (int(capacity) + 1023) // 1024 * 1024
    self.max_blocks = max_blocks
    self.epoch = 0
    self.heap = None
    if self.device.type == "cuda":
      from .. import native
      self._lib = native.library("op_comm")
      signal_bytes = 2 * _MAX_RANKS * 4
      self.heap = SymmetricHeap(SymmetricHeap.required(self.capacity, signal_bytes), self.device, group)
      self.heap.region("data", self.capacity)
      self.heap.region("signals", signal_bytes)
      self._counter = torch.zeros(1, dtype=torch.int32, device=self.device)
      self._ptrs = (ctypes.c_ulonglong * (2 * _MAX_RANKS + 3))()
      for r in range(self.world):
        self._ptrs[r] = self.heap.peer(r, "data")
        self._ptrs[_MAX_RANKS + r] = self.heap.peer(r, "signals")
      self._ptrs[2 * _MAX_RANKS] = self.heap.multicast("data") if self.world > 1 else 0
      self._ptrs[2 * _MAX_RANKS + 1] = self._counter.data_ptr()

  # -- introspection (MPISize / MPIRank / MPILocalRank) ------------------------- #
  def size(self):
    return self.world

  def local_rank(self):
    return int(os.environ.get("LOCAL_RANK", self.rank))

  @property
  def multicast(self):
    return bool(self.heap is not None and self._ptrs[2 * _MAX_RANKS])

  # -- zero-copy interface -------------------------------------------------------- #
  def buffer(self, numel, dtype=torch.float32):
    """Tensor view of the first `numel` elements of this rank's symmetric buffer: fill it, call `allreduce_`, read it."""
    nbytes = numel * torch.empty((), dtype=dtype).element_size()
    if nbytes > self.capacity:
      raise tools.UserException("Collective of %d bytes exceeds the communicator capacity (%d bytes)" % (nbytes, self.capacity))
    return self.heap.local("data", torch.uint8)[:nbytes].view(dtype)

  def _stream(self):
    return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

  def allreduce_(self, numel, dtype=torch.float32, mean=False):
    """In-place sum (or mean) across ranks of the first `numel` elements of the symmetric buffer (padded to 16 bytes)."""
    if dtype not in _DTYPES:
      raise tools.UserException("all-reduce supports float32, int32 and int64 tensors, got " + str(dtype))
    size = torch.empty((), dtype=dtype).element_size()
    vectors = (numel * size + 15) // 16
    self.epoch += 1
    status = self._lib.agb_comm_allreduce(self._ptrs, ctypes.c_longlong(vectors), ctypes.c_int(_DTYPES[dtype]), ctypes.c_int(1 if mean else 0), ctypes.c_int(self.world),
                                          ctypes.c_int(self.rank), ctypes.c_uint(self.epoch), ctypes.c_int(self.max_blocks), self._stream())
    if status != 0:
      raise RuntimeError("agb_comm_allreduce failed with status %d" % status)

  # -- tensor interface (MPIAllreduce / MPIAllgather) ------------------------------ #
  def allreduce(self, tensor, mean=False):
    """Sum (mean) of `tensor` over the ranks; same shape on every rank. Returns a new tensor."""
    if tensor.device.type != "cuda":
      out = tensor.clone()
      if self.world > 1:
        dist.all_reduce(out, group=self.group)
      return out / self.world if mean else out
    flat = tensor.contiguous().view(-1)
    staged = self.buffer((flat.numel() * flat.element_size() + 15) // 16 * 16 // flat.element_size(), flat.dtype)
    staged[:flat.numel()].copy_(flat)
    staged[flat.numel():].zero_()
    self.allreduce_(flat.numel(), flat.dtype, mean)
    return staged[:flat.numel()].clone().view(tensor.shape)

  def allgather(self, tensor, counts=None):
    """Concatenation along dimension 0 of every rank's `tensor` (first dimensions may differ, the others must match).
    `counts`: every rank's first dimension when the caller knows them — skips the size exchange, which is a host round trip."""
    if tensor.dim() == 0:
      raise tools.UserException("all-gather needs tensors of rank >= 1")
    rows = torch.tensor([tensor.shape[0]], dtype=torch.int64)
    if counts is not None:
      if len(counts) != self.world or int(counts[self.rank]) != tensor.shape[0]:
        raise tools.UserException("all-gather: `counts` must list every rank's first dimension")
      counts = [torch.tensor([int(c)]) for c in counts]
    elif self.world > 1:
      counts = [torch.zeros(1, dtype=torch.int64) for _ in range(self.world)]
      if tensor.device.type == "cuda":  # control plane: sizes travel through the process group of the job (NCCL)
        gathered = torch.zeros(self.world, dtype=torch.int64, device=tensor.device)
        dist.all_gather_into_tensor(gathered, rows.to(tensor.device), group=self.group)
        counts = [c.reshape(1) for c in gathered.cpu()]
      else:
        dist.all_gather(counts, rows, group=self.group)
    else:
      counts = [rows]
    counts = [int(c) for c in counts]
    if tensor.device.type != "cuda":
      if self.world == 1:
        return tensor.clone()
      pieces = [torch.empty((c,) + tuple(tensor.shape[1:]), dtype=tensor.dtype) for c in counts]
      if len(set(counts)) == 1:
        dist.all_gather(pieces, tensor.contiguous(), group=self.group)
      else:
        self._gather_uneven(pieces, tensor)
      return torch.cat(pieces, dim=0)
    row_bytes = tensor.element_size()
    for extent in tensor.shape[1:]:
      row_bytes *= extent
    # every block starts on a 16-byte boundary of the staging area and of the (padded) output
    block_vectors = [(c * row_bytes + 15) // 16 for c in counts]
    if max(block_vectors) * 16 > self.capacity:
      raise tools.UserException("Collective block of %d bytes exceeds the communicator capacity (%d bytes)" % (max(block_vectors) * 16, self.capacity))
    offs = (ctypes.c_longlong * (_MAX_RANKS + 1))()
    for r, v in enumerate(block_vectors):
      offs[r + 1] = offs[r] + v
    mine = tensor.contiguous().view(-1).view(torch.uint8)
    self.heap.local("data", torch.uint8)[:mine.numel()].copy_(mine)
    padded = torch.empty(offs[self.world] * 16, dtype=torch.uint8, device=self.device)
    self._ptrs[2 * _MAX_RANKS + 2] = padded.data_ptr()
    self.epoch += 1
    status = self._lib.agb_comm_allgather(self._ptrs, offs, ctypes.c_int(self.world), ctypes.c_int(self.rank), ctypes.c_uint(self.epoch), ctypes.c_int(self.max_blocks), self._stream())
    if status != 0:
      raise RuntimeError("agb_comm_allgather failed with status %d" % status)
    if all(c * row_bytes % 16 == 0 for c in counts):
      return padded.view(tensor.dtype).view((sum(counts),) + tuple(tensor.shape[1:]))
    pieces = [padded[offs[r] * 16:offs[r] * 16 + counts[r] * row_bytes] for r in range(self.world)]
    return torch.cat(pieces).view(tensor.dtype).view((sum(counts),) + tuple(tensor.shape[1:]))

  def _gather_uneven(self, pieces, tensor):
    """gloo all-gather with different first dimensions: one broadcast per rank."""
    for r, piece in enumerate(pieces):
      if r == self.rank:
        piece.copy_(tensor)
      dist.broadcast(piece, src=dist.get_global_rank(self.group, r) if self.group is not None else r, group=self.group)


# ---------------------------------------------------------------------------- #
# Module-level op set of the reference (one default communicator per process)

_default = None


def init(capacity=256 << 20, device=None, group=None):
  """`MPIInit`: build the default communicator (collective call: every rank of the group)."""
  global _default
  _default = Communicator(capacity, device, group)
  return _default


def _comm():
  if _default is None:
    raise tools.UserException("collectives.init() must be called first")
  return _default


def size():
  return _comm().size()


def rank():
  return _comm().rank


def local_rank():
  return _comm().local_rank()


def allreduce(tensor, mean=False):
  return _comm().allreduce(tensor, mean)


def allgather(tensor, counts=None):
  return _comm().allgather(tensor, counts)

"""Symmetric (peer-mapped) device memory for the fused aggregation path.

Every rank allocates one heap of identical size; after `rendezvous()` each rank knows the
address at which every peer's heap is mapped in *its own* address space, so kernels can
dereference peer pointers directly over NVLink (P2P loads/stores), plus — when the
platform supports NVLS — one multicast address that stores to all of them at once.

Two providers:
  * `torch`: `torch.distributed._symmetric_memory` (CUDA VMM + fabric/fd handles; gives the
    multicast mapping);
  * `ipc`: `cudaMalloc` + `cudaIpcGetMemHandle/OpenMemHandle` through `native/op_comm`
    (no multicast). Used when the torch provider is unavailable or `AGB_SYMM=ipc`.
With a single rank the heap is an ordinary local allocation.

This replaces the transport of the reference (gRPC / MPI tensor rendezvous between PS and
workers): there is no send/receive, only mapped memory and flags.
"""

import ctypes
import os

import torch
import torch.distributed as dist

from .. import tools

_ALIGN = 1024


class _RawCudaBuffer:
  """Minimal `__cuda_array_interface__` provider so torch can view a raw device pointer."""

  def __init__(self, ptr, nbytes, owner=None):
    self._owner = owner
    self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None}


def _view(ptr, nbytes, device, owner=None):
  with torch.cuda.device(device):
    return torch.as_tensor(_RawCudaBuffer(ptr, nbytes, owner), device=device)


class SymmetricHeap:
  """A named-region allocator over one symmetric buffer per rank."""

  def __init__(self, nbytes, device, group=None, provider=None):
    self.device = torch.device(device)
    self.group = group
    self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    self.rank = dist.get_rank(group) if self.world > 1 else 0
    self.nbytes = (int(nbytes) + _ALIGN - 1) // _ALIGN * _ALIGN
    self.provider = None
    self.multicast_ptr = 0
    self.peer_ptrs = []
    self._regions = {}
    self._cursor = 0
    self._keepalive = []
    provider = provider or os.environ.get("AGB_SYMM", "auto")
    if self.world == 1:
      self._init_local()
    else:
      errors = []
      for candidate in (("torch", "ipc") if provider == "auto" else (provider,)):
        try:
          getattr(self, "_init_" + candidate)()
          break
        except Exception as err:  # try the next provider
          errors.append(candidate + ": " + repr(err))
          self.provider = None
      if self.provider is None:
        raise tools.UserException("Unable to set up symmetric memory across %d ranks (%s)" % (self.world, "; ".join(errors)))

  # -- providers ------------------------------------------------------------ #
  def _init_local(self):
    self.buffer = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
    self.peer_ptrs = [self.buffer.data_ptr()]
    self.provider = "local"

  def _init_torch(self):
    import torch.distributed._symmetric_memory as symm_mem
    group = self.group if self.group is not None else dist.group.WORLD
    with torch.cuda.device(self.device):
      self.buffer = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.device)
      self.buffer.zero_()
      handle = symm_mem.rendezvous(self.buffer, group.group_name)
    self._keepalive.append(handle)
    self.peer_ptrs = [int(p) for p in handle.buffer_ptrs]
    if os.environ.get("AGB_NO_MULTICAST", "") == "":
      try:
        self.multicast_ptr = int(handle.multicast_ptr or 0)
      except Exception:
        self.multicast_ptr = 0
    torch.cuda.synchronize(self.device)
    dist.barrier(group=self.group)
    self.provider = "torch"

  def _init_ipc(self):
    from .. import native
    lib = native.library("op_comm")
    with torch.cuda.device(self.device):
      ptr = ctypes.c_ulonglong(0)
      if lib.agb_comm_alloc(ctypes.c_ulonglong(self.nbytes), ctypes.byref(ptr)) != 0:
        raise RuntimeError("cudaMalloc of the symmetric heap failed")
      handle = (ctypes.c_ubyte * 64)()
      if lib.agb_comm_ipc_handle(ptr, handle) != 0:
        raise RuntimeError("cudaIpcGetMemHandle failed")
      gathered = [None] * self.world
      dist.all_gather_object(gathered, (self.rank, bytes(handle), self.device.index), group=self.group)
      self.peer_ptrs = [0] * self.world
      for peer_rank, peer_handle, peer_device in gathered:
        if peer_rank == self.rank:
          self.peer_ptrs[peer_rank] = ptr.value
          continue
        mapped = ctypes.c_ulonglong(0)
        raw = (ctypes.c_ubyte * 64).from_buffer_copy(peer_handle)
        if lib.agb_comm_ipc_open(raw, ctypes.byref(mapped)) != 0:
          raise RuntimeError("cudaIpcOpenMemHandle failed for rank " + str(peer_rank))
        self.peer_ptrs[peer_rank] = mapped.value
      self.buffer = _view(ptr.value, self.nbytes, self.device)
    torch.cuda.synchronize(self.device)
    dist.barrier(group=self.group)
    self.provider = "ipc"

  # -- regions ---------------------------------------------------------------- #
  def region(self, name, nbytes):
    """Reserve `nbytes` (same call sequence on every rank => same offsets everywhere)."""
    if name in self._regions:
      raise AssertionError("Region " + repr(name) + " already exists")
    size = (int(nbytes) + _ALIGN - 1) // _ALIGN * _ALIGN
    if self._cursor + size > self.nbytes:
      raise tools.UserException("Symmetric heap exhausted while reserving " + repr(name))
    self._regions[name] = (self._cursor, int(nbytes))
    self._cursor += size
    return self._regions[name][0]

  def local(self, name, dtype, shape=None):
    """Local tensor view of a region."""
    offset, nbytes = self._regions[name]
    flat = self.buffer[offset:offset + nbytes].view(dtype)
    return flat if shape is None else flat.view(shape)

  def peer(self, rank, name):
    """Address (in this process) of `name` inside rank `rank`'s heap."""
    return self.peer_ptrs[rank] + self._regions[name][0]

  def multicast(self, name):
    """Multicast address of `name` (0 when NVLS is unavailable)."""
    return self.multicast_ptr + self._regions[name][0] if self.multicast_ptr else 0

  @staticmethod
  def required(*sizes):
    """Heap size needed for regions of the given byte sizes."""
    return sum((int(s) + _ALIGN - 1) // _ALIGN * _ALIGN for s in sizes)

"""Flat parameter/gradient layout (reference: `graph.py:144-199` flatten / mapflat / inflate).

The reference flattens every worker's gradient list into one 1-D tensor right before the
GAR and inflates the aggregate back into per-variable tensors. Here flatness is the *native*
storage: all parameters of a model are views into one fp32 buffer `[d_pad]`, and each
logical worker's gradient is a row of a `[w, d_pad]` buffer living in peer-mapped memory,
so wgrad kernels write straight into what the aggregation kernel reads — no flatten copy.

Every variable starts at a multiple of 8 elements (16-byte alignment of its bf16 compute
copy, 32 bytes in fp32, as TMA and 128-bit accesses need) and `d_pad` is a multiple of 1024.
Padding coordinates are zero in every worker's gradient, hence neutral for every rule.
"""

import math

import torch

from .. import tools

_VAR_ALIGN = 8
_TOTAL_ALIGN = 1024


class FlatLayout:
  """Ordered map `name -> (offset, shape)` over a flat buffer."""

  def __init__(self):
    self._entries = {}  # name -> (offset, shape, numel)
    self._order = []
    self._size = 0
    self._frozen = False

  def add(self, name, shape):
    if self._frozen:
      raise AssertionError("Layout is frozen")
    if name in self._entries:
      raise tools.UserException("Variable " + repr(name) + " declared twice")
    shape = tuple(int(s) for s in shape)
    numel = int(math.prod(shape)) if shape else 1
    offset = self._size
    self._entries[name] = (offset, shape, numel)
    self._order.append(name)
    self._size = (offset + numel + _VAR_ALIGN - 1) // _VAR_ALIGN * _VAR_ALIGN
    return offset

  def freeze(self):
    self._frozen = True
    return self

  @property
  def names(self):
    return list(self._order)

  @property
  def size(self):
    """Number of real coordinates d (sum of the variables' sizes)."""
    return sum(entry[2] for entry in self._entries.values())

  @property
  def padded_size(self):
    return (self._size + _TOTAL_ALIGN - 1) // _TOTAL_ALIGN * _TOTAL_ALIGN

  def offset(self, name):
    return self._entries[name][0]

  def shape(self, name):
    return self._entries[name][1]

  def view(self, flat, name):
    """View of variable `name` inside `flat` (a 1-D tensor of `padded_size` elements)."""
    offset, shape, numel = self._entries[name]
    return flat[offset:offset + numel].view(shape)

  def views(self, flat):
    return {name: self.view(flat, name) for name in self._order}

  def mask(self, device="cpu"):
    """Boolean [padded_size] tensor, True on real coordinates."""
    mask = torch.zeros(self.padded_size, dtype=torch.bool, device=device)
    for offset, _, numel in self._entries.values():
      mask[offset:offset + numel] = True
    return mask

  def slice_bounds(self, rank, world):
    """Coordinate range [lo, hi) of the flat buffer owned by `rank` (multiples of 4)."""
    quads = self.padded_size // 4
    return (quads * rank // world) * 4, (quads * (rank + 1) // world) * 4

  def describe(self):
    return {"variables": len(self._order), "d": self.size, "d_padded": self.padded_size}


def regularization(flat_params, l1, l2):
  """Value and gradient of the reference's regularisers (`graph.py:125-139`):
  l1 * sum|w| + l2 * sqrt(sum w^2) (the l2 term is the *un-squared* norm, kept on purpose).
  Returns (loss term as a 0-d tensor, gradient tensor or None)."""
  loss = torch.zeros((), dtype=torch.float32, device=flat_params.device)
  grad = None
  if l1 is not None and l1 > 0.:
    loss = loss + l1 * flat_params.abs().sum()
    grad = l1 * torch.sign(flat_params)
  if l2 is not None and l2 > 0.:
    norm = torch.linalg.vector_norm(flat_params)
    loss = loss + l2 * norm
    term = l2 * flat_params / torch.clamp(norm, min=1e-30)
    grad = term if grad is None else grad + term
  return loss, grad

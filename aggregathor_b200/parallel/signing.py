"""Gradient authentication: every rank signs what its workers publish, every consumer verifies before aggregating.

The reference's hardened transport does this per message: each worker generates an ed25519 key pair at start-up and sends the
public key to the parameter server over TCP (`tf_patches/patches/mpi_rendezvous_mgr.patch:253,304-306`); worker -> PS tensors
and UDP chunks carry a signature (`:514-523,588-591,606-614`) that the PS verifies (`:777-781,808-812,1057-1064`); a chunk
with a bad signature is dropped, i.e. becomes NaN coordinates (`:814-843`), which the NaN-aware GARs then tolerate.

Here gradients are not sent, they are *published* in peer-mapped memory and read in place by the aggregation kernel, so the
unit that is signed is the digest of a published row, split along the same coordinate slices the ranks consume:

* start-up: one ed25519 key pair per rank; public keys exchanged through the process group (the control plane);
* every step, for each local worker: digest of each of the R coordinate slices of its row (a SHA-256 tree computed by a device
  kernel on GPUs — `native/op_gar/digest.cu` — blake2b on CPU, 32 bytes either way), one signature over (step, worker, digests);
  records are all-gathered;
* every rank then checks each record's signature against the owner's public key and recomputes the digest of the slice(s) it
  is about to consume — through the peer mapping, i.e. over exactly the bytes its aggregation kernel will read — and fills a
  slice whose signature or digest does not match with NaN (the reference's "drop"): the forged part of that worker's gradient
  is lost, the rule's NaN handling does the rest (a NaN partial distance excludes the worker from Krum / Bulyan selections
  on every rank, the coordinate-wise rules ignore the NaN coordinates).

This costs a host round trip per step and is therefore opt-in (`runner.py --authenticate`), like the reference's transport.

Limit, stated plainly (weaker than per-message signatures over a socket):
* verify-then-use: the consumer verifies the peer-mapped row in place and its aggregation kernel re-reads the same memory
  afterwards. The owner of the row could rewrite it in between (it is its own memory). Ranks are separate processes of one
  job on one trusted box — the threat model the feature covers is a *corrupted* or *forged-in-flight* row (fault injection
  `--attack forge`), not a malicious co-resident process racing the verifier.
"""

import hashlib
import struct

import torch
import torch.distributed as dist

from .. import tools

try:
  import nacl.exceptions
  import nacl.signing

  class _Key:
    def __init__(self):
      self._key = nacl.signing.SigningKey.generate()
      self.public = bytes(self._key.verify_key.encode())

    def sign(self, message):
      return bytes(self._key.sign(message).signature)

  def _verify(public, message, signature):
    try:
      nacl.signing.VerifyKey(public).verify(message, signature)
      return True
    except (nacl.exceptions.BadSignatureError, ValueError, TypeError):
      return False

except ImportError:  # pragma: no cover - same primitive through `cryptography`
  from cryptography.exceptions import InvalidSignature
  from cryptography.hazmat.primitives import serialization
  from cryptography.hazmat.primitives.asymmetric import ed25519

  class _Key:
    def __init__(self):
      self._key = ed25519.Ed25519PrivateKey.generate()
      self.public = self._key.public_key().public_bytes(serialization.Encoding.Raw, serialization.PublicFormat.Raw)

    def sign(self, message):
      return self._key.sign(message)

  def _verify(public, message, signature):
    try:
      ed25519.Ed25519PublicKey.from_public_bytes(public).verify(signature, message)
      return True
    except (InvalidSignature, ValueError, TypeError):
      return False


def _message(step, worker, digests):
  return struct.pack("<qq", int(step), int(worker)) + b"".join(bytes(d) for d in digests)


class Authenticator:
  """Signs the local workers' rows and verifies everybody's before the aggregation step."""

  def __init__(self, layout, nbworkers, group=None):
    self.layout = layout
    self.n = nbworkers
    self.group = group
    self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    self.rank = dist.get_rank(group) if self.world > 1 else 0
    self.w = nbworkers // self.world
    self.bounds = [layout.slice_bounds(r, self.world) for r in range(self.world)]
    self.key = _Key()
    if self.world > 1:
      publics = [None] * self.world
      dist.all_gather_object(publics, self.key.public, group=self.group)
      self.publics = [bytes(p) for p in publics]
    else:
      self.publics = [self.key.public]
    self.rejected_total = 0
    tools.info("Gradient authentication on: ed25519, %d rank key(s), %d slice digest(s) per row" % (self.world, self.world), context="auth")

  # -- digests ---------------------------------------------------------------------- #
  @staticmethod
  def digest(piece):
    """32-byte cryptographic digest (uint8 tensor on the slice's device) of a flat fp32 slice: SHA-256 tree kernel on CUDA, blake2b
    on the host."""
    if piece.is_cuda:
      from ..ops import gar as gar_ops
      return gar_ops.sha256(piece)
    raw = hashlib.blake2b(piece.contiguous().numpy().tobytes(), digest_size=32).digest()
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8)

  def _digests(self, row, slices):
    values = torch.stack([self.digest(row[self.bounds[s][0]:self.bounds[s][1]]) for s in slices]).cpu()   # one device -> host copy
    return [bytes(v.numpy().tobytes()) for v in values]

  # -- protocol ----------------------------------------------------------------------- #
  def publish(self, step, local_rows, after_sign=None):
    """`local_rows`: [(row slot = rank * w + local row, flat fp32 row)] of this rank. Returns every rank's records {slot: (digests, signature)}.
    `after_sign()` runs between signing and the exchange (fault injection: tampering with an already signed row)."""
    mine = {}
    for worker, row in local_rows:
      digests = self._digests(row, range(self.world))
      mine[worker] = (digests, self.key.sign(_message(step, worker, digests)))
    if after_sign is not None:
      after_sign()
    if self.world == 1:
      return mine
    gathered = [None] * self.world
    dist.all_gather_object(gathered, mine, group=self.group)
    records = {}
    for owner, part in enumerate(gathered):
      for worker, record in part.items():
        if worker // self.w == owner:  # a rank may only speak for the workers it hosts
          records[worker] = record
    return records

  def verify(self, step, visible_rows, records, slices):
    """`visible_rows`: {worker: full-length flat row as this rank will read it}; `slices`: coordinate slices this rank consumes.
    NaN-fills every consumed slice that fails; returns the list of (worker, slice) rejected."""
    rejected = []
    for worker in range(self.n):
      row = visible_rows.get(worker)
      if row is None:
        continue
      record = records.get(worker)
      valid = record is not None and len(record[0]) == self.world and _verify(self.publics[worker // self.w], _message(step, worker, record[0]), record[1])
      seen = self._digests(row, slices) if valid else None
      for index, s in enumerate(slices):
        if not valid or seen[index] != record[0][s]:
          row[self.bounds[s][0]:self.bounds[s][1]].fill_(float("nan"))
          rejected.append((worker, s))
    if rejected:
      self.rejected_total += len(rejected)
      tools.warning("Step %d: dropped %d gradient slice(s) failing authentication: %r" % (step, len(rejected), rejected[:8]), context="auth")
    return rejected

"""sm_100a aggregation kernels vs the host C++ / torch fp32 oracles (single GPU)."""

import pytest
import torch

from aggregathor_b200 import aggregators
from aggregathor_b200.aggregators import FusedSpec, _ops
from aggregathor_b200.engine.flat import FlatLayout
from aggregathor_b200.engine.optimizers import optimizers
from aggregathor_b200.engine.schedules import build

pytestmark = pytest.mark.gpu


def _data(n, d, seed=0, outliers=2, nan_rows=0):
  gen = torch.Generator().manual_seed(seed)
  G = torch.randn(n, d, generator=gen)
  for k in range(outliers):
    G[n - 1 - k] = G[n - 1 - k] * 30 + 5
  for k in range(nan_rows):
    G[k, 7::13] = float("nan")
  return G


def _close(a, b, tol=2e-5):
  a, b = a.float().cpu(), b.float().cpu()
  same_nan = torch.isnan(a) == torch.isnan(b)
  assert bool(same_nan.all())
  a, b = torch.nan_to_num(a), torch.nan_to_num(b)
  assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max())), float((a - b).abs().max())


@pytest.mark.parametrize("n,d", [(8, 4096), (5, 1003), (16, 20000), (11, 333), (19, 7001), (32, 4099)])
@pytest.mark.parametrize("rule", ["average", "average-nan", "median", "averaged-median"])
def test_coordinate_rules(rule, n, d):
  from aggregathor_b200.ops import gar as gar_ops
  G = _data(n, d, seed=n * 7 + d, nan_rows=1 if rule in ("average-nan", "median") else 0)
  beta = n - 2
  spec = FusedSpec(rule, n, beta=beta)
  out = gar_ops.aggregate(spec, G.cuda())
  ref = {"average": _ops.host_average, "average-nan": _ops.host_average_nan, "median": _ops.host_median,
         "averaged-median": lambda M: _ops.host_averaged_median(M, beta)}[rule](G)
  _close(out, ref)


@pytest.mark.parametrize("n,f,d", [(8, 2, 100000), (5, 1, 1000), (16, 3, 30001), (11, 2, 4097), (16, 6, 555), (19, 4, 50003), (24, 5, 9001), (32, 7, 3001), (32, 2, 800)])
def test_krum(n, f, d):
  from aggregathor_b200.ops import gar as gar_ops
  G = _data(n, d, seed=d, outliers=f)
  m = n - f - 2
  out, dist, info = gar_ops.aggregate(FusedSpec("krum", n, f=f, m=m), G.cuda(), return_details=True)
  ref, selected = _ops.host_krum(G, f, m, return_selected=True)
  mask = sum(1 << int(i) for i in selected)
  assert int(info[1].item()) & 0xffffffff == mask, (bin(int(info[1].item()) & 0xffffffff), bin(mask))
  _close(dist, _ops.host_pairwise_distances(G), tol=1e-4)
  _close(out, ref)


def test_krum_nan_row_never_selected():
  from aggregathor_b200.ops import gar as gar_ops
  G = _data(8, 5000, seed=3, outliers=0)
  G[2, 100] = float("nan")
  G[5, :] = float("inf")
  out, dist, info = gar_ops.aggregate(FusedSpec("krum", 8, f=2, m=4), G.cuda(), return_details=True)
  mask = int(info[1].item())
  assert not mask & (1 << 2) and not mask & (1 << 5)
  assert bool(torch.isfinite(out).all())
  _close(out, _ops.host_krum(G, 2, 4))


@pytest.mark.parametrize("n,f,d", [(7, 1, 2000), (8, 1, 50000), (11, 2, 9999), (16, 3, 12345), (16, 2, 777), (19, 4, 6007), (24, 3, 4000), (32, 7, 1500)])
def test_bulyan(n, f, d):
  from aggregathor_b200.ops import gar as gar_ops
  G = _data(n, d, seed=n + d, outliers=f)
  m = n - f - 2
  out, dist, info = gar_ops.aggregate(FusedSpec("bulyan", n, f=f, m=m, beta=n - 4 * f - 2), G.cuda(), return_details=True)
  ref, weights = _ops.host_bulyan(G, f, m, return_weights=True)
  theta = n - 2 * f - 2
  assert int(info[0].item()) == theta
  for k in range(theta):
    mask = sum(1 << i for i in range(n) if float(weights[k, i]) != 0.0)
    assert int(info[1 + k].item()) & 0xffffffff == mask, (k, bin(int(info[1 + k].item()) & 0xffffffff), bin(mask))
  _close(out, ref)


@pytest.mark.parametrize("rule,f", [("average", 0), ("median", 0), ("averaged-median", 2), ("krum", 2), ("bulyan", 1)])
def test_double_precision_inputs_stay_double(rule, f):
  """The reference's ops take float and double: double gradients on the device are aggregated in double (no silent rounding to fp32)."""
  from aggregathor_b200.ops import gar as gar_ops
  n = 9
  G = _data(n, 3001, seed=4, outliers=f).double()
  G += torch.randn(G.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64) * 1e-9   # structure below fp32 resolution
  m = n - f - 2
  spec = FusedSpec(rule, n, f=f, m=m, beta=(n - 4 * f - 2) if rule == "bulyan" else n - 2)
  out = gar_ops.aggregate(spec, G.cuda())
  ref = {"average": _ops.host_average, "median": _ops.host_median, "averaged-median": lambda M: _ops.host_averaged_median(M, n - 2),
         "krum": lambda M: _ops.host_krum(M, f, m), "bulyan": lambda M: _ops.host_bulyan(M, f, m)}[rule](G)
  assert out.dtype == torch.float64
  assert float((out.cpu() - ref).abs().max()) <= 1e-12 * max(1.0, float(ref.abs().max()))


def test_plugin_dispatch_on_cuda():
  G = _data(8, 3000, seed=11)
  for name, f in (("average", 0), ("median", 0), ("krum-co", 2), ("krum-tf", 2), ("krum-py", 2), ("averaged-median", 2), ("average-nan", 0)):
    gar = aggregators.instantiate(name, 8, f, [])
    _close(gar.aggregate(list(G.cuda())), gar.aggregate(list(G)))


@pytest.mark.parametrize("opt,opt_args", [("sgd", []), ("adam", []), ("rmsprop", []), ("adagrad", []), ("adadelta", [])])
def test_fused_optimizers_single_rank(opt, opt_args):
  """Fused kernel (R = 1) with every optimizer vs HostAggregation running the same math in torch."""
  from aggregathor_b200.parallel.aggregation import FusedAggregation, HostAggregation
  layout = FlatLayout()
  layout.add("w", (1000, 37))
  layout.add("b", (37,))
  layout.freeze()
  spec = build(optimizers, "optimizer", opt, opt_args)
  gar = aggregators.instantiate("krum", 8, 2, [])
  fused = FusedAggregation(gar, layout, 8, spec, device="cuda", keep_aggregate=True)
  host = HostAggregation(gar, layout, 8, build(optimizers, "optimizer", opt, opt_args), device="cpu")
  gen = torch.Generator().manual_seed(5)
  init = torch.randn(layout.padded_size, generator=gen)
  fused.params.copy_(init)
  host.params.copy_(init)
  for step in range(3):
    G = torch.randn(8, layout.padded_size, generator=gen) * 0.1
    G[7] += 3.0
    fused.grads.copy_(G)
    host.grads.copy_(G)
    fused.step(0.05)
    host.step(0.05)
    torch.cuda.synchronize()
    _close(fused.last_aggregate, host.last_aggregate)
    _close(fused.params, host.params, tol=1e-4)


@pytest.mark.parametrize("rule,n,f", [("krum", 8, 2), ("bulyan", 11, 2), ("krum", 20, 4)])
def test_bucketed_phase_a_matches_the_single_launch(rule, n, f):
  """Pre-accumulating the distance pass bucket by bucket (`phase_a`, what runs under the backward pass) then the finish kernel selects
  the same workers and produces the same update as one launch over a contiguous slice; step state lives in device memory."""
  from aggregathor_b200.parallel.aggregation import FusedAggregation
  layout = FlatLayout()
  layout.add("a", (3000, 11))
  layout.add("b", (513,))
  layout.add("c", (77, 64))
  layout.freeze()
  d = layout.padded_size
  spec = build(optimizers, "optimizer", "sgd", [])
  gar = aggregators.instantiate(rule, n, f, [])
  cut1, cut2 = (2 * d // 3) // 8 * 8, (d // 4) // 8 * 8
  plain = FusedAggregation(gar, layout, n, spec, device="cuda", keep_aggregate=True)
  bucketed = FusedAggregation(gar, layout, n, build(optimizers, "optimizer", "sgd", []), device="cuda", keep_aggregate=True, buckets=[(cut1, d), (cut2, cut1), (0, cut2)], device_state=True)
  assert bucketed.overlappable and len(bucketed.segments) == 3
  gen = torch.Generator().manual_seed(9)
  init = torch.randn(d, generator=gen)
  plain.params.copy_(init)
  bucketed.params.copy_(init)
  side = torch.cuda.Stream()
  for step in range(3):
    G = torch.randn(n, d, generator=gen) * 0.1
    G[n - 1] += 2.0
    plain.grads.copy_(G)
    bucketed.grads.copy_(G)
    losses = torch.arange(1, n + 1, dtype=torch.float32, device="cuda") * (step + 1)
    plain.step(0.1, loss_in=losses)
    bucketed.prepare(0.1)
    side.wait_stream(torch.cuda.current_stream())
    bucketed.phase_a(0, stream=side)
    bucketed.phase_a(1, stream=side)
    torch.cuda.current_stream().wait_stream(side)
    bucketed.step(loss_in=losses, prepared=True)
    torch.cuda.synchronize()
    assert int(bucketed.epoch_dev.item()) == step + 1
    assert abs(float(bucketed.loss_out) - float(losses.sum())) < 1e-3 and abs(float(plain.loss_out) - float(losses.sum())) < 1e-3
    assert torch.equal(plain.launcher.info[:12], bucketed.launcher.info[:12])   # same selection
    _close(bucketed.launcher.dist_out, plain.launcher.dist_out, tol=1e-5)
    _close(bucketed.last_aggregate, plain.last_aggregate)
    _close(bucketed.params, plain.params, tol=1e-5)


def test_drop_chunks_and_checksum():
  from aggregathor_b200.ops import gar as gar_ops
  g = torch.ones(200000, device="cuda")
  gar_ops.drop_chunks_(g, 0.3, "nan", chunk_bytes=65000, seed=1)
  lost = torch.isnan(g).float().view(-1)
  frac = float(lost.mean())
  assert 0.05 < frac < 0.7
  chunk = 65000 // 4
  first = lost[:chunk * 12].view(12, chunk)
  assert bool(((first.mean(dim=1) == 0) | (first.mean(dim=1) == 1)).all())  # whole chunks are lost
  a = torch.randn(10000, device="cuda")
  assert int(gar_ops.checksum(a)) == int(gar_ops.checksum(a.clone()))
  b = a.clone()
  b[17] += 1e-3
  assert int(gar_ops.checksum(a)) != int(gar_ops.checksum(b))


@pytest.mark.parametrize("numel", [0, 1, 255, 256, 257, 8195, 262144, 1000003])
def test_sha256_tree_digest_matches_hashlib(numel):
  """The device digest signed by gradient authentication (`native/op_gar/digest.cu`: SHA-256 tree over 1024-byte leaves) vs the same
  tree built with hashlib on the host; 16-byte aligned and merely 4-byte aligned buffers; one flipped bit changes the digest."""
  from aggregathor_b200.ops import gar as gar_ops
  gen = torch.Generator(device="cuda").manual_seed(numel + 1)
  base = torch.randn(numel + 1, device="cuda", generator=gen)
  for view in (base[:numel], base[1:numel + 1]):                 # the second view starts 4 bytes into the allocation
    got = bytes(gar_ops.sha256(view).cpu().numpy().tobytes())
    assert got == gar_ops.sha256_tree_host(view.cpu().numpy().tobytes()), (numel, view.data_ptr() % 16)
  if numel:
    before = bytes(gar_ops.sha256(base[:numel]).cpu().numpy().tobytes())
    tampered = base[:numel].clone()
    tampered.view(torch.int32)[numel // 2] ^= 1
    assert bytes(gar_ops.sha256(tampered).cpu().numpy().tobytes()) != before


def test_authenticator_digests_on_cuda():
  """`Authenticator` on one rank with device rows: honest rows verify, a row modified after signing loses exactly its slice."""
  from aggregathor_b200.parallel.signing import Authenticator
  layout = FlatLayout()
  layout.add("w", (1000, 37))
  layout.freeze()
  auth = Authenticator(layout, 3)
  rows = {i: torch.randn(layout.padded_size, device="cuda") for i in range(3)}
  records = auth.publish(5, list(rows.items()))
  assert auth.verify(5, rows, records, [0]) == []
  rows[1][123] += 1.0
  assert auth.verify(5, rows, records, [0]) == [(1, 0)]
  assert bool(torch.isnan(rows[1][:layout.size]).all()) and not bool(torch.isnan(rows[0]).any())

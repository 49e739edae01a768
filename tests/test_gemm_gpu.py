"""tcgen05 GEMM (native/op_nn/gemm.cu) vs fp32 torch matmul on the same bf16-rounded operands."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
  gen = torch.Generator(device="cuda").manual_seed(seed)
  return torch.randn(shape, device="cuda", generator=gen).to(torch.bfloat16)


def _check(out, ref, tol=2e-3):
  err = float((out.float() - ref).abs().max())
  scale = max(1.0, float(ref.abs().max()))
  assert err <= tol * scale, (err, scale)


@pytest.fixture(params=[1, 0], ids=["persistent", "tile-per-cta"], autouse=True)
def gemm_variant(request):
  from aggregathor_b200.ops import nn_native as nat
  nat.set_gemm_persistent(request.param)
  yield
  nat.set_gemm_persistent(1)


SHAPES = [(128, 128, 64), (256, 64, 128), (300, 200, 136), (32, 10, 104), (1000, 192, 4096), (25088, 256, 64), (77, 1000, 2048), (4096, 384, 4096), (130, 70, 8)]


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_mm_nt(m, n, k):
  from aggregathor_b200.ops import nn_native as nat
  x, w = _rand((m, k), 1), _rand((n, k), 2)
  ref = x.float() @ w.float().t()
  _check(nat.mm_nt(x, w, out_dtype=torch.float32), ref)
  _check(nat.mm_nt(x, w), ref, tol=1e-2)
  bias = torch.randn(n, device="cuda")
  _check(nat.mm_nt(x, w, bias=bias, relu=True, out_dtype=torch.float32), torch.relu(ref + bias))
  for bn in (64, 128, 256):
    _check(nat.mm_nt(x, w, out_dtype=torch.float32, bn=bn), ref)


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_mm_nn(m, n, k):
  from aggregathor_b200.ops import nn_native as nat
  x, w = _rand((m, k), 3), _rand((k, n), 4)
  ref = x.float() @ w.float()
  _check(nat.mm_nn(x, w, out_dtype=torch.float32), ref)
  _check(nat.mm_nn(x, w, out_dtype=torch.float32, bn=64), ref)


@pytest.mark.parametrize("m,n,k", SHAPES + [(64, 64, 100352), (256, 2304, 6272)])
def test_mm_tn(m, n, k):
  from aggregathor_b200.ops import nn_native as nat
  x, y = _rand((k, m), 5), _rand((k, n), 6)
  ref = x.float().t() @ y.float()
  _check(nat.mm_tn(x, y, splits=1), ref, tol=3e-3)
  _check(nat.mm_tn(x, y), ref, tol=3e-3)          # automatic split-K with fp32 atomics
  out = torch.empty((m, n), device="cuda")
  _check(nat.mm_tn(x, y, out=out, splits=3), ref, tol=3e-3)


def test_strided_operands():
  from aggregathor_b200.ops import nn_native as nat
  big = _rand((200, 520), 7)
  x = big[:, :500]            # row stride 520 (multiple of 8), inner 500
  w = _rand((96, 500), 8)
  _check(nat.mm_nt(x, w, out_dtype=torch.float32), x.float() @ w.float().t())


def test_linear_and_pointwise_conv_layers():
  """Module-level check: native vs torch provider through models.core (forward, dgrad, wgrad)."""
  from aggregathor_b200.ops import nn as nn_ops
  x = _rand((64, 256, 14, 14), 9).contiguous(memory_format=torch.channels_last)
  w = (_rand((512, 1, 1, 256), 10).float() * 0.05).to(torch.bfloat16)
  dy = _rand((64, 512, 14, 14), 11).contiguous(memory_format=torch.channels_last)
  results = {}
  for backend in ("torch", "native"):
    y = nn_ops.conv2d_forward(backend, x, w, None, 1, (0, 0, 0, 0), False)
    gw = torch.zeros((512, 1, 1, 256), device="cuda")
    dx, _, _ = nn_ops.conv2d_backward(backend, dy, x, w, None, 1, (0, 0, 0, 0), False, False, True, gw, None)
    results[backend] = (y.float(), dx.float(), gw)
  for a, b in zip(results["torch"], results["native"]):
    _check(b, a, tol=2e-2)


@pytest.mark.parametrize("m,n,k", [(256, 256, 64), (512, 256, 128), (300, 200, 136), (1000, 192, 4096), (25088, 256, 64), (77, 1000, 2048), (4096, 384, 4096), (130, 70, 8), (2048, 1024, 1024)])
def test_pair_gemm(m, n, k):
  """CTA-pair kernel (tcgen05.mma.cta_group::2, 256 x 256 tiles) vs fp32 matmul, ragged edges included."""
  from aggregathor_b200.ops import nn_native as nat
  x, w = _rand((m, k), 21), _rand((n, k), 22)
  ref = x.float() @ w.float().t()
  _check(nat.mm_nt(x, w, out_dtype=torch.float32, bn=512), ref)
  _check(nat.mm_nt(x, w, bn=512), ref, tol=1e-2)
  bias = torch.randn(n, device="cuda")
  _check(nat.mm_nt(x, w, bias=bias, relu=True, out_dtype=torch.float32, bn=512), torch.relu(ref + bias))


# ---------------------------------------------------------------------------- #
# TF32 path (kind::tf32): fp32 operands straight from memory, fp32 accumulation and output — the parity precision of the fp32 reference

def _rand32(shape, seed):
  gen = torch.Generator(device="cuda").manual_seed(seed)
  return torch.randn(shape, device="cuda", generator=gen)


def _tf32_tol(k):
  return 2e-3 * max(1.0, (k / 64.0) ** 0.5)   # operands truncated to 10 mantissa bits: ~1e-3 relative per product, random-walk over k


TF32_SHAPES = [(128, 128, 32), (256, 64, 128), (300, 200, 136), (32, 12, 104), (1000, 192, 4096), (25088, 256, 64), (77, 1000, 2048), (130, 72, 8)]


@pytest.mark.parametrize("m,n,k", TF32_SHAPES)
def test_tf32_mm_nt(m, n, k):
  from aggregathor_b200.ops import nn_native as nat
  x, w = _rand32((m, k), 11), _rand32((n, k), 12)
  ref = x.double() @ w.double().t()
  out = nat.mm_nt(x, w)
  assert out.dtype == torch.float32
  _check(out, ref.float(), tol=_tf32_tol(k))
  bias = torch.randn(n, device="cuda")
  _check(nat.mm_nt(x, w, bias=bias, relu=True), torch.relu(ref.float() + bias), tol=_tf32_tol(k))
  _check(nat.mm_nt(x, w, bn=64), ref.float(), tol=_tf32_tol(k))


@pytest.mark.parametrize("m,n,k", TF32_SHAPES)
def test_tf32_mm_nn_and_tn(m, n, k):
  from aggregathor_b200.ops import nn_native as nat
  x, w = _rand32((m, k), 13), _rand32((k, n), 14)
  _check(nat.mm_nn(x, w), (x.double() @ w.double()).float(), tol=_tf32_tol(k))
  a, b = _rand32((k, m), 15), _rand32((k, n), 16)
  ref = (a.double().t() @ b.double()).float()
  _check(nat.mm_tn(a, b, splits=1), ref, tol=_tf32_tol(k))
  _check(nat.mm_tn(a, b), ref, tol=_tf32_tol(k))


def test_tf32_grouped_weight_gradient():
  from aggregathor_b200.ops import nn_native as nat
  groups, k, m, n = 4, 392, 96, 160
  a, b = _rand32((groups * k, m), 17), _rand32((groups * k, n), 18)
  rows = torch.zeros((groups, m * n + 64), device="cuda")
  out = rows[0, :m * n].view(m, n)
  nat.mm_tn(a, b, out=out, groups=groups, group_stride=rows.stride(0))
  for g in range(groups):
    ref = (a[g * k:(g + 1) * k].double().t() @ b[g * k:(g + 1) * k].double()).float()
    _check(rows[g, :m * n].view(m, n), ref, tol=_tf32_tol(k))

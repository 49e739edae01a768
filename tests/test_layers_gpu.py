"""Native sm_100a layer kernels vs the torch provider, op by op and end to end (ResNet-50 / cnnet / mnist gradients)."""

import pytest
import torch

pytestmark = pytest.mark.gpu
CL = torch.channels_last


def _rand(shape, seed, scale=1.0):
  gen = torch.Generator(device="cuda").manual_seed(seed)
  t = (torch.randn(shape, device="cuda", generator=gen) * scale).to(torch.bfloat16)
  return t.contiguous(memory_format=CL) if len(shape) == 4 else t


def _close(a, b, tol):
  a, b = a.float(), b.float()
  err = float((a - b).abs().max())
  scale = max(1e-3, float(b.abs().max()))
  assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize("shape", [(32, 64, 56, 56), (8, 256, 14, 14), (4, 2048, 7, 7), (16, 192, 5, 5), (32, 64, 112, 112), (2, 24, 3, 3)])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("fused", [True, False])
def test_batchnorm(shape, relu, fused):
  """`fused`: the single-launch resident-tile kernels (the 51 MB shape exceeds the on-chip stash and takes the re-read path) vs
  the statistics + apply kernel pair."""
  from aggregathor_b200.ops import nn as ops, nn_native
  nn_native.set_bn_fused(fused)
  x = _rand(shape, 1) * 2 + 0.5
  dy = _rand(shape, 2)
  c = shape[1]
  gamma = torch.rand(c, device="cuda") + 0.5
  beta = torch.randn(c, device="cuda") * 0.1
  out = {}
  mask = None
  for backend in ("native", "torch"):
    mm, mv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    y, mean, rstd = ops.batchnorm_forward(backend, x, gamma, beta, mm, mv, 0.9, 1e-5, relu)
    gg, gb = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    if mask is None:
      mask = y  # both backward passes use the same ReLU mask: an output rounding to 0 in one provider and to 1e-3 in the other is not an error
    dx = ops.batchnorm_backward(backend, dy, x, mask if relu else None, gamma, mean, rstd, relu, gg, gb)
    out[backend] = (y, mean, rstd, dx, gg, gb, mm, mv)
  nn_native.set_bn_fused(True)
  tols = (2e-2, 1e-3, 1e-3, 3e-2, 2e-2, 2e-2, 1e-3, 1e-3)
  for a, b, tol in zip(out["native"], out["torch"], tols):
    _close(a, b, tol)


@pytest.mark.parametrize("fused", [True, False])
def test_batchnorm_groups_and_repeats(fused):
  """Per-worker statistics (groups), many back-to-back launches of different widths (the single-launch kernels alternate between
  two workspace halves that must always be found zeroed), each checked against an fp32 run of the library provider."""
  from aggregathor_b200.ops import nn as ops, nn_native
  nn_native.set_bn_fused(fused)
  for rep in range(3):
    for c, hw, groups in ((64, 28, 4), (512, 7, 8), (256, 14, 2), (1024, 4, 1), (2048, 7, 1)):
      x = _rand((8 * groups, c, hw, hw), 50 + c) + 0.25
      dy = _rand((8 * groups, c, hw, hw), 60 + c)
      gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
      outs = {}
      for backend in ("native", "torch"):
        xin, dyin = (x, dy) if backend == "native" else (x.float(), dy.float())
        mm, mv = torch.full((c,), 0.2, device="cuda"), torch.ones(c, device="cuda")
        y, mean, rstd = ops.batchnorm_forward(backend, xin, gamma, beta, mm, mv, 0.9, 1e-5, True, groups)
        mask = y if backend == "native" else mask
        grads = torch.zeros((groups, 2, c), device="cuda")
        dx = ops.batchnorm_backward(backend, dyin, xin, mask.to(xin.dtype), gamma, mean, rstd, True, grads[0, 0], grads[0, 1], groups, grads.stride(0))
        outs[backend] = (y, mean, rstd, dx, grads, mm, mv)
      for a, b, tol in zip(outs["native"], outs["torch"], (1e-2, 1e-4, 1e-4, 2e-2, 1e-3, 1e-4, 1e-4)):
        _close(a, b, tol)
  nn_native.set_bn_fused(True)


@pytest.mark.parametrize("cin,cout,k,stride,hw,pads", [(64, 64, 3, 1, 56, (1, 1, 1, 1)), (128, 128, 3, 2, 28, (1, 1, 1, 1)), (3, 64, 7, 2, 224, (3, 3, 3, 3)),
                                                        (3, 64, 5, 1, 32, (2, 2, 2, 2)), (64, 64, 5, 1, 16, (2, 2, 2, 2)), (256, 512, 1, 1, 14, (0, 0, 0, 0)),
                                                        (64, 64, 3, 2, 57, (0, 1, 0, 1)), (64, 128, 3, 1, 28, (1, 1, 1, 1)), (256, 256, 3, 1, 14, (1, 1, 1, 1)),
                                                        (512, 512, 3, 1, 7, (1, 1, 1, 1)), (128, 64, 5, 1, 12, (2, 2, 2, 2)),
                                                        # stride-2 implicit GEMM (TMA element strides / parity-split data gradient): ResNet's three down-sampling 3x3
                                                        (64, 64, 3, 2, 56, (1, 1, 1, 1)), (256, 256, 3, 2, 14, (1, 1, 1, 1)), (64, 128, 3, 2, 16, (0, 1, 0, 1)), (128, 64, 5, 2, 24, (2, 2, 2, 2))])
@pytest.mark.parametrize("implicit", [True, False])
def test_conv(cin, cout, k, stride, hw, pads, implicit, monkeypatch):
  """Native convolution (forward, wgrad, bias grad, dgrad) vs fp32 autograd on the same bf16-rounded operands."""
  import torch.nn.functional as F
  from aggregathor_b200.ops import nn as ops
  from aggregathor_b200.ops import nn_native
  monkeypatch.setattr(nn_native, "_DISABLED", set() if implicit else {"implicit"})
  torch.backends.cudnn.allow_tf32 = False
  n = 4 if hw > 7 else 16
  x = _rand((n, cin, hw, hw), 3)
  w = _rand((cout, k, k, cin), 4, scale=(2.0 / (k * k * cin)) ** 0.5).contiguous()
  bias = torch.randn(cout, device="cuda") * 0.1
  xp = F.pad(x.float(), (pads[2], pads[3], pads[0], pads[1])).requires_grad_(True)
  wf = w.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
  bf = bias.clone().requires_grad_(True)
  pre = F.conv2d(xp, wf, bf, stride)
  y = ops.conv2d_forward("native", x, w, bias, stride, pads, True)
  _close(y, torch.relu(pre).detach(), 1e-2)
  dy = _rand(tuple(pre.shape), 5)
  # the ReLU mask is taken from the native (bf16) output: where |pre| is below the bf16 accumulation noise the fp32 and bf16
  # masks legitimately differ, and a single flipped element moves a weight-gradient entry by |dy * x|
  pre.backward(dy.float() * (y > 0).float())
  gw, gb = torch.zeros((cout, k, k, cin), device="cuda"), torch.zeros(cout, device="cuda")
  dx, _, _ = ops.conv2d_backward("native", dy, x, w, y, stride, pads, True, True, cin % 8 == 0, gw, gb)
  _close(gw, wf.grad.permute(0, 2, 3, 1), 1e-2)
  _close(gb, bf.grad, 1e-2)
  if dx is not None:
    _close(dx, xp.grad[:, :, pads[0]:pads[0] + hw, pads[2]:pads[2] + hw], 1e-2)


@pytest.mark.parametrize("c,h,w,k,stride,pads,dtype", [(3, 224, 224, 7, 2, (3, 3, 3, 3), torch.bfloat16), (3, 35, 37, 3, 2, (0, 0, 0, 0), torch.bfloat16), (1, 28, 28, 5, 1, (2, 2, 2, 2), torch.bfloat16),
                                                        (3, 33, 31, 7, 2, (2, 3, 2, 3), torch.float32), (3, 299, 299, 3, 2, (0, 0, 0, 0), torch.bfloat16)])
def test_stem_im2col(c, h, w, k, stride, pads, dtype):
  """The shared-memory staged im2col of few-channel inputs (column order kh, kw, c; zero padding, zero tail columns) vs `F.unfold`."""
  import torch.nn.functional as F
  from aggregathor_b200.ops import nn_native
  n = 3
  gen = torch.Generator(device="cuda").manual_seed(c * 100 + h)
  x = torch.randn((n, c, h, w), device="cuda", generator=gen).to(dtype).contiguous(memory_format=torch.channels_last)
  oh, ow = (h + pads[0] + pads[1] - k) // stride + 1, (w + pads[2] + pads[3] - k) // stride + 1
  col = nn_native._im2col(x, k, k, stride, pads, oh, ow)
  assert col.shape == (n * oh * ow, k * k * c)
  padded = F.pad(x.float(), (pads[2], pads[3], pads[0], pads[1]))
  want = F.unfold(padded, k, stride=stride).reshape(n, c, k, k, oh * ow).permute(0, 4, 2, 3, 1).reshape(n * oh * ow, k * k * c)   # (c, kh, kw) -> (kh, kw, c)
  assert torch.equal(col.float(), want)
  full = col.as_strided((col.shape[0], col.stride(0)), (col.stride(0), 1))
  assert float(full[:, k * k * c:].abs().max()) == 0.0 if col.stride(0) > k * k * c else True


@pytest.mark.parametrize("kh,kw,stride,padding", [(1, 7, 1, "SAME"), (7, 1, 1, "SAME"), (3, 1, 1, "SAME"), (1, 3, 2, "SAME"), (3, 1, 1, "VALID")])
def test_rectangular_conv(kh, kw, stride, padding):
  """The factorised 1x7 / 7x1 / 1x3 / 3x1 convolutions of the Inception families on the native im2col + tcgen05 GEMM path."""
  from aggregathor_b200.models.core import Context, Conv2d
  from aggregathor_b200.ops import nn as ops
  outs = {}
  for backend in ("native", "torch"):
    dtype = torch.bfloat16 if backend == "native" else torch.float32
    layer = Conv2d("c", 64, 96, (kh, kw), stride=stride, padding=padding, bias=True, relu=True)
    ctx = Context(backend, True, dtype, "cuda")
    w = (_rand((96, kh, kw, 64), 31, scale=0.08)).contiguous()
    b = torch.randn(96, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2)) * 0.1
    ctx.weights = {"c/weights": w.to(dtype)}
    ctx.master = {"c/weights": w.float(), "c/biases": b}
    gw, gb = torch.zeros((96, kh, kw, 64), device="cuda"), torch.zeros(96, device="cuda")
    ctx.grads = {"c/weights": gw, "c/biases": gb}
    x = _rand((4, 64, 17, 17), 32).to(dtype)
    before = dict(ops.fallbacks)
    y = layer.forward(x, ctx)
    dy = _rand(tuple(y.shape), 33).to(dtype)
    dx = layer.backward(dy, ctx)
    if backend == "native":
      assert ops.fallbacks == before, "served by the aten provider"
    outs[backend] = (y, dx, gw, gb)
  for got, want in zip(outs["native"], outs["torch"]):
    _close(got, want, 2e-2)


@pytest.mark.parametrize("cin,cout,k,stride,hw,pads", [(64, 64, 3, 1, 28, (1, 1, 1, 1)), (128, 96, 3, 2, 28, (1, 1, 1, 1)), (32, 64, 5, 1, 14, (2, 2, 2, 2)), (256, 128, 1, 1, 14, (0, 0, 0, 0)),
                                                        (3, 64, 7, 2, 64, (3, 3, 3, 3)), (96, 32, 3, 1, 7, (1, 1, 1, 1))])
def test_conv_tf32(cin, cout, k, stride, hw, pads):
  """TF32 convolution path (fp32 activations and weights, kind::tf32 products) vs fp64 autograd: forward, wgrad, bias grad, dgrad."""
  import torch.nn.functional as F
  from aggregathor_b200.ops import nn as ops
  n = 4
  gen = torch.Generator(device="cuda").manual_seed(7)
  x = torch.randn((n, cin, hw, hw), device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
  w = (torch.randn((cout, k, k, cin), device="cuda", generator=gen) * (2.0 / (k * k * cin)) ** 0.5).contiguous()
  bias = torch.randn(cout, device="cuda", generator=gen) * 0.1
  xp = F.pad(x.double(), (pads[2], pads[3], pads[0], pads[1])).requires_grad_(True)
  wf = w.double().permute(0, 3, 1, 2).clone().requires_grad_(True)
  bf = bias.double().clone().requires_grad_(True)
  pre = F.conv2d(xp, wf, bf, stride)
  before = dict(ops.fallbacks)
  y = ops.conv2d_forward("native", x, w, bias, stride, pads, True)
  assert y.dtype == torch.float32
  _close(y, torch.relu(pre).detach().float(), 4e-3)
  dy = torch.randn(tuple(pre.shape), device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
  pre.backward(dy.double() * (y > 0).double())
  gw, gb = torch.zeros((cout, k, k, cin), device="cuda"), torch.zeros(cout, device="cuda")
  dx, _, _ = ops.conv2d_backward("native", dy, x, w, y, stride, pads, True, True, cin % 8 == 0, gw, gb)
  assert ops.fallbacks == before, "served by the aten provider"
  _close(gw, wf.grad.permute(0, 2, 3, 1).float(), 4e-3)
  _close(gb, bf.grad.float(), 4e-3)
  if dx is not None:
    _close(dx, xp.grad[:, :, pads[0]:pads[0] + hw, pads[2]:pads[2] + hw].float(), 4e-3)


def test_resnet_step_tf32_matches_fp32_reference():
  """One ResNet training step (slim resnet_v1_18, 96 x 96, batch 8) on the TF32 native path vs the aten provider in strict fp32. A deep
  BN network amplifies any operand rounding, so the yardstick is what cuDNN / cuBLAS do with THEIR TF32 products on the same step:
  the native path must stay as close to the fp32 gradient (cosine of the whole flat gradient, loss) as the library TF32 path does,
  up to a small factor (tcgen05 truncates fp32 operands to TF32, the libraries round them)."""
  from aggregathor_b200.engine.flat import FlatLayout
  from aggregathor_b200.models import Context, nets_factory
  from aggregathor_b200.ops import nn as ops
  results = {}
  for label, backend, allow in (("fp32", "torch", False), ("lib-tf32", "torch", True), ("native-tf32", "native", False)):
    torch.backends.cudnn.allow_tf32 = allow
    torch.backends.cuda.matmul.allow_tf32 = allow
    model = nets_factory.get_network("resnet_v1_18", 16)
    layout, states = FlatLayout(), {}
    model.declare(layout, states)
    layout.freeze()
    host = torch.zeros(layout.padded_size)
    host_states = {k: torch.zeros(v) for k, v in states.items()}
    model.initialize(layout.views(host), host_states, torch.Generator().manual_seed(1))
    params = host.cuda()
    ctx = Context(backend, True, torch.float32, "cuda")
    ctx.master = ctx.weights = layout.views(params)
    ctx.state = {k: v.cuda() for k, v in host_states.items()}
    grads = torch.zeros_like(params)
    ctx.grads = layout.views(grads)
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((8, 3, 96, 96), device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
    labels = torch.randint(0, 16, (8,), device="cuda", generator=gen)
    before = dict(ops.fallbacks)
    loss = float(model.loss_and_backward(x, labels, ctx))
    if backend == "native":
      assert ops.fallbacks == before, {k: v - before.get(k, 0) for k, v in ops.fallbacks.items() if v != before.get(k, 0)}
    results[label] = (loss, grads)
  torch.backends.cudnn.allow_tf32 = False
  torch.backends.cuda.matmul.allow_tf32 = False
  ref_loss, ref = results["fp32"]
  cos = lambda g: float(torch.nn.functional.cosine_similarity(g.double(), ref.double(), dim=0))
  lib_gap, native_gap = 1.0 - cos(results["lib-tf32"][1]), 1.0 - cos(results["native-tf32"][1])
  lib_loss, native_loss = abs(results["lib-tf32"][0] - ref_loss), abs(results["native-tf32"][0] - ref_loss)
  print("1 - cos(gradient, fp32 gradient): library TF32 %.3e, native TF32 %.3e; |loss - fp32 loss|: %.3e, %.3e" % (lib_gap, native_gap, lib_loss, native_loss))
  assert native_gap <= 8.0 * lib_gap + 1e-3, (native_gap, lib_gap)
  assert native_loss <= 8.0 * lib_loss + 5e-3 * max(1.0, abs(ref_loss)), (native_loss, lib_loss)


def test_launch_overlap_keeps_gradients():
  """Programmatic dependent launch + weight gradients on a side stream (what single-worker ranks run) must not change the result:
  one slim resnet_v1_50 step (batch 32, 224 x 224: the single-launch batch norm is on its envelope), weight gradients without split-K,
  serial vs overlapped launches, eager and replayed from a CUDA graph, compared up to the run-to-run spread of the serial path; every
  kernel started early has to wait for its producer."""
  from aggregathor_b200.engine.flat import FlatLayout
  from aggregathor_b200.models import Context, nets_factory
  from aggregathor_b200.ops import nn_native
  model = nets_factory.get_network("resnet_v1_50", 1000)
  layout, states = FlatLayout(), {}
  model.declare(layout, states)
  layout.freeze()
  host = torch.zeros(layout.padded_size)
  host_states = {k: torch.zeros(v) for k, v in states.items()}
  model.initialize(layout.views(host), host_states, torch.Generator().manual_seed(1))
  params = host.cuda()
  gen = torch.Generator(device="cuda").manual_seed(3)
  x = torch.randn((32, 3, 224, 224), device="cuda", generator=gen).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
  labels = torch.randint(0, 1000, (32,), device="cuda", generator=gen)

  def step(graph):
    ctx = Context("native", True, torch.bfloat16, "cuda")
    ctx.master = layout.views(params)
    ctx.weights = layout.views(params.to(torch.bfloat16))
    ctx.state = {k: v.cuda() for k, v in host_states.items()}
    grads = torch.zeros_like(params)
    ctx.grads = layout.views(grads)
    if not graph:
      loss = model.loss_and_backward(x, labels, ctx)
      torch.cuda.synchronize()
      return float(loss), grads
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
      model.loss_and_backward(x, labels, ctx)           # warm-up (workspaces, tensor maps)
      stream.synchronize()
      captured = torch.cuda.CUDAGraph()
      with torch.cuda.graph(captured, stream=stream):
        grads.zero_()
        loss = model.loss_and_backward(x, labels, ctx)
      for _ in range(3):
        captured.replay()
      stream.synchronize()
    return float(loss), grads.clone()

  nn_native.set_deterministic(True)
  try:
    for graph in (False, True):
      # each mode against its own serial run: the warm-up and the replays of the graph mode move the moving mean the forward pass
      # pivots its variance sums on, a rounding-level change that a randomly initialised 50-layer bf16 network amplifies
      nn_native.set_launch_overlap(False)
      loss_serial, grads_serial = step(graph)
      _, again = step(graph)
      assert loss_serial == loss_serial and float(grads_serial.abs().max()) > 0
      norm = float(grads_serial.norm())
      noise = float((again - grads_serial).norm()) / norm     # run-to-run spread of the serial path
      nn_native.set_launch_overlap(True)
      loss, grads = step(graph)
      nn_native.set_launch_overlap(False)
      error = float((grads - grads_serial).norm()) / norm
      assert abs(loss - loss_serial) <= 1e-3 * abs(loss_serial), (graph, loss, loss_serial)
      assert error <= max(5 * noise, 1e-3), (graph, error, noise)
  finally:
    nn_native.set_launch_overlap(False)
    nn_native.set_deterministic(False)


def test_deterministic_weight_gradients():
  """`set_deterministic(True)`: no split-K => the weight gradient of a GEMM-shaped and of an implicit-GEMM layer is bit-identical from run to run."""
  from aggregathor_b200.ops import nn as ops
  from aggregathor_b200.ops import nn_native
  nn_native.set_deterministic(True)
  try:
    x, dy = _rand((8, 128, 28, 28), 41), _rand((8, 128, 28, 28), 42)
    for k in (1, 3):
      w = _rand((128, k, k, 128), 43, scale=0.05).contiguous()
      pads = ((k - 1) // 2,) * 4
      runs = []
      for _ in range(3):
        gw = torch.zeros((128, k, k, 128), device="cuda")
        ops.conv2d_backward("native", dy, x, w, None, 1, pads, False, False, True, gw, None)
        runs.append(gw.clone())
      assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
  finally:
    nn_native.set_deterministic(False)


def test_pools_and_eltwise():
  from aggregathor_b200.ops import nn as ops
  x = _rand((8, 64, 112, 112), 6)
  for pads, k, s in (((0, 1, 0, 1), 3, 2), ((0, 0, 0, 0), 2, 2), ((1, 1, 1, 1), 3, 2)):
    res = {}
    for backend in ("torch", "native"):
      y, index = ops.maxpool_forward(backend, x, k, s, pads)
      dy = _rand(tuple(y.shape), 7)
      res[backend] = (y, ops.maxpool_backward(backend, dy, x.shape, index, k, s, pads, x, y))
    _close(res["native"][0], res["torch"][0], 1e-6)
    _close(res["native"][1], res["torch"][1], 1e-2)
  for shape, pads in (((3, 16, 15, 17), (1, 1, 1, 1)), ((3, 16, 15, 17), (0, 0, 0, 0)), ((2, 24, 35, 35), (0, 0, 0, 0)), ((2, 8, 6, 9), (0, 1, 1, 0)), ((2, 8, 5, 4), (1, 0, 0, 1))):
    xs = _rand(shape, 12)        # odd maps, every parity of the padding: the 2 x 2-block backward of the 3x3/2 pools
    res = {}
    for backend in ("torch", "native"):
      y, index = ops.maxpool_forward(backend, xs, 3, 2, pads)
      dy = _rand(tuple(y.shape), 13)
      res[backend] = (y, ops.maxpool_backward(backend, dy, xs.shape, index, 3, 2, pads, xs, y))
    _close(res["native"][0], res["torch"][0], 1e-6)
    _close(res["native"][1], res["torch"][1], 1e-2)
  z = _rand((8, 2048, 7, 7), 8)
  _close(ops.global_avgpool_forward("native", z), ops.global_avgpool_forward("torch", z), 1e-2)
  d = _rand((8, 2048, 1, 1), 9)
  _close(ops.global_avgpool_backward("native", d, z.shape), ops.global_avgpool_backward("torch", d, z.shape), 1e-2)
  a, b = _rand((8, 256, 14, 14), 10), _rand((8, 256, 14, 14), 11)
  _close(ops.add_relu_forward("native", a, b, True), ops.add_relu_forward("torch", a, b, True), 1e-2)
  _close(ops.relu_backward("native", a, b), ops.relu_backward("torch", a, b), 1e-6)


@pytest.mark.parametrize("rows,c", [(4096, 1024), (333, 384), (32, 4096), (50000, 64)])
def test_layernorm(rows, c):
  from aggregathor_b200.ops import nn as ops
  x = (_rand((rows, c), 31) * 1.5 + 0.3)
  dy = _rand((rows, c), 32)
  gamma = torch.rand(c, device="cuda") + 0.5
  beta = torch.randn(c, device="cuda") * 0.1
  out = {}
  for backend in ("torch", "native"):
    y, mean, rstd = ops.layernorm_forward(backend, x, gamma, beta, 1e-5)
    gg, gb = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    dx = ops.layernorm_backward(backend, dy, x, gamma, mean, rstd, gg, gb)
    out[backend] = (y, mean, rstd, dx, gg, gb)
  for a, b, tol in zip(out["native"], out["torch"], (2e-2, 1e-3, 1e-3, 3e-2, 2e-2, 2e-2)):
    _close(a, b, tol)


def test_softmax_xent_and_image_normalize():
  from aggregathor_b200.ops import nn as ops
  logits = _rand((32, 1000), 12) * 3
  labels = torch.randint(0, 1000, (32,), device="cuda")
  for smoothing in (0.0, 0.1):
    l0, d0 = ops.softmax_xent("torch", logits, labels, smoothing)
    l1, d1 = ops.softmax_xent("native", logits, labels, smoothing)
    assert abs(float(l0) - float(l1)) < 2e-3 * max(1.0, abs(float(l0)))
    _close(d1, d0, 2e-2)
  images = torch.randint(0, 256, (4, 32, 32, 3), device="cuda", dtype=torch.uint8)
  for mode in ("vgg", "inception"):
    _close(ops.image_normalize("native", images, mode, torch.bfloat16), ops.image_normalize("torch", images, mode, torch.bfloat16), 1e-2)


@pytest.mark.parametrize("name,classes,batch,image", [("resnet_v1_50", 1000, 16, 128), ("cnnet", 10, 16, 32), ("mlp", 10, 32, None)])
def test_model_gradients_native_vs_fp32(name, classes, batch, image):
  """Whole-model check against an fp32 (TF32 off) run of the library provider: the bf16 native path must agree with
  it at least as well as the bf16 library provider does."""
  from aggregathor_b200.engine.flat import FlatLayout
  from aggregathor_b200.models import Context, get_network
  torch.backends.cudnn.allow_tf32 = False
  torch.backends.cuda.matmul.allow_tf32 = False
  model = get_network(name, classes)
  layout, shapes = FlatLayout(), {}
  model.declare(layout, shapes)
  layout.freeze()
  gen = torch.Generator().manual_seed(0)
  init = torch.zeros(layout.padded_size)
  init_states = {k: torch.zeros(v) for k, v in shapes.items()}
  model.initialize(layout.views(init), init_states, gen)
  params = init.cuda()
  weights = params.to(torch.bfloat16)
  x = _rand((batch, 784), 20).abs() if image is None else _rand((batch, model.input_shape[0], image, image), 20)
  labels = torch.randint(0, classes, (batch,), device="cuda")
  grads, losses = {}, {}
  for tag, backend, dtype in (("fp32", "torch", torch.float32), ("torch", "torch", torch.bfloat16), ("native", "native", torch.bfloat16)):
    ctx = Context(backend, True, dtype, "cuda")
    ctx.master = layout.views(params)
    ctx.weights = ctx.master if dtype == torch.float32 else layout.views(weights)
    ctx.state = {k: v.clone().cuda() for k, v in init_states.items()}
    g = torch.zeros(layout.padded_size, device="cuda")
    ctx.grads = layout.views(g)
    losses[tag] = float(model.loss_and_backward(x.to(dtype), labels, ctx))
    grads[tag] = g
  cos = lambda a, b: float(torch.nn.functional.cosine_similarity(grads[a], grads[b], dim=0))
  report = {"loss": losses, "cos_native_fp32": cos("native", "fp32"), "cos_torch_fp32": cos("torch", "fp32"), "cos_native_torch": cos("native", "torch")}
  print(report)
  assert abs(losses["native"] - losses["fp32"]) < 3e-2 * max(1.0, abs(losses["fp32"])), report
  assert report["cos_native_fp32"] > min(0.97, report["cos_torch_fp32"] - 0.05), report
  ratio = float(grads["native"].norm() / grads["fp32"].norm())
  assert 0.85 < ratio < 1.15, (ratio, report)


@pytest.mark.parametrize("bn", ["pair", "single-launch"])
@pytest.mark.parametrize("name,classes,batch,image", [("resnet_v1_50", 1000, 8, 64), ("cnnet", 10, 8, 32), ("mlp", 10, 16, None)])
def test_batched_workers_native(name, classes, batch, image, bn, request):
  """Native kernels with `ctx.groups` = 4 workers in one pass (grouped wgrad GEMM / conv, per-group BN, per-group loss) vs four
  sequential native passes: same per-worker losses and gradient rows (up to atomics' summation order). A randomly initialised
  ResNet at batch 8 amplifies one-ulp differences of the batch statistics into very different gradients, so both passes must
  take the same batch-norm kernels whatever the tensor size: the kernel pair, or the single-launch kernels for every size."""
  from aggregathor_b200.engine.flat import FlatLayout
  from aggregathor_b200.models import Context, get_network
  from aggregathor_b200.ops import nn_native
  nn_native.set_bn_fused(bn != "pair")
  nn_native.set_bn_fused_limits(1 << 20, 1 << 20, 1)
  request.addfinalizer(lambda: (nn_native.set_bn_fused(True), nn_native.set_bn_fused_limits()))
  workers = 4
  model = get_network(name, classes)
  layout, shapes = FlatLayout(), {}
  model.declare(layout, shapes)
  layout.freeze()
  init = torch.zeros(layout.padded_size)
  init_states = {k: torch.zeros(v) for k, v in shapes.items()}
  model.initialize(layout.views(init), init_states, torch.Generator().manual_seed(0))
  params = init.cuda()
  weights = params.to(torch.bfloat16)
  xs = [(_rand((batch, 784), 40 + i).abs() if image is None else _rand((batch, model.input_shape[0], image, image), 40 + i)) for i in range(workers)]
  ys = [torch.randint(0, classes, (batch,), device="cuda") for _ in range(workers)]

  def context(rows):
    ctx = Context("native", True, torch.bfloat16, "cuda")
    ctx.master, ctx.weights = layout.views(params), layout.views(weights)
    ctx.state = {k: v.clone().cuda() for k, v in init_states.items()}
    ctx.grads = layout.views(rows)
    return ctx

  seq = torch.zeros((workers, layout.padded_size), device="cuda")
  seq_losses = [float(model.loss_and_backward(x, y, context(seq[i]))) for i, (x, y) in enumerate(zip(xs, ys))]
  bat = torch.zeros((workers, layout.padded_size), device="cuda")
  ctx = context(bat[0])
  ctx.groups, ctx.group_stride = workers, bat.stride(0)
  x_all = torch.cat(xs, dim=0)
  if image is not None:
    x_all = x_all.contiguous(memory_format=torch.channels_last)
  losses = model.loss_and_backward(x_all, torch.cat(ys), ctx)
  assert losses.shape == (workers,)
  for a, b in zip(losses.tolist(), seq_losses):
    assert abs(a - b) <= 2e-2 * max(1.0, abs(b)), (losses.tolist(), seq_losses)
  for i in range(workers):
    cos = float(torch.nn.functional.cosine_similarity(bat[i], seq[i], dim=0))
    assert cos > 0.995, (i, cos)
    other = float(torch.nn.functional.cosine_similarity(bat[i], seq[(i + 1) % workers], dim=0))
    assert other < 0.9, (i, other)   # rows are really per-worker, not a shared/summed gradient


def _strip_randomness(module):
  from aggregathor_b200.models.core import DropPath, Dropout
  if isinstance(module, (Dropout, DropPath)):
    module.keep_prob = 1.0
  for child in module.children():
    _strip_randomness(child)


def _model_run(model, layout, params, weights, init_states, x, labels, backend, dtype):
  from aggregathor_b200.models import Context
  ctx = Context(backend, True, dtype, "cuda")
  ctx.master = layout.views(params)
  ctx.weights = ctx.master if dtype == torch.float32 else layout.views(weights)
  ctx.state = {k: v.clone().cuda() for k, v in init_states.items()}
  g = torch.zeros(layout.padded_size, device="cuda")
  ctx.grads = layout.views(g)
  return float(model.loss_and_backward(x.to(dtype), labels, ctx)), g


@pytest.mark.parametrize("name,batch", [("inception_v3", 4), ("mobilenet_v2", 8), ("nasnet_mobile", 4), ("inception_resnet_v2", 2)])
def test_searched_and_inception_families_native(name, batch):
  """The Inception / MobileNet-v2 / NASNet graphs (branches, DAG cells, auxiliary heads, rectangular and depthwise kernels) run on
  the native provider at their default resolution and agree with an fp32 library run as well as the bf16 library provider does."""
  from aggregathor_b200.engine.flat import FlatLayout
  from aggregathor_b200.models import get_network
  torch.backends.cudnn.allow_tf32 = False
  torch.backends.cuda.matmul.allow_tf32 = False
  model = get_network(name, 101)
  _strip_randomness(model.root)
  layout, shapes = FlatLayout(), {}
  model.declare(layout, shapes)
  layout.freeze()
  init = torch.zeros(layout.padded_size)
  init_states = {k: torch.zeros(v) for k, v in shapes.items()}
  model.initialize(layout.views(init), init_states, torch.Generator().manual_seed(0))
  params = init.cuda()
  weights = params.to(torch.bfloat16)
  x = _rand((batch,) + tuple(model.input_shape), 30)
  labels = torch.randint(0, 101, (batch,), device="cuda")
  losses, grads = {}, {}
  for tag, backend, dtype in (("fp32", "torch", torch.float32), ("torch", "torch", torch.bfloat16), ("native", "native", torch.bfloat16)):
    losses[tag], grads[tag] = _model_run(model, layout, params, weights, init_states, x, labels, backend, dtype)
  cos = lambda a, b: float(torch.nn.functional.cosine_similarity(grads[a], grads[b], dim=0))
  report = {"loss": losses, "cos_native_fp32": cos("native", "fp32"), "cos_torch_fp32": cos("torch", "fp32")}
  print(name, report)
  assert all(v == v for v in losses.values()), report
  assert abs(losses["native"] - losses["fp32"]) < 5e-2 * max(1.0, abs(losses["fp32"])), report
  if report["cos_torch_fp32"] >= 0.5:  # below that the bf16 library run itself has decorrelated from fp32 (tiny batch, deep BN stack): nothing to compare
    assert report["cos_native_fp32"] > min(0.95, report["cos_torch_fp32"] - 0.1), report


def test_wgrad_side_stream_matches(monkeypatch):
  """Weight gradients computed on the side stream (fork / join around each layer's data gradient) equal the single-stream ones."""
  from aggregathor_b200.engine.flat import FlatLayout
  from aggregathor_b200.models import get_network
  from aggregathor_b200.ops import nn_native
  model = get_network("resnet_v1_50", 100)
  layout, shapes = FlatLayout(), {}
  model.declare(layout, shapes)
  layout.freeze()
  init = torch.zeros(layout.padded_size)
  init_states = {k: torch.zeros(v) for k, v in shapes.items()}
  model.initialize(layout.views(init), init_states, torch.Generator().manual_seed(0))
  params = init.cuda()
  weights = params.to(torch.bfloat16)
  x = _rand((8, 3, 64, 64), 31)
  labels = torch.randint(0, 100, (8,), device="cuda")
  monkeypatch.setattr(nn_native, "_WGRAD_STREAM", False)
  loss_a, grad_a = _model_run(model, layout, params, weights, init_states, x, labels, "native", torch.bfloat16)
  monkeypatch.setattr(nn_native, "_WGRAD_STREAM", True)
  loss_b, grad_b = _model_run(model, layout, params, weights, init_states, x, labels, "native", torch.bfloat16)
  torch.cuda.synchronize()
  assert abs(loss_a - loss_b) < 1e-6
  assert float(torch.nn.functional.cosine_similarity(grad_a, grad_b, dim=0)) > 0.9999


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("c,hw,groups", [(256, 16, 4), (64, 56, 1), (512, 8, 2), (2048, 2, 1)])
def test_batchnorm_add_relu(fused, c, hw, groups):
  """Closing layer of a residual unit, y = relu(bn(x) + shortcut), and its backward (dx, masked dy) — in the single-launch kernel
  and in the statistics + apply pair — against the composition of the library ops in fp32."""
  from aggregathor_b200.ops import nn as ops, nn_native
  nn_native.set_bn_fused(fused)
  nn_native.set_bn_fused_limits(1 << 20, 1 << 20, 1)
  try:
    x, res, dy = _rand((8 * groups, c, hw, hw), 70) + 0.25, _rand((8 * groups, c, hw, hw), 71), _rand((8 * groups, c, hw, hw), 72)
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    outs = {}
    for backend in ("native", "torch"):
      cast = (lambda t: t) if backend == "native" else (lambda t: t.float())
      mm, mv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
      y, mean, rstd = ops.batchnorm_add_relu_forward(backend, cast(x), gamma, beta, mm, mv, 0.9, 1e-5, cast(res), groups)
      mask = y if backend == "native" else mask
      grads = torch.zeros((groups, 2, c), device="cuda")
      dx, g = ops.batchnorm_add_relu_backward(backend, cast(dy), cast(x), cast(mask), gamma, mean, rstd, grads[0, 0], grads[0, 1], groups, grads.stride(0))
      outs[backend] = (y, mean, rstd, dx, g, grads)
    for a, b, tol in zip(outs["native"], outs["torch"], (1e-2, 1e-4, 1e-4, 2e-2, 1e-6, 1e-3)):
      _close(a, b, tol)
  finally:
    nn_native.set_bn_fused(True)
    nn_native.set_bn_fused_limits()


@pytest.mark.parametrize("shape,stride", [((8, 256, 56, 56), 2), ((4, 64, 7, 7), 2), ((2, 128, 9, 10), 3)])
def test_subsample(shape, stride):
  from aggregathor_b200.ops import nn as ops
  x = _rand(shape, 80)
  y = ops.subsample_forward("native", x, stride)
  ref = ops.subsample_forward("torch", x, stride)
  assert y.shape == ref.shape and torch.equal(y.contiguous(), ref.contiguous())
  dy = _rand(tuple(ref.shape), 81)
  dx = ops.subsample_backward("native", dy, shape, stride)
  assert dx.shape == x.shape and torch.equal(dx.contiguous(), ops.subsample_backward("torch", dy, shape, stride).contiguous())


_PREVIEW = __import__("os").environ.get("AGB_NATIVE_DEPTHWISE", __import__("os").environ.get("AGB_NATIVE_PREVIEW", "1")) not in ("", "0")
_preview = pytest.mark.skipif(not _PREVIEW, reason="depthwise / SAME average pool / ReLU6 kernels switched off (AGB_NATIVE_DEPTHWISE=0)")


@_preview
@pytest.mark.parametrize("c,hw,k,stride,groups", [(64, 28, 3, 1, 1), (128, 14, 3, 2, 2), (88, 21, 5, 2, 1), (176, 11, 7, 1, 4), (32, 9, 7, 2, 1)])
def test_depthwise_native(c, hw, k, stride, groups):
  """Depthwise forward / data gradient / per-worker weight gradient kernels vs the aten grouped convolution in fp32."""
  from aggregathor_b200.models.core import same_padding
  from aggregathor_b200.ops import nn as ops
  x, weight = _rand((4 * groups, c, hw, hw), 90), (_rand((c, k, k, 1), 91) * 0.2).contiguous()
  pads = same_padding(hw, k, stride) + same_padding(hw, k, stride)
  y = ops.depthwise_forward("native", x, weight, stride, pads)
  ref = ops.depthwise_forward("torch", x.float(), weight.float(), stride, pads)
  _close(y, ref, 1e-2)
  dy = _rand(tuple(ref.shape), 92)
  grads = {}
  for backend in ("native", "torch"):
    rows = torch.zeros((groups, c, k, k, 1), device="cuda")
    cast = (lambda t: t) if backend == "native" else (lambda t: t.float())
    dx = ops.depthwise_backward(backend, cast(dy), cast(x), cast(weight), stride, pads, rows[0], groups, rows.stride(0))
    grads[backend] = (dx, rows)
  _close(grads["native"][0], grads["torch"][0], 2e-2)
  _close(grads["native"][1], grads["torch"][1], 2e-2)


@_preview
@pytest.mark.parametrize("shape,k,stride,padding", [((4, 64, 35, 35), 3, 1, "SAME"), ((2, 128, 17, 18), 3, 2, "SAME"), ((2, 768, 17, 17), 5, 3, "VALID"), ((3, 32, 8, 8), 2, 2, "SAME")])
def test_avgpool2d_and_relu6_native(shape, k, stride, padding):
  """SAME / VALID average pooling (divisor = in-image window size) and ReLU6, forward and backward, native vs the aten path of the modules."""
  from aggregathor_b200.models.core import AvgPool, Context, ReLU6
  x, outs = _rand(shape, 95) * 4, {}
  for backend in ("native", "torch"):
    ctx = Context(backend, True, torch.bfloat16 if backend == "native" else torch.float32, "cuda")
    pool, act = AvgPool("p", k, stride, padding), ReLU6("r")
    xin = x if backend == "native" else x.float()
    y = pool.forward(xin, ctx)
    dy = _rand(tuple(y.shape), 96)
    dx = pool.backward(dy if backend == "native" else dy.float(), ctx)
    a = act.forward(xin, ctx)
    da = act.backward(_rand(shape, 97) if backend == "native" else _rand(shape, 97).float(), ctx)
    outs[backend] = (y, dx, a, da)
  for got, want, tol in zip(outs["native"], outs["torch"], (1e-2, 1e-2, 1e-2, 1e-2)):
    _close(got, want, tol)


def _factory_names():
  from aggregathor_b200.models import nets_factory
  return sorted(nets_factory.networks_map)


@pytest.mark.parametrize("name", _factory_names())
def test_every_factory_net_trains_one_step_on_native_kernels(name):
  """All 33 names of the slim factory (reference: `external/slim/nets/nets_factory.py:39-72`): one training step at a reduced
  resolution with the native provider; reports which ops (if any) were served by the aten provider instead."""
  from aggregathor_b200.engine.flat import FlatLayout
  from aggregathor_b200.models import Context, nets_factory
  from aggregathor_b200.ops import nn as ops
  model = nets_factory.get_network(name, 16)
  size = model.input_shape[-1]   # full resolution: the VGG / AlexNet / OverFeat heads are sized for it
  layout, states = FlatLayout(), {}
  model.declare(layout, states)
  layout.freeze()
  params = torch.zeros(layout.padded_size, device="cuda")
  host = torch.zeros(layout.padded_size)
  host_states = {k: torch.zeros(v) for k, v in states.items()}
  model.initialize(layout.views(host), host_states, torch.Generator().manual_seed(1))
  params.copy_(host)
  ctx = Context("native", True, torch.bfloat16, "cuda")
  ctx.master = layout.views(params)
  ctx.weights = layout.views(params.to(torch.bfloat16))
  ctx.state = {k: v.cuda() for k, v in host_states.items()}
  grads = torch.zeros_like(params)
  ctx.grads = layout.views(grads)
  ctx.generator = torch.Generator(device="cuda").manual_seed(2)
  channels = model.input_shape[0]
  x = torch.randn((2, channels, size, size), device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
  labels = torch.randint(0, 16, (2,), device="cuda")
  before = dict(ops.fallbacks)
  loss = float(model.loss_and_backward(x, labels, ctx))
  torch.cuda.synchronize()
  served_by_aten = {k: v - before.get(k, 0) for k, v in ops.fallbacks.items() if v - before.get(k, 0) > 0}
  assert loss == loss and float(grads.abs().sum()) > 0
  print("%s: aten fallbacks %r" % (name, served_by_aten))
  # NASNet / PNASNet widths (44, 54, 42 ... channels) are not multiples of 8: their layers cannot take the 16-byte vector kernels and the
  # TMA row-stride rule and are served by the aten provider; Inception-v2's first separable convolution is depthwise on 3 channels
  if not (name.startswith(("nasnet", "pnasnet")) or name == "inception_v2"):
    assert not served_by_aten, served_by_aten

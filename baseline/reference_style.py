"""The same-box *baseline* arm: what a competent PyTorch user builds to reproduce AggregaThor's data path on B200s
with library kernels only. NOT the product, and none of the product's nn kernels / engine / fused aggregation is on
this path:

  per logical worker  torch.nn ResNet-50 (slim v1 layout) in channels_last, cuDNN / cuBLAS kernels (bf16, or fp32 with
                      TF32 tensor cores), autograd, the whole forward + loss + backward + gradient flattening of a
                      worker captured ONCE into a CUDA graph and replayed every step (cudnn.benchmark picked the algorithms)
  gather              NCCL `all_gather_into_tensor` of the flat fp32 gradients (the reference's worker -> PS transfer,
                      `graph.py:267-273`, with NCCL instead of gRPC)
  aggregate           ONE stand-alone `[n, d] -> [d]` GAR kernel launch (the reference's `native/op_krum` role)
  apply               separate SGD kernel on the fp32 master copy + cast of the compute copy (`graph.py:281`)

`BASELINE.json`: "A path that only calls NCCL all-gather + a standalone GAR kernel is the baseline, not the product."
Logical workers sharing a GPU run as sequential batch-32 passes (one graph each): per-worker gradients out of a single
batched autograd pass would need per-sample-group backward, which torch does not offer without `vmap`-ing the convolutions
into grouped ones (slower than the sequential passes).
"""

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

_VGG_MEANS = (123.68, 116.78, 103.94)


def _conv_same(cin, cout, k, stride):
  """slim `conv2d_same`: SAME for stride 1, explicit (k-1)//2 .. padding + VALID otherwise (symmetric for odd k)."""
  return nn.Conv2d(cin, cout, k, stride=stride, padding=(k - 1) // 2, bias=False)


class _Bottleneck(nn.Module):
  def __init__(self, cin, depth, inner, stride):
    super().__init__()
    self.stride = stride
    self.shortcut = None
    if depth != cin:
      self.shortcut = nn.Sequential(nn.Conv2d(cin, depth, 1, stride=stride, bias=False), nn.BatchNorm2d(depth, eps=1e-5, momentum=0.003))
    self.conv1, self.bn1 = nn.Conv2d(cin, inner, 1, bias=False), nn.BatchNorm2d(inner, eps=1e-5, momentum=0.003)
    self.conv2, self.bn2 = _conv_same(inner, inner, 3, stride), nn.BatchNorm2d(inner, eps=1e-5, momentum=0.003)
    self.conv3, self.bn3 = nn.Conv2d(inner, depth, 1, bias=False), nn.BatchNorm2d(depth, eps=1e-5, momentum=0.003)

  def forward(self, x):
    if self.shortcut is not None:
      a = self.shortcut(x)
    else:
      a = x if self.stride == 1 else x[:, :, ::self.stride, ::self.stride]
    b = F.relu(self.bn1(self.conv1(x)))
    b = F.relu(self.bn2(self.conv2(b)))
    b = self.bn3(self.conv3(b))
    return F.relu(a + b)


class ResNet50V1(nn.Module):
  """slim `resnet_v1_50`: stride on the 3x3 convolution of the LAST unit of blocks 1-3 (`external/slim/nets/resnet_v1.py:258-279`),
  7x7/2 stem, 3x3/2 SAME max-pool, 1x1-conv logits with bias (a Linear on the pooled features): 25 557 032 parameters."""

  def __init__(self, num_classes=1000):
    super().__init__()
    self.conv1, self.bn1 = _conv_same(3, 64, 7, 2), nn.BatchNorm2d(64, eps=1e-5, momentum=0.003)
    units, cin = [], 64
    for base, count, stride in zip((64, 128, 256, 512), (3, 4, 6, 3), (2, 2, 2, 1)):
      for u in range(count):
        units.append(_Bottleneck(cin, base * 4, base, stride if u == count - 1 else 1))
        cin = base * 4
    self.units = nn.Sequential(*units)
    self.logits = nn.Linear(cin, num_classes)

  def forward(self, x):
    x = F.relu(self.bn1(self.conv1(x)))
    x = F.max_pool2d(x, 3, 2, ceil_mode=True)   # TF SAME for even sizes: the extra window hangs over the bottom/right border
    x = self.units(x)
    return self.logits(x.mean(dim=(2, 3)))


class ReferenceStyleTrainer:
  """`train(batches)` = one synchronous step of n logical workers (w = n / world of them on this rank)."""

  def __init__(self, n, gar_spec, batch, image_size, device, *, precision="bf16", lr=0.01, num_classes=1000, group=None, graphs=True, seed=0):
    from aggregathor_b200.ops import gar as gar_ops   # only the stand-alone aggregation / SGD / cast kernels
    self.gar_ops, self.spec = gar_ops, gar_spec
    self.device, self.group = torch.device(device), group
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    self.rank = dist.get_rank(group) if dist.is_initialized() else 0
    self.n, self.w, self.lr = n, n // self.world, lr
    self.precision = precision
    torch.backends.cudnn.benchmark = True
    tf32 = precision == "tf32"
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    self.compute = torch.bfloat16 if precision == "bf16" else torch.float32
    torch.manual_seed(seed)   # identical initial parameters on every rank
    model = ResNet50V1(num_classes).to(self.device)
    # flat storage: [conv + linear weights | everything else], compute copy + fp32 master + per-worker fp32 gradient rows
    heavy = [p for m in model.modules() if isinstance(m, (nn.Conv2d, nn.Linear)) for p in m.parameters(recurse=False)]
    heavy_ids = {id(p) for p in heavy}
    light = [p for p in model.parameters() if id(p) not in heavy_ids]
    self.d_heavy, self.d_light = sum(p.numel() for p in heavy), sum(p.numel() for p in light)
    self.d = self.d_heavy + self.d_light
    self.d_padded = (self.d + 7) // 8 * 8
    self.master = torch.zeros(self.d_padded, dtype=torch.float32, device=self.device)
    self.flat_heavy = torch.zeros(self.d_heavy, dtype=self.compute, device=self.device)
    self.grad_heavy = torch.zeros(self.d_heavy, dtype=self.compute, device=self.device)
    self.flat_light = torch.zeros(self.d_light, dtype=torch.float32, device=self.device)
    self.grad_light = torch.zeros(self.d_light, dtype=torch.float32, device=self.device)

    def rebind(params, flat, grad):
      offset = 0
      for p in params:
        count = p.numel()
        if p.dim() == 4:   # physically channels_last (OHWI): what cuDNN's NHWC kernels consume without a per-call transpose
          o, i, h, w = p.shape
          view = flat[offset:offset + count].view(o, h, w, i).permute(0, 3, 1, 2)
          gview = grad[offset:offset + count].view(o, h, w, i).permute(0, 3, 1, 2)
        else:
          view, gview = flat[offset:offset + count].view(p.shape), grad[offset:offset + count].view(p.shape)
        view.copy_(p.data)
        p.data = view
        p.grad = gview
        offset += count

    rebind(heavy, self.flat_heavy, self.grad_heavy)
    rebind(light, self.flat_light, self.grad_light)
    self.master[:self.d_heavy].copy_(self.flat_heavy)
    self.master[self.d_heavy:self.d].copy_(self.flat_light)
    self.model = model   # 4-D parameters are already channels_last views of the flat buffers: no `.to(memory_format=...)` (it could re-allocate)
    self.model.train()
    self.rows = torch.zeros((self.w, self.d_padded), dtype=torch.float32, device=self.device)
    self.gathered = torch.zeros((self.n, self.d_padded), dtype=torch.float32, device=self.device) if self.world > 1 else self.rows
    self.mean = torch.tensor(_VGG_MEANS, dtype=torch.float32, device=self.device).view(1, 3, 1, 1)
    self.static = [(torch.zeros((batch, image_size, image_size, 3), dtype=torch.uint8, device=self.device),
                    torch.zeros(batch, dtype=torch.int64, device=self.device)) for _ in range(self.w)]
    self.losses = torch.zeros(self.w, dtype=torch.float32, device=self.device)
    self.graphs = [None] * self.w
    self.use_graphs = graphs
    self.launches_per_step = None
    self.step = 0

  # ------------------------------------------------------------------------------------------------ #
  def _worker_pass(self, j):
    images, labels = self.static[j]
    x = (images.permute(0, 3, 1, 2).float() - self.mean).to(self.compute).contiguous(memory_format=torch.channels_last)
    self.grad_heavy.zero_()
    self.grad_light.zero_()
    loss = F.cross_entropy(self.model(x).float(), labels)
    loss.backward()
    self.rows[j, :self.d_heavy].copy_(self.grad_heavy)
    self.rows[j, self.d_heavy:self.d].copy_(self.grad_light)
    self.losses[j] = loss.detach()

  def _capture(self):
    side = torch.cuda.Stream(self.device)
    side.wait_stream(torch.cuda.current_stream(self.device))
    with torch.cuda.stream(side):   # warm-up outside capture: cuDNN algorithm search, autograd buffers
      for _ in range(3):
        self._worker_pass(0)
    torch.cuda.current_stream(self.device).wait_stream(side)
    torch.cuda.synchronize(self.device)
    pool = None
    for j in range(self.w):
      graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(graph, pool=pool):
        self._worker_pass(j)
      pool = graph.pool()
      self.graphs[j] = graph

  def train(self, batches):
    """`batches`: w (uint8 NHWC images, int64 labels) device tensors. Returns the total loss (0-d device tensor)."""
    for (sx, sy), (x, y) in zip(self.static, batches):
      sx.copy_(x, non_blocking=True)
      sy.copy_(y, non_blocking=True)
    if self.use_graphs and self.graphs[0] is None:
      self._capture()
    for j in range(self.w):
      if self.use_graphs:
        self.graphs[j].replay()
      else:
        self._worker_pass(j)
    if self.world > 1:
      dist.all_gather_into_tensor(self.gathered, self.rows, group=self.group)
    aggregated = self.gar_ops.aggregate(self.spec, self.gathered)
    self.gar_ops.sgd_(self.master, aggregated, self.lr)
    self.flat_heavy.copy_(self.master[:self.d_heavy])
    self.flat_light.copy_(self.master[self.d_heavy:self.d])
    total = self.losses.sum()
    if self.world > 1:
      dist.all_reduce(total, group=self.group)
    self.step += 1
    return total

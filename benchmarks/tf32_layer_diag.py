import sys; sys.path.insert(0, "/root/repo")
import torch
from aggregathor_b200.ops import nn as ops
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
gen = torch.Generator(device="cuda").manual_seed(3)
def cl(t): return t.contiguous(memory_format=torch.channels_last)
def rel(a, b): return float((a.float() - b.float()).abs().max()) / max(float(b.abs().max()), 1e-6)
x = cl(torch.randn((8, 64, 32, 32), device="cuda", generator=gen)); dy = cl(torch.randn((8, 64, 32, 32), device="cuda", generator=gen))
gamma = torch.rand(64, device="cuda") + 0.5; beta = torch.randn(64, device="cuda")
res = {}
for backend in ("native", "torch"):
  mm, mv = torch.zeros(64, device="cuda"), torch.ones(64, device="cuda")
  y, mean, rstd = ops.batchnorm_forward(backend, x, gamma, beta, mm, mv, 0.997, 1e-5, True, 1)
  gg, gb = torch.zeros(64, device="cuda"), torch.zeros(64, device="cuda")
  dx = ops.batchnorm_backward(backend, dy, x, y, gamma, mean, rstd, True, gg, gb, 1, 0)
  yp, idx = ops.maxpool_forward(backend, x, 3, 2, (0, 1, 0, 1))
  dyp = cl(torch.randn(tuple(yp.shape), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)))
  dxp = ops.maxpool_backward(backend, dyp, x.shape, idx, 3, 2, (0, 1, 0, 1), x, yp)
  ar = ops.add_relu_forward(backend, x, dy, True)
  rb = ops.relu_backward(backend, dy, ar)
  ga = ops.global_avgpool_forward(backend, x)
  gab = ops.global_avgpool_backward(backend, cl(torch.ones((8, 64, 1, 1), device="cuda")), x.shape)
  sub = ops.subsample_forward(backend, x, 2)
  subb = ops.subsample_backward(backend, sub, x.shape, 2)
  logits = torch.randn((8, 16), device="cuda", generator=torch.Generator(device="cuda").manual_seed(6))
  loss, dl = ops.softmax_xent(backend, logits, torch.arange(8, device="cuda") % 16, 0.0, 1)
  res[backend] = dict(y=y, mean=mean, rstd=rstd, dx=dx, gg=gg, gb=gb, yp=yp, dxp=dxp, ar=ar, rb=rb, ga=ga, gab=gab, sub=sub, subb=subb, loss=loss, dl=dl)
for k in res["native"]:
  print(k, rel(res["native"][k], res["torch"][k]), res["native"][k].dtype)
print("fallbacks", ops.fallbacks)

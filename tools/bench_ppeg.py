"""time the PPEG stencil kernels (mil.py:190-205: identity + 7x7 + 5x5 + 3x3 depthwise convs folded into one 7x7 stencil) at the
bench size (B x (1 + 128 x 128) x 512) and check them against F.conv2d.  usage: python tools/bench_ppeg.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200 import functional as Fn

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H = W = 128; C = 512
torch.manual_seed(0)
x = torch.randn(B, 1 + H * W, C, device="cuda", requires_grad=True)
w7 = (torch.randn(49, C, device="cuda") * 0.1).requires_grad_()
bias = torch.randn(C, device="cuda", requires_grad=True)
y = Fn.ppeg(x, w7, bias, H, W)
g = torch.randn_like(y)
y.backward(g)
# reference
xr = x.detach().clone().requires_grad_(); wr = w7.detach().clone().requires_grad_(); br = bias.detach().clone().requires_grad_()
img = xr[:, 1:].transpose(1, 2).reshape(B, C, H, W)
yr = torch.nn.functional.conv2d(img, wr.t().reshape(C, 1, 7, 7), br, padding=3, groups=C)
yr = torch.cat([xr[:, :1], yr.flatten(2).transpose(1, 2)], 1)
yr.backward(g)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
print("fwd %.2e  dx %.2e  dw %.2e  db %.2e" % (rel(y, yr), rel(x.grad, xr.grad), rel(w7.grad, wr.grad), rel(bias.grad, br.grad)))


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


xd = x.detach()
t_f = timeit(lambda: Fn.ppeg(xd, w7.detach(), bias.detach(), H, W))
def fb():
    xx = xd.requires_grad_()
    yy = Fn.ppeg(xx, w7, bias, H, W)
    yy.backward(g)
t_fb = timeit(fb)
gb = x.numel() * 4 / 1e9
print("B=%d: forward %.0f us (%.2f of HBM peak for read + write), forward + backward (dx + dw) %.0f us" % (B, t_f, 2 * gb / (t_f * 1e-6) / 6543.7, t_fb))

"""Stand-alone launches of the HBM-bound row kernels at the B=8 TransMIL sizes, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200 import ops
from sml_b200._lib import lib, ptr, stream, check, c_i32

dev = "cuda"
B, npad, h, d, m = 8, 16640, 8, 64, 256
qkv = torch.randn(B, npad, 3, h, d, device=dev)
merged = torch.randn(B, npad, h * d, device=dev)
mh = torch.empty(B, npad, h * d, device=dev, dtype=torch.bfloat16)
wc = torch.randn(h, 33, device=dev)
x = torch.randn(B, 1 + 128 * 128, 512, device=dev); y = torch.empty_like(x)
w7 = torch.randn(49, 512, device=dev); b7 = torch.randn(512, device=dev)
s3 = torch.randn(B * h * m, npad, device=dev); s3h = torch.empty(B * h * m, npad, device=dev, dtype=torch.bfloat16)
dA = torch.randn(B * h * m, npad, device=dev); dsh = torch.empty_like(s3h)
xln = torch.randn(B * 16385, 512, device=dev); dyln = torch.randn(B, npad, 512, device=dev)
lw = torch.randn(512, device=dev); lb = torch.randn(512, device=dev)
yln = torch.empty(B, npad, 512, device=dev)
mean, rstd = ops.layernorm_fwd(xln, lw, lb, yln, 16385, npad, 255)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for r in range(2):
    ev[0].record()
    check(lib.sml_nystrom_resconv_fwd(ptr(qkv), ptr(wc), ptr(merged), c_i32(B), c_i32(npad), c_i32(h), c_i32(d), c_i32(33), ptr(mh), None, None, stream()))
    ev[1].record()
    check(lib.sml_ppeg_fwd(ptr(x), ptr(w7), ptr(b7), ptr(y), c_i32(B), c_i32(128), c_i32(128), c_i32(512), stream()))
    ev[2].record()
    ops.softmax_(s3, s3h, None, write_f32=False)
    ev[3].record()
    ops.softmax_bwd_(s3h, dA, dsh, None, write_f32=False)
    ev[4].record()
    ops.layernorm_bwd(dyln.view(B * npad, 512), xln, lw, mean, rstd, 16385, npad, 255)
    ev[5].record()
torch.cuda.synchronize()
print("resconv_fwd %.0f us | ppeg_fwd %.0f us | softmax_fwd(bf16 out) %.0f us | softmax_bwd %.0f us | ln_bwd(+colsums) %.0f us"
      % tuple(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(5)))

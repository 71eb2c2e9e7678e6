"""One pass over the round's dominant kernels at the bench sizes (B = 8 TransMIL shapes, B = 1 deformable attention),
each launched `reps` times: the target of the round's `ncu --set full` capture and of plain CUDA-event timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200.ops import Split, gemm, to_bf16
from sml_b200._lib import lib, ptr, stream, check, c_i32

torch.manual_seed(0)
dev = "cuda"; bf = torch.bfloat16
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, npad, h, d, m = 8, 16640, 8, 64, 256
ql = torch.randn(64, 256, 64, device=dev).to(bf); k = torch.randn(64, 16640, 64, device=dev).to(bf)
q = torch.randn(64, 16640, 64, device=dev).to(bf) * 0.3
out_a = torch.empty(64, 256, 16640, device=dev)                                  # (a) sim3 (fp32 out, K = 64)
a1 = torch.empty(64, 16640, 256, device=dev, dtype=bf)                           # (b) sim1 + fused row softmax -> bf16
x = torch.randn(133120, 512, device=dev).to(bf); w = torch.randn(1536, 512, device=dev).to(bf)
out_c = torch.empty(133120, 1536, device=dev); out_ch = torch.empty(133120, 1536, device=dev, dtype=bf)   # (c) qkv projection
a2 = to_bf16(torch.softmax(torch.randn(64, 256, 256, device=dev), -1), split=True)
z = to_bf16(torch.randn(64, 256, 256, device=dev) * 0.01, split=True)
out_d = torch.empty(64, 256, 256, device=dev); out_dh = torch.empty(64, 256, 256, device=dev, dtype=bf); out_dl = torch.empty_like(out_dh)
# fused A1-path backward
dout = torch.randn(B, npad, h * d, device=dev).to(bf); wm = torch.randn(B, h, m, d, device=dev).to(bf)
a1p = torch.softmax(torch.randn(B, h, npad, m, device=dev), -1).to(bf); kl = torch.randn(B, h, m, d, device=dev).to(bf)
qkv = torch.randn(B, npad, 3, h, d, device=dev).to(bf); dql = torch.randn(B, h, m, d, device=dev)
dqkv = torch.empty(B, npad, 3, h, d, device=dev); dqkv_hi = torch.empty(B, npad, 3, h, d, device=dev, dtype=bf)
dkl = torch.zeros(B, h, m, d, device=dev)
# deformable attention, 2-D, B = 1
mod = sml_b200.DeformCrossAttention2D(dim=128).to(dev).eval()
mod.precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
x1 = torch.randn(1, 128, 16384, device=dev, requires_grad=True); x2 = torch.randn(1, 128, 16384, device=dev, requires_grad=True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
for r in range(reps):
    ev[0].record()
    gemm(Split(ql), Split(k.transpose(-1, -2)), out=out_a)
    ev[1].record()
    gemm(Split(q), Split(ql.transpose(-1, -2)), out=None, out_bf16=a1, want_f32=False, softmax=True)
    ev[2].record()
    gemm(Split(x), Split(w.t()), out=out_c, out_bf16=out_ch)
    ev[3].record()
    gemm(a2, z, out=out_d, out_bf16=out_dh, out_bf16_lo=out_dl)
    ev[4].record()
    check(lib.sml_nystrom_a1_bwd(ptr(dout), ptr(wm), ptr(a1p), ptr(kl), ptr(qkv), ptr(dql), ptr(dqkv), ptr(dqkv_hi), ptr(dkl), c_i32(B),
                                 c_i32(npad), c_i32(h), c_i32(d), c_i32(m), stream()))
    ev[5].record()
    y, vg = mod(x1, x2, return_vgrid=True)
    ev[6].record()
    (y.sum() + vg.sum()).backward()
    ev[7].record()
torch.cuda.synchronize()
names = ["sim3 gemm", "sim1+softmax gemm", "qkv gemm", "pinv-step gemm (bf16x3)", "a1_bwd fused", "deform2d module fwd (B=1)", "deform2d module bwd (B=1)"]
print(" | ".join("%s %.1f us" % (n, ev[i].elapsed_time(ev[i + 1]) * 1e3) for i, n in enumerate(names)))
print("timeout flag", lib.sml_tc_timeout_flag())

"""Stand-alone launches of the three dominant tcgen05 GEMM shapes of the TransMIL step (B=8), for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200.ops import Split, gemm, to_bf16

torch.manual_seed(0)
dev = "cuda"
bf = torch.bfloat16
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
# (a) sim3: ql [64,256,64] x k^T -> [64,256,16640] fp32 (K = 64, epilogue/store bound)
ql = torch.randn(64, 256, 64, device=dev).to(bf); k = torch.randn(64, 16640, 64, device=dev).to(bf)
out_a = torch.empty(64, 256, 16640, device=dev)
# (b) qkv projection: [133120,512] x [512,1536] -> fp32 + bf16
x = torch.randn(133120, 512, device=dev).to(bf); w = torch.randn(1536, 512, device=dev).to(bf)
out_b = torch.empty(133120, 1536, device=dev); out_bh = torch.empty(133120, 1536, device=dev, dtype=bf)
# (c) pinv step: 64 x (256x256x256), bf16x3 split, fp32 + hi/lo out
a2 = to_bf16(torch.softmax(torch.randn(64, 256, 256, device=dev), -1), split=True)
z = to_bf16(torch.randn(64, 256, 256, device=dev) * 0.01, split=True)
out_c = torch.empty(64, 256, 256, device=dev); out_ch = torch.empty(64, 256, 256, device=dev, dtype=bf); out_cl = torch.empty_like(out_ch)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for r in range(reps):
    ev[0].record()
    gemm(Split(ql), Split(k.transpose(-1, -2)), out=out_a)
    ev[1].record()
    gemm(Split(x), Split(w.t()), out=out_b, out_bf16=out_bh)
    ev[2].record()
    gemm(a2, z, out=out_c, out_bf16=out_ch, out_bf16_lo=out_cl)
    ev[3].record()
torch.cuda.synchronize()
print("sim3 %.1f us | qkv %.1f us | pinv %.1f us" % tuple(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(3)))
print("timeout flag", sml_b200._lib.lib.sml_tc_timeout_flag())

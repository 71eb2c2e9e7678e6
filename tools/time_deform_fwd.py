"""time the fused 2-D deformable attention forward kernels at the bench size (B slides x 16 384 queries x 1024 keys x 8 heads):
warp-level mma.sync (bf16 / bf16x3) vs the tcgen05 / TMEM kernel.  usage: python tools/time_deform_fwd.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200._lib import c_f32, c_i32, c_u64, check, lib, ptr, stream

b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
heads = G = 8; T = 16384; nkv = 1024; side = 128
torch.manual_seed(0)
q = torch.randn(b, T, heads * 64, device="cuda") * 0.3
k = torch.randn(b, nkv, heads * 64, device="cuda"); v = torch.randn(b, nkv, heads * 64, device="cuda")
vs = torch.rand(b * G, nkv, 2, device="cuda") * 2.4 - 1.2
cpb = torch.randn(lib.sml_deform_cpb_nparams(c_i32(2), c_i32(1)), device="cuda") * 0.4
cfg = (c_i32 * 8)(b, heads, T, nkv, G, 2, side, 64)
out = torch.empty(b, T, heads * 64, device="cuda"); lse = torch.empty(b, heads, T, device="cuda")
pairs = float(b) * heads * T * nkv
flops = pairs * (2 * (2 * 32 + 32 * 32 + 32) + 4 * 64)


def run(name):
    if name == "tc5":
        check(lib.sml_deform_attn_fwd_tc5(cfg, c_f32(0.1), c_u64(1), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse), stream()))
    else:
        check(lib.sml_deform_attn_fwd_tc(cfg, c_f32(0.1), c_u64(1), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse),
                                         c_i32(1 if name == "split" else 0), stream()))


for name in ("mma", "tc5", "split"):
    for _ in range(2):
        run(name)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run(name)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%-6s B=%d: %.3f ms  %.1f TFLOP/s algorithmic  %.0f cycles @1.965 GHz per 128-pair unit per SM" % (
        name, b, ms, flops / ms / 1e9, ms * 1e-3 * 1.965e9 / (pairs / 128 / 148)))

"""time the fused 2-D deformable attention kernels at the bench size (B slides x 16 384 queries x 1024 keys x 8 heads):
lean bf16 kernels (and their variants) vs the round-1 bf16 kernels vs the fp32-grade split.
usage: python tools/time_deform_fwd.py [B] [variants...]   variant = name:splitcode"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200._lib import c_f32, c_i32, c_u64, check, lib, ptr, stream

b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
variants = [a.split(":") for a in sys.argv[2:]] or [["lean", "0"], ["old", "2"], ["split", "1"]]
heads = G = 8; T = 16384; nkv = 1024; side = 128
torch.manual_seed(0)
q = torch.randn(b, T, heads * 64, device="cuda") * 0.3
k = torch.randn(b, nkv, heads * 64, device="cuda"); v = torch.randn(b, nkv, heads * 64, device="cuda")
vs = torch.rand(b * G, nkv, 2, device="cuda") * 2.4 - 1.2
cpb = torch.randn(lib.sml_deform_cpb_nparams(c_i32(2), c_i32(1)), device="cuda") * 0.4
cfg = (c_i32 * 8)(b, heads, T, nkv, G, 2, side, 64)
out = torch.empty(b, T, heads * 64, device="cuda"); lse = torch.empty(b, heads, T, device="cuda")
dout = torch.randn_like(out); dq = torch.empty_like(q); dk = torch.zeros_like(k); dv = torch.zeros_like(v); dvs = torch.zeros_like(vs)
part = torch.empty(lib.sml_deform_attn_bwd_blocks(cfg), lib.sml_deform_cpb_nparams(c_i32(2), c_i32(1)), device="cuda")
pairs = float(b) * heads * T * nkv
flops = pairs * (2 * (2 * 32 + 32 * 32 + 32) + 4 * 64)


def fwd(code):
    check(lib.sml_deform_attn_fwd_tc(cfg, c_f32(0.1), c_u64(1), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse),
                                     c_i32(code), stream()))


def bwd(code):
    check(lib.sml_deform_attn_bwd_tc(cfg, c_f32(0.1), c_u64(1), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse), ptr(dout),
                                     ptr(dq), ptr(dk), ptr(dv), ptr(dvs), ptr(part), c_i32(code), stream()))


def timeit(fn, code, reps=3):
    for _ in range(2):
        fn(code)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn(code)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, code in variants:
    code = int(code)
    f = timeit(fwd, code)
    g = timeit(bwd, code)
    cyc = lambda ms: ms * 1e-3 * 1.965e9 / (pairs / 128 / 148)
    print("%-6s B=%d: fwd %.3f ms (%.0f TFLOP/s, %.0f cyc/unit)   bwd %.3f ms (%.0f TFLOP/s, %.0f cyc/unit)" % (
        name, b, f, flops / f / 1e9, cyc(f), g, 2 * flops / g / 1e9, cyc(g)))

"""time the 1-D fused deformable attention kernels at the bench size (B x 16385 queries x 4096 keys x 8 heads, 4 offset groups):
MLP-per-pair lean kernels vs the piecewise-linear-table kernels.  usage: python tools/time_deform1d.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200._lib import c_f32, c_i32, c_u64, check, lib, ptr, stream
from sml_b200.deform import cpb1d_table, key_ranges

b = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
heads, G, T, nkv = 8, 4, 16385, 4096
torch.manual_seed(0)
q = torch.randn(b, T, heads * 64, device="cuda") * 0.3
k = torch.randn(b, nkv, heads * 64, device="cuda"); v = torch.randn(b, nkv, heads * 64, device="cuda")
vs = (torch.linspace(-1, 1, nkv, device="cuda").repeat(b * G, 1) + torch.randn(b * G, nkv, device="cuda") * 2e-3).unsqueeze(-1).contiguous()
npar = lib.sml_deform_cpb_nparams(c_i32(1), c_i32(2))
cpb = torch.randn(npar, device="cuda") * 0.4
if "--dense" in sys.argv:           # layer-1 kinks clustered in |t| < 0.1: many (row, key tile) ranges hold >= 2 breakpoints
    w1 = torch.randn(32, device="cuda") * 5.0
    cpb[:32] = w1; cpb[32:64] = -w1 * torch.randn(32, device="cuda") * 0.03
if "--wide" in sys.argv:            # large first-layer weights: ~1000 breakpoints spread over the whole distance range
    cpb[:32] *= 8.0; cpb[64:64 + 1024] *= 3.0
cfg = (c_i32 * 8)(b, heads, T, nkv, G, 1, 0, 64)
out = torch.empty(b, T, heads * 64, device="cuda"); lse = torch.empty(b, heads, T, device="cuda")
dout = torch.randn_like(out); dq = torch.empty_like(q); dk = torch.zeros_like(k); dv = torch.zeros_like(v); dvs = torch.zeros_like(vs)
part = torch.empty(lib.sml_deform_attn_bwd_blocks(cfg), npar, device="cuda")
brk, coef, nseg = cpb1d_table(cpb, 2)
print("segments", int(nseg))
gseg = torch.zeros(2, lib.sml_cpb1d_seg_rows(), 2, device="cuda")


def fwd_lean(): check(lib.sml_deform_attn_fwd_tc(cfg, c_f32(0.0), c_u64(1), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse), c_i32(0), stream()))
kh, vh, kr = k.bfloat16(), v.bfloat16(), key_ranges(cfg, vs)
def fwd_pwl(): check(lib.sml_deform_attn_fwd_pwl(cfg, c_f32(0.0), c_u64(1), ptr(q), ptr(kh), ptr(vh), ptr(vs), ptr(kr), ptr(brk), ptr(coef), ptr(nseg), ptr(out), ptr(lse), stream()))
def bwd_lean(): check(lib.sml_deform_attn_bwd_tc(cfg, c_f32(0.0), c_u64(1), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse), ptr(dout), ptr(dq), ptr(dk), ptr(dv), ptr(dvs), ptr(part), c_i32(0), stream()))
fns = [("fwd table", fwd_pwl)]
if "--lean" in sys.argv: fns += [("fwd lean (MLP per pair)", fwd_lean), ("bwd lean (MLP per pair)", bwd_lean)]
if hasattr(lib, "sml_deform_attn_bwd_pwl"):
    def bwd_pwl(): check(lib.sml_deform_attn_bwd_pwl(cfg, c_f32(0.0), c_u64(1), ptr(q), ptr(kh), ptr(vh), ptr(vs), ptr(kr), ptr(brk), ptr(coef), ptr(nseg), ptr(out), ptr(lse), ptr(dout), ptr(dq), ptr(dk), ptr(dv), ptr(dvs), ptr(gseg), stream()))
    fns.append(("bwd table", bwd_pwl))
pairs = float(b) * heads * T * nkv
for name, fn in fns:
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%-26s B=%d: %.3f ms  (%.0f cycles @1.965 GHz per 128-pair unit per SM)" % (name, b, ms, ms * 1e-3 * 1.965e9 / (pairs / 128 / 148)))

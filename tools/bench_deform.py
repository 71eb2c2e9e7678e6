"""Per-kernel timing of the deformable path at N = 16 384 (2-D, 128x128 lattice, n_kv = 1024) and N = 2 500:
achieved GB/s of the gather / scatter pair and TFLOP/s of the fused CPB attention (SURVEY §8d formulas)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200.deform import gather_with_taps
from sml_b200._lib import lib

dev = "cuda"


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


out = {}
for B, side in ((8, 128), (8, 50)):
    N = side * side
    torch.manual_seed(0)
    mod = sml_b200.DeformCrossAttention2D(dim=128).to(dev).eval()
    x1 = torch.randn(B, 128, N, device=dev, requires_grad=True); x2 = torch.randn(B, 128, N, device=dev, requires_grad=True)
    hk = (side + 2 - 6) // 4 + 1
    nkv = hk * hk
    # gather alone (B*8 groups, 16 channels each): bytes per sample = 4 taps*16*4 read + 16*4 write + 8 coords
    vs = (torch.rand(B * 8, nkv, 2, device=dev) * 2 - 1)
    x2t = x2.detach().transpose(1, 2).contiguous()
    t_g = timeit(lambda: gather_with_taps(x2t, vs, 8, side, side), 20)
    gbytes = B * 8 * nkv * (4 * 16 * 4 + 16 * 4 + 8)
    # full forward / forward+backward of the attention module
    t_f = timeit(lambda: mod(x1, x2), 3)
    def fb():
        y, vg = mod(x1, x2, return_vgrid=True)
        (y.sum() + vg.sum()).backward()
    t_fb = timeit(fb, 3)
    pairs = B * 8 * N * nkv
    flops_fwd = pairs * 2240 + 4 * B * 8 * N * nkv * 64
    out["N=%d" % N] = dict(B=B, n_kv=nkv, gather_us=t_g * 1e6, gather_GBps=gbytes / t_g / 1e9, gather_alg_bytes=gbytes,
                           attn_module_fwd_ms=t_f * 1e3, attn_module_fwd_bwd_ms=t_fb * 1e3,
                           fwd_algorithmic_TFLOPs=flops_fwd / t_f / 1e12, pairs=pairs)
print(json.dumps(out, indent=1))

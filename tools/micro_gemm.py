"""Graph-replayed micro-timings of small tcgen05 GEMM launches (pure GPU time per launch, no host enqueue cost)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200.ops import Split, gemm, to_bf16

dev = "cuda"
bf = torch.bfloat16


def graph_time(fn, n=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for nb in (8, 64):
    A = to_bf16(torch.softmax(torch.randn(nb, 256, 256, device=dev), -1), split=True)
    Z = to_bf16(torch.randn(nb, 256, 256, device=dev) * 0.01, split=True)
    C = torch.empty(nb, 256, 256, device=dev); H = torch.empty(nb, 256, 256, device=dev, dtype=bf); L = torch.empty_like(H)
    res = {}
    res["split N-major B, f32+hi+lo"] = graph_time(lambda: gemm(A, Z, out=C, out_bf16=H, out_bf16_lo=L))
    res["split N-major B, f32 only"] = graph_time(lambda: gemm(A, Z, out=C))
    res["split K-major B (Z^T view), f32+hi+lo"] = graph_time(lambda: gemm(A, Z.transpose(-1, -2), out=C, out_bf16=H, out_bf16_lo=L))
    res["single-pass bf16, f32 only"] = graph_time(lambda: gemm(Split(A.hi), Split(Z.hi), out=C))
    A64 = A.map(lambda t: t[:, :, :64].contiguous()); Z64 = Z.map(lambda t: t[:, :64, :].contiguous())
    res["split K=64 (1 k-block), f32+hi+lo"] = graph_time(lambda: gemm(A64, Z64, out=C, out_bf16=H, out_bf16_lo=L))
    res["split, N=128 half (BN=128 tiles)"] = graph_time(lambda: gemm(A, Z.map(lambda t: t[:, :, :128]), out=C[:, :, :128]))
    print("batch", nb, {k: "%.1f us" % v for k, v in res.items()})
x = torch.randn(8, 512, device=dev); w = torch.randn(4, 512, device=dev)
print("simt 8x4x512: %.1f us" % graph_time(lambda: gemm(x, w.t())))
torch.cuda.synchronize()
print("timeout flag", sml_b200._lib.lib.sml_tc_timeout_flag())

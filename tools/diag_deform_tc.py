"""Tensor-core deformable attention (bf16x3 / bf16) against the SIMT fp32 kernels: outputs, every gradient, timings."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200

dev = "cuda"


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(mod, x1, x2, gy, prec):
    mod.precision = prec
    for p in mod.parameters():
        p.grad = None
    a = x1.clone().requires_grad_(); b = x2.clone().requires_grad_()
    y, vg = mod(a, b, return_vgrid=True)
    (y * gy).sum().backward()
    return y.detach(), a.grad, b.grad, {k: p.grad.clone() for k, p in mod.named_parameters()}


def case(kind, B, n, drop=0.0):
    torch.manual_seed(0)
    if kind == "2d":
        mod = sml_b200.DeformCrossAttention2D(dim=128, dropout=drop).to(dev)
    else:
        mod = sml_b200.DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6, dropout=drop).to(dev)
    mod.train(drop > 0)
    with torch.no_grad():
        for p in mod.rel_pos_bias.parameters():
            p.mul_(2.0)
    x1 = torch.randn(B, 128, n, device=dev); x2 = torch.randn(B, 128, n, device=dev); gy = torch.randn(B, 128, n, device=dev)
    from sml_b200 import functional as Fn
    outs = {}
    for prec in ("fp32", "bf16x3", "bf16"):
        Fn.manual_seed(5)
        outs[prec] = run(mod, x1, x2, gy, prec)
    for prec in ("bf16x3", "bf16"):
        y0, a0, b0, g0 = outs["fp32"]; y, a, b, g = outs[prec]
        worst = max((rel(g[k], g0[k]), k) for k in g0)
        print("%s B=%d n=%d drop=%.1f %-6s y %.2e dx1 %.2e dx2 %.2e worst-param %.2e (%s)" %
              (kind, B, n, drop, prec, rel(y, y0), rel(a, a0), rel(b, b0), worst[0], worst[1]))
        bad = sorted(((rel(g[k], g0[k]), k) for k in g0), reverse=True)[:4]
        print("      ", " ".join("%s=%.1e" % (k.replace("rel_pos_bias.", "cpb."), v) for v, k in bad))


def timing(kind, B, n):
    torch.manual_seed(0)
    if kind == "2d":
        mod = sml_b200.DeformCrossAttention2D(dim=128).to(dev).eval()
    else:
        mod = sml_b200.DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6).to(dev).eval()
    x1 = torch.randn(B, 128, n, device=dev, requires_grad=True); x2 = torch.randn(B, 128, n, device=dev, requires_grad=True)
    for prec in ("bf16x3", "bf16") + (("fp32",) if os.environ.get("TIME_FP32") else ()):
        mod.precision = prec
        def f():
            return mod(x1, x2)
        def fb():
            y, vg = mod(x1, x2, return_vgrid=True)
            (y.sum() + vg.sum()).backward()
        res = []
        for fn in (f, fb):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                fn()
            e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 3)
        print("time %s B=%d n=%d %-6s fwd %.2f ms  fwd+bwd %.2f ms" % (kind, B, n, prec, res[0], res[1]))


if __name__ == "__main__":
    case("2d", 2, 2500)
    case("2d", 1, 64 * 64)
    case("1d", 2, 2501)
    case("1d", 1, 1000)
    case("2d", 1, 2500, drop=0.1)
    if len(sys.argv) > 1:
        timing("2d", 8, 16384)
        timing("1d", 8, 16385)
    print("launches", sml_b200._lib.launch_count())

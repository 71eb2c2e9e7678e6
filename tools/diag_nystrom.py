"""GPU diagnostic (not a test): per-intermediate error of the Nystrom pipeline vs the fp64 oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import oracle, sml_b200
from sml_b200.nystrom import pinv_fwd, pinv_bwd
from synth import randn


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


g = torch.load(os.path.join(ROOT, "tests/golden/pinv.pt"), weights_only=False)
x64 = g["x"].double().requires_grad_()
z64 = oracle.nystrom.pinv_iter(x64, 6); z64.backward(g["gz"].double())
print("pinv golden(fp32 ref) vs fp64 oracle: z %.2e gx %.2e" % (rel(g["z"], z64), rel(g["gx"], x64.grad)))
for mode in ("fp32", "bf16x3", "bf16"):
    z, saved = pinv_fwd(g["x"].cuda(), 6, mode)
    gx = pinv_bwd(g["x"].cuda(), saved, g["gz"].cuda(), mode)
    print("pinv %-6s vs fp64: z %.2e gx %.2e | scal %s arg %s" % (mode, rel(z.f, z64), rel(gx, x64.grad), saved[1].tolist(), saved[2].tolist()))
ax = g["x"].abs(); print("   oracle scal", float(ax.sum(-1).max()), float(ax.sum(-2).max()))

for (dim, dh, hh, m, n, b) in [(64, 8, 8, 16, 37, 2), (512, 64, 8, 256, 1001, 2)]:
    torch.manual_seed(0)
    mod = sml_b200.NystromAttention(dim=dim, dim_head=dh, heads=hh, num_landmarks=m, pinv_iterations=6).cuda().eval()
    p64 = {k: v.detach().cpu().double().requires_grad_() for k, v in mod.state_dict().items()}
    p32 = {k: v.detach().cpu().clone().requires_grad_() for k, v in mod.state_dict().items()}
    x = randn((b, n, dim), 5); gy = randn((b, n, dim), 6)
    x64 = x.double().requires_grad_(); x32 = x.clone().requires_grad_()
    y64 = oracle.nystrom.nystrom_attention(x64, p64, heads=hh, num_landmarks=m); y64.backward(gy.double())
    y32 = oracle.nystrom.nystrom_attention(x32, p32, heads=hh, num_landmarks=m); y32.backward(gy)
    print("cfg", (dim, dh, hh, m, n, b), "fp32-oracle vs fp64: y %.2e gx %.2e gWqkv %.2e" % (rel(y32, y64), rel(x32.grad, x64.grad), rel(p32["to_qkv.weight"].grad, p64["to_qkv.weight"].grad)))
    for mode in ("fp32", "bf16x3", "bf16"):
        try:
            mod.precision = mode
            mod.zero_grad()
            xg = x.cuda().requires_grad_()
            y = mod(xg); y.backward(gy.cuda())
            gp = dict(mod.named_parameters())
            print("   %-6s vs fp64: y %.2e gx %.2e " % (mode, rel(y, y64), rel(xg.grad, x64.grad)) +
                  " ".join("%s %.2e" % (k.split(".")[0] + k[-4:], rel(gp[k].grad, p64[k].grad)) for k in gp))
        except Exception as e:
            print("   %-6s FAILED: %s" % (mode, str(e)[:200]))
torch.cuda.synchronize()
print("timeout flag", sml_b200._lib.lib.sml_tc_timeout_flag())

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, oracle, sml_b200
g = torch.load(os.path.join(ROOT, "tests/golden/deform1d.pt"), weights_only=False)
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
for order in ("reference", "upstream"):
    mod = sml_b200.DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6, pad_order=order).cuda().eval()
    mod.load_state_dict(g["state"])
    p32 = {k: v.clone() for k, v in g["state"].items()}
    p64 = {k: v.double() for k, v in g["state"].items()}
    y32, parts32 = oracle.deform.deform_attn_1d(g["x1"], g["x2"], p32, pad_order=order, return_parts=True)
    y64, parts64 = oracle.deform.deform_attn_1d(g["x1"].double(), g["x2"].double(), p64, pad_order=order, return_parts=True)
    for rep in range(3):
        y, vg = mod(g["x1"].cuda(), g["x2"].cuda(), return_vgrid=True)
        print(order, rep, "kernel vs fp64 %.2e | oracle32 vs fp64 %.2e | vgrid vs fp64 %.2e" % (rel(y, y64), rel(y32, y64), rel(vg, parts64["vgrid"])),
              "taps32==taps64:", bool(torch.equal(parts32["taps"][1], parts64["taps"][1])))

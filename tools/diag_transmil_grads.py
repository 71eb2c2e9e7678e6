"""bf16 vs bf16x3 gradients of the TransMIL step at bench size: per-parameter max-norm and L2 relative differences."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from argparse import Namespace
import sml_b200
from sml_b200 import nystrom as ny

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.manual_seed(0)
net = sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).cuda().eval()
x = torch.randn(1, N, 1024, device="cuda")
y = torch.tensor([2], device="cuda")


def grads(mode, fuse):
    ny.FUSE_A1_BWD = fuse
    net.set_precision(mode)
    for p in net.parameters():
        p.grad = None
    xx = x.clone().requires_grad_()
    enc, logits, _ = net(xx)
    loss = torch.nn.functional.cross_entropy(logits, y) + 0.01 * enc.square().sum()
    loss.backward()
    g = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    g["input"] = xx.grad.clone()
    return float(loss), g


l0, ref = grads("bf16x3", False)
for mode, fuse in (("bf16", True), ("bf16", False), ("fp32", False)) if N <= 4096 else (("bf16", True), ("bf16", False)):
    l1, g = grads(mode, fuse)
    print("== %s fuse_a1=%s  loss %.6f vs %.6f" % (mode, fuse, l1, l0))
    rows = []
    for k in ref:
        d = (g[k] - ref[k]).double()
        rows.append((float(d.abs().max() / ref[k].abs().max().clamp_min(1e-30)), float(d.norm() / ref[k].double().norm().clamp_min(1e-30)), k))
    for mx, l2, k in sorted(rows, reverse=True)[:8]:
        print("   max-rel %.3e  l2-rel %.3e  %s" % (mx, l2, k))

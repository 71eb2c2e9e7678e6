"""Stand-alone launches of the tensor-core deformable attention fwd/bwd kernels (2-D, 128x128 lattice, n_kv = 1024), for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = "cuda"
torch.manual_seed(0)
mod = sml_b200.DeformCrossAttention2D(dim=128).to(dev).eval()
mod.precision = prec
x1 = torch.randn(B, 128, 16384, device=dev, requires_grad=True); x2 = torch.randn(B, 128, 16384, device=dev, requires_grad=True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for _ in range(reps):
    ev[0].record()
    y, vg = mod(x1, x2, return_vgrid=True)
    ev[1].record()
    (y.sum() + vg.sum()).backward()
    ev[2].record()
torch.cuda.synchronize()
print("B=%d %s module fwd %.2f ms bwd %.2f ms" % (B, prec, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))

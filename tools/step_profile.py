"""Kernel-time breakdown of the bench train step with torch.profiler (CUPTI activity records, no replays)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from argparse import Namespace
import sml_b200
from sml_b200 import functional as Fn
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dev = "cuda"
torch.manual_seed(42)
workload = sys.argv[3] if len(sys.argv) > 3 else "transmil"
if workload == "transmil":
    model = sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).to(dev).train()
else:
    model = sml_b200.DeformCrossTransMIL(Namespace(path_dim=128, attn_dim=2 if workload == "deform2d" else 1, return_vgrid=False)).to(dev).train()
omic = torch.randn(B, 128, device=dev).abs()
model.set_precision(prec)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0.1, fused=True)
x = torch.randn(B, 16384, 1024, device=dev); y = torch.randint(0, 4, (B,), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    logits = model(x)[1] if workload == "transmil" else model(x, omic)[1]
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name.split("(")[0][:90]
        agg[name][0] += 1; agg[name][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(workload, "B=%d %s: %.2f ms of kernel time per step (%d kernels/step)" % (B, prec, tot / N / 1e3, sum(v[0] for v in agg.values()) // N))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%9.1f us/step %5.1f%% n=%4d avg %7.1f  %s" % (v[1] / N, 100 * v[1] / tot, v[0] // N, v[1] / v[0], k))

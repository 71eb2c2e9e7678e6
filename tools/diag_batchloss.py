import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import sml_b200
from sml_b200.utils.gather import PackedGather

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

for Ng in (8, 16, 64):
    o = torch.randn(Ng, 128, device=dev, requires_grad=True); vg = torch.randn(8 * Ng, 2, 32, 32, device=dev, requires_grad=True)
    bl = sml_b200.BatchLoss(Ng, 1)
    def f():
        bl(o.unsqueeze(1).expand(Ng, 16384, 128), vg).sum().backward()
    r = timeit(f)
    if rank == 0: print("BatchLoss single-process Ng=%d: gpu %.2f ms, wall %.2f ms" % (Ng, *r))
if world > 1:
    o = torch.randn(8, 128, device=dev, requires_grad=True); vg = torch.randn(64, 2, 32, 32, device=dev, requires_grad=True)
    def g():
        a, b = PackedGather.apply(o, vg)
        (a.sum() + b.sum()).backward()
    r = timeit(g)
    if rank == 0: print("PackedGather fwd+bwd: gpu %.2f ms, wall %.2f ms" % r)
    bl = sml_b200.BatchLoss(8, world)
    def h():
        bl(o.unsqueeze(1).expand(8, 16384, 128), vg).sum().backward()
    r = timeit(h)
    if rank == 0: print("BatchLoss world %d: gpu %.2f ms, wall %.2f ms" % (world, *r))
    dist.destroy_process_group()

"""Phase timing (CUDA events) of the DeformPathomicNet train step at world size 1 or 2 (torchrun)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from argparse import Namespace
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import sml_b200
margs = Namespace(init_type="max", input_size_omic_tumor=59, input_size_omic_immune=361, omic_dim=128, path_dim=128, mmhid=128,
                  return_grad="False", dropout_rate=0.1, label_dim=4, fusion_type="concat", cut_fuse_grad=False, return_vgrid=True,
                  attn_dim=2, task_type="diag2021")
torch.manual_seed(0)
model = sml_b200.DeformPathomicNet(margs).to(dev).train(); model.set_precision("bf16")
net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True) if world > 1 else model
bl = sml_b200.BatchLoss(8, world)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
x = torch.randn(8, 16384, 1024, device=dev); y = torch.randint(0, 4, (8,), device=dev)
xt = torch.randn(8, 59, device=dev).abs(); xi = torch.randn(8, 361, device=dev).abs()
use_bl = os.environ.get("NO_BL", "0") != "1"
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
acc = [0.0] * 4
for it in range(8):
    opt.zero_grad(set_to_none=True)
    import time
    t0 = time.perf_counter()
    ev[0].record()
    lg = net(x_path=x, x_omic_tumor=xt, x_omic_immune=xi)[3]
    ev[1].record()
    t1 = time.perf_counter()
    loss = torch.nn.functional.cross_entropy(lg[2], y)
    if use_bl:
        loss = loss + 0.5 * bl(lg[3], lg[4]).sum() + 0.5 * bl(lg[5], lg[6]).sum()
    ev[2].record()
    t2 = time.perf_counter()
    loss.backward()
    ev[3].record()
    t3 = time.perf_counter()
    opt.step()
    ev[4].record()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    if it == 6:
        print("rank %d CPU: fwd issued %.1f ms, loss issued %.1f, bwd issued %.1f, step done %.1f | GPU: fwd %.1f loss %.1f bwd %.1f"
              % (rank, (t1 - t0) * 1e3, (t2 - t0) * 1e3, (t3 - t0) * 1e3, (t4 - t0) * 1e3, ev[0].elapsed_time(ev[1]),
                 ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])), flush=True)
    if it >= 3:
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1]) / 5
if rank == 0:
    print("world %d batchloss %s: fwd %.1f  loss %.1f  bwd %.1f  opt %.1f ms" % (world, use_bl, *acc))
if world > 1:
    dist.destroy_process_group()

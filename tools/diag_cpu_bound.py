"""Is the TransMIL train step CPU- (launch-) bound?  CPU issue time per step vs GPU time per step at two batch sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from argparse import Namespace
import sml_b200
dev = "cuda"
torch.manual_seed(0)
model = sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).to(dev).train(); model.set_precision("bf16")
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0.1, fused=True)
for B in (4, 8, 16):
    x = torch.randn(B, 16384, 1024, device=dev); y = torch.randint(0, 4, (B,), device=dev)
    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(x)[1], y)
        loss.backward(); opt.step()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10): step()
    t_issue = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    print("B=%2d: CPU issue %.2f ms/step, GPU %.2f ms/step -> %.1f slides/s" % (B, t_issue * 100, e0.elapsed_time(e1) / 10, B * 1e4 / e0.elapsed_time(e1)))
import cProfile, pstats
x = torch.randn(8, 16384, 1024, device=dev); y = torch.randint(0, 4, (8,), device=dev)
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(model(x)[1], y)
    loss.backward(); opt.step()
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative")
import io; s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])

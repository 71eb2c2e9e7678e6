"""BASELINE.json configs[4]: TransMIL inference on one 100 489-patch bag (317^2), 512 landmarks, bf16, no-grad (replicas
only at N > 1).  Prints slides/s and the agreement with the fp32-grade split."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from argparse import Namespace
import sml_b200

dev = "cuda"
torch.manual_seed(42)
net = sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).to(dev).eval()
for lyr in (net.layer1, net.layer2):
    lyr.attn.num_landmarks = 512
x = torch.randn(1, 317 * 317, 1024, device=dev)
res = {}
with torch.no_grad():
    for mode in ("bf16x3", "bf16"):
        net.set_precision(mode)
        for _ in range(2):
            enc, logits, _ = net(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            enc, logits, _ = net(x)
        e1.record(); torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / 5, logits.clone(), enc.clone())
d = float((res["bf16"][1] - res["bf16x3"][1]).abs().max() / res["bf16x3"][1].abs().max())
print(json.dumps({"workload": "TransMIL inference, 100 489-patch bag, 512 landmarks, no-grad", "bf16_ms": res["bf16"][0],
                  "bf16_slides_per_s": 1e3 / res["bf16"][0], "bf16x3_ms": res["bf16x3"][0],
                  "logits_rel_diff_bf16_vs_bf16x3": d, "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}))

"""BASELINE configs[0]'s co-attention: 6 genomic group tokens attending over a 16 384-patch bag (MCAT model.py:627, E = 256, 1 head),
forward + backward, B slides: the re-associated form (K / V never materialised) against the projected form.
usage: python tools/bench_coattn.py [B] [precision]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
torch.manual_seed(0)
mha = sml_b200.MultiheadAttention(256, 1).cuda().train()
mha.precision = prec
q = torch.randn(6, B, 256, device="cuda", requires_grad=True)
kv = (torch.randn(16384, B, 256, device="cuda") * 0.5).requires_grad_()
out = {"workload": "6 genomic queries x 16384 patch tokens, E=256, 1 head, fwd+bwd", "B": B, "precision": prec}
for name, fold in (("projected", False), ("folded", True)):
    mha.fold_kv = fold
    for _ in range(3):
        o, raw = mha(q, kv, kv); (o.sum() + raw.sum() * 1e-3).backward()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o, raw = mha(q, kv, kv); (o.sum() + raw.sum() * 1e-3).backward()
    e1.record(); torch.cuda.synchronize()
    out[name + "_ms"] = e0.elapsed_time(e1) / 10
    out[name + "_peak_MB"] = torch.cuda.max_memory_allocated() / 2 ** 20
print(json.dumps(out))

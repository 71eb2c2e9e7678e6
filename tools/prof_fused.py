"""Stand-alone launch of the fused A1-path backward kernel at the B=8 TransMIL sizes (timing + ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200._lib import lib, ptr, stream, check, c_i32
dev = "cuda"; bf = torch.bfloat16
B, npad, h, d, m = 8, 16640, 8, 64, 256
dout = torch.randn(B, npad, h * d, device=dev).to(bf)
wm = torch.randn(B, h, m, d, device=dev).to(bf)
a1 = torch.softmax(torch.randn(B, h, npad, m, device=dev), -1).to(bf)
kl = torch.randn(B, h, m, d, device=dev).to(bf)
qkv = torch.randn(B, npad, 3, h, d, device=dev).to(bf)
dql = torch.randn(B, h, m, d, device=dev)
dqkv = torch.empty(B, npad, 3, h, d, device=dev); dqkv_hi = torch.empty(B, npad, 3, h, d, device=dev, dtype=bf)
dkl = torch.zeros(B, h, m, d, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ev[0].record()
    check(lib.sml_nystrom_a1_bwd(ptr(dout), ptr(wm), ptr(a1), ptr(kl), ptr(qkv), ptr(dql), ptr(dqkv), ptr(dqkv_hi), ptr(dkl), c_i32(B),
                                 c_i32(npad), c_i32(h), c_i32(d), c_i32(m), stream()))
    ev[1].record()
torch.cuda.synchronize()
print("a1_bwd fused: %.0f us ; timeout flag %d" % (ev[0].elapsed_time(ev[1]) * 1e3, lib.sml_tc_timeout_flag()))

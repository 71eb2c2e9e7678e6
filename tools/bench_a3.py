"""time the fused landmark-by-token product of the Nystrom attention (csrc/nystrom_a3.cu: A3 V = softmax(ql k^T) v, forward and
recomputing backward) at the bench size: B x 8 heads, 256 landmarks, n' = 16 640 tokens, head dim 64.
usage: python tools/bench_a3.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200._lib import c_i32, check, lib, ptr, stream
from sml_b200.nystrom import _a3_cfg, _a3_chunks

b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
h, m, d, npad = 8, 256, 64, 16640
dev = "cuda"
torch.manual_seed(0)
qkv = (torch.randn(b, npad, 3, h, d, device=dev) * 0.7).bfloat16()
ql = (torch.randn(b, h, m, d, device=dev) * 0.5).bfloat16()
nf, nb = _a3_chunks(b, h, npad, 64, 3 * 148, rows_per_cta=m // 64), _a3_chunks(b, h, npad, 32, 148)
if len(sys.argv) > 2: nf = int(sys.argv[2])
if len(sys.argv) > 3: nb = int(sys.argv[3])
cf, cb = _a3_cfg(b, h, m, d, npad, nf), _a3_cfg(b, h, m, d, npad, nb)
a3v = torch.empty(b, h, m, d, device=dev); a3v_hi = torch.empty(b, h, m, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(b, h, m, device=dev)
opart = torch.empty(b * h, nf, m, d, device=dev); ml = torch.empty(b * h, nf, m, 2, device=dev)
g = torch.randn(b, h, m, d, device=dev); dkl = torch.randn(b, h, m, d, device=dev)
dql = torch.zeros(b, h, m, d, device=dev)
dqkv_hi = torch.zeros(b, npad, 3, h, d, device=dev, dtype=torch.bfloat16); dqkv_f = torch.zeros(b, npad, 3, h, d, device=dev)


def fwd(): check(lib.sml_nystrom_a3_fwd(cf, ptr(ql), ptr(qkv), ptr(opart), ptr(ml), ptr(a3v), ptr(a3v_hi), ptr(lse), stream()), "fwd")
def bwd(): check(lib.sml_nystrom_a3_bwd(cb, ptr(ql), ptr(qkv), ptr(g), ptr(a3v), ptr(lse), None, c_i32(npad // m), ptr(dql), ptr(dqkv_hi), ptr(dqkv_f), stream()), "bwd")
def post(): check(lib.sml_nystrom_a3_dk_landmarks(cb, ptr(dkl), c_i32(npad // m), ptr(dqkv_hi), stream()), "post")


pairs = float(b) * h * m * npad
for name, fn, flop in (("forward (+ combine), %d chunks" % nf, fwd, 4 * d), ("backward, %d chunks" % nb, bwd, 10 * d), ("dk landmark term", post, 0)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print("%-32s B=%d: %7.1f us  %6.1f TFLOP/s" % (name, b, us, pairs * flop / us / 1e6))

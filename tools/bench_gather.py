"""HBM roofline of the deformable gather (forward) and its backward (coordinate gradients + sorted atomic-free scatter) on a launch
large enough to leave the launch-latency regime (one slide touches < 3 MB).  Algorithmic bytes per sample as SURVEY 8(d):
forward taps*C_g*4 read + C_g*4 write + coords; backward C_g*4 (d kv) + taps*C_g*4 (x2 for the coordinate gradient) read,
taps*C_g*4 (d x2) + coords written.   usage: python tools/bench_gather.py [B]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
from sml_b200._lib import c_i32, check, lib, ptr, stream

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
peak = 6543.7
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(name, b, G, H, W, cg, nkv, CD, mode, taps):
    torch.manual_seed(0)
    x2 = torch.randn(b, H * W, G * cg, device="cuda")
    vs = (torch.randn(b * G, nkv, CD, device="cuda") * 0.5).clamp(-1.2, 1.2)
    kv = torch.empty(b, nkv, G * cg, device="cuda"); dkv = torch.randn_like(kv)
    dx2 = torch.zeros_like(x2); dvs = torch.empty_like(vs)
    cfg = (c_i32 * 8)(b, G, cg, H, W, nkv, CD, mode)
    ns = b * G * nkv
    fwd_bytes = ns * (taps * cg * 4 + cg * 4 + 4 * CD)
    bwd_bytes = ns * (cg * 4 + 2 * taps * cg * 4 + 8 * CD)

    def t(fn, reps=5):
        best = 1e9
        for _ in range(reps):
            flush.zero_()                      # L2 flush between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    f = t(lambda: check(lib.sml_deform_gather_fwd(cfg, ptr(x2), ptr(vs), ptr(kv), None, stream())))
    g = t(lambda: check(lib.sml_deform_gather_bwd(cfg, ptr(x2), ptr(vs), ptr(dkv), ptr(dx2), None, ptr(dvs), stream())))
    out = {"case": name, "B": b, "samples": ns, "fwd_us": f * 1e3, "fwd_GBs": fwd_bytes / f / 1e6, "fwd_frac": fwd_bytes / f / 1e6 / peak,
           "bwd_us": g * 1e3, "bwd_GBs": bwd_bytes / g / 1e6, "bwd_frac": bwd_bytes / g / 1e6 / peak, "peak_GBs": peak}
    print(json.dumps(out))


run("2-D 128x128 lattice, n_kv 1024, 8 groups x 16 ch", B, 8, 128, 128, 16, 1024, 2, 0, 4)
run("1-D 16385 tokens (upstream pad order), n_kv 4096, 4 groups x 32 ch", B, 4, 16385, 1, 32, 4096, 1, 1, 2)
run("2-D 128x128, one slide batch of 8 (bench size)", 8, 8, 128, 128, 16, 1024, 2, 0, 4)

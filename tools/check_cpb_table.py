"""GPU check of the piecewise-linear table of the 1-D bias MLP (csrc/cpb_table.cu) against the MLP itself and oracle/pwl.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sml_b200
import oracle
from sml_b200._lib import c_i32, c_i64, check, lib, ptr, stream

ok = True
for seed, OUT in ((0, 2), (1, 2), (2, 1)):
    torch.manual_seed(seed)
    p = {"mlp.0.0.weight": torch.randn(32, 1) * 1.5, "mlp.0.0.bias": torch.randn(32) * 0.7, "mlp.1.0.weight": torch.randn(32, 32) * 0.4,
         "mlp.1.0.bias": torch.randn(32) * 0.3, "mlp.2.weight": torch.randn(OUT, 32) * 0.5, "mlp.2.bias": torch.randn(OUT) * 0.1}
    cpb = torch.cat([p[k].reshape(-1) for k in ("mlp.0.0.weight", "mlp.0.0.bias", "mlp.1.0.weight", "mlp.1.0.bias", "mlp.2.weight", "mlp.2.bias")]).cuda()
    slots = lib.sml_cpb1d_table_slots()
    brk = torch.empty(slots, device="cuda"); coef = torch.empty(slots + 1, OUT, 2, device="cuda")
    nseg = torch.zeros(1, device="cuda", dtype=torch.int32)
    check(lib.sml_cpb1d_table_build(ptr(cpb), c_i32(OUT), ptr(brk), ptr(coef), ptr(nseg), stream()), "table_build")
    dx = (torch.rand(1 << 20, device="cuda") * 2 - 1) * 2.4
    t = torch.sign(dx) * torch.log(dx.abs() + 1)
    out = torch.empty(t.numel(), OUT, device="cuda")
    check(lib.sml_cpb1d_table_eval(ptr(t), c_i64(t.numel()), ptr(brk), ptr(coef), ptr(nseg), c_i32(OUT), ptr(out), stream()), "table_eval")
    want = oracle.deform.cpb_mlp(t.double().cpu().unsqueeze(1), {k: v.double() for k, v in p.items()}, "")
    err = float((out.double().cpu() - want).abs().max() / want.abs().max())
    ref_tab = oracle.pwl.build_table(p)
    ns = int(nseg.item())
    print("seed %d OUT %d: segments gpu %d / oracle %d, sorted %s, max rel err %.2e" %
          (seed, OUT, ns, ref_tab[0].numel() + 1, bool((brk[:ns - 1][1:] >= brk[:ns - 1][:-1]).all()), err))
    ok = ok and err < 1e-5 and ns == ref_tab[0].numel() + 1
print("CPB_TABLE_OK" if ok else "CPB_TABLE_MISMATCH")

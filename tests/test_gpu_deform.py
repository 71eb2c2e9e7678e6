"""GPU parity: deformable cross-attention 1-D / 2-D and DeformCrossTransMIL on libsml_b200 vs the golden
fixtures of the unmodified reference (N = 2500) and the CPU oracle (other N).  fp32: 1e-4 relative on outputs,
1e-3 on gradients (atomics / different summation order); tap indices bit-exact for identical coordinates."""
from argparse import Namespace

import pytest
import torch

import oracle
from conftest import grads_close, load_golden, rel_err
from synth import randn, synth_state_dict

pytestmark = pytest.mark.gpu

# Tolerances per arithmetic mode of the fused attention kernel: (outputs, input gradients and projection / offset-network
# parameter gradients, parameter gradients of the bias MLP).  "bf16x3" is the default and meets the fp32 bar of BASELINE.json
# on outputs and input gradients (measured 3e-6 / 4e-5); "bf16" meets its 2e-2 bar.  The last column is wider in the
# tensor-core modes for a structural reason: every bias-MLP gradient is sum_ij dS_ij f(i, j) with sum_j dS_ij = 0, i.e. a
# ~100x cancellation, and f contains the ReLU masks 1[a > 0]: a pre-activation error eps flips the mask of the ~eps/sigma
# pairs with |a| < eps, each flip changing its term by 100 % - an error linear in the pair count, like the signal.  With
# eps = 1e-5 (hi/lo bf16 operands) that is 2e-3 .. 7e-3 of the gradient, with 4e-3 (plain bf16) 5e-2 .. 2e-1; the exact-fp32
# SIMT kernels (precision = "fp32", eps = 1e-7) stay inside 1e-3.
TOLS = {"fp32": (1e-4, 1e-3, 1e-3), "bf16x3": (1e-4, 1e-3, 1e-2), "bf16": (2e-2, 2e-2, 3e-1)}


def check_param_grads(got, want, prec, scale=1.0, what=""):
    _, ti, tc = TOLS[prec]
    # besides the bias MLP, the _fc1 bias gradient is a cancellation-dominated sum (column sum over all tokens of a
    # LayerNorm-backward output): same wider column
    cpb = {k for k in want if "rel_pos_bias" in k or k.endswith("_fc1.0.bias")}
    gmax = max(float(v.abs().max()) for v in want.values())
    grads_close({k: got[k] for k in want if k not in cpb}, {k: v for k, v in want.items() if k not in cpb}, scale * ti, what, gmax)
    if cpb:
        grads_close({k: got[k] for k in cpb}, {k: want[k] for k in cpb}, scale * tc, what, gmax)


@pytest.fixture(params=["fp32", "bf16x3", "bf16"])
def prec(request):
    from sml_b200.models import _deform_common as dc
    old = dc.DEFAULT_PRECISION
    dc.DEFAULT_PRECISION = request.param
    yield request.param
    dc.DEFAULT_PRECISION = old


def test_deform1d_golden(prec):
    import sml_b200
    to, ti, tp = TOLS[prec]
    g = load_golden("deform1d")
    mod = sml_b200.DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6).cuda().eval()
    mod.load_state_dict(g["state"])
    x1 = g["x1"].cuda().requires_grad_(); x2 = g["x2"].cuda().requires_grad_()
    y, vg = mod(x1, x2, return_vgrid=True)
    assert rel_err(vg, g["vgrid"]) < 1e-5
    assert rel_err(y, g["y"]) < to
    y.backward(g["gy"].cuda())
    assert rel_err(x1.grad, g["gx1"]) < ti
    assert rel_err(x2.grad, g["gx2"]) < ti
    check_param_grads({k: p.grad for k, p in mod.named_parameters()}, g["gparams"], prec)


def test_deform1d_upstream_order_vs_oracle(prec):
    import sml_b200
    g = load_golden("deform1d")
    mod = sml_b200.DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6,
                                          pad_order="upstream").cuda().eval()
    mod.load_state_dict(g["state"])
    # the oracle is evaluated in fp64: some GPU-box host CPUs run fp32 matmul/conv through reduced-precision fast paths
    # (observed: the fp32 CPU oracle itself 3.5e-4 off its fp64 evaluation on one box, the kernel 2.4e-6)
    p = {k: v.double().requires_grad_() for k, v in g["state"].items()}
    x1c = g["x1"].double().requires_grad_(); x2c = g["x2"].double().requires_grad_()
    yo = oracle.deform.deform_attn_1d(x1c, x2c, p, pad_order="upstream")
    yo.backward(g["gy"].double())
    x1 = g["x1"].cuda().requires_grad_(); x2 = g["x2"].cuda().requires_grad_()
    to, ti, tp = TOLS[prec]
    y = mod(x1, x2)
    assert rel_err(y, yo) < to
    y.backward(g["gy"].cuda())
    assert rel_err(x1.grad, x1c.grad) < ti and rel_err(x2.grad, x2c.grad) < ti
    check_param_grads({k: q.grad for k, q in mod.named_parameters()}, {k: v.grad for k, v in p.items()}, prec)


def test_gather_taps_bit_exact():
    """Same normalised coordinates in -> identical integer tap indices and values as the oracle sampler."""
    from sml_b200.deform import gather_with_taps
    b, G, H, W, cg, nkv = 2, 8, 50, 50, 16, 144
    x2 = randn((b, H * W, G * cg), 1)
    vs = (randn((b * G, nkv, 2), 2) * 0.7).clamp(-1.4, 1.4)
    feat = x2.view(b, H, W, G, cg).permute(0, 3, 4, 1, 2).reshape(b * G, cg, H, W)
    want, (x0, y0) = oracle.deform.bilinear_sample(feat, vs[..., 0], vs[..., 1])
    kv, taps = gather_with_taps(x2.cuda(), vs.cuda(), G, H, W)
    assert torch.equal(taps[..., 0].cpu().long(), x0) and torch.equal(taps[..., 1].cpu().long(), y0)
    got = kv.view(b, nkv, G, cg).permute(0, 2, 3, 1).reshape(b * G, cg, nkv)
    assert rel_err(got, want) < 1e-6
    # 1-D, both pad orders (H = n, W = 1)
    n = 41
    x1d = randn((b, n, 4 * 32), 3)
    v1 = (randn((b * 4, 10, 1), 4)).clamp(-1.9, 1.9)
    f1 = x1d.view(b, n, 4, 32).permute(0, 2, 3, 1).reshape(b * 4, 32, n, 1)
    for mode in (0, 1):
        z = torch.zeros_like(v1[..., 0])
        want, (x0, y0) = oracle.deform.bilinear_sample(f1, v1[..., 0] if mode == 0 else z, z if mode == 0 else v1[..., 0])
        kv, taps = gather_with_taps(x1d.cuda(), v1.cuda(), 4, n, 1, mode=mode)
        assert torch.equal(taps[..., 0].cpu().long(), x0) and torch.equal(taps[..., 1].cpu().long(), y0)
        assert rel_err(kv.view(b, 10, 4, 32).permute(0, 2, 3, 1).reshape(b * 4, 32, 10), want) < 1e-6


def test_deform2d_golden(prec):
    import sml_b200
    to, ti, tp = TOLS[prec]
    g = load_golden("deform2d")
    mod = sml_b200.DeformCrossAttention2D(dim=128, dim_head=64, heads=8, dropout=0.1, downsample_factor=4, offset_scale=4,
                                          offset_groups=8, offset_kernel_size=6).cuda().eval()
    mod.load_state_dict(g["state"])
    x1 = randn((1, 128, 2500), g["x1_seed"]).cuda().requires_grad_(); x2 = randn((1, 128, 2500), g["x2_seed"]).cuda().requires_grad_()
    y, vg = mod(x1, x2, return_vgrid=True)
    assert rel_err(vg, g["vgrid"]) < 1e-5
    assert rel_err(y, g["y"]) < to
    y.backward(randn(y.shape, g["gy_seed"]).cuda())
    assert rel_err(x1.grad[:, :, ::7], g["gx1"]) < ti
    assert rel_err(x2.grad[:, :, ::7], g["gx2"]) < ti
    check_param_grads({k: p.grad for k, p in mod.named_parameters()}, g["gparams"], prec)


def test_deform2d_vs_oracle_64x64_batch2(prec):
    import sml_b200
    to, ti, tp = TOLS[prec]
    torch.manual_seed(3)
    mod = sml_b200.DeformCrossAttention2D(dim=128).cuda().eval()
    sd = mod.state_dict()
    sd["to_offsets.2.weight"] = sd["to_offsets.2.weight"] * 8.0
    mod.load_state_dict(sd)
    p = {k: v.detach().cpu().double().requires_grad_() for k, v in mod.state_dict().items()}      # fp64 oracle (see above)
    x1c = randn((2, 128, 4096), 11).double().requires_grad_(); x2c = randn((2, 128, 4096), 12).double().requires_grad_()
    gy = randn((2, 128, 4096), 13); gv = randn((16, 2, 16, 16), 14)
    yo, vo = oracle.deform.deform_attn_2d(x1c, x2c, p, side=64, q_chunk=512)
    (yo * gy.double()).sum().add((vo * gv.double()).sum()).backward()
    x1 = x1c.detach().float().cuda().requires_grad_(); x2 = x2c.detach().float().cuda().requires_grad_()
    y, v = mod(x1, x2, return_vgrid=True)
    assert rel_err(y, yo) < to and rel_err(v, vo) < 1e-5
    (y * gy.cuda()).sum().add((v * gv.cuda()).sum()).backward()
    assert rel_err(x1.grad, x1c.grad) < ti and rel_err(x2.grad, x2c.grad) < ti
    check_param_grads({k: q.grad for k, q in mod.named_parameters()}, {k: t.grad for k, t in p.items()}, prec)


@pytest.mark.parametrize("ad", [2, 1])
def test_deform_cross_transmil_golden(ad, prec):
    import sml_b200
    to, ti, tp = TOLS[prec]
    g = load_golden(f"deformtransmil_ad{ad}")
    net = sml_b200.DeformCrossTransMIL(Namespace(path_dim=128, attn_dim=ad, return_vgrid=(ad == 2))).cuda().eval()
    st = synth_state_dict(g["shapes"], g["seed"])
    for k in st:
        if k.endswith("to_offsets.2.weight"):
            st[k] = st[k] * g["offw_scale"]
    net.load_state_dict(st)
    path = randn((1, 2500, 1024), 83).cuda().requires_grad_(); omic = randn((1, 128), 84).abs().cuda().requires_grad_()
    outs = net(path, omic)
    enc, logits = outs[0], outs[1]
    assert rel_err(logits, g["logits"]) < to and rel_err(enc, g["enc"]) < to
    loss = (logits * randn(logits.shape, 85).cuda()).sum() + (enc * randn(enc.shape, 86).cuda()).sum()
    if ad == 2:
        assert rel_err(outs[4], g["vgrid"]) < 1e-4                 # offsets do not depend on the attention arithmetic
        assert outs[3].shape == (1, 2500, 128) and torch.equal(outs[3][0, 17], omic.detach()[0])
        loss = loss + (outs[4] * randn(outs[4].shape, 87).cuda()).sum() * 0.01
    loss.backward()
    assert rel_err(omic.grad, g["gomic"]) < ti
    assert rel_err(path.grad[:, ::50, ::16], g["gpath_sub"]) < ti
    gp = {k: p.grad for k, p in net.named_parameters()}
    check_param_grads(gp, {k: v for k, v in g["gparams"].items() if gp[k] is not None}, prec, scale=2.0)
    for k, nv in g["gnorms"].items():
        if gp[k] is None:
            assert nv == 0.0 or "attn1d" in k or "attn2d" in k or k in ("cls_token",) or "pooler" in k, k
            continue
        tk = 2 * (tp if "rel_pos_bias" in k else ti)
        floor = 1e-5 + 1e-3 * tk * max(g["gnorms"].values())     # mathematically-zero gradients (last CPB bias) are noise
        assert abs(float(gp[k].norm()) - nv) <= tk * nv + floor, (k, float(gp[k].norm()), nv)


def test_deform_attention_dropout_is_unbiased():
    import sml_b200
    from sml_b200 import functional as Fn
    torch.manual_seed(5)
    mod = sml_b200.DeformCrossAttention2D(dim=128, dropout=0.1).cuda()
    x1 = randn((1, 128, 1024), 21).cuda(); x2 = randn((1, 128, 1024), 22).cuda()
    mod.eval(); y0 = mod(x1, x2)
    mod.train(); Fn.manual_seed(3)
    acc = torch.zeros_like(y0)
    for _ in range(16):
        acc += mod(x1, x2).detach()
    assert rel_err(acc / 16, y0) < 0.1                   # dropout(p) / (1-p) keeps the expectation
    assert rel_err(mod(x1, x2), y0) > 1e-3               # and it is really applied in training mode


def test_deform_pathomic_net_golden(prec):
    """DeformPathomicNet (models/model.py:440-544): two MaxNet encoders + two DeformCrossTransMIL branches on the same bag +
    concat + three heads, against the unmodified reference (tests/golden/make_golden_pathomic.py), batch 2, N = 2500."""
    import sml_b200
    from make_golden_pathomic import pathomic_args
    to, ti, tp = TOLS[prec]
    g = load_golden("pathomic")
    net = sml_b200.DeformPathomicNet(pathomic_args()).cuda().eval()
    st = synth_state_dict(g["shapes"], g["seed"])
    for k in st:
        if k.endswith("to_offsets.2.weight"):
            st[k] = st[k] * g["offw_scale"]
        if k.endswith("output_range"):
            st[k] = torch.tensor([6.0])
        if k.endswith("output_shift"):
            st[k] = torch.tensor([-3.0])
    net.load_state_dict(st)                                   # reference state_dict keys load unchanged
    x_path = randn((2, 2500, 1024), 112).cuda()
    xt = randn((2, 59), 113).abs().cuda().requires_grad_(); xi = randn((2, 361), 114).abs().cuda().requires_grad_()
    feats, vt, vi, logits, fg, _, _ = net(x_path=x_path, x_omic_tumor=xt, x_omic_immune=xi)
    assert fg is None and len(logits) == 7
    assert rel_err(feats, g["feats"]) < to and rel_err(vt, g["vt"]) < to and rel_err(vi, g["vi"]) < to
    # head logits: [2, 4] tensors after LayerNorm + pooler + two stacked linear heads; the hi/lo split measures 1.1e-4 here
    # against the fp32-CPU reference (exact-fp32 kernels: 2e-5), so the composite model gets 2x the single-module bar
    for a, b in zip(logits[:3], g["logits"]):
        assert rel_err(a, b) < 2 * to
    assert rel_err(logits[4], g["vgrid_tumor"]) < 1e-4 and rel_err(logits[6], g["vgrid_immune"]) < 1e-4
    assert rel_err(logits[3][:, 5], g["omic_tumor_row"]) < 1e-5
    loss = sum((l * randn(l.shape, 120 + j).cuda()).sum() for j, l in enumerate(logits[:3])) + (feats * randn(feats.shape, 125).cuda()).sum()
    loss = loss + 0.01 * ((logits[4] * randn(logits[4].shape, 126).cuda()).sum() + (logits[6] * randn(logits[6].shape, 127).cuda()).sum())
    loss.backward()
    assert rel_err(xt.grad, g["gxt"]) < 2 * ti and rel_err(xi.grad, g["gxi"]) < 2 * ti          # composite model: 2x (measured 1.1e-3)
    gp = {k: p.grad for k, p in net.named_parameters()}
    # 3x the single-module bar: the _fc1 bias / weight gradients sit behind the LayerNorm backward's ~10x cancellation (see
    # tests/test_gpu_fullsize.py) and sum two slides here; hi/lo split measures 2.4e-3 on _fc1.0.bias, exact fp32 4e-4
    check_param_grads(gp, {k: v for k, v in g["gparams"].items() if gp[k] is not None}, prec, scale=3.0)
    for k, nv in g["gnorms"].items():
        if gp[k] is None:
            continue
        tk = 3 * (tp if ("rel_pos_bias" in k or k.endswith("_fc1.0.bias")) else ti)
        floor = 1e-5 + 1e-3 * tk * max(g["gnorms"].values())
        assert abs(float(gp[k].norm()) - nv) <= tk * nv + floor, (k, float(gp[k].norm()), nv)


def test_maxnet_alpha_dropout_statistics():
    """AlphaDropout keeps mean and variance of a standard-normal input (torch.nn.AlphaDropout contract) and its backward
    regenerates the same mask."""
    from sml_b200 import functional as Fn
    Fn.manual_seed(11)
    x = torch.randn(1 << 20, device="cuda", requires_grad=True)
    y = Fn.alpha_dropout(x, 0.1, True)
    assert abs(float(y.detach().mean())) < 5e-3 and abs(float(y.detach().var()) - 1.0) < 1e-2
    y.sum().backward()
    kept = x.grad != 0
    assert abs(float(kept.float().mean()) - 0.9) < 2e-3
    a = float(x.grad[kept][0])
    assert torch.allclose(y.detach()[kept], a * x.detach()[kept] + (y.detach()[kept][0] - a * x.detach()[kept][0]), atol=1e-5)
    assert Fn.alpha_dropout(x, 0.1, False) is x


def test_gather_backward_sorted_scatter_is_deterministic_and_matches_autograd():
    """d x2 of the bilinear gather: sorted, atomic-free scatter (2-D, 64x64 lattice, n_kv = 256) against torch autograd of
    F.grid_sample on the same coordinates, bitwise identical across repeated launches (no atomics)."""
    from sml_b200._lib import c_i32, check, lib, ptr, stream
    b, G, H, W, cg, nkv = 2, 8, 64, 64, 16, 256
    torch.manual_seed(9)
    x2 = torch.randn(b, H * W, G * cg, device="cuda")
    vs = (torch.randn(b * G, nkv, 2, device="cuda") * 0.6).clamp(-1.3, 1.3)
    dkv = torch.randn(b, nkv, G * cg, device="cuda")
    cfg = (c_i32 * 8)(b, G, cg, H, W, nkv, 2, 0)
    outs = []
    for _ in range(3):
        dx2 = torch.zeros_like(x2); dvs = torch.empty_like(vs)
        check(lib.sml_deform_gather_bwd(cfg, ptr(x2), ptr(vs), ptr(dkv), ptr(dx2), None, ptr(dvs), stream()), "gather_bwd")
        outs.append((dx2.clone(), dvs.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][0], outs[2][0])
    feat = x2.view(b, H, W, G, cg).permute(0, 3, 4, 1, 2).reshape(b * G, cg, H, W).clone().requires_grad_()
    grid = vs.view(b * G, nkv, 1, 2).clone().requires_grad_()
    samp = torch.nn.functional.grid_sample(feat, grid, mode="bilinear", padding_mode="zeros", align_corners=False)   # [bG, cg, nkv, 1]
    gout = dkv.view(b, nkv, G, cg).permute(0, 2, 3, 1).reshape(b * G, cg, nkv, 1)
    samp.backward(gout)
    want_dx2 = feat.grad.view(b, G, cg, H, W).permute(0, 3, 4, 1, 2).reshape(b, H * W, G * cg)
    assert rel_err(outs[0][0], want_dx2) < 1e-5
    assert rel_err(outs[0][1], grid.grad.view(b * G, nkv, 2)) < 1e-4


@pytest.mark.parametrize("seed,OUT", [(0, 2), (1, 2), (2, 1)])
def test_cpb1d_table_matches_mlp(seed, OUT):
    """The 1-D bias MLP as an exact piecewise-linear table (csrc/cpb_table.cu; specification oracle/pwl.py): same number of
    segments as the fp64 oracle table and the MLP's values to fp32 rounding on a million random signed-log distances."""
    from sml_b200._lib import c_i32, c_i64, check, lib, ptr, stream
    torch.manual_seed(seed)
    p = {"mlp.0.0.weight": torch.randn(32, 1) * 1.5, "mlp.0.0.bias": torch.randn(32) * 0.7, "mlp.1.0.weight": torch.randn(32, 32) * 0.4,
         "mlp.1.0.bias": torch.randn(32) * 0.3, "mlp.2.weight": torch.randn(OUT, 32) * 0.5, "mlp.2.bias": torch.randn(OUT) * 0.1}
    cpb = torch.cat([p[k].reshape(-1) for k in ("mlp.0.0.weight", "mlp.0.0.bias", "mlp.1.0.weight", "mlp.1.0.bias", "mlp.2.weight",
                                                "mlp.2.bias")]).cuda()
    slots = lib.sml_cpb1d_table_slots()
    brk = torch.empty(slots, device="cuda"); coef = torch.empty(slots + 1, OUT, 2, device="cuda")
    nseg = torch.zeros(1, device="cuda", dtype=torch.int32)
    check(lib.sml_cpb1d_table_build(ptr(cpb), c_i32(OUT), ptr(brk), ptr(coef), ptr(nseg), stream()), "table_build")
    dx = (torch.rand(1 << 20, device="cuda") * 2 - 1) * 2.4
    t = torch.sign(dx) * torch.log(dx.abs() + 1)
    out = torch.empty(t.numel(), OUT, device="cuda")
    check(lib.sml_cpb1d_table_eval(ptr(t), c_i64(t.numel()), ptr(brk), ptr(coef), ptr(nseg), c_i32(OUT), ptr(out), stream()), "table_eval")
    want = oracle.deform.cpb_mlp(t.double().cpu().unsqueeze(1), {k: v.double() for k, v in p.items()}, "")
    assert rel_err(out, want) < 1e-5
    ns = int(nseg.item())
    assert ns == oracle.pwl.build_table(p)[0].numel() + 1
    assert bool((brk[:ns - 1][1:] >= brk[:ns - 1][:-1]).all())


@pytest.mark.parametrize("shape", [(1, 2500, 144, 50), (2, 16384, 1024, 128), (1, 1024, 49, 32)])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_deform_fwd_tcgen05_matches_warp_mma(shape, p_drop):
    """the tcgen05 / TMEM forward (csrc/deform_tc5.cu) against the warp-level kernel it replaces in bf16 mode: same inputs, same
    dropout hash -> outputs and log-sum-exp within bf16 rounding of each other, and against the fp32-grade split kernel at 2e-2"""
    import sml_b200
    from sml_b200._lib import c_f32, c_i32, c_u64, check, lib, ptr, stream
    b, T, nkv, side = shape
    heads = G = 8
    torch.manual_seed(3)
    q = torch.randn(b, T, heads * 64, device="cuda") * 0.3
    k = torch.randn(b, nkv, heads * 64, device="cuda")
    v = torch.randn(b, nkv, heads * 64, device="cuda")
    vs = torch.rand(b * G, nkv, 2, device="cuda") * 2.4 - 1.2
    cpb = torch.randn(lib.sml_deform_cpb_nparams(c_i32(2), c_i32(1)), device="cuda") * 0.4
    cfg = (c_i32 * 8)(b, heads, T, nkv, G, 2, side, 64)
    outs = {}
    for name in ("tc5", "mma", "split"):
        out = torch.empty(b, T, heads * 64, device="cuda"); lse = torch.empty(b, heads, T, device="cuda")
        if name == "tc5":
            check(lib.sml_deform_attn_fwd_tc5(cfg, c_f32(p_drop), c_u64(1234), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse),
                                              stream()), "deform_attn_fwd_tc5")
        else:
            check(lib.sml_deform_attn_fwd_tc(cfg, c_f32(p_drop), c_u64(1234), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse),
                                             c_i32(1 if name == "split" else 0), stream()), "deform_attn_fwd_tc")
        torch.cuda.synchronize()
        outs[name] = (out, lse)
    assert torch.isfinite(outs["tc5"][0]).all() and torch.isfinite(outs["tc5"][1]).all()
    assert rel_err(outs["tc5"][1], outs["split"][1]) < 5e-3, rel_err(outs["tc5"][1], outs["split"][1])          # log-sum-exp
    assert rel_err(outs["tc5"][0], outs["split"][0]) < 2e-2, rel_err(outs["tc5"][0], outs["split"][0])
    # no worse than the kernel it replaces (both single-pass bf16 inside), with slack for different rounding points
    assert rel_err(outs["tc5"][0], outs["split"][0]) < 2.0 * rel_err(outs["mma"][0], outs["split"][0]) + 2e-3


@pytest.mark.parametrize("shape", [(1, 2500, 144, 50, 2), (2, 16384, 1024, 128, 2), (1, 1024, 49, 32, 2), (2, 2501, 625, 0, 1), (1, 777, 130, 0, 1)])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_deform_fwd_lean_matches_split(shape, p_drop):
    """the lean bf16 forward (table-driven distances, layer 3 on the MMA pipe) against the fp32-grade split kernel and the round-1
    bf16 kernel: same inputs, same dropout hash; 1-D (two heads per offset group) and 2-D, ragged tiles"""
    import sml_b200
    from sml_b200._lib import c_f32, c_i32, c_u64, check, lib, ptr, stream
    b, T, nkv, side, cd = shape
    heads = 8
    G = 8 if cd == 2 else 4
    torch.manual_seed(5)
    q = torch.randn(b, T, heads * 64, device="cuda") * 0.3
    k = torch.randn(b, nkv, heads * 64, device="cuda")
    v = torch.randn(b, nkv, heads * 64, device="cuda")
    vs = torch.rand(b * G, nkv, cd, device="cuda") * 2.4 - 1.2
    cpb = torch.randn(lib.sml_deform_cpb_nparams(c_i32(cd), c_i32(heads // G)), device="cuda") * 0.4
    cfg = (c_i32 * 8)(b, heads, T, nkv, G, cd, side, 64)
    outs = {}
    for name, split in (("lean", 0), ("split", 1), ("old", 2)):
        out = torch.empty(b, T, heads * 64, device="cuda"); lse = torch.empty(b, heads, T, device="cuda")
        check(lib.sml_deform_attn_fwd_tc(cfg, c_f32(p_drop), c_u64(99), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse),
                                         c_i32(split), stream()), "deform_attn_fwd_tc")
        torch.cuda.synchronize()
        outs[name] = (out, lse)
    assert torch.isfinite(outs["lean"][0]).all() and torch.isfinite(outs["lean"][1]).all()
    e_lse, e_out = rel_err(outs["lean"][1], outs["split"][1]), rel_err(outs["lean"][0], outs["split"][0])
    o_lse, o_out = rel_err(outs["old"][1], outs["split"][1]), rel_err(outs["old"][0], outs["split"][0])
    assert e_lse < 5e-3 and e_out < 2e-2, (e_lse, e_out)
    assert e_out < 2.0 * o_out + 2e-3 and e_lse < 2.0 * o_lse + 1e-3, (e_lse, o_lse, e_out, o_out)


@pytest.mark.parametrize("case", ["1d_16k_upstream", "1d_16k_reference", "2d_256x256", "2d_50x50"])
def test_gather_sorted_scatter_large(case):
    """the sorted, atomic-free d x2 scatter beyond n_kv = 1024: the 1-D path at 16 385 tokens (n_kv = 4096, both pad orders) and a
    256 x 256 lattice (n_kv = 4096, 16 384 sorted entries): bitwise identical across launches, equal to autograd of
    F.grid_sample; the vectorised forward gather against F.grid_sample as well"""
    from sml_b200._lib import c_i32, check, lib, ptr, stream
    if case.startswith("1d"):
        b, G, H, W, cg, nkv, CD, mode = 2, 4, 16385, 1, 32, 4096, 1, (1 if case.endswith("upstream") else 0)
    elif case == "2d_256x256":
        b, G, H, W, cg, nkv, CD, mode = 1, 8, 256, 256, 16, 4096, 2, 0
    else:
        b, G, H, W, cg, nkv, CD, mode = 2, 8, 50, 50, 16, 144, 2, 0
    torch.manual_seed(10)
    x2 = torch.randn(b, H * W, G * cg, device="cuda")
    vs = (torch.randn(b * G, nkv, CD, device="cuda") * 0.6).clamp(-1.3, 1.3)
    dkv = torch.randn(b, nkv, G * cg, device="cuda")
    cfg = (c_i32 * 8)(b, G, cg, H, W, nkv, CD, mode)
    kv = torch.empty(b, nkv, G * cg, device="cuda")
    check(lib.sml_deform_gather_fwd(cfg, ptr(x2), ptr(vs), ptr(kv), None, stream()), "gather_fwd")
    outs = []
    for _ in range(3):
        dx2 = torch.zeros_like(x2); dvs = torch.empty_like(vs)
        check(lib.sml_deform_gather_bwd(cfg, ptr(x2), ptr(vs), ptr(dkv), ptr(dx2), None, ptr(dvs), stream()), "gather_bwd")
        outs.append((dx2.clone(), dvs.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][0], outs[2][0])
    # F.grid_sample on the CPU.  For the 16 385-row 1-D image ATen itself is not unique: its scalar path and the CUDA kernel
    # unnormalise as ((c + 1) * size - 1) / 2 (this repo's order, bit-exact taps in test_gather_taps_bit_exact), the vectorised CPU
    # path as (c + 1) * (size / 2) - 0.5, one ulp of the pixel coordinate (1e-3 px) apart -> 1e-3 on interpolation weights
    feat = x2.cpu().view(b, H, W, G, cg).permute(0, 3, 4, 1, 2).reshape(b * G, cg, H, W).clone().requires_grad_()
    if CD == 2:
        gxy = vs.cpu().view(b * G, nkv, 1, 2).clone().requires_grad_()
        grid = gxy
    else:
        gxy = vs.cpu().view(b * G, nkv, 1, 1).clone().requires_grad_()
        z = torch.zeros_like(gxy)
        grid = torch.cat([gxy, z], -1) if mode == 0 else torch.cat([z, gxy], -1)
    samp = torch.nn.functional.grid_sample(feat, grid, mode="bilinear", padding_mode="zeros", align_corners=False)   # [bG, cg, nkv, 1]
    want_kv = samp.detach().view(b, G, cg, nkv).permute(0, 3, 1, 2).reshape(b, nkv, G * cg)
    tol = 2e-3 if case == "1d_16k_upstream" else 1e-5
    assert rel_err(kv, want_kv) < tol
    gout = dkv.cpu().view(b, nkv, G, cg).permute(0, 2, 3, 1).reshape(b * G, cg, nkv, 1)
    samp.backward(gout)
    want_dx2 = feat.grad.view(b, G, cg, H, W).permute(0, 3, 4, 1, 2).reshape(b, H * W, G * cg)
    assert rel_err(outs[0][0], want_dx2) < tol
    assert rel_err(outs[0][1], gxy.grad.view(b * G, nkv, CD)) < max(tol, 1e-4)


@pytest.mark.parametrize("shape", [(1, 2500, 144, 50, 2), (1, 16384, 1024, 128, 2), (2, 2501, 625, 0, 1)])
def test_deform_bwd_warp_specialised_matches_lean(shape):
    """the producer / consumer backward (split code 5) against the lean backward (split code 0): same inputs, same dropout hash,
    every output (d q, d k, d v, key-coordinate gradients, summed bias-MLP partials)"""
    from sml_b200._lib import c_f32, c_i32, c_u64, check, lib, ptr, stream
    b, T, nkv, side, cd = shape
    heads = 8
    G = 8 if cd == 2 else 4
    torch.manual_seed(6)
    q = torch.randn(b, T, heads * 64, device="cuda") * 0.3
    k = torch.randn(b, nkv, heads * 64, device="cuda"); v = torch.randn(b, nkv, heads * 64, device="cuda")
    vs = torch.rand(b * G, nkv, cd, device="cuda") * 2.4 - 1.2
    npar = lib.sml_deform_cpb_nparams(c_i32(cd), c_i32(heads // G))
    cpb = torch.randn(npar, device="cuda") * 0.4
    cfg = (c_i32 * 8)(b, heads, T, nkv, G, cd, side, 64)
    out = torch.empty(b, T, heads * 64, device="cuda"); lse = torch.empty(b, heads, T, device="cuda")
    check(lib.sml_deform_attn_fwd_tc(cfg, c_f32(0.1), c_u64(7), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse), c_i32(0), stream()))
    dout = torch.randn_like(out)
    res = {}
    for code in (0, 5):
        dq = torch.empty_like(q); dk = torch.zeros_like(k); dv = torch.zeros_like(v); dvs = torch.zeros_like(vs)
        part = torch.empty(lib.sml_deform_attn_bwd_blocks(cfg), npar, device="cuda")
        check(lib.sml_deform_attn_bwd_tc(cfg, c_f32(0.1), c_u64(7), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse), ptr(dout),
                                         ptr(dq), ptr(dk), ptr(dv), ptr(dvs), ptr(part), c_i32(code), stream()))
        torch.cuda.synchronize()
        res[code] = (dq, dk, dv, dvs, part.sum(0))
    for a, bb, tol in zip(res[0], res[5], (1e-5, 1e-4, 1e-4, 1e-3, 2e-3)):
        assert rel_err(bb, a) < tol, rel_err(bb, a)


def _cpb1d_params(npar, dense):
    """random bias-MLP parameters; ``dense`` clusters the 32 layer-1 kinks inside |t| < 0.1, so that the range of one
    (query row, key tile) holds several breakpoints and the table kernels take their per-pair search path"""
    cpb = torch.randn(npar, device="cuda") * 0.4
    if dense:
        w1 = torch.randn(32, device="cuda") * 5.0
        cpb[:32] = w1
        cpb[32:64] = -w1 * torch.randn(32, device="cuda") * 0.03
    return cpb


@pytest.mark.parametrize("shape", [(2, 2501, 625), (1, 16385, 4096), (1, 777, 130)])
@pytest.mark.parametrize("p_drop,dense", [(0.0, False), (0.1, False), (0.0, True)])
def test_deform1d_fwd_pwl_matches_split(shape, p_drop, dense):
    """1-D forward with the bias MLP tabulated as a piecewise-linear function (sml_deform_attn_fwd_pwl) against the fp32-grade
    split kernel that evaluates the MLP per pair: same inputs, same dropout hash"""
    from sml_b200._lib import c_f32, c_i32, c_u64, check, lib, ptr, stream
    from sml_b200.deform import cpb1d_table, key_ranges
    b, T, nkv = shape
    heads, G = 8, 4
    torch.manual_seed(8)
    q = torch.randn(b, T, heads * 64, device="cuda") * 0.3
    k = torch.randn(b, nkv, heads * 64, device="cuda"); v = torch.randn(b, nkv, heads * 64, device="cuda")
    vs = (torch.linspace(-1, 1, nkv, device="cuda").repeat(b * G, 1) + torch.randn(b * G, nkv, device="cuda") * 2e-3).unsqueeze(-1).contiguous()
    cpb = _cpb1d_params(lib.sml_deform_cpb_nparams(c_i32(1), c_i32(2)), dense)
    cfg = (c_i32 * 8)(b, heads, T, nkv, G, 1, 0, 64)
    outs = {}
    for name in ("pwl", "split", "lean"):
        out = torch.empty(b, T, heads * 64, device="cuda"); lse = torch.empty(b, heads, T, device="cuda")
        if name == "pwl":
            brk, coef, nseg = cpb1d_table(cpb, 2)
            kh, vh, kr = k.bfloat16(), v.bfloat16(), key_ranges(cfg, vs)
            check(lib.sml_deform_attn_fwd_pwl(cfg, c_f32(p_drop), c_u64(5), ptr(q), ptr(kh), ptr(vh), ptr(vs), ptr(kr), ptr(brk), ptr(coef),
                                              ptr(nseg), ptr(out), ptr(lse), stream()), "fwd_pwl")
        else:
            check(lib.sml_deform_attn_fwd_tc(cfg, c_f32(p_drop), c_u64(5), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse),
                                             c_i32(1 if name == "split" else 0), stream()), "fwd_tc")
        torch.cuda.synchronize()
        outs[name] = (out, lse)
    e_lse, e_out = rel_err(outs["pwl"][1], outs["split"][1]), rel_err(outs["pwl"][0], outs["split"][0])
    o_lse, o_out = rel_err(outs["lean"][1], outs["split"][1]), rel_err(outs["lean"][0], outs["split"][0])
    assert e_lse < 5e-3 and e_out < 2e-2, (e_lse, e_out)
    assert e_out < 2.0 * o_out + 2e-3 and e_lse < 2.0 * o_lse + 1e-3, (e_lse, o_lse, e_out, o_out)


@pytest.mark.parametrize("shape", [(2, 2501, 625), (1, 16385, 4096), (1, 777, 130)])
@pytest.mark.parametrize("p_drop,dense", [(0.0, False), (0.1, False), (0.0, True)])
def test_deform1d_bwd_pwl_matches_split(shape, p_drop, dense):
    """1-D backward with the tabulated bias (sml_deform_attn_bwd_pwl + sml_cpb1d_table_bwd: per-segment sums mapped to the MLP
    parameters) against the fp32-grade split kernel that differentiates the MLP per pair; the plain-bf16 per-pair kernel on the
    same inputs sets the scale of the allowed error."""
    from sml_b200._lib import c_f32, c_i32, c_u64, check, lib, ptr, stream
    from sml_b200.deform import cpb1d_table, key_ranges
    b, T, nkv = shape
    heads, G = 8, 4
    torch.manual_seed(9)
    q = torch.randn(b, T, heads * 64, device="cuda") * 0.3
    k = torch.randn(b, nkv, heads * 64, device="cuda"); v = torch.randn(b, nkv, heads * 64, device="cuda")
    vs = (torch.linspace(-1, 1, nkv, device="cuda").repeat(b * G, 1) + torch.randn(b * G, nkv, device="cuda") * 2e-3).unsqueeze(-1).contiguous()
    npar = lib.sml_deform_cpb_nparams(c_i32(1), c_i32(2))
    cpb = _cpb1d_params(npar, dense)
    cfg = (c_i32 * 8)(b, heads, T, nkv, G, 1, 0, 64)
    dout = torch.randn(b, T, heads * 64, device="cuda")
    res = {}
    for name in ("pwl", "split", "lean"):
        out = torch.empty(b, T, heads * 64, device="cuda"); lse = torch.empty(b, heads, T, device="cuda")
        dq = torch.empty_like(q); dk = torch.zeros_like(k); dv = torch.zeros_like(v); dvs = torch.zeros_like(vs)
        if name == "pwl":
            brk, coef, nseg = cpb1d_table(cpb, 2)
            kh, vh, kr = k.bfloat16(), v.bfloat16(), key_ranges(cfg, vs)
            check(lib.sml_deform_attn_fwd_pwl(cfg, c_f32(p_drop), c_u64(5), ptr(q), ptr(kh), ptr(vh), ptr(vs), ptr(kr), ptr(brk), ptr(coef),
                                              ptr(nseg), ptr(out), ptr(lse), stream()), "fwd_pwl")
            gseg = torch.zeros(2, lib.sml_cpb1d_seg_rows(), 2, device="cuda")
            dcpb = torch.empty(npar, device="cuda")
            check(lib.sml_deform_attn_bwd_pwl(cfg, c_f32(p_drop), c_u64(5), ptr(q), ptr(kh), ptr(vh), ptr(vs), ptr(kr), ptr(brk), ptr(coef), ptr(nseg),
                                              ptr(out), ptr(lse), ptr(dout), ptr(dq), ptr(dk), ptr(dv), ptr(dvs), ptr(gseg), stream()), "bwd_pwl")
            check(lib.sml_cpb1d_table_bwd(ptr(cpb), c_i32(2), ptr(brk), ptr(nseg), ptr(gseg), ptr(dcpb), stream()), "table_bwd")
        else:
            split = c_i32(1 if name == "split" else 0)
            check(lib.sml_deform_attn_fwd_tc(cfg, c_f32(p_drop), c_u64(5), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse),
                                             split, stream()), "fwd_tc")
            part = torch.empty(lib.sml_deform_attn_bwd_blocks(cfg), npar, device="cuda")
            check(lib.sml_deform_attn_bwd_tc(cfg, c_f32(p_drop), c_u64(5), ptr(q), ptr(k), ptr(v), ptr(vs), ptr(cpb), ptr(out), ptr(lse),
                                             ptr(dout), ptr(dq), ptr(dk), ptr(dv), ptr(dvs), ptr(part), split, stream()), "bwd_tc")
            dcpb = part.sum(0)
        torch.cuda.synchronize()
        res[name] = (dq, dk, dv, dvs, dcpb)
    for i, (nm, floor) in enumerate((("dq", 2e-2), ("dk", 2e-2), ("dv", 2e-2), ("dvs", 3e-2), ("dcpb", 3e-2))):
        e_pwl, e_lean = rel_err(res["pwl"][i], res["split"][i]), rel_err(res["lean"][i], res["split"][i])
        assert e_pwl < max(2.0 * e_lean, floor), (nm, e_pwl, e_lean)

"""GPU parity: Nystrom attention / TransLayer / PPEG / TransMIL on libsml_b200 vs the CPU oracle and the
golden fixtures made from the unmodified reference.  Tolerances are BASELINE.json's: 1e-4 relative in the
fp32 modes ("fp32" SIMT and "bf16x3" tensor-core split), 2e-2 in bf16; landmark ids bit-exact."""
import pytest
import torch

import oracle
from conftest import grads_close, load_golden, rel_err
from synth import randn, synth_state_dict

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16x3": 1e-4, "bf16": 2e-2}
MODES = ["fp32", "bf16x3", "bf16"]


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("mode", MODES)
def test_pinv_golden(mode):
    import sml_b200
    from sml_b200.nystrom import pinv_fwd, pinv_bwd
    g = load_golden("pinv")
    x = g["x"].cuda()
    z, saved = pinv_fwd(x, g["iters"], mode)
    assert rel_err(z.f, g["z"]) < 1e-4
    gx = pinv_bwd(x, saved, g["gz"].cuda(), mode)
    # The init scale is max(row sums) * max(col sums) of a softmax output: every row sum is 1 up to an ulp, so
    # WHICH row is the arg-max (and receives the scalar's gradient as a row-constant shift) is decided by
    # rounding — the reference's own fp32 gradient differs from an fp64 evaluation by 5e-2 here.  On the path x
    # is always a softmax output and a row-constant shift of dx is annihilated by the softmax backward, so
    # parity is stated on the gradient w.r.t. the logits: ds = x * (dx - sum(dx * x)).
    def through_softmax(gxx):
        xx = g["x"].double()
        gxx = gxx.detach().double().cpu()
        return xx * (gxx - (gxx * xx).sum(-1, keepdim=True))
    assert rel_err(through_softmax(gx), through_softmax(g["gx"])) < 1e-3


@pytest.mark.parametrize("mode", MODES)
def test_nystrom_golden(mode):
    import sml_b200
    for c in load_golden("nystrom"):
        cfg = c["cfg"]
        mod = sml_b200.NystromAttention(dim=cfg["dim"], dim_head=cfg["dim_head"], heads=cfg["heads"],
                                        num_landmarks=cfg["num_landmarks"], pinv_iterations=cfg["pinv_iterations"],
                                        dropout=0.1).cuda().eval()
        mod.load_state_dict(c["state"])
        mod.precision = mode
        x = c["x"].cuda().requires_grad_()
        y = mod(x)
        assert rel_err(y, c["y"]) < TOL[mode], cfg
        y.backward(c["gy"].cuda())
        assert rel_err(x.grad, c["gx"]) < 10 * TOL[mode], cfg
        grads_close({k: p.grad for k, p in mod.named_parameters()}, c["gparams"], 10 * TOL[mode], what=str(cfg))
        n, m = x.shape[1], cfg["num_landmarks"]
        ids = mod.landmark_ids(n).cpu().long()
        npad = n + (m - n % m) % m
        assert torch.equal(ids, oracle.nystrom.landmark_ids(npad, m))        # bit-exact index contract


@pytest.mark.parametrize("mode", MODES)
def test_nystrom_vs_oracle_mid_size(mode):
    """dim 512 / 8 heads / d 64 / m 256 (the TransMIL configuration) on a 1 000-token bag, fwd + bwd."""
    import sml_b200
    torch.manual_seed(0)
    mod = sml_b200.NystromAttention(dim=512, dim_head=64, heads=8, num_landmarks=256, pinv_iterations=6).cuda().eval()
    mod.precision = mode
    # fp64 evaluation of the oracle (host-CPU independent; the fp32 reference path agrees with it to ~1e-6)
    p = {k: v.detach().cpu().double().requires_grad_() for k, v in mod.state_dict().items()}
    xc = randn((2, 1001, 512), 5).double().requires_grad_()
    gy = randn((2, 1001, 512), 6)
    yo = oracle.nystrom.nystrom_attention(xc, p, heads=8, num_landmarks=256, pinv_iterations=6)
    yo.backward(gy.double())
    x = xc.detach().float().cuda().requires_grad_()
    y = mod(x)
    assert rel_err(y, yo) < TOL[mode]
    y.backward(gy.cuda())
    assert rel_err(x.grad, xc.grad) < 10 * TOL[mode]
    grads_close({k: q.grad for k, q in mod.named_parameters()}, {k: v.grad for k, v in p.items()}, 10 * TOL[mode])


def test_pinv_chain_as_gemm_program():
    """the fused persistent-launch variant of the pinv chain (sml_gemm_tc_program) gives the same result"""
    import sml_b200
    from sml_b200 import nystrom as ny
    g = load_golden("pinv")
    x = g["x"].cuda()
    ny.FUSE_PINV_CHAIN = True
    try:
        z, saved = ny.pinv_fwd(x, g["iters"], "bf16x3")
        gx = ny.pinv_bwd(x, saved, g["gz"].cuda(), "bf16x3")
    finally:
        ny.FUSE_PINV_CHAIN = False
    z2, saved2 = ny.pinv_fwd(x, g["iters"], "bf16x3")
    gx2 = ny.pinv_bwd(x, saved2, g["gz"].cuda(), "bf16x3")
    assert rel_err(z.f, g["z"]) < 1e-4 and rel_err(z.f, z2.f) < 1e-4
    assert rel_err(gx, gx2) < 1e-3
    torch.cuda.synchronize()
    assert sml_b200._lib.lib.sml_tc_timeout_flag() == 0


def test_translayer_ppeg_golden():
    import sml_b200
    g = load_golden("translayer_ppeg")
    tl = sml_b200.TransLayer(dim=64).cuda().eval()
    tl.load_state_dict(g["tl_state"])
    x = g["x"].cuda().requires_grad_()
    y = tl(x)
    assert rel_err(y, g["y"]) < 1e-4
    y.backward(g["gy"].cuda())
    assert rel_err(x.grad, g["gx"]) < 1e-3
    grads_close({k: p.grad for k, p in tl.named_parameters()}, g["tl_gparams"], 1e-3)
    pp = sml_b200.PPEG(dim=64).cuda().eval()
    pp.load_state_dict(g["pp_state"])
    x = g["px"].cuda().requires_grad_()
    y = pp(x, 6, 6)
    assert rel_err(y, g["py"]) < 1e-5
    y.backward(g["pgy"].cuda())
    assert rel_err(x.grad, g["pgx"]) < 1e-5
    grads_close({k: p.grad for k, p in pp.named_parameters()}, g["pp_gparams"], 1e-4)


@pytest.mark.parametrize("mode", MODES)
def test_transmil_golden(mode):
    import sml_b200
    from argparse import Namespace
    g = load_golden("transmil")
    tm = sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).cuda().eval()
    tm.load_state_dict(synth_state_dict(g["shapes"], g["seed"]))
    tm.set_precision(mode)
    x = randn(g["x_shape"], g["x_seed"]).cuda()
    enc, logits, _ = tm(x)
    assert rel_err(logits, g["logits"]) < TOL[mode]
    assert rel_err(enc, g["enc"]) < TOL[mode]
    loss = (logits * randn(logits.shape, g["glogits_seed"]).cuda()).sum() + (enc * randn(enc.shape, g["genc_seed"]).cuda()).sum()
    loss.backward()
    gp = {k: p.grad for k, p in tm.named_parameters()}
    grads_close({k: gp[k] for k in g["gparams"]}, g["gparams"], 20 * TOL[mode])
    for k, nv in g["gnorms"].items():
        if gp[k] is None:
            continue
        assert abs(float(gp[k].norm()) - nv) <= 20 * TOL[mode] * nv + 1e-5, (k, float(gp[k].norm()), nv)


def test_dropout_statistics():
    import sml_b200
    from sml_b200 import functional as Fn
    Fn.manual_seed(1)
    u = torch.ones(1 << 20, device="cuda", requires_grad=True)
    y = Fn.dropout_add(None, u, 0.1, True)
    keep = (y != 0).float().mean().item()
    assert abs(keep - 0.9) < 2e-3
    assert abs(y.mean().item() - 1.0) < 5e-3
    y.sum().backward()
    assert torch.equal(u.grad != 0, y != 0)          # same mask regenerated in backward


def test_translayer_training_mode_skip_and_dropout():
    """TransLayer in training mode = x + dropout(attn(norm(x)), 0.1) as one autograd node: every element of (out - x) is
    either 0 or the eval-mode value / 0.9, the kept fraction is ~0.9, and the backward regenerates the same mask: with
    d out = 1 the gradient w.r.t. the attention bias (to_out.0.bias) is the per-column count of kept elements / 0.9."""
    import sml_b200
    from sml_b200 import functional as Fn
    torch.manual_seed(0)
    tl = sml_b200.TransLayer(dim=64).cuda()
    x = randn((2, 300, 64), 31).cuda()
    tl.eval()
    y_eval = tl(x).detach() - x
    tl.train(); Fn.manual_seed(77)
    xg = x.clone().requires_grad_()
    out = tl(xg)
    u = (out.detach() - x)
    kept = u != 0
    assert abs(float(kept.float().mean()) - 0.9) < 2e-2
    assert rel_err(u[kept], y_eval[kept] / 0.9) < 1e-4
    out.sum().backward()
    want_db = kept.float().sum(dim=(0, 1)) / 0.9
    assert rel_err(tl.attn.to_out[0].bias.grad, want_db) < 1e-4
    assert torch.isfinite(xg.grad).all() and float((xg.grad - 1).abs().max()) > 0          # skip gradient (1) + attention path


def test_return_attn_map():
    """NystromAttention(..., return_attn=True): second output = A1 A2^+ A3 over the padded tokens (NystromAttention.py:153-155)"""
    import sml_b200
    torch.manual_seed(4)
    mod = sml_b200.NystromAttention(dim=128, dim_head=16, heads=8, num_landmarks=32, pinv_iterations=6).cuda().eval()
    x = randn((2, 200, 128), 611)
    p = {k: v.detach().cpu().double() for k, v in mod.state_dict().items()}
    yo, parts = oracle.nystrom.nystrom_attention(x.double(), p, heads=8, num_landmarks=32, pinv_iterations=6, return_parts=True)
    want = (parts["a1"] @ parts["z"]) @ parts["a3"]
    y, attn = mod(x.cuda(), return_attn=True)
    assert attn.shape == (2, 8, 224, 224)
    assert rel_err(attn, want) < 1e-4 and rel_err(y, yo) < 1e-4
    with pytest.raises(NotImplementedError):
        mod(x.cuda(), mask=torch.ones(2, 200, dtype=torch.bool, device="cuda"))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 8, 1280, 3), (1, 2, 16640, 5), (1, 1, 256, 1)])
def test_nystrom_a3_fused_matches_torch(shape):
    """sml_nystrom_a3_fwd / _bwd (flash form of softmax(ql k^T) v over key chunks, NystromAttention.py:124, 130-137, 140) against
    fp32 torch on the same bf16-rounded operands: a3v, row log-sum-exp, dql, dk (+ landmark-mean term), dv."""
    from sml_b200._lib import c_i32, check, lib, ptr, stream
    from sml_b200.nystrom import _a3_cfg
    b, h, npad, nch = shape
    m, d = 256, 64
    torch.manual_seed(3)
    dev = "cuda"
    qkv = (torch.randn(b, npad, 3, h, d, device=dev) * 0.7).bfloat16()
    ql = (torch.randn(b, h, m, d, device=dev) * 0.5).bfloat16()
    cfg = _a3_cfg(b, h, m, d, npad, nch)
    a3v = torch.empty(b, h, m, d, device=dev); a3v_hi = torch.empty(b, h, m, d, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(b, h, m, device=dev)
    opart = torch.empty(b * h, nch, m, d, device=dev); ml = torch.empty(b * h, nch, m, 2, device=dev)
    check(lib.sml_nystrom_a3_fwd(cfg, ptr(ql), ptr(qkv), ptr(opart), ptr(ml), ptr(a3v), ptr(a3v_hi), ptr(lse), stream()), "a3_fwd")
    # reference
    qf = ql.float().requires_grad_()
    kf = qkv[:, :, 1].permute(0, 2, 1, 3).float().requires_grad_()
    vf = qkv[:, :, 2].permute(0, 2, 1, 3).float().requires_grad_()
    sref = qf @ kf.transpose(-1, -2)
    want = torch.softmax(sref, -1) @ vf
    assert rel_err(a3v, want.detach()) < 1e-2, rel_err(a3v, want.detach())
    assert rel_err(a3v_hi.float(), want.detach()) < 1.5e-2
    assert (lse - torch.logsumexp(sref.detach(), -1)).abs().max() < 2e-3
    # backward
    g = torch.randn(b, h, m, d, device=dev)
    want.backward(g)
    lsz = max(npad // m, 1)
    dkl = torch.randn(b, h, m, d, device=dev)
    dql = torch.full((b, h, m, d), 0.25, device=dev)                 # accumulated into
    dqkv_hi = torch.zeros(b, npad, 3, h, d, device=dev, dtype=torch.bfloat16)
    dqkv_f = torch.zeros(b, npad, 3, h, d, device=dev)
    nchb = min(nch + 1, npad // 32)
    cfgb = _a3_cfg(b, h, m, d, npad, nchb)
    check(lib.sml_nystrom_a3_bwd(cfgb, ptr(ql), ptr(qkv), ptr(g), ptr(a3v), ptr(lse), None, c_i32(lsz), ptr(dql), ptr(dqkv_hi), ptr(dqkv_f),
                                 stream()), "a3_bwd")
    assert rel_err(dql - 0.25, qf.grad) < 2e-2, rel_err(dql - 0.25, qf.grad)
    dk = dqkv_hi[:, :, 1].permute(0, 2, 1, 3).float()
    assert rel_err(dk, kf.grad) < 2e-2, rel_err(dk, kf.grad)
    dv = dqkv_f[:, :, 2].permute(0, 2, 1, 3)
    assert rel_err(dv, vf.grad) < 2e-2, rel_err(dv, vf.grad)
    assert float(dqkv_hi[:, :, 0].abs().max()) == 0 and float(dqkv_f[:, :, :2].abs().max()) == 0      # other slices untouched
    check(lib.sml_nystrom_a3_dk_landmarks(cfgb, ptr(dkl), c_i32(lsz), ptr(dqkv_hi), stream()), "a3_dk_landmarks")
    idx = torch.arange(npad, device=dev) // lsz
    want_dk = kf.grad + dkl[:, :, idx.clamp(max=m - 1)] / lsz
    dk2 = dqkv_hi[:, :, 1].permute(0, 2, 1, 3).float()
    assert rel_err(dk2, want_dk) < 2e-2, rel_err(dk2, want_dk)
    # the landmark term inside the backward kernel gives the same result
    dqkv_hi2 = torch.zeros_like(dqkv_hi); dql2 = torch.zeros_like(dql)
    check(lib.sml_nystrom_a3_bwd(cfgb, ptr(ql), ptr(qkv), ptr(g), ptr(a3v), ptr(lse), ptr(dkl), c_i32(lsz), ptr(dql2), ptr(dqkv_hi2),
                                 ptr(dqkv_f), stream()), "a3_bwd")
    assert rel_err(dqkv_hi2[:, :, 1].float(), want_dk.permute(0, 2, 1, 3)) < 2e-2


@pytest.mark.gpu
def test_nystrom_a3_forward_512_landmarks():
    """the fused forward with 512 landmark rows (BASELINE configs[4]: 100 k-patch inference) against fp32 torch"""
    from sml_b200._lib import check, lib, ptr, stream
    from sml_b200.nystrom import _a3_cfg
    b, h, npad, nch, m, d = 1, 3, 2560, 4, 512, 64
    torch.manual_seed(5)
    dev = "cuda"
    qkv = (torch.randn(b, npad, 3, h, d, device=dev) * 0.7).bfloat16()
    ql = (torch.randn(b, h, m, d, device=dev) * 0.5).bfloat16()
    a3v = torch.empty(b, h, m, d, device=dev); lse = torch.empty(b, h, m, device=dev)
    opart = torch.empty(b * h, nch, m, d, device=dev); ml = torch.empty(b * h, nch, m, 2, device=dev)
    check(lib.sml_nystrom_a3_fwd(_a3_cfg(b, h, m, d, npad, nch), ptr(ql), ptr(qkv), ptr(opart), ptr(ml), ptr(a3v), None, ptr(lse), stream()),
          "a3_fwd")
    s = ql.float() @ qkv[:, :, 1].permute(0, 2, 1, 3).float().transpose(-1, -2)
    want = torch.softmax(s, -1) @ qkv[:, :, 2].permute(0, 2, 1, 3).float()
    assert rel_err(a3v, want) < 1e-2, rel_err(a3v, want)
    assert (lse - torch.logsumexp(s, -1)).abs().max() < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("m", [256, 512])
def test_nystrom_inference_fused_matches_training_path(m):
    """under torch.no_grad() the bf16 mode runs both long-axis products as flash kernels (no A1 / A3 tensors); the result must
    agree with the grad-enabled forward of the same module (materialised A1; A3 fused only for 256 landmarks)"""
    import sml_b200
    torch.manual_seed(2)
    mod = sml_b200.NystromAttention(dim=512, dim_head=64, heads=8, num_landmarks=m, pinv_iterations=6).to("cuda").eval()
    mod.precision = "bf16"
    x = torch.randn(2, 3000, 512, device="cuda")
    with torch.no_grad():
        y_inf = mod(x)
    y_tr = mod(x.clone().requires_grad_())
    assert rel_err(y_inf, y_tr) < 2e-2, rel_err(y_inf, y_tr)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 13, 9, 96), (1, 50, 50, 512), (2, 17, 33, 512), (1, 128, 128, 512)])
def test_ppeg_stencil_matches_conv2d(shape):
    """the folded 7x7 PPEG stencil (mil.py:190-205) on ragged lattices and both kernel families (compile-time 512 channels,
    run-time channel count): forward, dx, dw, db against F.conv2d"""
    from sml_b200 import functional as Fn
    B, H, W, C = shape
    torch.manual_seed(4)
    x = torch.randn(B, 1 + H * W, C, device="cuda", requires_grad=True)
    w7 = (torch.randn(49, C, device="cuda") * 0.1).requires_grad_()
    bias = torch.randn(C, device="cuda", requires_grad=True)
    y = Fn.ppeg(x, w7, bias, H, W)
    g = torch.randn_like(y)
    y.backward(g)
    xr, wr, br = (t.detach().clone().requires_grad_() for t in (x, w7, bias))
    img = xr[:, 1:].transpose(1, 2).reshape(B, C, H, W)
    yr = torch.nn.functional.conv2d(img, wr.t().reshape(C, 1, 7, 7), br, padding=3, groups=C)
    yr = torch.cat([xr[:, :1], yr.flatten(2).transpose(1, 2)], 1)
    yr.backward(g)
    for a, b_, tol in ((y, yr, 1e-5), (x.grad, xr.grad, 1e-5), (w7.grad, wr.grad, 1e-4), (bias.grad, br.grad, 1e-4)):
        assert rel_err(a, b_) < tol, rel_err(a, b_)

"""GPU parity: co-attention, fusion heads and subspace losses vs golden fixtures of the unmodified reference."""
import pytest
import torch

import oracle
from conftest import grads_close, load_golden, rel_err
from synth import randn, synth_state_dict

pytestmark = pytest.mark.gpu


def test_mha_golden():
    import sml_b200
    g = load_golden("mha")
    mha = sml_b200.MultiheadAttention(256, 1).cuda().eval()
    mha.load_state_dict(g["state"])
    q = g["q"].cuda().requires_grad_(); kv = g["kv"].cuda().requires_grad_()
    o, raw = mha(q, kv, kv)
    assert o.shape == (6, 2, 256) and raw.shape == (2, 1, 6, 70)
    assert rel_err(o, g["out"]) < 1e-4 and rel_err(raw, g["raw"]) < 1e-4
    (o * g["go"].cuda()).sum().backward()
    assert rel_err(q.grad, g["gq"]) < 1e-3 and rel_err(kv.grad, g["gkv"]) < 1e-3
    grads_close({k: p.grad for k, p in mha.named_parameters()}, g["gparams"], 1e-3)
    mha4 = sml_b200.MultiheadAttention(256, 4).cuda().eval()           # CMTA direction: N queries x G keys, 4 heads
    mha4.load_state_dict(g["state"])
    o4, raw4 = mha4(g["kv"].cuda(), g["q"].cuda(), g["q"].cuda())
    assert rel_err(o4, g["out4"]) < 1e-4 and rel_err(raw4, g["raw4"]) < 1e-4


@pytest.mark.parametrize("heads", [1, 4])
def test_mha_folded_kv_equals_projected(heads):
    """the re-associated co-attention (K / V projections never materialised) against the golden fixture of the
    reference (forced on its 70-key case) and against the projected form at 6 genomic queries x 16 384 patch tokens: outputs,
    raw scores and every gradient"""
    import sml_b200
    g = load_golden("mha")
    if heads == 1:
        mha = sml_b200.MultiheadAttention(256, 1).cuda().eval()
        mha.load_state_dict(g["state"])
        mha.fold_kv = True
        q = g["q"].cuda().requires_grad_(); kv = g["kv"].cuda().requires_grad_()
        o, raw = mha(q, kv, kv)
        assert rel_err(o, g["out"]) < 1e-4 and rel_err(raw, g["raw"]) < 1e-4
        (o * g["go"].cuda()).sum().backward()
        assert rel_err(q.grad, g["gq"]) < 1e-3 and rel_err(kv.grad, g["gkv"]) < 1e-3
        grads_close({k: p.grad for k, p in mha.named_parameters()}, g["gparams"], 1e-3)
    torch.manual_seed(2)
    mha = sml_b200.MultiheadAttention(256, heads).cuda().eval()
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.1)
    qx = randn((6, 2, 256), 501).cuda(); kx = (randn((16384, 2, 256), 502) * 0.5).cuda(); go = randn((6, 2, 256), 503).cuda()
    res = {}
    for fold in (False, True):
        mha.fold_kv = fold
        mha.zero_grad(set_to_none=True)
        q = qx.clone().requires_grad_(); kv = kx.clone().requires_grad_()
        o, raw = mha(q, kv, kv)
        (o * go).sum().backward()
        res[fold] = (o.detach(), raw.detach(), q.grad, kv.grad, {k: p.grad.clone() for k, p in mha.named_parameters()})
    mha.fold_kv = None
    assert mha(qx, kx, kx)[0].shape == (6, 2, 256)                                  # automatic choice takes the folded form here
    a, b = res[False], res[True]
    assert rel_err(b[0], a[0]) < 1e-4 and rel_err(b[1], a[1]) < 1e-4
    assert rel_err(b[2], a[2]) < 1e-3 and rel_err(b[3], a[3]) < 1e-3
    grads_close(b[4], a[4], 1e-3)


def test_bilinear_fusion_golden():
    import sml_b200
    g = load_golden("bilinear_fusion")
    bf = sml_b200.BilinearFusion(skip=0, use_bilinear=1, gate1=1, gate2=1, dim1=32, dim2=32, scale_dim1=1, scale_dim2=1,
                                 mmhid=64, dropout_rate=0.0).cuda()
    bf.load_state_dict(synth_state_dict(g["shapes"], g["seed"]))
    bf.eval()
    assert rel_err(bf(g["v1"].cuda(), g["v2"].cuda()), g["y_eval"]) < 1e-4
    bf.train()
    v1 = g["v1"].cuda().requires_grad_(); v2 = g["v2"].cuda().requires_grad_()
    y = bf(v1, v2)
    assert rel_err(y, g["y_train"]) < 1e-4
    y.backward(g["gy"].cuda())
    assert rel_err(v1.grad, g["gv1"]) < 1e-3 and rel_err(v2.grad, g["gv2"]) < 1e-3
    gp = {k: p.grad for k, p in bf.named_parameters()}
    # encoder{1,2}.0.bias feed a BatchNorm: their true gradient is 0 (BN removes the mean) and both sides hold rounding
    # noise only, hence the looser bound relative to the largest parameter gradient
    grads_close({k: gp[k] for k in g["gparams"]}, g["gparams"], 5e-3)


def test_losses_golden():
    import sml_b200
    g = load_golden("losses")
    om = g["omic"].cuda().requires_grad_(); vg = g["vgrid"].cuda().requires_grad_()
    bl = sml_b200.BatchLoss(om.shape[0], 1)(om, vg)
    assert rel_err(bl, g["batch_loss"]) < 1e-4
    bl.backward(g["gbl"].cuda())
    assert rel_err(om.grad, g["gomic"]) < 1e-3 and rel_err(vg.grad, g["gvgrid"]) < 1e-3
    ts = [g[k].cuda().requires_grad_() for k in ("P", "Ph", "G", "Gh")]
    ol = sml_b200.OrthogonalLoss(gamma=0.5)(*ts)
    assert rel_err(ol, g["orth"]) < 1e-5
    ol.sum().backward()
    for t, k in zip(ts, ("gP", "gPh", "gG", "gGh")):
        assert rel_err(t.grad, g[k]) < 1e-4
    bn = sml_b200.SynchronizedBatchNorm1d(8).cuda().train()
    y = bn(g["bn_x"].cuda())
    want = (g["bn_x"] - g["bn_mean"]) * g["bn_inv_std"]
    assert rel_err(y, want) < 1e-5
    assert rel_err(bn.running_var, g["bn_running_var"]) < 1e-5


def test_batchloss_collapsed_gram_equals_repeated():
    """omic repeated over N tokens (DeformCrossTransMIL.py:104) vs the tagged [B,128] base: same loss and gradients."""
    import sml_b200
    B, N = 6, 400
    base = randn((B, 128), 1).abs()
    vg = randn((8 * B, 2, 5, 5), 2)
    bc = base.double().requires_grad_(); vc = vg.double().requires_grad_()          # fp64 oracle
    want = oracle.loss.batch_loss(bc.unsqueeze(1).repeat(1, N, 1), vc, B)
    gl = randn(want.shape, 3)
    want.backward(gl.double())
    b1 = base.cuda().requires_grad_(); v1 = vg.cuda().requires_grad_()
    rep = b1.unsqueeze(1).repeat(1, N, 1)
    rep._sml_base = b1
    got = sml_b200.BatchLoss(B, 1)(rep, v1)
    assert rel_err(got, want) < 1e-4
    got.backward(gl.cuda())
    assert rel_err(b1.grad, bc.grad) < 1e-3 and rel_err(v1.grad, vc.grad) < 1e-3
    b2 = base.cuda().requires_grad_()
    got2 = sml_b200.BatchLoss(B, 1)(b2.unsqueeze(1).repeat(1, N, 1), vg.cuda())     # untagged: generic [Ng, N*128] Gram
    assert rel_err(got2, want) < 5e-4       # 51 200-long fp32 dot products on both sides, then a cancelling difference


def test_batchloss_long_rows_match_collapsed_base():
    """The token-repeated omic tensor (what DDP hands to the loss after cloning the outputs) and its [B,128] base give the
    same loss and the same gradient w.r.t. the base (row normalisation cancels the repetition, utils/loss.py:26-38)."""
    import sml_b200
    torch.manual_seed(4)
    B, N = 4, 4096
    base = torch.randn(B, 128, device="cuda").abs()
    vg = torch.randn(8 * B, 2, 12, 12, device="cuda", requires_grad=True)
    bl = sml_b200.BatchLoss(B, 1)
    b1 = base.clone().requires_grad_()
    l1 = bl(b1.unsqueeze(1).expand(B, N, 128), vg)               # stride-0 view -> collapsed path
    l1.sum().backward()
    g_v1 = vg.grad.clone(); vg.grad = None
    b2 = base.clone().requires_grad_()
    l2 = bl(b2.unsqueeze(1).repeat(1, N, 1), vg)                   # materialised rows -> tensor-core Gram with split-K
    l2.sum().backward()
    assert rel_err(l2, l1) < 1e-4
    assert rel_err(b2.grad, b1.grad) < 1e-3 and rel_err(vg.grad, g_v1) < 1e-3


def test_gradient_modulation_kernel_matches_reference_lines():
    """sml_grad_modulate (one kernel, no host sync) vs the reference's gradient-modulation block executed verbatim
    (train_test.py:87-184, fixture tests/golden/modulation.pt) and vs the fp64 oracle on a fresh case."""
    import sml_b200
    for c in load_golden("modulation"):
        cls = torch.nn.Linear(c["W"].shape[1], c["W"].shape[0]).cuda()
        with torch.no_grad():
            cls.weight.copy_(c["W"]); cls.bias.copy_(c["b"])
        cls.weight.grad = c["G"].clone().cuda()
        info = sml_b200.modulate_classifier_grads(cls, c["ft"].cuda(), c["fi"].cuda(), c["labels"].cuda(), return_info=True)
        assert rel_err(cls.weight.grad, c["G_out"]) < 1e-5
        assert abs(float(info[0] / info[1]) - c["ratio_t"]) < 1e-4 * abs(c["ratio_t"])
        assert int(info[2] + info[3]) == int((c["G_out"] != c["G"]).any(dim=1).sum())
    torch.manual_seed(12)
    W = torch.randn(3, 64) * 0.2; b = torch.randn(3) * 0.1; G = torch.randn(3, 64); G[:, 32:] = -G[:, :32] + 0.2 * torch.randn(3, 32)
    ft = torch.randn(5, 32); fi = torch.randn(5, 32); lab = torch.randint(0, 3, (5,))
    want = oracle.modulation.modulate(W.double(), b.double(), G.double(), ft.double(), fi.double(), lab)
    cls = torch.nn.Linear(64, 3).cuda()
    with torch.no_grad():
        cls.weight.copy_(W); cls.bias.copy_(b)
    cls.weight.grad = G.clone().cuda()
    sml_b200.modulate_classifier_grads(cls, ft.cuda(), fi.cuda(), lab.cuda())
    assert rel_err(cls.weight.grad, want) < 1e-5

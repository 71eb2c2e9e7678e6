"""CPU: the oracle restatement must reproduce the unmodified reference (golden fixtures)."""
import torch

import oracle
from conftest import load_golden, rel_err, grads_close
from synth import randn, synth_state_dict

TOL = 2e-5   # fp32 vs fp32, different op order


def _leaf(d):
    return {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in d.items()}


def test_pinv():
    g = load_golden("pinv")
    x = g["x"].clone().requires_grad_()
    z = oracle.nystrom.pinv_iter(x, g["iters"])
    assert rel_err(z, g["z"]) < TOL
    z.backward(g["gz"])
    assert rel_err(x.grad, g["gx"]) < 1e-4


def test_nystrom_attention():
    for c in load_golden("nystrom"):
        p = _leaf(c["state"])
        x = c["x"].clone().requires_grad_()
        cfg = c["cfg"]
        y, parts = oracle.nystrom.nystrom_attention(x, p, heads=cfg["heads"], num_landmarks=cfg["num_landmarks"],
                                                    pinv_iterations=cfg["pinv_iterations"], return_parts=True)
        assert rel_err(y, c["y"]) < TOL
        y.backward(c["gy"])
        assert rel_err(x.grad, c["gx"]) < 1e-4
        grads_close({k: p[k].grad for k in c["gparams"]}, c["gparams"], 1e-4)
        n = x.shape[1]; m = cfg["num_landmarks"]
        npad = n + parts["pad"]
        assert torch.equal(parts["landmark_ids"], torch.arange(npad) // (npad // m))


def test_translayer_ppeg():
    g = load_golden("translayer_ppeg")
    p = _leaf(g["tl_state"]); x = g["x"].clone().requires_grad_()
    y = oracle.nystrom.trans_layer(x, p, "", dim=64)
    assert rel_err(y, g["y"]) < TOL
    y.backward(g["gy"])
    assert rel_err(x.grad, g["gx"]) < 1e-4
    p = _leaf(g["pp_state"]); x = g["px"].clone().requires_grad_()
    y = oracle.nystrom.ppeg(x, p, "", 6, 6)
    assert rel_err(y, g["py"]) < TOL
    y.backward(g["pgy"])
    assert rel_err(x.grad, g["pgx"]) < 1e-4
    grads_close({k: p[k].grad for k in g["pp_gparams"]}, g["pp_gparams"], 1e-4)


def test_transmil():
    g = load_golden("transmil")
    p = _leaf(synth_state_dict(g["shapes"], g["seed"]))
    x = randn(g["x_shape"], g["x_seed"])
    enc, logits = oracle.nystrom.transmil(x, p)
    assert rel_err(logits, g["logits"]) < 1e-4
    assert rel_err(enc, g["enc"]) < 1e-4
    loss = (logits * randn(logits.shape, g["glogits_seed"])).sum() + (enc * randn(enc.shape, g["genc_seed"])).sum()
    loss.backward()
    grads_close({k: p[k].grad for k in g["gparams"]}, g["gparams"], 1e-3)
    for k, nv in g["gnorms"].items():
        assert abs(float(p[k].grad.norm()) - nv) <= 1e-3 * nv + 1e-6, k


def test_deform1d():
    g = load_golden("deform1d")
    p = _leaf(g["state"])
    x1 = g["x1"].clone().requires_grad_(); x2 = g["x2"].clone().requires_grad_()
    y, parts = oracle.deform.deform_attn_1d(x1, x2, p, return_parts=True)
    assert rel_err(y, g["y"]) < TOL
    assert rel_err(parts["vgrid"], g["vgrid"]) < TOL
    y.backward(g["gy"])
    assert rel_err(x1.grad, g["gx1"]) < 1e-4
    assert rel_err(x2.grad, g["gx2"]) < 1e-4
    grads_close({k: p[k].grad for k in g["gparams"]}, g["gparams"], 1e-4)
    # chunked evaluation is the same function
    y2 = oracle.deform.deform_attn_1d(g["x1"], g["x2"], g["state"], q_chunk=7)
    assert rel_err(y2, g["y"]) < TOL
    # F5: only the middle token of x2 matters in the reference pad order
    x2b = g["x2"].clone(); x2b[:, :, :20] = 7.0; x2b[:, :, 21:] = -3.0
    y3 = oracle.deform.deform_attn_1d(g["x1"], x2b, g["state"])
    assert rel_err(y3, g["y"]) < TOL


def test_deform1d_upstream_taps_match_grid_sample():
    """The non-default pad order equals F.grid_sample on the transposed axes; taps are integer-exact."""
    import torch.nn.functional as F
    feat = randn((3, 5, 23, 1), 7)
    gy = randn((3, 11), 8).clamp(-1.3, 1.3)
    out, (x0, y0) = oracle.deform.bilinear_sample(feat, torch.zeros_like(gy), gy)
    grid = torch.stack([torch.zeros_like(gy), gy], -1).view(3, 11, 1, 2)
    ref = F.grid_sample(feat, grid, mode="bilinear", padding_mode="zeros", align_corners=False)[..., 0]
    assert rel_err(out, ref) < 1e-6
    assert torch.equal(y0, torch.floor(((gy + 1) * 23 - 1) / 2).long())


def test_deform2d():
    g = load_golden("deform2d")
    p = _leaf(g["state"])
    x1 = randn((1, 128, 2500), g["x1_seed"]).requires_grad_(); x2 = randn((1, 128, 2500), g["x2_seed"]).requires_grad_()
    y, vgrid = oracle.deform.deform_attn_2d(x1, x2, p, side=50, q_chunk=500)
    assert rel_err(y, g["y"]) < TOL
    assert rel_err(vgrid, g["vgrid"]) < TOL
    y.backward(randn(y.shape, g["gy_seed"]))
    assert rel_err(x1.grad[:, :, ::7], g["gx1"]) < 1e-4
    assert rel_err(x2.grad[:, :, ::7], g["gx2"]) < 1e-4
    grads_close({k: p[k].grad for k in g["gparams"]}, g["gparams"], 1e-4)


def test_deform_cross_transmil():
    for ad in (2, 1):
        g = load_golden(f"deformtransmil_ad{ad}")
        st = synth_state_dict(g["shapes"], g["seed"])
        for k in st:
            if k.endswith("to_offsets.2.weight"):
                st[k] = st[k] * g["offw_scale"]
        p = _leaf(st)
        path = randn((1, 2500, 1024), 83).requires_grad_(); omic = randn((1, 128), 84).abs().requires_grad_()
        o = oracle.deform.deform_cross_transmil(path, omic, p, attn_dim=ad, q_chunk=500)
        assert rel_err(o["logits"], g["logits"]) < 1e-4
        assert rel_err(o["encoded"], g["enc"]) < 1e-4
        loss = (o["logits"] * randn(o["logits"].shape, 85)).sum() + (o["encoded"] * randn(o["encoded"].shape, 86)).sum()
        if ad == 2:
            assert rel_err(o["vgrid"], g["vgrid"]) < 1e-4
            loss = loss + (o["vgrid"] * randn(o["vgrid"].shape, 87)).sum() * 0.01
        loss.backward()
        assert rel_err(omic.grad, g["gomic"]) < 1e-3
        assert rel_err(path.grad[:, ::50, ::16], g["gpath_sub"]) < 1e-3
        for k, nv in g["gnorms"].items():
            assert abs(float(p[k].grad.norm()) - nv) <= 2e-3 * nv + 1e-6, k


def test_mha():
    g = load_golden("mha")
    p = _leaf(g["state"])
    q = g["q"].clone().requires_grad_(); kv = g["kv"].clone().requires_grad_()
    o, raw = oracle.heads.mha_coattn(q, kv, kv, p, 1)
    assert rel_err(o, g["out"]) < TOL and rel_err(raw, g["raw"]) < TOL
    (o * g["go"]).sum().backward()
    assert rel_err(q.grad, g["gq"]) < 1e-4 and rel_err(kv.grad, g["gkv"]) < 1e-4
    grads_close({k: p[k].grad for k in g["gparams"]}, g["gparams"], 1e-4)
    o4, raw4 = oracle.heads.mha_coattn(g["kv"], g["q"], g["q"], g["state"], 4)
    assert rel_err(o4, g["out4"]) < TOL and rel_err(raw4, g["raw4"]) < TOL


def test_bilinear_fusion():
    g = load_golden("bilinear_fusion")
    p = _leaf(synth_state_dict(g["shapes"], g["seed"]))
    ye = oracle.heads.bilinear_fusion(g["v1"], g["v2"], p, training_bn=False)
    assert rel_err(ye, g["y_eval"]) < 1e-4
    v1 = g["v1"].clone().requires_grad_(); v2 = g["v2"].clone().requires_grad_()
    yt = oracle.heads.bilinear_fusion(v1, v2, p, training_bn=True)
    assert rel_err(yt, g["y_train"]) < 1e-4
    yt.backward(g["gy"])
    assert rel_err(v1.grad, g["gv1"]) < 1e-3 and rel_err(v2.grad, g["gv2"]) < 1e-3


def test_losses():
    g = load_golden("losses")
    om = g["omic"].clone().requires_grad_(); vg = g["vgrid"].clone().requires_grad_()
    bl = oracle.loss.batch_loss(om, vg, om.shape[0])
    assert rel_err(bl, g["batch_loss"]) < TOL
    bl.backward(g["gbl"])
    assert rel_err(om.grad, g["gomic"]) < 1e-4 and rel_err(vg.grad, g["gvgrid"]) < 1e-4
    ts = [g[k].clone().requires_grad_() for k in ("P", "Ph", "G", "Gh")]
    ol = oracle.loss.orthogonal_loss(*ts)
    assert rel_err(ol, g["orth"]) < TOL
    ol.sum().backward()
    for t, k in zip(ts, ("gP", "gPh", "gG", "gGh")):
        assert rel_err(t.grad, g[k]) < 1e-4
    mean, inv_std, unb = oracle.loss.sync_bn_stats([g["bn_x"][:4], g["bn_x"][4:]])
    assert rel_err(mean, g["bn_mean"]) < TOL and rel_err(inv_std, g["bn_inv_std"]) < TOL
    assert rel_err(0.9 * torch.ones(8) + 0.1 * unb, g["bn_running_var"]) < TOL


def test_modulation_oracle_matches_reference_source_lines():
    """oracle.modulation.modulate vs the reference's own lines executed verbatim (tests/golden/make_golden_modulation.py)."""
    import torch
    import oracle
    from conftest import load_golden
    for c in load_golden("modulation"):
        got = oracle.modulation.modulate(c["W"], c["b"], c["G"], c["ft"], c["fi"], c["labels"])
        assert torch.allclose(got, c["G_out"], rtol=1e-5, atol=1e-6)
        assert (got != c["G"]).any()                      # the fixture exercises the projection


def test_deform_pathomic_net():
    """oracle.deform.deform_pathomic_net vs the unmodified reference (tests/golden/make_golden_pathomic.py), batch 2."""
    import torch
    g = load_golden("pathomic")
    st = synth_state_dict(g["shapes"], g["seed"])
    for k in st:
        if k.endswith("to_offsets.2.weight"):
            st[k] = st[k] * g["offw_scale"]
    p = _leaf({k: v for k, v in st.items() if v.is_floating_point() and not k.endswith(("output_range", "output_shift"))})
    x_path = randn((2, 2500, 1024), 112)
    xt = randn((2, 59), 113).abs().requires_grad_(); xi = randn((2, 361), 114).abs().requires_grad_()
    o = oracle.deform.deform_pathomic_net(x_path, xt, xi, p, q_chunk=500)
    assert rel_err(o["features"], g["feats"]) < 1e-4 and rel_err(o["vec_tumor"], g["vt"]) < 1e-4
    for a, b in zip(o["logits"], g["logits"]):
        assert rel_err(a, b) < 1e-4
    assert rel_err(o["vgrid_tumor"], g["vgrid_tumor"]) < 1e-4 and rel_err(o["vgrid_immune"], g["vgrid_immune"]) < 1e-4
    assert rel_err(o["omic_tumor"][:, 5], g["omic_tumor_row"]) < 1e-5
    loss = sum((l * randn(l.shape, 120 + j)).sum() for j, l in enumerate(o["logits"])) + (o["features"] * randn(o["features"].shape, 125)).sum()
    loss = loss + 0.01 * ((o["vgrid_tumor"] * randn(o["vgrid_tumor"].shape, 126)).sum() + (o["vgrid_immune"] * randn(o["vgrid_immune"].shape, 127)).sum())
    loss.backward()
    assert rel_err(xt.grad, g["gxt"]) < 1e-3 and rel_err(xi.grad, g["gxi"]) < 1e-3
    for k, nv in g["gnorms"].items():
        if k in p and p[k].grad is not None:
            assert abs(float(p[k].grad.norm()) - nv) <= 2e-3 * nv + 1e-6, k


def _cpb1d_params(seed):
    import torch
    torch.manual_seed(seed)
    return {"mlp.0.0.weight": torch.randn(32, 1) * 1.5, "mlp.0.0.bias": torch.randn(32) * 0.7, "mlp.1.0.weight": torch.randn(32, 32) * 0.4,
            "mlp.1.0.bias": torch.randn(32) * 0.3, "mlp.2.weight": torch.randn(2, 32) * 0.5, "mlp.2.bias": torch.randn(2) * 0.1}


def test_pwl_table_equals_bias_mlp():
    """The 1-D bias MLP is exactly a piecewise-linear table of the signed-log distance (oracle/pwl.py)."""
    import torch
    for seed in (0, 1, 2):
        p = _cpb1d_params(seed)
        table = oracle.pwl.build_table(p)
        assert 32 <= table[0].numel() <= 32 + 33 * 32
        dx = (torch.rand(20000, dtype=torch.float64) * 2 - 1) * 2.4
        t = oracle.deform.slog(dx)
        want = oracle.deform.cpb_mlp(t.unsqueeze(1), {k: v.double() for k, v in p.items()}, "")
        got = oracle.pwl.eval_table(t, table)
        assert float((got - want).abs().max()) < 1e-10


def test_pwl_table_backward_equals_mlp_backward():
    """Per-segment sums of (d bias, d bias * t) reproduce the MLP's parameter and position gradients."""
    import torch
    p = _cpb1d_params(3)
    table = oracle.pwl.build_table(p)
    t = oracle.deform.slog((torch.rand(5000, dtype=torch.float64) * 2 - 1) * 2.2).requires_grad_()
    g = torch.randn(5000, 2, dtype=torch.float64)
    q = {k: v.double().requires_grad_() for k, v in p.items()}
    oracle.deform.cpb_mlp(t.unsqueeze(1), q, "").backward(g)
    G0, G1, dt = oracle.pwl.segment_sums(t.detach(), g, table)
    assert float((dt - t.grad).abs().max()) < 1e-10
    got = oracle.pwl.param_grads_from_segments(p, "", table, G0, G1)
    for k in q:
        assert float((got[k] - q[k].grad).abs().max()) < 1e-8 * max(1.0, float(q[k].grad.abs().max())), k


def test_cmta():
    """oracle/cmta.py vs the unmodified reference CMTA (tests/golden/make_golden_cmta.py), outputs and gradients"""
    g = load_golden("cmta")["cmta"]
    p = _leaf(synth_state_dict(g["shapes"], g["seed"]))
    with torch.no_grad():
        for k in p:
            if k.endswith("cls_token"):
                p[k].mul_(0.1)
    x_path = randn((2, 150, 1024), 212); x_omic = randn((2, 431), 213).abs()
    outs = oracle.cmta.cmta(x_path, x_omic, p)
    for j, (o, w) in enumerate(zip(outs, g["outs"])):
        assert rel_err(o, w) < 1e-4, j
    sum((o * randn(o.shape, 220 + j)).sum() for j, o in enumerate(outs)).backward()
    grads_close({k: p[k].grad for k in g["gparams"]}, g["gparams"], 2e-3)
    for k, nv in g["gnorms"].items():
        assert abs(float(p[k].grad.norm()) - nv) <= 2e-3 * nv + 1e-6, k


def test_mcat_front():
    """the patch-token part of MCAT_Surv (projection + SNNs + co-attention) vs the logits of the unmodified reference is
    covered on the GPU (torch glue downstream); here: the oracle's co-attention on the fixture's parameters is finite and
    has the reference's shapes (G queries x N keys, raw scores, F13)"""
    for tag in ("mcat4", "mcat6"):
        g = load_golden("cmta")[tag]
        p = synth_state_dict(g["shapes"], g["seed"])
        seed = g["seed"] - 1
        x_path = randn((2, 150, 1024), seed + 2); x_omic = randn((2, 431), seed + 3).abs()
        o, raw = oracle.cmta.mcat_front(x_path, x_omic, p, g["sizes"])
        G = len(g["sizes"])
        assert o.shape == (G, 2, 256) and raw.shape == (2, 1, G, 150) and torch.isfinite(o).all()

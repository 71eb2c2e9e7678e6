"""GPU parity: CMTA / MCAT_Surv end to end (models/model.py:559-853, cmta_utils.py:858-948) on libsml_b200 vs golden fixtures
of the unmodified reference (tests/golden/make_golden_cmta.py; CMTA evaluated in fp64, see that file) and vs the CPU oracle
at a 16 384-patch bag.  Tolerances are BASELINE.json's: 1e-4 relative in the fp32-grade modes, 2e-2 in bf16."""
from argparse import Namespace

import pytest
import torch

import oracle
from conftest import grads_close, load_golden, rel_err
from synth import randn, synth_state_dict

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "bf16x3": 1e-4, "bf16": 2e-2}


def _cmta_state(g):
    st = synth_state_dict(g["shapes"], g["seed"])
    for k in st:
        if k.endswith("cls_token"):
            st[k] = st[k] * 0.1
    return st


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "bf16"])
def test_cmta_golden(mode):
    import sml_b200
    g = load_golden("cmta")["cmta"]
    net = sml_b200.CMTA(Namespace(label_dim=4)).cuda().eval()
    net.load_state_dict(_cmta_state(g))                       # reference-keyed state_dict loads unchanged
    net.set_precision(mode)
    outs = net(x_path=randn((2, 150, 1024), 212).cuda(), x_omic=randn((2, 431), 213).abs().cuda())
    assert len(outs) == 7
    for j, (o, w) in enumerate(zip(outs, g["outs"])):
        assert o.shape == w.shape and rel_err(o, w) < TOL[mode], (j, rel_err(o, w))
    sum((o * randn(o.shape, 220 + j).cuda()).sum() for j, o in enumerate(outs)).backward()
    gp = {k: p.grad for k, p in net.named_parameters()}
    grads_close({k: gp[k] for k in g["gparams"]}, g["gparams"], 20 * TOL[mode])
    for k, nv in g["gnorms"].items():
        assert abs(float(gp[k].norm()) - nv) <= 20 * TOL[mode] * nv + 1e-5, (k, float(gp[k].norm()), nv)


@pytest.mark.parametrize("tag", ["mcat4", "mcat6"])
def test_mcat_golden(tag):
    """MCAT_Surv: 4 live genomic groups and the 6-group co-attention shape of BASELINE.json configs[0]"""
    import sml_b200
    g = load_golden("cmta")[tag]
    net = sml_b200.MCAT_Surv(Namespace(label_dim=4), omic_sizes=g["sizes"]).cuda().eval()
    net.load_state_dict(synth_state_dict(g["shapes"], g["seed"]))
    seed = g["seed"] - 1
    x_path = randn((2, 150, 1024), seed + 2).cuda(); x_omic = randn((2, 431), seed + 3).abs().cuda()
    outs = net(x_path=x_path, x_omic=x_omic)
    for j, (o, w) in enumerate(zip(outs, g["outs"])):
        assert rel_err(o, w) < 1e-4, (j, rel_err(o, w))
    sum((o * randn(o.shape, seed + 10 + j).cuda()).sum() for j, o in enumerate(outs)).backward()
    gp = {k: p.grad for k, p in net.named_parameters()}
    grads_close({k: gp[k] for k in g["gparams"]}, g["gparams"], 2e-3)
    # the raw co-attention scores (second output of the operator, F13) against the oracle
    p = {k: v for k, v in synth_state_dict(g["shapes"], g["seed"]).items()}
    o_ref, raw_ref = oracle.cmta.mcat_front(x_path.cpu(), x_omic.cpu(), p, g["sizes"])
    with torch.no_grad():
        h_path, h_omic = net._front(dict(x_path=x_path, x_omic=x_omic))
        o, raw = net.coattn(torch.stack(h_omic), h_path.transpose(0, 1), h_path.transpose(0, 1))
    assert raw.shape == raw_ref.shape and rel_err(raw, raw_ref) < 1e-4 and rel_err(o, o_ref) < 1e-4


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_cmta_16k_vs_oracle(mode):
    """BASELINE size: one 16 384-patch bag through the whole CMTA model against the fp32 CPU oracle (logits, the four cls
    embeddings, loss and the gradients of the patch projection / classifier)."""
    import sml_b200
    torch.manual_seed(5)
    net = sml_b200.CMTA(Namespace(label_dim=4)).cuda().eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    st = synth_state_dict(shapes, 311)
    for k in st:
        if k.endswith("cls_token"):
            st[k] = st[k] * 0.1
    net.load_state_dict(st)
    net.set_precision(mode)
    x_path = randn((1, 16384, 1024), 312); x_omic = randn((1, 431), 313).abs()
    y = torch.tensor([2])
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in st.items()}
    ref = oracle.cmta.cmta(x_path, x_omic, p)
    loss_ref = torch.nn.functional.cross_entropy(ref[0], y) + 0.5 * ((ref[3].detach() - ref[4]).abs().mean() + (ref[5].detach() - ref[6]).abs().mean())
    loss_ref.backward()                                       # train_test.py:347-373 (CE + L1 alignment)
    outs = net(x_path=x_path.cuda(), x_omic=x_omic.cuda())
    loss = torch.nn.functional.cross_entropy(outs[0], y.cuda()) + 0.5 * ((outs[3].detach() - outs[4]).abs().mean() + (outs[5].detach() - outs[6]).abs().mean())
    loss.backward()
    tol = TOL[mode]
    for j in (0, 3, 4, 5, 6):
        assert rel_err(outs[j], ref[j]) < tol, (j, rel_err(outs[j], ref[j]))
    assert abs(float(loss) - float(loss_ref)) <= tol * abs(float(loss_ref))
    gp = {k: q.grad for k, q in net.named_parameters()}
    keys = ["wsi_net.0.weight", "classifier.weight", "mm.0.weight", "P_in_G_Att.in_proj_weight", "G_in_P_Att.in_proj_weight",
            "pathomics_encoder.layer1.attn.to_qkv.weight", "pathomics_decoder.layer2.attn.to_out.0.weight"]
    grads_close({k: gp[k] for k in keys}, {k: p[k].grad for k in keys}, 20 * tol)

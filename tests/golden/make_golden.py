"""Generate golden fixtures from the UNMODIFIED reference (authoring container only).

    python tests/golden/make_golden.py

Imports /root/reference with the stubs of SURVEY.md §8(c) (matplotlib, lifelines,
sksurv, imblearn; ``nystrom_attention`` aliased to the in-tree copy
models/NystromAttention.py; ``.cuda()`` made a no-op), runs the reference
modules on seeded inputs and stores inputs / parameters / outputs / gradients in
tests/golden/*.pt.  Large tensors are regenerated from (shape, seed) by
tests/golden/synth.py instead of being stored.  Nothing here runs on the GPU box.
"""
import os
import sys
import types
from argparse import Namespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from synth import randn, synth_state_dict  # noqa: E402

REF = "/root/reference"


def install_stubs():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    stub("matplotlib"); stub("matplotlib.pyplot")
    stub("lifelines"); stub("lifelines.utils", concordance_index=None)
    stub("sksurv"); stub("sksurv.metrics", concordance_index_censored=None)
    stub("imblearn"); stub("imblearn.over_sampling", RandomOverSampler=None)
    sys.path.insert(0, REF)
    import models.NystromAttention as na
    sys.modules["nystrom_attention"] = na
    torch.Tensor.cuda = lambda self, *a, **k: self           # reference calls .cuda() unconditionally
    torch.nn.Module.cuda = lambda self, *a, **k: self

    class _FT:                                                # torch.cuda.FloatTensor(n, 1).fill_(1)
        def __new__(cls, *shape):
            return torch.empty(*shape, dtype=torch.float32)
    torch.cuda.FloatTensor = _FT


def shapes_of(mod):
    return {k: tuple(v.shape) for k, v in mod.state_dict().items()}


def grads_of(mod):
    return {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}


def save(name, obj):
    path = os.path.join(HERE, name + ".pt")
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def main():
    install_stubs()
    torch.set_num_threads(8)
    from models.NystromAttention import NystromAttention, moore_penrose_iter_pinv
    from models.DeformableAttention1D import DeformCrossAttention1D
    from models.DeformableAttention2D import DeformCrossAttention2D
    from models.MultiheadAttention import MultiheadAttention
    from models.DeformCrossTransMIL import DeformCrossTransMIL
    from models.mil import TransMIL, TransLayer, PPEG
    from models.fusion import BilinearFusion
    from models.cmta_utils import OrthogonalLoss
    from utils.loss import BatchLoss
    from utils.sync_batchnorm import SynchronizedBatchNorm1d

    # ---- pinv -------------------------------------------------------------------------
    x = torch.softmax(randn((2, 3, 16, 16), 1, 2.0), dim=-1).requires_grad_()
    z = moore_penrose_iter_pinv(x, 6)
    g = randn(z.shape, 2)
    z.backward(g)
    save("pinv", dict(x=x.detach(), iters=6, z=z.detach(), gz=g, gx=x.grad))

    # ---- NystromAttention: small (padding, l>1) and head-dim-64 configs -----------------
    cases = []
    for ci, (dim, dh, hh, m, n, b) in enumerate([(64, 8, 8, 16, 37, 2), (128, 64, 2, 32, 100, 1), (64, 16, 4, 8, 64, 2)]):
        torch.manual_seed(10 + ci)
        mod = NystromAttention(dim=dim, dim_head=dh, heads=hh, num_landmarks=m, pinv_iterations=6,
                               residual=True, dropout=0.1).eval()
        xin = randn((b, n, dim), 20 + ci).requires_grad_()
        y = mod(xin)
        gy = randn(y.shape, 30 + ci)
        y.backward(gy)
        cases.append(dict(cfg=dict(dim=dim, dim_head=dh, heads=hh, num_landmarks=m, pinv_iterations=6),
                          state=mod.state_dict(), x=xin.detach(), y=y.detach(), gy=gy, gx=xin.grad,
                          gparams=grads_of(mod)))
    save("nystrom", cases)

    # ---- TransLayer + PPEG (mil.py) at dim 64 ----------------------------------------------
    torch.manual_seed(40)
    tl = TransLayer(dim=64).eval()
    xin = randn((2, 50, 64), 41).requires_grad_()
    y = tl(xin); gy = randn(y.shape, 42); y.backward(gy)
    pp = PPEG(dim=64).eval()
    xin2 = randn((2, 1 + 36, 64), 43).requires_grad_()
    y2 = pp(xin2, 6, 6); gy2 = randn(y2.shape, 44); y2.backward(gy2)
    save("translayer_ppeg", dict(tl_state=tl.state_dict(), x=xin.detach(), y=y.detach(), gy=gy, gx=xin.grad,
                                 tl_gparams=grads_of(tl),
                                 pp_state=pp.state_dict(), px=xin2.detach(), py=y2.detach(), pgy=gy2,
                                 pgx=xin2.grad, pp_gparams=grads_of(pp)))

    # ---- TransMIL (mil.py:209-259), params from synth seed -------------------------------
    args = Namespace(label_dim=4, path_dim=128)
    tm = TransMIL(args).eval()
    shp = shapes_of(tm)
    tm.load_state_dict(synth_state_dict(shp, 50))
    xin = randn((1, 300, 1024), 51)                       # 300 -> padded to 18*18=324, +cls
    enc, logits, _ = tm(xin)
    loss = (logits * randn(logits.shape, 52)).sum() + (enc * randn(enc.shape, 53)).sum()
    loss.backward()
    gp = grads_of(tm)
    keep = ["_fc2.weight", "cls_token", "layer1.attn.res_conv.weight", "layer2.norm.weight", "pos_layer.proj.bias",
            "layer1.attn.to_out.0.bias"]
    save("transmil", dict(shapes=shp, seed=50, x_shape=(1, 300, 1024), x_seed=51, enc=enc.detach(),
                          logits=logits.detach(), glogits_seed=52, genc_seed=53,
                          gparams={k: gp[k] for k in keep},
                          gnorms={k: float(v.norm()) for k, v in gp.items()}))

    # ---- DeformCrossAttention1D (ctor as DeformCrossTransMIL.py:55-60) -----------------------
    torch.manual_seed(60)
    a1 = DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6).eval()
    sd = a1.state_dict()
    sd["to_offsets.2.weight"] = sd["to_offsets.2.weight"] * 8.0      # make offsets non-trivial
    a1.load_state_dict(sd)
    x1 = randn((2, 128, 41), 61).requires_grad_(); x2 = randn((2, 128, 41), 62).requires_grad_()
    y, vg = a1(x1, x2, return_vgrid=True)
    gy = randn(y.shape, 63); y.backward(gy)
    save("deform1d", dict(state=a1.state_dict(), x1=x1.detach(), x2=x2.detach(), y=y.detach(), vgrid=vg.detach(),
                          gy=gy, gx1=x1.grad, gx2=x2.grad, gparams=grads_of(a1)))

    # ---- DeformCrossAttention2D (ctor as DeformCrossTransMIL.py:45-54), N = 2500 fixed -------
    torch.manual_seed(70)
    a2 = DeformCrossAttention2D(dim=128, dim_head=64, heads=8, dropout=0.1, downsample_factor=4, offset_scale=4,
                                offset_groups=8, offset_kernel_size=6).eval()
    sd = a2.state_dict()
    sd["to_offsets.2.weight"] = sd["to_offsets.2.weight"] * 8.0
    a2.load_state_dict(sd)
    x1 = randn((1, 128, 2500), 71).requires_grad_(); x2 = randn((1, 128, 2500), 72).requires_grad_()
    y, vg = a2(x1, x2, return_vgrid=True)
    gy = randn(y.shape, 73); y.backward(gy)
    save("deform2d", dict(state=a2.state_dict(), x1_seed=71, x2_seed=72, gy_seed=73, y=y.detach(), vgrid=vg.detach(),
                          gx1=x1.grad[:, :, ::7].clone(), gx2=x2.grad[:, :, ::7].clone(), gparams=grads_of(a2)))

    # ---- DeformCrossTransMIL, attn_dim 2 (+vgrid) and attn_dim 1 ------------------------------
    for ad in (2, 1):
        args = Namespace(path_dim=128, attn_dim=ad, return_vgrid=(ad == 2))
        dm = DeformCrossTransMIL(args).eval()
        shp = shapes_of(dm)
        st = synth_state_dict(shp, 80 + ad)
        for k in st:
            if k.endswith("to_offsets.2.weight"):
                st[k] = st[k] * 4.0
        dm.load_state_dict(st)
        path = randn((1, 2500, 1024), 83).requires_grad_(); omic = randn((1, 128), 84).abs().requires_grad_()
        outs = dm(path, omic)
        enc, logits = outs[0], outs[1]
        loss = (logits * randn(logits.shape, 85)).sum() + (enc * randn(enc.shape, 86)).sum()
        if ad == 2:
            loss = loss + (outs[4] * randn(outs[4].shape, 87)).sum() * 0.01
        loss.backward()
        gp = grads_of(dm)
        save(f"deformtransmil_ad{ad}", dict(shapes=shp, seed=80 + ad, offw_scale=4.0, enc=enc.detach(), logits=logits.detach(),
                                            vgrid=(outs[4].detach() if ad == 2 else None), gomic=omic.grad,
                                            gpath_sub=path.grad[:, ::50, ::16].clone(),
                                            gnorms={k: float(v.norm()) for k, v in gp.items()},
                                            gparams={k: v for k, v in gp.items() if v.numel() <= 4096}))

    # ---- MultiheadAttention co-attention (MCAT: G queries x N keys; CMTA: N x G) ---------------
    torch.manual_seed(90)
    mha = MultiheadAttention(embed_dim=256, num_heads=1).eval()
    qg = randn((6, 2, 256), 91).requires_grad_(); kv = randn((70, 2, 256), 92).requires_grad_()
    o, raw = mha(qg, kv, kv)
    go = randn(o.shape, 93); graw = randn(raw.shape, 94, 0.1)
    (o * go).sum().backward(retain_graph=True)
    g1 = dict(gq=qg.grad.clone(), gkv=kv.grad.clone(), gparams=grads_of(mha))
    mha4 = MultiheadAttention(embed_dim=256, num_heads=4).eval()
    mha4.load_state_dict(mha.state_dict())
    o4, raw4 = mha4(kv.detach(), qg.detach(), qg.detach())
    save("mha", dict(state=mha.state_dict(), q=qg.detach(), kv=kv.detach(), out=o.detach(), raw=raw.detach(), go=go,
                     **g1, out4=o4.detach(), raw4=raw4.detach()))

    # ---- BilinearFusion (fusion.py) train-mode BN + eval-mode BN --------------------------------
    bf = BilinearFusion(skip=0, use_bilinear=1, gate1=1, gate2=1, dim1=32, dim2=32, scale_dim1=1, scale_dim2=1,
                        mmhid=64, dropout_rate=0.0)
    shp = shapes_of(bf)
    bf.load_state_dict(synth_state_dict(shp, 100))
    v1 = randn((6, 32), 101).requires_grad_(); v2 = randn((6, 32), 102).requires_grad_()
    bf.eval(); ye = bf(v1, v2).detach()
    bf.train(); yt = bf(v1, v2); gy = randn(yt.shape, 103); yt.backward(gy)
    save("bilinear_fusion", dict(shapes=shp, seed=100, v1=v1.detach(), v2=v2.detach(), y_eval=ye, y_train=yt.detach(),
                                 gy=gy, gv1=v1.grad, gv2=v2.grad,
                                 gparams={k: v for k, v in grads_of(bf).items() if v.numel() <= 8192}))

    # ---- BatchLoss (world 1) + OrthogonalLoss + sync-BN statistics --------------------------------
    B = 4
    om = randn((B, 30, 16), 110).requires_grad_(); vg = randn((8 * B, 2, 3, 3), 111).requires_grad_()
    bl = BatchLoss(B, 1)(om, vg)
    gbl = randn(bl.shape, 112); bl.backward(gbl)
    P, Ph, G, Gh = (randn((5, 256), 120 + i).requires_grad_() for i in range(4))
    ol = OrthogonalLoss(gamma=0.5)(P, Ph, G, Gh)
    ol.sum().backward()
    sbn = SynchronizedBatchNorm1d(8)
    xb = randn((10, 8), 130)
    mean, inv_std = sbn._compute_mean_std(xb.sum(0), (xb * xb).sum(0), 10)
    save("losses", dict(omic=om.detach(), vgrid=vg.detach(), batch_loss=bl.detach(), gbl=gbl, gomic=om.grad, gvgrid=vg.grad,
                        P=P.detach(), Ph=Ph.detach(), G=G.detach(), Gh=Gh.detach(), orth=ol.detach(),
                        gP=P.grad, gPh=Ph.grad, gG=G.grad, gGh=Gh.grad,
                        bn_x=xb, bn_mean=mean, bn_inv_std=inv_std, bn_running_var=sbn.running_var.clone()))


if __name__ == "__main__":
    main()

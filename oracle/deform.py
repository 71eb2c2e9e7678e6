"""Oracle (test infrastructure only): deformable cross-attention 1-D / 2-D and DeformCrossTransMIL.

Restates /root/reference/models/DeformableAttention1D.py:36-240,
models/DeformableAttention2D.py:88-325 and models/DeformCrossTransMIL.py:28-202
in plain torch ops (own bilinear sampler that also returns the integer tap
indices; no einops; no F.grid_sample).  The hard-coded 2500 / 50x50 of the
reference (DeformCrossTransMIL.py:104, DeformableAttention2D.py:239-240,318) is
parameterised as N / sqrt(N); at N=2500 the functions equal the unmodified
reference (pinned by tests/golden).  ``q_chunk`` evaluates attention by
query-row tiles (rows are independent) so 16k bags fit in host memory.
"""
import math

import torch
import torch.nn.functional as F

from .nystrom import layer_norm


def bilinear_sample(feat, gx, gy):
    """grid_sample(mode=bilinear, padding_mode=zeros, align_corners=False) restated.

    feat [G,C,H,W]; gx, gy [G,P] normalised coords.  Returns out [G,C,P] and the
    integer top-left tap (x0, y0) [G,P] (bit-exact contract for the CUDA gather).
    """
    G, C, H, W = feat.shape
    ix = ((gx + 1.0) * W - 1.0) / 2.0
    iy = ((gy + 1.0) * H - 1.0) / 2.0
    x0f, y0f = torch.floor(ix), torch.floor(iy)
    x0, y0 = x0f.long(), y0f.long()
    wx1, wy1 = ix - x0f, iy - y0f
    wx0, wy0 = 1.0 - wx1, 1.0 - wy1
    flat = feat.reshape(G, C, H * W)
    out = feat.new_zeros(G, C, gx.shape[1])
    for dx, dy, w in ((0, 0, wx0 * wy0), (1, 0, wx1 * wy0), (0, 1, wx0 * wy1), (1, 1, wx1 * wy1)):
        xx, yy = x0 + dx, y0 + dy
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1))
        val = torch.gather(flat, 2, idx.unsqueeze(1).expand(G, C, -1))
        out = out + val * (w * ok.to(feat.dtype)).unsqueeze(1)
    return out, (x0, y0)


def slog(x):
    return torch.sign(x) * torch.log(x.abs() + 1.0)


def cpb_mlp(pos, p, prefix):
    """CPB MLP: Linear(c,32) ReLU Linear(32,32) ReLU Linear(32,o) (1D.py:60-102, 2D.py:120-157)."""
    h = torch.relu(pos @ p[prefix + "mlp.0.0.weight"].t() + p[prefix + "mlp.0.0.bias"])
    h = torch.relu(h @ p[prefix + "mlp.1.0.weight"].t() + p[prefix + "mlp.1.0.bias"])
    return h @ p[prefix + "mlp.2.weight"].t() + p[prefix + "mlp.2.bias"]


def _attend(q, k, v, bias_fn, q_chunk):
    """softmax(q k^T + bias) v by query-row chunks.  q [b,h,n,d], k/v [b,h,j,d]."""
    n = q.shape[2]
    outs = []
    step = n if not q_chunk else q_chunk
    for s in range(0, n, step):
        e = min(n, s + step)
        sim = q[:, :, s:e] @ k.transpose(-1, -2) + bias_fn(s, e)
        sim = sim - sim.amax(dim=-1, keepdim=True).detach()
        outs.append(sim.softmax(dim=-1) @ v)
    return torch.cat(outs, dim=2)


def deform_attn_1d(x1, x2, p, heads=8, dim_head=64, groups=4, ds=4, ksize=6, offset_scale=2.0,
                   pad_order="reference", q_chunk=0, return_parts=False, rows=None):
    """DeformCrossAttention1D.forward, dropout 0 (DeformableAttention1D.py:156-240).

    pad_order="reference" reproduces the degenerate grid_sample_1d of the
    reference (grid padded as (x=vgrid, y=0) on an [n,1] image, :36-43);
    "upstream" is the intended 1-D linear interpolation along the sequence.
    ``rows`` (LongTensor of query indices): evaluate the attention only for those queries (softmax rows are independent)
    and return y[:, :, rows] - how BASELINE-size bags (n = 16 385, n_kv = 4096) are checked in seconds.
    """
    b, C, n = x1.shape
    inner = heads * dim_head
    g = groups
    gd = inner // g
    q = torch.einsum("oc,bcn->bon", p["to_q.weight"][:, :, 0], x1)            # :175
    gq = q.reshape(b * g, gd, n)                                              # :179-181
    pad = (ksize - ds) // 2
    t = F.conv1d(gq, p["to_offsets.0.weight"], p["to_offsets.0.bias"], stride=ds, padding=pad, groups=gd)
    t = 0.5 * t * (1.0 + torch.erf(t / math.sqrt(2.0)))                       # exact GELU
    off = torch.tanh(torch.einsum("c,gcn->gn", p["to_offsets.2.weight"][0, :, 0], t)) * offset_scale
    nkv = off.shape[-1]
    vgrid = torch.arange(nkv, dtype=x1.dtype) + off                           # :186-187
    vs = 2.0 * vgrid / max(nkv - 1, 1) - 1.0                                  # :188, :45-48
    feat = x2.reshape(b * g, C // g, n, 1)                                    # H=n, W=1
    zeros = torch.zeros_like(vs)
    if pad_order == "reference":
        kv, taps = bilinear_sample(feat, vs, zeros)                           # :36-43 as written
    else:
        kv, taps = bilinear_sample(feat, zeros, vs)
    kv = kv.reshape(b, C, nkv)                                                # :195
    k = torch.einsum("oc,bcn->bon", p["to_k.weight"][:, :, 0], kv)            # :199
    v = torch.einsum("oc,bcn->bon", p["to_v.weight"][:, :, 0], kv)
    qs = q * dim_head ** -0.5                                                 # :203
    sp = lambda t_: t_.reshape(b, heads, dim_head, -1).transpose(-1, -2)      # b h n d
    qh, kh, vh = sp(qs), sp(k), sp(v)
    seq = 2.0 * torch.arange(n, dtype=x1.dtype) / max(n - 1, 1) - 1.0         # :215-216
    hpg = heads // g
    if rows is not None:
        qh, seq = qh[:, :, rows], seq[rows]

    def bias_fn(s, e):                                                        # :84-102
        pos = slog(seq[s:e].view(1, -1, 1, 1) - vs.view(b * g, 1, nkv, 1))
        bb = cpb_mlp(pos, p, "rel_pos_bias.")                                 # [(b g), i, j, o]
        return bb.reshape(b, g, e - s, nkv, hpg).permute(0, 1, 4, 2, 3).reshape(b, heads, e - s, nkv)

    out = _attend(qh, kh, vh, bias_fn, q_chunk)                               # :211-231
    out = out.transpose(-1, -2).reshape(b, inner, -1)
    y = torch.einsum("oc,bcn->bon", p["to_out.weight"][:, :, 0], out) + p["to_out.bias"].view(1, -1, 1)
    if return_parts:
        return y, dict(vgrid=vgrid, vgrid_scaled=vs, taps=taps, kv=kv, offsets=off)
    return y


def deform_attn_2d(x1, x2, p, side, heads=8, dim_head=64, groups=8, ds=4, ksize=6, offset_scale=4.0,
                   q_chunk=0, return_parts=False, rows=None):
    """DeformCrossAttention2D.forward, eval-mode dropout (DeformableAttention2D.py:223-325).

    x1, x2 [b,C,N] with N = side*side (the reference hard-codes side=50).
    Returns (out [b,C,N], vgrid [(b g),2,hk,wk]) like return_vgrid=True.  ``rows``: as in deform_attn_1d.
    """
    b, C, N = x1.shape
    H = W = side
    assert H * W == N
    inner = heads * dim_head
    g = groups
    gd = inner // g                      # 64
    cg = C // g                          # 16

    def gconv(w, x):                     # grouped 1x1 conv, weight [inner, C/g, 1, 1]
        return torch.einsum("goc,bgcn->bgon", w[:, :, 0, 0].reshape(g, gd, cg),
                            x.reshape(b, g, cg, -1)).reshape(b, inner, -1)

    q = gconv(p["to_q.weight"], x1)                                           # :246
    gq = q.reshape(b * g, gd, H, W)                                           # :252-254
    pad = (ksize - ds) // 2
    t = F.conv2d(gq, p["to_offsets.0.weight"], p["to_offsets.0.bias"], stride=ds, padding=pad, groups=gd)
    t = 0.5 * t * (1.0 + torch.erf(t / math.sqrt(2.0)))
    off = torch.tanh(torch.einsum("oc,gchw->gohw", p["to_offsets.2.weight"][:, :, 0, 0], t)) * offset_scale
    hk, wk = off.shape[-2:]
    gx0 = torch.arange(wk, dtype=x1.dtype).view(1, 1, wk).expand(1, hk, wk)    # :88-98 channel0 = x
    gy0 = torch.arange(hk, dtype=x1.dtype).view(1, hk, 1).expand(1, hk, wk)
    vgrid = torch.stack([gx0 + off[:, 0], gy0 + off[:, 1]], dim=1)            # [(bg),2,hk,wk]  :260
    # normalize_grid (:100-108): channel 0 is divided by (h-1), channel 1 by (w-1) of the kv grid
    vsx = 2.0 * vgrid[:, 0] / max(hk - 1, 1) - 1.0
    vsy = 2.0 * vgrid[:, 1] / max(wk - 1, 1) - 1.0
    nkv = hk * wk
    feat = x2.reshape(b * g, cg, H, W)
    kv, taps = bilinear_sample(feat, vsx.reshape(b * g, nkv), vsy.reshape(b * g, nkv))   # :268-271
    kv = kv.reshape(b, C, nkv)                                                # :274
    k = gconv(p["to_k.weight"], kv)                                           # :279
    v = gconv(p["to_v.weight"], kv)
    qs = q * dim_head ** -0.5
    sp = lambda t_: t_.reshape(b, heads, dim_head, -1).transpose(-1, -2)
    qh, kh, vh = sp(qs), sp(k), sp(v)
    # query lattice, normalised; create_grid_like(x2) + normalize_grid(dim=0) (:296-297)
    qx = 2.0 * torch.arange(W, dtype=x1.dtype) / max(H - 1, 1) - 1.0
    qy = 2.0 * torch.arange(H, dtype=x1.dtype) / max(W - 1, 1) - 1.0
    qgrid = torch.stack([qx.view(1, W).expand(H, W), qy.view(H, 1).expand(H, W)], dim=-1).reshape(N, 2)
    kgrid = torch.stack([vsx.reshape(b * g, nkv), vsy.reshape(b * g, nkv)], dim=-1)     # [(bg), nkv, 2]
    hpg = heads // g
    if rows is not None:
        qh, qgrid = qh[:, :, rows], qgrid[rows]

    def bias_fn(s, e):                                                        # :143-157
        pos = slog(qgrid[s:e].view(1, e - s, 1, 2) - kgrid.view(b * g, 1, nkv, 2))
        bb = cpb_mlp(pos, p, "rel_pos_bias.")
        return bb.reshape(b, g, e - s, nkv, hpg).permute(0, 1, 4, 2, 3).reshape(b, heads, e - s, nkv)

    out = _attend(qh, kh, vh, bias_fn, q_chunk)                               # :292-312
    out = out.transpose(-1, -2).reshape(b, inner, -1)
    y = torch.einsum("oc,bcn->bon", p["to_out.weight"][:, :, 0, 0], out) + p["to_out.bias"].view(1, -1, 1)
    if return_parts:
        return y, vgrid, dict(vgrid_scaled=kgrid, taps=taps, kv=kv, offsets=off)
    return y, vgrid


def sub(p, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


def deform_cross_transmil(path, omic, p, attn_dim=2, pad_order="reference", q_chunk=0):
    """DeformCrossTransMIL.forward, eval mode (DeformCrossTransMIL.py:97-160, Pooler :169-202).

    path [B,N,1024], omic [B,128].  Returns dict(encoded, logits, omic_rep_base, vgrid).
    N is free (the reference hard-codes 2500).
    """
    B, N, _ = path.shape
    pt = torch.relu(path.float() @ p["_fc1.0.weight"].t() + p["_fc1.0.bias"])           # :99-100
    om = omic.float().unsqueeze(1).expand(B, N, omic.shape[-1])                         # :104
    h = torch.cat([pt, om], dim=-1) @ p["fusion_layer.fusion_layer.weight"].t() \
        + p["fusion_layer.fusion_layer.bias"]                                           # :110, :33-38 (path first)
    vgrid = None
    nw, nb = p["layer3.norm.weight"], p["layer3.norm.bias"]
    if attn_dim == 1:
        cls = p["cls_token"].expand(B, -1, -1)
        h = torch.cat([cls, h], dim=1)                                                  # :114-121
        pt = torch.cat([cls, pt], dim=1)
        a = deform_attn_1d(layer_norm(h, nw, nb).transpose(1, 2), layer_norm(pt, nw, nb).transpose(1, 2),
                           sub(p, "layer3.attn1d."), pad_order=pad_order, q_chunk=q_chunk)
        h = h + a.transpose(1, 2)                                                       # :66-67
        z = layer_norm(h, p["norm.weight"], p["norm.bias"])[:, 0]                       # :127
    else:
        side = int(round(math.sqrt(N)))
        a, vgrid = deform_attn_2d(layer_norm(h, nw, nb).transpose(1, 2), layer_norm(pt, nw, nb).transpose(1, 2),
                                  sub(p, "layer3.attn2d."), side=side, q_chunk=q_chunk)
        h = h + a.transpose(1, 2)                                                       # :71-75
        z = layer_norm(h, p["norm.weight"], p["norm.bias"]).mean(dim=1)                 # :144, :193
        z = torch.tanh(z @ p["pooler.dense.weight"].t() + p["pooler.dense.bias"])       # :199-200
    logits = z @ p["_fc2.weight"].t() + p["_fc2.bias"]
    encoded = z @ p["multimodal_projection.weight"].t() + p["multimodal_projection.bias"]
    return dict(encoded=encoded, logits=logits, omic=om, vgrid=vgrid)


def maxnet(x, p, prefix):
    """MaxNet.forward, eval mode (models/model.py:142-187): 4 x (Linear, ELU, AlphaDropout = identity), ReLU; classifier."""
    for i in range(4):
        x = torch.nn.functional.elu(x @ p[f"{prefix}encoder.{i}.0.weight"].t() + p[f"{prefix}encoder.{i}.0.bias"])
    feats = torch.relu(x)
    return feats, feats @ p[f"{prefix}classifier.0.weight"].t() + p[f"{prefix}classifier.0.bias"]


def deform_pathomic_net(x_path, x_omic_tumor, x_omic_immune, p, attn_dim=2, q_chunk=0):
    """DeformPathomicNet.forward, eval mode, fusion_type "concat", task_type != "survival" (models/model.py:481-544):
    two genomic encoders, two DeformCrossTransMIL branches on the same bag, concat, three linear heads."""
    om_t, _ = maxnet(x_omic_tumor.float(), p, "omic_net_tumor.")                          # :486-487
    om_i, _ = maxnet(x_omic_immune.float(), p, "omic_net_immune.")                        # :489-490
    bt = deform_cross_transmil(x_path, om_t, sub(p, "pathomic_net_tumor."), attn_dim=attn_dim, q_chunk=q_chunk)
    bi = deform_cross_transmil(x_path, om_i, sub(p, "pathomic_net_immune."), attn_dim=attn_dim, q_chunk=q_chunk)
    vt, vi = bt["encoded"], bi["encoded"]
    feats = torch.cat((vt, vi), 1)                                                        # :511-515
    hazard = feats @ p["classifier.weight"].t() + p["classifier.bias"]                    # :520
    hz_t = vt @ p["classifier_tumor.0.weight"].t() + p["classifier_tumor.0.bias"]         # :522
    hz_i = vi @ p["classifier_immune.0.weight"].t() + p["classifier_immune.0.bias"]       # :523
    return dict(features=feats, vec_tumor=vt, vec_immune=vi, logits=[hz_t, hz_i, hazard], omic_tumor=bt["omic"],
                vgrid_tumor=bt["vgrid"], omic_immune=bi["omic"], vgrid_immune=bi["vgrid"])

"""Oracle (test infrastructure only): Nystrom attention and the TransMIL stack.

Restates /root/reference/models/NystromAttention.py:20-157 (identical copy at
models/cmta_utils.py:147-281) and models/mil.py:171-259 in plain torch ops.
Parameters are passed as dicts keyed by the reference ``state_dict`` names so
that golden fixtures made from the reference modules load directly.
"""
import math

import torch
import torch.nn.functional as F


def landmark_ids(n_padded: int, m: int) -> torch.Tensor:
    """Integer landmark id of every (padded) token: t // l (NystromAttention.py:102-105)."""
    l = n_padded // m
    return torch.arange(n_padded, dtype=torch.int64) // l


def pinv_iter(x: torch.Tensor, iters: int = 6, return_trace: bool = False):
    """Moore-Penrose iterative pseudo-inverse (NystromAttention.py:20-35).

    The initial scale uses ONE scalar for the whole [b,h] batch: the global max
    of the row sums times the global max of the column sums of |x|.
    """
    ax = x.abs()
    row_sums = ax.sum(dim=-1)          # reference calls this ``col``
    col_sums = ax.sum(dim=-2)          # reference calls this ``row``
    z = x.transpose(-1, -2) / (row_sums.max() * col_sums.max())
    eye = torch.eye(x.shape[-1], dtype=x.dtype, device=x.device)
    trace = [z]
    for _ in range(iters):
        xz = x @ z
        t1 = 7.0 * eye - xz
        t2 = 15.0 * eye - xz @ t1
        t3 = 13.0 * eye - xz @ t2
        z = 0.25 * (z @ t3)
        trace.append(z)
    return (z, trace) if return_trace else z


def nystrom_attention(x, p, heads=8, num_landmarks=256, pinv_iterations=6,
                      residual=True, prefix="", return_parts=False):
    """NystromAttention.forward with mask=None, eval-mode dropout (NystromAttention.py:74-157).

    p: ``to_qkv.weight`` [3hd, D], ``to_out.0.weight`` [D, hd], ``to_out.0.bias`` [D],
       ``res_conv.weight`` [h,1,K,1].
    """
    b, n, dim = x.shape
    m = num_landmarks
    w_qkv = p[prefix + "to_qkv.weight"]
    inner = w_qkv.shape[0] // 3
    d = inner // heads
    pad = (m - n % m) % m
    if pad:                                                    # :79-82 (front pad)
        x = torch.cat([x.new_zeros(b, pad, dim), x], dim=1)
    npad = n + pad
    qkv = x @ w_qkv.t()                                        # :89
    q, k, v = (qkv[..., i * inner:(i + 1) * inner].reshape(b, npad, heads, d).permute(0, 2, 1, 3)
               for i in range(3))                              # :90
    q = q * (d ** -0.5)                                        # :98
    l = math.ceil(n / m)                                       # :102
    assert l * m == npad
    ql = q.reshape(b, heads, m, l, d).sum(dim=3) / l           # :103-118
    kl = k.reshape(b, heads, m, l, d).sum(dim=3) / l
    s1 = q @ kl.transpose(-1, -2)                              # :122-125
    s2 = ql @ kl.transpose(-1, -2)
    s3 = ql @ k.transpose(-1, -2)
    a1, a2, a3 = s1.softmax(-1), s2.softmax(-1), s3.softmax(-1)  # :137
    z = pinv_iter(a2, pinv_iterations)                         # :138
    out = (a1 @ z) @ (a3 @ v)                                  # :140
    if residual:                                               # :144-145
        wc = p[prefix + "res_conv.weight"]                     # [h,1,K,1]
        kk = wc.shape[2]
        vp = F.pad(v, (0, 0, kk // 2, kk // 2))                # zero pad tokens
        conv = torch.zeros_like(v)
        for tap in range(kk):
            conv = conv + wc[:, 0, tap, 0].view(1, heads, 1, 1) * vp[:, :, tap:tap + npad]
        out = out + conv
    merged = out.permute(0, 2, 1, 3).reshape(b, npad, inner)   # :149
    y = merged @ p[prefix + "to_out.0.weight"].t() + p[prefix + "to_out.0.bias"]  # :150
    y = y[:, -n:]                                              # :151
    if return_parts:
        return y, dict(q=q, k=k, v=v, ql=ql, kl=kl, a1=a1, a2=a2, a3=a3, z=z,
                       landmark_ids=landmark_ids(npad, m), pad=pad)
    return y


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def trans_layer(x, p, prefix, dim=512):
    """mil.py:171-189 — x + attn(norm(x)), m = dim/2, d = dim/8, 8 heads, 6 iters."""
    xn = layer_norm(x, p[prefix + "norm.weight"], p[prefix + "norm.bias"])
    return x + nystrom_attention(xn, p, heads=8, num_landmarks=dim // 2, pinv_iterations=6,
                                 prefix=prefix + "attn.")


def ppeg(x, p, prefix, H, W):
    """mil.py:192-206 — identity + depthwise 7x7 + 5x5 + 3x3 on the [C,H,W] view of tokens 1.."""
    B, _, C = x.shape
    cls, feat = x[:, :1], x[:, 1:]
    img = feat.transpose(1, 2).reshape(B, C, H, W)
    out = img.clone()
    for name, ks in (("proj", 7), ("proj1", 5), ("proj2", 3)):
        out = out + F.conv2d(img, p[prefix + name + ".weight"], p[prefix + name + ".bias"],
                             padding=ks // 2, groups=C)
    return torch.cat([cls, out.flatten(2).transpose(1, 2)], dim=1)


def transmil(x, p, dim=512):
    """mil.py:225-259 TransMIL.forward (eval mode).  Returns (encoded, logits)."""
    h = torch.relu(x.to(p["_fc1.0.weight"].dtype) @ p["_fc1.0.weight"].t() + p["_fc1.0.bias"])   # .float() in the reference; fp64 when the parameters are
    n = h.shape[1]
    side = int(math.ceil(math.sqrt(n)))
    add = side * side - n
    h = torch.cat([h, h[:, :add]], dim=1)                      # :232-235
    h = torch.cat([p["cls_token"].expand(h.shape[0], -1, -1), h], dim=1)
    h = trans_layer(h, p, "layer1.", dim)
    h = ppeg(h, p, "pos_layer.", side, side)
    h = trans_layer(h, p, "layer2.", dim)
    h = layer_norm(h, p["norm.weight"], p["norm.bias"])[:, 0]
    logits = h @ p["_fc2.weight"].t() + p["_fc2.bias"]
    encoded = h @ p["multimodal_projection.weight"].t() + p["multimodal_projection.bias"]
    return encoded, logits

"""Oracle (test infrastructure only): genomic x WSI co-attention and fusion heads.

Restates /root/reference/models/MultiheadAttention.py:7-321 (packed in-proj,
raw pre-softmax scores as the second output, F13) and
/root/reference/models/fusion.py:36-63 (BilinearFusion, eval-mode dropout,
BatchNorm1d in eval or train mode) in plain torch ops.
"""
import torch


def mha_coattn(query, key, value, p, num_heads, prefix=""):
    """multi_head_attention_forward, no masks, no bias_kv, eval dropout (MultiheadAttention.py:117-312).

    query [L,B,E], key/value [S,B,E].  Returns (out [L,B,E], raw_scores [B,h,L,S]).
    """
    L, B, E = query.shape
    S = key.shape[0]
    hd = E // num_heads
    w, bias = p[prefix + "in_proj_weight"], p[prefix + "in_proj_bias"]
    q = query @ w[:E].t() + bias[:E]                                   # :125-182
    k = key @ w[E:2 * E].t() + bias[E:2 * E]
    v = value @ w[2 * E:].t() + bias[2 * E:]
    q = q * float(hd) ** -0.5                                          # :205
    q = q.reshape(L, B * num_heads, hd).transpose(0, 1)                # :245-249
    k = k.reshape(S, B * num_heads, hd).transpose(0, 1)
    v = v.reshape(S, B * num_heads, hd).transpose(0, 1)
    raw = q @ k.transpose(1, 2)                                        # :282
    attn = raw.softmax(dim=-1)                                         # :300
    o = (attn @ v).transpose(0, 1).reshape(L, B, E)                    # :303-305
    o = o @ p[prefix + "out_proj.weight"].t() + p[prefix + "out_proj.bias"]
    return o, raw.reshape(B, num_heads, L, S)                          # :308-312


def batch_norm_1d(x, w, b, running_mean, running_var, training, eps=1e-5):
    if training:
        mu = x.mean(0)
        var = ((x - mu) ** 2).mean(0)
    else:
        mu, var = running_mean, running_var
    return (x - mu) / torch.sqrt(var + eps) * w + b


def bilinear_fusion(v1, v2, p, training_bn=False, skip=0):
    """BilinearFusion.forward with gates on, use_bilinear=1, dropout off (fusion.py:36-63)."""
    v1, v2 = torch.relu(v1), torch.relu(v2)

    def bil(w, bb):   # nn.Bilinear: y_o = x1^T W_o x2 + b_o
        return torch.einsum("bi,oij,bj->bo", v1, w, v2) + bb

    h1 = torch.relu(v1 @ p["linear_h1.0.weight"].t() + p["linear_h1.0.bias"])
    z1 = bil(p["linear_z1.weight"], p["linear_z1.bias"])
    o1 = torch.relu((torch.sigmoid(z1) * h1) @ p["linear_o1.0.weight"].t() + p["linear_o1.0.bias"])
    h2 = torch.relu(v2 @ p["linear_h2.0.weight"].t() + p["linear_h2.0.bias"])
    z2 = bil(p["linear_z2.weight"], p["linear_z2.bias"])
    o2 = torch.relu((torch.sigmoid(z2) * h2) @ p["linear_o2.0.weight"].t() + p["linear_o2.0.bias"])
    one = v1.new_ones(v1.shape[0], 1)
    o1, o2 = torch.cat([o1, one], 1), torch.cat([o2, one], 1)          # :56-57
    o12 = (o1.unsqueeze(2) * o2.unsqueeze(1)).flatten(1)               # :58
    out = o12 @ p["encoder1.0.weight"].t() + p["encoder1.0.bias"]
    out = torch.relu(batch_norm_1d(out, p["encoder1.1.weight"], p["encoder1.1.bias"],
                                   p["encoder1.1.running_mean"], p["encoder1.1.running_var"], training_bn))
    if skip:
        out = torch.cat([out, o1, o2], 1)
    out = out @ p["encoder2.0.weight"].t() + p["encoder2.0.bias"]
    out = torch.relu(batch_norm_1d(out, p["encoder2.1.weight"], p["encoder2.1.bias"],
                                   p["encoder2.1.running_mean"], p["encoder2.1.running_var"], training_bn))
    return out

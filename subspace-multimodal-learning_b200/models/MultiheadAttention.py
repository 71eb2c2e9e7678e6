"""Drop-in for /root/reference/models/MultiheadAttention.py:333-489 (+ multi_head_attention_forward :7-321; identical
copy at models/cmta_utils.py:354-847): same constructor, forward signature, parameter names; the second output is
the RAW pre-softmax score tensor [B, heads, L, S] (need_raw=True, :299-312).  CUDA only.

Only the paths the reference call sites use are implemented (MCAT model.py:587,627; CMTA model.py:748-750,809-818):
packed in-projection with bias, no masks, no bias_kv / zero_attn, same embedding dim for q/k/v.

Genomic x WSI direction (a handful of genomic queries over tens of thousands of patch tokens: MCAT model.py:627, CMTA
model.py:809): the K and V projections of the patch tokens are never materialised.  With the per-head slices
Wk_h, Wv_h of the packed in-projection,
    scores_h = q_h (Wk_h x + bk_h)^T = (q_h Wk_h) x^T + q_h . bk_h            -> one [h L, E] x [E, S] product on the raw tokens
    out_h    = sum_j a_hj (Wv_h x_j + bv_h) = (a_h x) Wv_h^T + (sum_j a_hj) bv_h -> one [h L, S] x [S, E] product, then [L, E] x [E, hd]
which reads the token matrix twice instead of projecting it (2 x 2 S E^2 FLOPs, two [S, E] tensors written and re-read):
at S = 16 384, E = 256, 6 queries that is 0.4 instead of 17 GFLOP per slide.  Exact re-association: same outputs, raw scores and
gradients to rounding (tests/test_gpu_heads.py)."""
import torch
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from .. import functional as Fn


class MultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None, vdim=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.kdim = kdim if kdim is not None else embed_dim
        self.vdim = vdim if vdim is not None else embed_dim
        self._qkv_same_embed_dim = self.kdim == embed_dim and self.vdim == embed_dim
        if not self._qkv_same_embed_dim or add_bias_kv or add_zero_attn:
            raise NotImplementedError("sml_b200 MultiheadAttention: kdim/vdim/add_bias_kv/add_zero_attn are not used by the reference")
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.register_parameter('q_proj_weight', None)
        self.register_parameter('k_proj_weight', None)
        self.register_parameter('v_proj_weight', None)
        if bias:
            self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        else:
            self.register_parameter('in_proj_bias', None)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.bias_k = self.bias_v = None
        self.add_zero_attn = add_zero_attn
        self.precision = None      # engine of the token-level projections: None / "fp32" (SIMT), "bf16x3", "bf16" (tcgen05)
        self.fold_kv = None        # None: automatic (few queries over many keys); True / False force the re-associated form
        self._reset_parameters()

    def _reset_parameters(self):
        xavier_uniform_(self.in_proj_weight)
        if self.in_proj_bias is not None:
            constant_(self.in_proj_bias, 0.)
            constant_(self.out_proj.bias, 0.)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, need_raw=True, attn_mask=None):
        if key_padding_mask is not None or attn_mask is not None:
            raise NotImplementedError("sml_b200 MultiheadAttention: masks are never passed by the reference call sites")
        L, B, E = query.shape
        S = key.shape[0]
        h, hd = self.num_heads, self.head_dim
        assert key.shape == value.shape and E == self.embed_dim
        w, bias = self.in_proj_weight, self.in_proj_bias
        bq, bk, bv = (None, None, None) if bias is None else (bias[:E], bias[E:2 * E], bias[2 * E:])
        mode = self.precision or "fp32"
        mq = mode if L >= 1024 else "fp32"            # the long side (N patch tokens) runs on the tensor-core engine
        mk = mode if S >= 1024 else "fp32"
        q = Fn.scale(Fn.linear(query.float(), w[:E], bq, mode=mq), float(hd) ** -0.5)   # :125-205
        fold = self.fold_kv if self.fold_kv is not None else (h * L <= 64 and S >= 1024)
        if fold:
            return self._forward_folded(q, key.float(), value.float(), w, bk, bv, need_weights, need_raw)
        k = Fn.linear(key.float(), w[E:2 * E], bk, mode=mk)
        v = Fn.linear(value.float(), w[2 * E:], bv, mode=mk)
        qh = q.view(L, B, h, hd).permute(1, 2, 0, 3)                                    # [B,h,L,hd] views (:245-249)
        kh = k.view(S, B, h, hd).permute(1, 2, 0, 3)
        vh = v.view(S, B, h, hd).permute(1, 2, 0, 3)
        raw = Fn.matmul(qh, kh.transpose(-1, -2))                                       # [B,h,L,S]  :282
        attn = Fn.softmax_lastdim(raw)                                                  # :300
        attn = Fn.dropout_add(None, attn, self.dropout, self.training)                  # :301
        o = Fn.matmul(attn, vh)                                                         # [B,h,L,hd] :303
        o = o.permute(2, 0, 1, 3).reshape(L, B, E)                                      # :305
        out = Fn.linear(o, self.out_proj.weight, self.out_proj.bias, mode=mq)           # :306
        if need_weights:
            if need_raw:
                return out, raw                                                         # :308-312
            return out, Fn.scale(attn.sum(dim=1), 1.0 / h)
        return out, None

    def _forward_folded(self, q, key, value, w, bk, bv, need_weights, need_raw):
        """few queries over many keys: K / V projections folded into the query side and the output side (module docstring)"""
        L, B, E = q.shape
        S = key.shape[0]
        h, hd = self.num_heads, self.head_dim
        qh = q.view(L, B, h, hd).permute(1, 2, 0, 3)                                    # [B,h,L,hd]
        wk_h = w[E:2 * E].view(1, h, hd, E).expand(B, h, hd, E)                         # per-head slices, broadcast over B
        wv_h = w[2 * E:].view(1, h, hd, E).expand(B, h, hd, E)
        qk = Fn.matmul(qh, wk_h)                                                        # [B,h,L,E] = q_h Wk_h
        kt = key.permute(1, 2, 0)                                                       # [B,E,S] view of the raw tokens
        raw = Fn.matmul(qk.reshape(B, h * L, E), kt).view(B, h, L, S)                   # (q_h Wk_h) x^T
        if bk is not None:
            cb = Fn.matmul(qh, bk.view(1, h, hd, 1).expand(B, h, hd, 1))                # [B,h,L,1] = q_h . bk_h
            raw = raw + cb
        attn = Fn.softmax_lastdim(raw)
        attn = Fn.dropout_add(None, attn, self.dropout, self.training)
        ctx = Fn.matmul(attn.reshape(B, h * L, S), value.permute(1, 0, 2)).view(B, h, L, E)     # a_h x
        o = Fn.matmul(ctx, wv_h.transpose(-1, -2))                                      # [B,h,L,hd] = (a_h x) Wv_h^T
        if bv is not None:
            o = o + attn.sum(dim=-1, keepdim=True) * bv.view(1, h, 1, hd)               # (sum_j a_hj) bv_h (= bv_h without dropout)
        o = o.permute(2, 0, 1, 3).reshape(L, B, E)
        out = Fn.linear(o, self.out_proj.weight, self.out_proj.bias)
        if need_weights:
            if need_raw:
                return out, raw
            return out, Fn.scale(attn.sum(dim=1), 1.0 / h)
        return out, None

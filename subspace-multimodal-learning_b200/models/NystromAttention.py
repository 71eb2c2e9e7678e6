"""Drop-in for /root/reference/models/NystromAttention.py:39-157 (and the PyPI ``nystrom_attention``
package the reference imports at models/mil.py:24): same constructor, forward signature and
``state_dict`` keys (to_qkv.weight, to_out.0.weight, to_out.0.bias, res_conv.weight); the arithmetic runs
on libsml_b200 (CUDA only — a CPU tensor raises)."""
from math import ceil

import torch
from torch import nn

from .. import functional as Fn
from ..nystrom import nystrom_attention, landmark_ids

DEFAULT_PRECISION = "fp32"


def moore_penrose_iter_pinv(x, iters=6, precision=None):
    """NystromAttention.py:20-35 on the device kernels (forward only helper)."""
    from ..nystrom import pinv_fwd
    return pinv_fwd(x.contiguous(), iters, precision or DEFAULT_PRECISION)[0].f


class NystromAttention(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8, num_landmarks=256, pinv_iterations=6, residual=True,
                 residual_conv_kernel=33, eps=1e-8, dropout=0.):
        super().__init__()
        self.eps = eps
        inner_dim = heads * dim_head
        self.num_landmarks = num_landmarks
        self.pinv_iterations = pinv_iterations
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))
        self.residual = residual
        if residual:
            self.res_conv = nn.Conv2d(heads, heads, (residual_conv_kernel, 1), padding=(residual_conv_kernel // 2, 0),
                                      groups=heads, bias=False)
        self.precision = None      # None -> module-level DEFAULT_PRECISION

    def forward(self, x, mask=None, return_attn=False):
        if mask is not None:
            raise NotImplementedError("sml_b200 NystromAttention: mask is never passed by the reference call sites "
                                      "(mil.py:186, cmta_utils.py:872) and is not implemented")
        if not self.residual:
            raise NotImplementedError("sml_b200 NystromAttention: residual=False is not on the reference path")
        y = nystrom_attention(x.float(), self.to_qkv.weight, self.to_out[0].weight, self.to_out[0].bias,
                              self.res_conv.weight, self.heads, self.num_landmarks, self.pinv_iterations,
                              self.precision or DEFAULT_PRECISION)
        out = Fn.dropout_add(None, y, self.to_out[1].p, self.training)
        if return_attn:                                     # :153-155 (detached, exact-fp32 engine; inspection only)
            from ..nystrom import nystrom_attn_map
            return out, nystrom_attn_map(x.detach().float(), self.to_qkv.weight, self.heads, self.num_landmarks, self.pinv_iterations)
        return out

    def landmark_ids(self, n, device="cuda"):
        """int32 landmark id of each (front-padded) token: t // ceil(n/m) — the bit-exact index contract."""
        m = self.num_landmarks
        npad = n + (m - n % m) % m
        assert npad // m == ceil(n / m)
        return landmark_ids(npad, m, device)

"""Drop-ins for TransLayer / PPEG / TransMIL of /root/reference/models/mil.py:171-259 (same attribute names
and state_dict keys; identical copies live in models/cmta_utils.py:858-891).  CUDA only."""
import math

import torch
from torch import nn

from .. import functional as Fn
from .NystromAttention import NystromAttention


class TransLayer(nn.Module):
    def __init__(self, norm_layer=nn.LayerNorm, dim=512):
        super().__init__()
        self.norm = norm_layer(dim)
        self.attn = NystromAttention(dim=dim, dim_head=dim // 8, heads=8, num_landmarks=dim // 2, pinv_iterations=6,
                                     residual=True, dropout=0.1)

    def forward(self, x):
        a = self.attn
        from ..nystrom import nystrom_attention
        from . import NystromAttention as NA
        # LayerNorm fused into the attention function: it writes the normalised rows (and their bf16 copy) straight
        # into the front-padded buffer the qkv GEMM reads
        # x + dropout(attn(norm(x))) as ONE autograd node: the skip gradient is added inside the LayerNorm-backward kernel
        p = float(a.to_out[1].p) if self.training else 0.0
        x = x.float()
        skip = (p, Fn.current_seed() if p > 0 else 0, Fn._next_offset(x.numel()) if p > 0 else 0)
        return nystrom_attention(x, a.to_qkv.weight, a.to_out[0].weight, a.to_out[0].bias, a.res_conv.weight, a.heads,
                                 a.num_landmarks, a.pinv_iterations, a.precision or NA.DEFAULT_PRECISION,
                                 pre_norm=(self.norm.weight, self.norm.bias, self.norm.eps), skip=skip)


class PPEG(nn.Module):
    def __init__(self, dim=512):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim, 7, 1, 7 // 2, groups=dim)
        self.proj1 = nn.Conv2d(dim, dim, 5, 1, 5 // 2, groups=dim)
        self.proj2 = nn.Conv2d(dim, dim, 3, 1, 3 // 2, groups=dim)

    def folded(self):
        """identity + 7x7 + 5x5 + 3x3 depthwise kernels summed into one tap-major [49,C] stencil (exact), one kernel."""
        return Fn.ppeg_fold(self.proj.weight, self.proj.bias, self.proj1.weight, self.proj1.bias, self.proj2.weight,
                            self.proj2.bias)

    def forward(self, x, H, W):
        w7, bias = self.folded()
        return Fn.ppeg(x, w7, bias, H, W)


class TransMIL(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.pos_layer = PPEG(dim=512)
        self._fc1 = nn.Sequential(nn.Linear(1024, 512), nn.ReLU())
        self.cls_token = nn.Parameter(torch.randn(1, 1, 512))
        self.n_classes = self.args.label_dim
        self.layer1 = TransLayer(dim=512)
        self.layer2 = TransLayer(dim=512)
        self.norm = nn.LayerNorm(512)
        self._fc2 = nn.Linear(512, self.n_classes)
        self.multimodal_projection = nn.Linear(512, self.args.path_dim)
        self.precision = None

    def set_precision(self, p):
        self.precision = p
        self.layer1.attn.precision = p
        self.layer2.attn.precision = p

    def forward(self, x):
        from . import NystromAttention as NA
        mode = self.precision or NA.DEFAULT_PRECISION
        n = x.shape[1]
        side = int(math.ceil(math.sqrt(n)))
        add = side * side - n
        # _fc1 + ReLU, pad-to-square and cls token in one pass: the GEMM writes into the [B, 1+n+add, 512] buffer
        h = Fn.fc1_cls(x if x.dtype == torch.bfloat16 else x.float(), self._fc1[0].weight, self._fc1[0].bias, self.cls_token, add, mode)   # mil.py:227-240
        h = self.layer1(h)
        h = self.pos_layer(h, side, side)
        h = self.layer2(h)
        # only the cls row of the final LayerNorm is used (mil.py:252)
        z = Fn.layer_norm(h[:, 0].contiguous(), self.norm.weight, self.norm.bias, self.norm.eps)
        logits = Fn.linear(z, self._fc2.weight, self._fc2.bias)
        encoded = Fn.linear(z, self.multimodal_projection.weight, self.multimodal_projection.bias)
        return encoded, logits, None

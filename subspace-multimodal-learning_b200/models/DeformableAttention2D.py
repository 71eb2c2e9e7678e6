"""Drop-in for /root/reference/models/DeformableAttention2D.py:161-325 (class DeformCrossAttention2D; same
constructor, forward(x1, x2, return_vgrid) and state_dict keys).  CUDA only.

The reference hard-codes a 50x50 lattice (:239-240, :318); here the side is sqrt(N) of the incoming bag, so N = 2500
reproduces the reference and 16 384 (128x128) / any square N run through the same kernels."""
import math

from torch import nn

from ._deform_common import CPB, Scale, default, divisible_by, run_deform


class DeformCrossAttention2D(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8, dropout=0., downsample_factor=4, offset_scale=4, offset_groups=8,
                 offset_kernel_size=6, group_queries=True, group_key_values=True):
        super().__init__()
        offset_scale = default(offset_scale, downsample_factor)
        assert offset_kernel_size >= downsample_factor, 'offset kernel size must be greater than or equal to the downsample factor'
        assert divisible_by(offset_kernel_size - downsample_factor, 2)
        offset_groups = default(offset_groups, heads)
        assert divisible_by(heads, offset_groups)
        assert dim_head == 64, "sml_b200 deformable attention kernels are built for dim_head = 64"
        inner_dim = dim_head * heads
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.offset_groups = offset_groups
        offset_dims = inner_dim // offset_groups
        self.downsample_factor = downsample_factor
        self.offset_kernel_size = offset_kernel_size
        self.offset_scale = offset_scale
        self.group_queries, self.group_key_values = group_queries, group_key_values
        self.precision = None          # None -> _deform_common.DEFAULT_PRECISION ("bf16x3"); "bf16" / "fp32"
        self.log_distance_ok = True
        self.pad_order_mode = 0
        self.to_offsets = nn.Sequential(
            nn.Conv2d(offset_dims, offset_dims, offset_kernel_size, groups=offset_dims, stride=downsample_factor,
                      padding=(offset_kernel_size - downsample_factor) // 2),
            nn.GELU(),
            nn.Conv2d(offset_dims, 2, 1, bias=False),
            nn.Tanh(),
            Scale(offset_scale))
        self.rel_pos_bias = CPB(dim // 4, offset_groups=offset_groups, heads=heads, depth=2, in_dim=2)
        self.dropout = nn.Dropout(dropout)
        self.to_q = nn.Conv2d(dim, inner_dim, 1, groups=offset_groups if group_queries else 1, bias=False)
        self.to_k = nn.Conv2d(dim, inner_dim, 1, groups=offset_groups if group_key_values else 1, bias=False)
        self.to_v = nn.Conv2d(dim, inner_dim, 1, groups=offset_groups if group_key_values else 1, bias=False)
        self.to_out = nn.Conv2d(inner_dim, dim, 1)

    def forward(self, x1, x2, return_vgrid=False):
        n = x1.shape[-1]
        side = int(round(math.sqrt(n)))
        assert side * side == n, "DeformCrossAttention2D needs a square number of tokens (reference: 50*50)"
        return run_deform(self, x1, x2, return_vgrid, two_d=True, side=side)

"""Drop-in for /root/reference/models/DeformableAttention1D.py:106-240 (class DeformCrossAttention1D; same
constructor, forward(x1, x2, return_vgrid) and state_dict keys).  CUDA only.

``pad_order="reference"`` (default) reproduces the reference's grid_sample_1d exactly as written (:36-43: the grid
is padded as (x=vgrid, y=0) on an [n,1] image, so every key/value token is a scaled copy of the middle token of
x2 — SURVEY F5); ``pad_order="upstream"`` is the intended 1-D linear interpolation along the sequence."""
from torch import nn

from ._deform_common import CPB, Scale, default, divisible_by, run_deform


class DeformCrossAttention1D(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8, dropout=0., downsample_factor=4, offset_scale=None, offset_groups=4,
                 offset_kernel_size=6, cpb_log_distance=True, group_queries=False, group_key_values=False,
                 pad_order="reference"):
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
        self.log_distance_ok = bool(cpb_log_distance)
        assert pad_order in ("reference", "upstream")
        self.pad_order_mode = 0 if pad_order == "reference" else 1
        self.to_offsets = nn.Sequential(
            nn.Conv1d(offset_dims, offset_dims, offset_kernel_size, groups=offset_dims, stride=downsample_factor,
                      padding=(offset_kernel_size - downsample_factor) // 2),
            nn.GELU(),
            nn.Conv1d(offset_dims, 1, 1, bias=False),
            nn.Identity(),            # Rearrange('b 1 n -> b n') in the reference
            nn.Tanh(),
            Scale(offset_scale))
        self.rel_pos_bias = CPB(dim // 4, offset_groups=offset_groups, heads=heads, depth=2, in_dim=1, log_distance=cpb_log_distance)
        self.dropout = nn.Dropout(dropout)
        self.to_q = nn.Conv1d(dim, inner_dim, 1, groups=offset_groups if group_queries else 1, bias=False)
        self.to_k = nn.Conv1d(dim, inner_dim, 1, groups=offset_groups if group_key_values else 1, bias=False)
        self.to_v = nn.Conv1d(dim, inner_dim, 1, groups=offset_groups if group_key_values else 1, bias=False)
        self.to_out = nn.Conv1d(inner_dim, dim, 1)

    def forward(self, x1, x2, return_vgrid=False):
        return run_deform(self, x1, x2, return_vgrid, two_d=False)

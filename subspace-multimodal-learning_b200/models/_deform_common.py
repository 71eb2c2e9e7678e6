"""Shared pieces of the DeformCrossAttention1D/2D drop-ins."""
import torch
from torch import nn

from .. import functional as Fn
from ..deform import DeformAttnFn


# Arithmetic of the fused attention kernel (per module: ``mod.precision``; None -> this default):
#   "bf16x3"  warp-level bf16 MMAs with hi/lo 3-pass products (fp32-grade: 1e-4 on outputs)       [default]
#   "bf16"    plain bf16 MMA operands (2e-2)
#   "fp32"    the SIMT fp32 kernels (exact fp32 arithmetic, ~7x slower)
DEFAULT_PRECISION = "bf16x3"


def gemm_mode(precision):
    """engine of the projection GEMMs around the attention kernel: SIMT fp32 for "fp32", else the tcgen05 engine with the
    fp32-grade hi/lo split (see deform._op)."""
    return "fp32" if (precision or DEFAULT_PRECISION) == "fp32" else "bf16x3"


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else d


def divisible_by(numer, denom):
    return (numer % denom) == 0


class Scale(nn.Module):
    def __init__(self, scale):
        super().__init__()
        self.scale = scale

    def forward(self, x):
        return x * self.scale


class CPB(nn.Module):
    """Parameter container with the reference's names (rel_pos_bias.mlp.{0.0,1.0,2}.{weight,bias});
    the MLP itself is evaluated per (query, key) pair inside the fused attention kernel."""

    def __init__(self, dim, *, heads, offset_groups, depth, in_dim, log_distance=True):
        super().__init__()
        assert depth == 2 and dim == 32, "fused CPB kernel is specialised to the reference configuration (32 hidden, depth 2)"
        self.heads, self.offset_groups, self.log_distance = heads, offset_groups, log_distance
        self.mlp = nn.ModuleList([])
        self.mlp.append(nn.Sequential(nn.Linear(in_dim, dim), nn.ReLU()))
        for _ in range(depth - 1):
            self.mlp.append(nn.Sequential(nn.Linear(dim, dim), nn.ReLU()))
        self.mlp.append(nn.Linear(dim, heads // offset_groups))

    def params(self):
        return (self.mlp[0][0].weight, self.mlp[0][0].bias, self.mlp[1][0].weight, self.mlp[1][0].bias,
                self.mlp[2].weight, self.mlp[2].bias)

    def forward(self, grid_q, grid_kv):
        raise NotImplementedError("sml_b200: CPB is fused into the deformable attention kernel (never materialised)")


def run_deform(mod, x1, x2, return_vgrid, two_d, side=None, want_taps=False):
    """x1, x2 [b, C, n] as the reference passes them (transposed views of token-major tensors)."""
    assert mod.log_distance_ok, "cpb_log_distance=False is not on the reference path"
    x1t, x2t = x1.transpose(1, 2), x2.transpose(1, 2)
    p = mod.dropout.p if mod.training else 0.0
    seed = Fn.current_seed() + 7919 * Fn._next_offset(1) if p > 0 else 0
    cfg = dict(dim2d=two_d, side=side, heads=mod.heads, groups=mod.offset_groups, ds=mod.downsample_factor,
               ksize=mod.offset_kernel_size, scale=float(mod.offset_scale), gq=mod.group_queries, gkv=mod.group_key_values,
               p_drop=float(p), seed=int(seed), mode=mod.pad_order_mode, want_taps=want_taps,
               precision=getattr(mod, "precision", None) or DEFAULT_PRECISION)
    y, vgrid = DeformAttnFn.apply(x1t.float(), x2t.float(), mod.to_q.weight, mod.to_k.weight, mod.to_v.weight, mod.to_out.weight,
                                  mod.to_out.bias, mod.to_offsets[0].weight, mod.to_offsets[0].bias, mod.to_offsets[2].weight,
                                  *mod.rel_pos_bias.params(), cfg)
    out = y.transpose(1, 2)
    return (out, vgrid) if return_vgrid else out

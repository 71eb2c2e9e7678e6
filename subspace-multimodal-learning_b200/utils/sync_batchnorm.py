"""Synchronized batch norm for [B, C] features: drop-in for the role of torch.nn.SyncBatchNorm (main.py:118) and of
the vendored utils/sync_batchnorm (SynchronizedBatchNorm1d, convert_model; batchnorm.py:56-195, 399-444).

Each rank reduces its rows to (sum x, sum x^2) with one kernel, the 2C-float message is all-reduced over
torch.distributed (NCCL / NVLink; gloo in CPU tests of the host logic), and the normalise kernel uses the global
statistics.  Backward exchanges (sum dy, sum dy*xhat) the same way.  ``var_mode``: "add_eps" = 1/sqrt(var+eps)
(torch semantics), "clamp" = clamp(var, eps)^-1/2 (the vendored library, batchnorm.py:195)."""
import torch
import torch.distributed as dist
from torch import nn

from .. import ops
from .._lib import c_f32, c_i32, check, lib, ptr, stream
from ..ops import F32


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class _SyncBNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps, var_mode, module):
        ops.require_cuda(x)
        x = x.contiguous()
        R, C = x.shape
        sums = torch.empty(2 * C + 1, device=x.device, dtype=F32)
        check(lib.sml_bn_stats(ptr(x), ptr(sums), c_i32(R), c_i32(C), stream()), "bn_stats")
        sums[2 * C] = float(R)
        if _world() > 1:
            dist.all_reduce(sums)
        n = sums[2 * C]
        mean = sums[:C] / n
        sumvar = sums[C:2 * C] - sums[:C] * mean
        bias_var = sumvar / n
        invstd = (bias_var.clamp(min=eps) ** -0.5) if var_mode == "clamp" else torch.rsqrt(bias_var + eps)
        if module is not None and module.track_running_stats:
            with torch.no_grad():
                m = module.momentum
                module.running_mean.mul_(1 - m).add_(m * mean)
                module.running_var.mul_(1 - m).add_(m * sumvar / (n - 1))
                module.num_batches_tracked += 1
        mean, invstd = mean.contiguous(), invstd.contiguous()
        y = torch.empty_like(x)
        check(lib.sml_bn_fwd(ptr(x), ptr(mean), ptr(invstd), ptr(w), ptr(b), ptr(y), c_i32(R), c_i32(C), stream()), "bn_fwd")
        ctx.save_for_backward(x, w, mean, invstd, n.reshape(1))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, invstd, n = ctx.saved_tensors
        R, C = x.shape
        dy = dy.contiguous()
        sums = torch.empty(2 * C, device=x.device, dtype=F32)
        check(lib.sml_bn_bwd_stats(ptr(dy), ptr(x), ptr(mean), ptr(invstd), ptr(sums), c_i32(R), c_i32(C), stream()), "bn_bwd_stats")
        local = sums.clone()
        if _world() > 1:
            dist.all_reduce(sums)
        dx = torch.empty_like(x)
        check(lib.sml_bn_bwd(ptr(dy), ptr(x), ptr(mean), ptr(invstd), ptr(w), ptr(sums), c_f32(1.0 / float(n)), ptr(dx), c_i32(R),
                             c_i32(C), stream()), "bn_bwd")
        dw = local[C:] if w is not None else None           # per-rank parameter grads (DDP averages them)
        db = local[:C] if w is not None else None
        return dx, dw, db, None, None, None


class SyncBatchNorm1d(nn.BatchNorm1d):
    """state_dict-compatible with nn.BatchNorm1d (weight, bias, running_mean, running_var, num_batches_tracked)."""

    var_mode = "add_eps"

    def forward(self, x):
        if x.dim() != 2:
            raise NotImplementedError("sml_b200 SyncBatchNorm1d handles [B, C] inputs (the only BN site: fusion.py:29-32)")
        if self.training:
            return _SyncBNFn.apply(x.float(), self.weight, self.bias, self.eps, self.var_mode, self)
        mean = self.running_mean
        invstd = (self.running_var.clamp(min=self.eps) ** -0.5) if self.var_mode == "clamp" else torch.rsqrt(self.running_var + self.eps)
        return _EvalBNFn.apply(x.float(), self.weight, self.bias, mean.contiguous(), invstd.contiguous())


class _EvalBNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, mean, invstd):
        ops.require_cuda(x)
        x = x.contiguous()
        R, C = x.shape
        y = torch.empty_like(x)
        check(lib.sml_bn_fwd(ptr(x), ptr(mean), ptr(invstd), ptr(w), ptr(b), ptr(y), c_i32(R), c_i32(C), stream()), "bn_fwd")
        ctx.save_for_backward(x, w, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, invstd = ctx.saved_tensors
        xhat = (x - mean) * invstd
        return dy * w * invstd, (dy * xhat).sum(0), dy.sum(0), None, None


class SynchronizedBatchNorm1d(SyncBatchNorm1d):
    """Vendored-library semantics (clamp(var, eps)^-1/2; utils/sync_batchnorm/batchnorm.py:168-195)."""
    var_mode = "clamp"


def convert_model(module):
    """Replace every nn.BatchNorm1d by SynchronizedBatchNorm1d (utils/sync_batchnorm/batchnorm.py:399-444)."""
    mod = module
    if isinstance(module, nn.BatchNorm1d) and not isinstance(module, SyncBatchNorm1d):
        mod = SynchronizedBatchNorm1d(module.num_features, module.eps, module.momentum, module.affine, module.track_running_stats)
        mod.load_state_dict(module.state_dict())
        mod.to(module.weight.device if module.affine else module.running_mean.device)
    for name, child in module.named_children():
        new = convert_model(child)
        if new is not child:
            setattr(mod, name, new)
    return mod

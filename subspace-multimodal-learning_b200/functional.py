"""Autograd functions for the glue ops of the path (Linear(+ReLU), LayerNorm, PPEG, dropout+residual).

Every forward/backward is a C-ABI kernel call; there is no torch-eager arithmetic fallback.
"""
import torch

from . import ops
from ._lib import c_i32, check, lib, ptr, stream
from .ops import BF16, F32, Split, gemm, to_bf16

# Counter-based dropout state (Dropout, AlphaDropout, attention dropout, the TransLayer skip): every mask is a pure function of
# (seed, offset).  Like the reference, which draws from torch's generator, the stream follows ``torch.manual_seed``: on first
# use (and whenever torch's seed changes) the seed is derived from ``torch.initial_seed()`` and the distributed rank, so seed
# sweeps differ, DDP ranks draw different masks, and a run resumed with ``set_rng_state(get_rng_state())`` continues the stream.
_seed_state = {"seed": None, "offset": 0, "torch_seed": None, "explicit": False}


def _rank():
    try:
        import torch.distributed as dist
        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    except Exception:
        return 0


def _sync_with_torch():
    if _seed_state["explicit"]:
        return
    ts = int(torch.initial_seed())
    if _seed_state["torch_seed"] != ts:
        _seed_state["torch_seed"] = ts
        _seed_state["seed"] = (ts * 0x9E3779B97F4A7C15 + 0x5EED + 7919 * _rank()) & 0xFFFFFFFFFFFF
        _seed_state["offset"] = 0


def manual_seed(seed):
    """pin the dropout stream explicitly (tests, benchmarks); ``manual_seed(None)`` returns to following torch.manual_seed"""
    if seed is None:
        _seed_state.update(explicit=False, torch_seed=None)
        return
    _seed_state.update(seed=int(seed) & 0xFFFFFFFFFFFF, offset=0, explicit=True)


def current_seed():
    _sync_with_torch()
    return _seed_state["seed"]


def get_rng_state():
    """checkpointable state of the dropout stream (save next to the optimizer state)"""
    _sync_with_torch()
    return dict(_seed_state)


def set_rng_state(state):
    _seed_state.update(state)


def _next_offset(n):
    _sync_with_torch()
    o = _seed_state["offset"]
    _seed_state["offset"] = o + int(n)
    return o


def _prep(t, mode):
    """GEMM operand of `t` for the arithmetic mode.  A bf16 tensor (bf16-packed bags, data/bags.py) is its own operand in
    "bf16" mode -- the rounding the mode would apply to fp32 data has already happened at packing time -- and is widened
    (exactly) for the fp32-grade modes."""
    if t.dtype == BF16:
        if mode == "bf16":
            return Split(t.contiguous())
        t = t.float()
    return t if mode == "fp32" else to_bf16(t.contiguous(), split=(mode == "bf16x3"))


def _T(op):
    return op.transpose(-1, -2)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b, optional fused ReLU.  x [..., K] -> [..., N]."""

    @staticmethod
    def forward(ctx, x, w, b, relu, mode):
        ops.require_cuda(x, w, b)
        shp = x.shape
        x2 = x.contiguous().view(-1, shp[-1])
        y = gemm(_prep(x2, mode), _T(_prep(w.detach().contiguous(), mode)), bias=None if b is None else b.detach(), relu=relu)
        ctx.save_for_backward(x2, w, y if relu else None)
        ctx.cfg = (relu, mode, shp, b is not None)
        return y.view(*shp[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        relu, mode, shp, has_b = ctx.cfg
        x2, w, y = ctx.saved_tensors
        dy2 = dy.contiguous().view(-1, w.shape[0])
        if relu:
            g = torch.empty_like(dy2)
            check(lib.sml_relu_bwd(ptr(y), ptr(dy2), ptr(g), ops.c_i64(g.numel()), stream()), "relu_bwd")
            dy2 = g
        dy_op = _prep(dy2, mode)
        dx = gemm(dy_op, _prep(w.detach().contiguous(), mode)).view(shp) if ctx.needs_input_grad[0] else None
        dw = gemm(_T(dy_op), _prep(x2, mode))
        db = ops.colsum(dy2) if has_b else None
        return dx, dw, db, None, None


def linear(x, w, b=None, relu=False, mode="fp32"):
    return LinearFn.apply(x, w, b, relu, mode)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        ops.require_cuda(x, w, b)
        shp = x.shape
        x2 = x.contiguous().view(-1, shp[-1])
        y = torch.empty_like(x2)
        R = x2.shape[0]
        mean, rstd = ops.layernorm_fwd(x2, w.detach(), b.detach(), y, R, R, 0, eps)
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.shp = shp
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, w, mean, rstd = ctx.saved_tensors
        R = x2.shape[0]
        dx, dw, db = ops.layernorm_bwd(dy.contiguous().view(R, -1), x2, w.detach(), mean, rstd, R, R, 0)
        return dx.view(ctx.shp), dw, db, None


def layer_norm(x, w, b, eps=1e-5):
    return LayerNormFn.apply(x, w, b, eps)


class DropoutAddFn(torch.autograd.Function):
    """y = x + dropout(u, p)  (x may be None).  The mask is regenerated from (seed, offset) in backward."""

    @staticmethod
    def forward(ctx, x, u, p, seed, offset):
        ops.require_cuda(u)
        y = ops.dropout_add(None if x is None else x.contiguous(), u.contiguous(), p, seed, offset)
        ctx.cfg = (p, seed, offset, x is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        p, seed, offset, has_x = ctx.cfg
        dy = dy.contiguous()
        du = ops.dropout_add(None, dy, p, seed, offset) if p > 0 else dy
        return (dy if has_x else None), du, None, None, None


def dropout_add(x, u, p, training):
    p = float(p) if training else 0.0
    if p == 0.0 and x is None:
        return u
    off = _next_offset(u.numel()) if p > 0 else 0
    return DropoutAddFn.apply(x, u, p, current_seed(), off)


class AlphaDropoutFn(torch.autograd.Function):
    """torch.nn.AlphaDropout (MaxNet encoders, model.py:152-170); the mask is regenerated from (seed, offset) in backward."""

    @staticmethod
    def forward(ctx, x, p, seed, offset):
        ops.require_cuda(x)
        x = x.contiguous()
        y = torch.empty_like(x)
        check(lib.sml_alpha_dropout(ptr(x), ptr(y), ops.c_i64(x.numel()), ops.c_f32(p), c_i32(0), ops.c_u64(seed), ops.c_u64(offset),
                                    stream()), "alpha_dropout")
        ctx.cfg = (p, seed, offset)
        return y

    @staticmethod
    def backward(ctx, dy):
        p, seed, offset = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        check(lib.sml_alpha_dropout(ptr(dy), ptr(dx), ops.c_i64(dy.numel()), ops.c_f32(p), c_i32(1), ops.c_u64(seed),
                                    ops.c_u64(offset), stream()), "alpha_dropout_bwd")
        return dx, None, None, None


def alpha_dropout(x, p, training):
    if not training or p == 0.0:
        return x
    return AlphaDropoutFn.apply(x, float(p), current_seed(), _next_offset(x.numel()))


class PPEGFn(torch.autograd.Function):
    """x [B,1+H*W,C] -> cls row copied, tokens get the folded 7x7 depthwise stencil (mil.py:192-206).
    w7 [49,C] (tap-major) and bias [C] are the folded parameters (see models/mil.py:PPEG.folded)."""

    @staticmethod
    def forward(ctx, x, w7, bias, H, W):
        ops.require_cuda(x, w7, bias)
        x = x.contiguous()
        B, T, C = x.shape
        assert T == 1 + H * W
        y = torch.empty_like(x)
        w7c = w7.detach().contiguous()
        bc = bias.detach().contiguous()
        check(lib.sml_ppeg_fwd(ptr(x), ptr(w7c), ptr(bc), ptr(y), c_i32(B), c_i32(H), c_i32(W), c_i32(C), stream()),
              "ppeg_fwd")
        ctx.save_for_backward(x, w7c)
        ctx.cfg = (B, H, W, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, H, W, C = ctx.cfg
        x, w7 = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        nb = lib.sml_ppeg_bwd_blocks(c_i32(B), c_i32(H), c_i32(W))
        dwp = torch.empty(nb, 49 * C, device=x.device, dtype=F32)
        dbp = torch.empty(nb, C, device=x.device, dtype=F32)
        check(lib.sml_ppeg_bwd(ptr(x), ptr(w7), ptr(dy), ptr(dx), ptr(dwp), ptr(dbp), c_i32(B), c_i32(H), c_i32(W), c_i32(C),
                               stream()), "ppeg_bwd")
        return dx, ops.colsum(dwp).view(49, C), ops.colsum(dbp), None, None


class PPEGFoldFn(torch.autograd.Function):
    """(w7, b7, w5, b5, w3, b3) of the three depthwise convs -> folded tap-major stencil [49,C] (+ identity) and bias [C];
    one kernel each way instead of ~15 small torch ops in forward and as many in backward."""

    @staticmethod
    def forward(ctx, w7, b7, w5, b5, w3, b3):
        ops.require_cuda(w7, w5, w3)
        C = w7.shape[0]
        wf = torch.empty(49, C, device=w7.device, dtype=F32)
        bf = torch.empty(C, device=w7.device, dtype=F32)
        w7c, w5c, w3c = w7.detach().contiguous(), w5.detach().contiguous(), w3.detach().contiguous()   # locals outlive the launch
        b7c, b5c, b3c = b7.detach().contiguous(), b5.detach().contiguous(), b3.detach().contiguous()
        check(lib.sml_ppeg_fold(ptr(w7c), ptr(w5c), ptr(w3c), ptr(b7c), ptr(b5c), ptr(b3c), ptr(wf), ptr(bf), c_i32(C), stream()),
              "ppeg_fold")
        ctx.shapes = (w7.shape, w5.shape, w3.shape)
        return wf, bf

    @staticmethod
    def backward(ctx, dwf, dbf):
        s7, s5, s3 = ctx.shapes
        C = s7[0]
        dw7 = torch.empty(s7, device=dwf.device, dtype=F32)
        dw5 = torch.empty(s5, device=dwf.device, dtype=F32)
        dw3 = torch.empty(s3, device=dwf.device, dtype=F32)
        dwfc = dwf.contiguous()
        check(lib.sml_ppeg_unfold(ptr(dwfc), ptr(dw7), ptr(dw5), ptr(dw3), c_i32(C), stream()), "ppeg_unfold")
        return dw7, dbf, dw5, dbf, dw3, dbf


def ppeg_fold(w7, b7, w5, b5, w3, b3):
    return PPEGFoldFn.apply(w7, b7, w5, b5, w3, b3)


def ppeg(x, w7, bias, H, W):
    return PPEGFn.apply(x, w7, bias, H, W)


class AddRowVecFn(torch.autograd.Function):
    """y[b,t,:] = x[b,t,:] + vec[b,:]  (x [B,T,C], vec [B,C])."""

    @staticmethod
    def forward(ctx, x, vec):
        ops.require_cuda(x, vec)
        x, vec = x.contiguous(), vec.contiguous()
        B, T, C = x.shape
        y = torch.empty_like(x)
        check(lib.sml_add_rowvec(ptr(y), ptr(x), ptr(vec), ops.c_f32(1.0), ops.c_i64(B), ops.c_i64(T), c_i32(C), stream()), "add_rowvec")
        ctx.shape = (B, T, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, T, C = ctx.shape
        dy = dy.contiguous()
        dvec = torch.empty(B, C, device=dy.device, dtype=F32)
        for b in range(B):
            ops.colsum(dy[b], out=dvec[b])
        return dy, dvec


def add_rowvec(x, vec):
    return AddRowVecFn.apply(x, vec)


class MeanTokensFn(torch.autograd.Function):
    """x [B,T,C] -> mean over T [B,C]."""

    @staticmethod
    def forward(ctx, x):
        ops.require_cuda(x)
        x = x.contiguous()
        B, T, C = x.shape
        y = torch.empty(B, C, device=x.device, dtype=F32)
        for b in range(B):
            ops.colsum(x[b], out=y[b])
        ops.axpby_(y, y.clone(), 1.0 / T, 0.0)
        ctx.shape = (B, T, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, T, C = ctx.shape
        dx = torch.empty(B, T, C, device=dy.device, dtype=F32)
        dyc = dy.contiguous()
        check(lib.sml_add_rowvec(ptr(dx), None, ptr(dyc), ops.c_f32(1.0 / T), ops.c_i64(B), ops.c_i64(T), c_i32(C),
                                 stream()), "add_rowvec")
        return dx


def mean_tokens(x):
    return MeanTokensFn.apply(x)


_UNARY = {"tanh": 0, "sigmoid": 1, "relu": 2, "elu": 3}


class UnaryFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, op):
        ops.require_cuda(x)
        x = x.contiguous()
        y = torch.empty_like(x)
        check(lib.sml_unary_fwd(c_i32(op), ptr(x), ptr(y), ops.c_i64(x.numel()), stream()), "unary_fwd")
        ctx.save_for_backward(x, y)
        ctx.op = op
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dx = torch.empty_like(x)
        dyc = dy.contiguous()
        check(lib.sml_unary_bwd(c_i32(ctx.op), ptr(x), ptr(y), ptr(dyc), ptr(dx), ops.c_i64(x.numel()), stream()), "unary_bwd")
        return dx, None


def unary(x, name):
    return UnaryFn.apply(x, _UNARY[name])


class MatmulFn(torch.autograd.Function):
    """C = A @ B on strided fp32 views ([..., M, K] @ [..., K, N], <= 2 equal batch dims)."""

    @staticmethod
    def forward(ctx, A, B, alpha):
        ops.require_cuda(A, B)
        assert A.shape[:-2] == B.shape[:-2], "MatmulFn needs equal batch dims"
        ctx.save_for_backward(A, B)
        ctx.alpha = alpha
        return gemm(A, B, alpha=alpha)

    @staticmethod
    def backward(ctx, dC):
        A, B = ctx.saved_tensors
        dC = dC.contiguous()
        dA = gemm(dC, B.transpose(-1, -2), alpha=ctx.alpha) if ctx.needs_input_grad[0] else None
        dB = gemm(A.transpose(-1, -2), dC, alpha=ctx.alpha) if ctx.needs_input_grad[1] else None
        return dA, dB, None


def matmul(A, B, alpha=1.0):
    return MatmulFn.apply(A, B, alpha)


class SoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ops.require_cuda(x)
        y = x.contiguous().clone()
        ops.softmax_(y.view(-1, y.shape[-1]))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        g = dy.contiguous().clone()
        ops.softmax_bwd_(y.view(-1, y.shape[-1]), g.view(-1, y.shape[-1]))
        return g


def softmax_lastdim(x):
    return SoftmaxFn.apply(x)


class MulFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ops.require_cuda(a, b)
        a, b = a.contiguous(), b.contiguous()
        ctx.save_for_backward(a, b)
        return _mul(a, b)

    @staticmethod
    def backward(ctx, dy):
        a, b = ctx.saved_tensors
        dy = dy.contiguous()
        return _mul(dy, b), _mul(dy, a)


def _mul(a, b):
    y = torch.empty_like(a)
    check(lib.sml_mul(ptr(a), ptr(b), ptr(y), ops.c_i64(a.numel()), stream()), "mul")
    return y


def mul(a, b):
    return MulFn.apply(a, b)


class ScaleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a):
        ctx.a = a
        x = x.contiguous()
        return ops.axpby_(torch.empty_like(x), x, a, 0.0)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        return ops.axpby_(torch.empty_like(dy), dy, ctx.a, 0.0), None


def scale(x, a):
    return ScaleFn.apply(x, float(a))


class Fc1ClsFn(torch.autograd.Function):
    """TransMIL front end in one pass (mil.py:227-240): h = [cls | relu(x W^T + b) | first `add` rows again].
    The GEMM writes straight into the [B, 1+n+add, D] buffer through a strided output view (no torch.cat copy)."""

    @staticmethod
    def forward(ctx, x, w, b, cls, add, mode):
        ops.require_cuda(x, w, b, cls)
        B, n, K = x.shape
        D = w.shape[0]
        x = x.contiguous()
        if x.dtype == BF16 and mode != "bf16":
            x = x.float()
        h = torch.empty(B, 1 + n + add, D, device=x.device, dtype=F32)
        xo = _prep(x.view(B * n, K), mode)
        wo = _prep(w.detach().contiguous(), mode)
        xo3 = xo if mode == "fp32" else xo.map(lambda t: t.view(B, n, K))
        gemm(xo3 if mode != "fp32" else x, _T(wo), out=h[:, 1:1 + n], bias=b.detach(), relu=True)
        h[:, 0] = cls.detach().reshape(1, D)
        if add:
            h[:, 1 + n:] = h[:, 1:1 + add]
        ctx.save_for_backward(x, w, h)
        ctx.cfg = (B, n, K, D, add, mode, xo if mode != "fp32" else None)
        return h

    @staticmethod
    def backward(ctx, dh):
        B, n, K, D, add, mode, xo = ctx.cfg
        x, w, h = ctx.saved_tensors
        dh = dh.contiguous()
        dcls = dh[:, 0].sum(0).reshape(1, 1, D)
        dmain = dh[:, 1:1 + n]
        if add:
            dmain = dmain.clone()
            dmain[:, :add] += dh[:, 1 + n:]
        g = torch.empty(B, n, D, device=dh.device, dtype=F32)
        for bi in range(B):
            hb, db_ = h[bi, 1:1 + n], dmain[bi].contiguous()       # bound to a local: alive until the launch is enqueued
            check(lib.sml_relu_bwd(ptr(hb), ptr(db_), ptr(g[bi]), ops.c_i64(n * D), stream()), "relu_bwd")
        g2 = g.view(B * n, D)
        g_op = _prep(g2, mode)
        x_op = xo if xo is not None else x.view(B * n, K)
        dw = gemm(_T(g_op), x_op)
        db = ops.colsum(g2)
        dx = gemm(g_op, _prep(w.detach().contiguous(), mode)).view(B, n, K) if ctx.needs_input_grad[0] else None
        return dx, dw, db, dcls, None, None


def fc1_cls(x, w, b, cls, add, mode="fp32"):
    return Fc1ClsFn.apply(x, w, b, cls, add, mode)

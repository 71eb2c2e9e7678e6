"""Nystrom attention forward/backward on the C-ABI kernels (torch.autograd.Function).

Implements /root/reference/models/NystromAttention.py:74-157 (+ pinv :20-35) with mask=None.
Differences in *schedule* only (results match the reference within the tolerances in tests/):
  * q-scale folded into the q rows of W_qkv (exact for dim_head = 64: scale = 2^-3);
  * (A1 @ Z) @ (A3 @ V) is evaluated as A1 @ (Z @ (A3 @ V))  (4x fewer FLOPs);
  * the three similarity softmaxes over the landmark axis are fused into the GEMM epilogue on the
    tensor-core engine; identity terms of the pinv iteration are GEMM-epilogue diagonals.

Precision modes (``precision`` argument):
  "fp32"   — SIMT fp32 FFMA engine everywhere (device-side exact-fp32 path);
  "bf16x3" — tcgen05 engine with the 3-pass bf16 split on every contraction (fp32-grade, ~2^-16);
  "bf16"   — tcgen05 single-pass bf16 operands, fp32 accumulate; the pinv chain stays bf16x3.
"""
import torch

from . import ops
from ._lib import c_i32, c_vp, check, lib, ptr, stream
from .ops import F32, Split, gemm, to_bf16

PRECISIONS = ("fp32", "bf16x3", "bf16")


def _prep(t, mode, hi_precision=False):
    """contiguous fp32 tensor -> GEMM operand for `mode`."""
    if mode == "fp32":
        return t
    return to_bf16(t.contiguous(), split=(mode == "bf16x3" or hi_precision))


def _view(op, f):
    return op.map(f) if isinstance(op, Split) else f(op)


def _T(op):
    return _view(op, lambda t: t.transpose(-1, -2))


def landmarks_fwd(qkv, b, npad, h, d, m):
    ql = torch.empty(b, h, m, d, device=qkv.device, dtype=F32)
    kl = torch.empty_like(ql)
    check(lib.sml_nystrom_landmarks_fwd(ptr(qkv), ptr(ql), ptr(kl), c_i32(b), c_i32(npad), c_i32(h), c_i32(d), c_i32(m),
                                        stream()), "landmarks_fwd")
    return ql, kl


def landmark_ids(npad, m, device):
    ids = torch.empty(npad, device=device, dtype=torch.int32)
    check(lib.sml_nystrom_landmark_ids(ptr(ids), c_i32(npad), c_i32(m), stream()), "landmark_ids")
    return ids


def pinv_fwd(a2, iters, mode):
    """Moore-Penrose iteration (NystromAttention.py:20-35).  Returns z and what backward needs."""
    bh, m = a2.numel() // (a2.shape[-1] ** 2), a2.shape[-1]
    dev = a2.device
    scal = torch.empty(2, device=dev, dtype=F32)
    arg = torch.empty(2, device=dev, dtype=torch.int32)
    ws = torch.empty(2, device=dev, dtype=torch.int64)
    check(lib.sml_pinv_scale(ptr(a2), ptr(scal), ptr(arg), ptr(ws), c_i32(bh), c_i32(m), stream()), "pinv_scale")
    z = torch.empty_like(a2)
    check(lib.sml_pinv_init(ptr(a2), ptr(scal), ptr(z), c_i32(bh), c_i32(m), stream()), "pinv_init")
    zs = [z]
    hp = mode != "fp32"
    pa2 = _prep(a2, mode, True)
    for _ in range(iters):
        z = zs[-1]
        z_op = _prep(z, mode, True)
        xz = gemm(pa2, z_op)                                        # x z
        xz_op = _prep(xz, mode, True)
        t1 = ops.diag_axpby(xz, -1.0, 7.0)                          # 7I - xz
        t2 = gemm(xz_op, _prep(t1, mode, True), alpha=-1.0, diag=15.0)   # 15I - xz t1
        t3 = gemm(xz_op, _prep(t2, mode, True), alpha=-1.0, diag=13.0)   # 13I - xz t2
        zs.append(gemm(z_op, _prep(t3, mode, True), alpha=0.25))    # z <- 1/4 z t3
    del hp
    return zs[-1], (zs[:-1], scal, arg)


def pinv_bwd(a2, saved, dz, mode):
    """d a2 given d z_final; recomputes xz, t1..t3 per iteration from the saved z_k."""
    zs, scal, arg = saved
    bh, m = a2.numel() // (a2.shape[-1] ** 2), a2.shape[-1]
    pa2 = _prep(a2, mode, True)
    da2 = torch.zeros_like(a2)
    g = dz
    for z in reversed(zs):
        z_op = _prep(z, mode, True)
        xz = gemm(pa2, z_op)
        xz_op = _prep(xz, mode, True)
        t1 = ops.diag_axpby(xz, -1.0, 7.0)
        t1_op = _prep(t1, mode, True)
        t2 = gemm(xz_op, t1_op, alpha=-1.0, diag=15.0)
        t2_op = _prep(t2, mode, True)
        t3 = gemm(xz_op, t2_op, alpha=-1.0, diag=13.0)
        g_op = _prep(g, mode, True)
        dzk = gemm(g_op, _T(_prep(t3, mode, True)), alpha=0.25)             # 1/4 G t3^T
        dt3 = gemm(_T(z_op), g_op, alpha=0.25)                              # 1/4 z^T G
        dt3_op = _prep(dt3, mode, True)
        dxz = gemm(dt3_op, _T(t2_op), alpha=-1.0)                           # -dt3 t2^T
        dt2 = gemm(_T(xz_op), dt3_op, alpha=-1.0)                           # -xz^T dt3
        dt2_op = _prep(dt2, mode, True)
        gemm(dt2_op, _T(t1_op), out=dxz, alpha=-1.0, beta=1.0)              # -dt2 t1^T
        dt1 = gemm(_T(xz_op), dt2_op, alpha=-1.0)                           # -xz^T dt2
        ops.axpby_(dxz, dt1, -1.0, 1.0)                                     # t1 = 7I - xz
        dxz_op = _prep(dxz, mode, True)
        gemm(dxz_op, _T(z_op), out=da2, beta=1.0)                           # dx += dxz z^T
        gemm(_T(pa2), dxz_op, out=dzk, beta=1.0)                            # dz += x^T dxz
        g = dzk
    tmp = torch.empty(1, device=a2.device, dtype=F32)
    check(lib.sml_pinv_init_bwd(ptr(a2), ptr(scal), ptr(arg), ptr(g), ptr(da2), ptr(tmp), c_i32(bh), c_i32(m), stream()),
          "pinv_init_bwd")
    return da2


class NystromAttentionFn(torch.autograd.Function):
    """y = NystromAttention(x) without the trailing dropout.  x [b,n,D] fp32 CUDA."""

    @staticmethod
    def forward(ctx, x, w_qkv, w_out, b_out, w_conv, heads, m, iters, mode):
        assert mode in PRECISIONS
        ops.require_cuda(x, w_qkv, w_out, b_out, w_conv)
        b, n, D = x.shape
        inner = w_qkv.shape[0] // 3
        h, d = heads, inner // heads
        pad = (m - n % m) % m
        npad = n + pad
        dev = x.device
        x = x.contiguous()
        if pad:
            xp = torch.zeros(b, npad, D, device=dev, dtype=F32)        # front pad (:79-82)
            xp[:, pad:] = x
        else:
            xp = x
        weff = w_qkv.detach().clone()
        ops.axpby_(weff[:inner], w_qkv.detach()[:inner].contiguous(), d ** -0.5, 0.0)    # fold q * scale (:98)
        tc = mode != "fp32"
        xp_op = _prep(xp.view(b * npad, D), mode)
        weff_op = _prep(weff, mode)
        qkv = torch.empty(b, npad, 3, h, d, device=dev, dtype=F32)
        gemm(xp_op, _T(weff_op), out=qkv.view(b * npad, 3 * inner))                  # :89
        ql, kl = landmarks_fwd(qkv, b, npad, h, d, m)                                # :102-118
        qkv_op = _prep(qkv, mode)
        hv = lambda t, i: t[:, :, i].permute(0, 2, 1, 3)                             # [b,h,npad,d] view
        q_op, k_op, v_op = (_view(qkv_op, lambda t, i=i: hv(t, i)) for i in range(3))
        ql_op, kl_op = _prep(ql, mode), _prep(kl, mode)
        a1 = torch.empty(b, h, npad, m, device=dev, dtype=F32)
        a2 = torch.empty(b, h, m, m, device=dev, dtype=F32)
        a3 = torch.empty(b, h, m, npad, device=dev, dtype=F32)
        fuse = tc and m <= 256
        gemm(q_op, _T(kl_op), out=a1, softmax=fuse)                                  # sim1 (+softmax)
        gemm(ql_op, _T(kl_op), out=a2, softmax=fuse)                                 # sim2
        gemm(ql_op, _T(k_op), out=a3)                                                # sim3
        if not fuse:
            ops.softmax_(a1.view(-1, m))
            ops.softmax_(a2.view(-1, m))
        ops.softmax_(a3.view(-1, npad))                                              # softmax over tokens
        z, pinv_saved = pinv_fwd(a2, iters, mode)                                    # :138
        a3_op = _prep(a3, mode)
        a3v = gemm(a3_op, v_op)                                                      # [b,h,m,d]
        wm = gemm(_prep(z, mode), _prep(a3v, mode))                                  # Z (A3 V)
        merged = torch.empty(b, npad, inner, device=dev, dtype=F32)
        mview = merged.view(b, npad, h, d).permute(0, 2, 1, 3)
        a1_op = _prep(a1, mode)
        gemm(a1_op, _prep(wm, mode), out=mview)                                      # A1 (Z A3 V)  :140
        ksize = w_conv.shape[2]
        wc = w_conv.detach().reshape(h, ksize).contiguous()
        check(lib.sml_nystrom_resconv_fwd(ptr(qkv), ptr(wc), ptr(merged), c_i32(b), c_i32(npad), c_i32(h), c_i32(d),
                                          c_i32(ksize), stream()), "resconv_fwd")   # :144-145
        y = torch.empty(b, n, D, device=dev, dtype=F32)
        m_op = _prep(merged, mode)
        gemm(_view(m_op, lambda t: t[:, pad:]), _T(_prep(w_out.detach().contiguous(), mode)), out=y, bias=b_out.detach())  # :150-151
        ctx.save_for_backward(xp, weff, w_out, qkv, ql, kl, a1, a2, a3, z, a3v, wm, merged, wc, *pinv_saved[0],
                              pinv_saved[1], pinv_saved[2])
        ctx.cfg = (b, n, D, h, d, m, iters, pad, npad, inner, ksize, mode)
        return y

    @staticmethod
    def backward(ctx, dy):
        b, n, D, h, d, m, iters, pad, npad, inner, ksize, mode = ctx.cfg
        sv = ctx.saved_tensors
        xp, weff, w_out, qkv, ql, kl, a1, a2, a3, z, a3v, wm, merged, wc = sv[:14]
        zs, scal, arg = list(sv[14:14 + iters]), sv[14 + iters], sv[15 + iters]
        dev = dy.device
        dy = dy.contiguous()
        if pad:
            dyp = torch.zeros(b, npad, D, device=dev, dtype=F32)
            dyp[:, pad:] = dy
        else:
            dyp = dy
        dyp2 = dyp.view(b * npad, D)
        dy_op = _prep(dyp2, mode)
        hv = lambda t, i: t[:, :, i].permute(0, 2, 1, 3)
        # ---- to_out
        db_out = ops.colsum(dyp2)
        dw_out = gemm(_T(dy_op), _prep(merged.view(b * npad, inner), mode))                  # [D, inner]
        dmerged = gemm(dy_op, _prep(w_out.detach().contiguous(), mode)).view(b, npad, inner)
        dout_h = dmerged.view(b, npad, h, d).permute(0, 2, 1, 3)
        dm_op = _prep(dmerged, mode)
        dout_op = _view(dm_op, lambda t: t.view(b, npad, h, d).permute(0, 2, 1, 3))
        qkv_op = _prep(qkv, mode)
        q_op, k_op, v_op = (_view(qkv_op, lambda t, i=i: hv(t, i)) for i in range(3))
        ql_op, kl_op = _prep(ql, mode), _prep(kl, mode)
        dqkv = torch.empty_like(qkv)
        dq, dk, dv = (hv(dqkv, i) for i in range(3))
        # ---- out = A1 Wm ; Wm = Z A3V ; A3V = A3 V
        a1_op, a3_op = _prep(a1, mode), _prep(a3, mode)
        da1 = gemm(dout_op, _T(_prep(wm, mode)))                                             # [b,h,npad,m]
        dwm = gemm(_T(a1_op), dout_op)                                                       # [b,h,m,d]
        dwm_op = _prep(dwm, mode)
        dz = gemm(dwm_op, _T(_prep(a3v, mode)))                                              # [b,h,m,m]
        da3v = gemm(_T(_prep(z, mode)), dwm_op)                                              # [b,h,m,d]
        da3v_op = _prep(da3v, mode)
        da3 = gemm(da3v_op, _T(v_op))                                                        # [b,h,m,npad]
        gemm(_T(a3_op), da3v_op, out=dv)                                                     # dv = A3^T dA3V
        dwc_part = torch.empty(lib.sml_nystrom_resconv_bwd_blocks(c_i32(b), c_i32(npad), c_i32(h * d)), h * ksize, device=dev, dtype=F32)
        check(lib.sml_nystrom_resconv_bwd(ptr(qkv), ptr(wc), ptr(dmerged), ptr(dqkv), ptr(dwc_part), c_i32(b), c_i32(npad),
                                          c_i32(h), c_i32(d), c_i32(ksize), stream()), "resconv_bwd")   # dv += conv^T
        dw_conv = ops.colsum(dwc_part).view(h, 1, ksize, 1)
        # ---- pinv + softmaxes
        da2 = pinv_bwd(a2, (zs, scal, arg), dz, mode)
        ds1 = ops.softmax_bwd_(a1.view(-1, m), da1.view(-1, m)).view(b, h, npad, m)
        ds2 = ops.softmax_bwd_(a2.view(-1, m), da2.view(-1, m)).view(b, h, m, m)
        ds3 = ops.softmax_bwd_(a3.view(-1, npad), da3.view(-1, npad)).view(b, h, m, npad)
        ds1_op, ds2_op, ds3_op = _prep(ds1, mode), _prep(ds2, mode), _prep(ds3, mode)
        # ---- similarities
        gemm(ds1_op, kl_op, out=dq)                                                          # dq = dS1 kl
        dkl = gemm(_T(ds1_op), q_op)                                                         # dkl = dS1^T q
        dql = gemm(ds2_op, kl_op)
        gemm(_T(ds2_op), ql_op, out=dkl, beta=1.0)
        gemm(ds3_op, k_op, out=dql, beta=1.0)
        gemm(_T(ds3_op), ql_op, out=dk)                                                      # dk = dS3^T ql
        check(lib.sml_nystrom_landmarks_bwd(ptr(dqkv), ptr(dql), ptr(dkl), c_i32(b), c_i32(npad), c_i32(h), c_i32(d),
                                            c_i32(m), stream()), "landmarks_bwd")
        # ---- qkv projection
        dqkv2 = dqkv.view(b * npad, 3 * inner)
        dqkv_op = _prep(dqkv2, mode)
        dxp = gemm(dqkv_op, _prep(weff, mode)).view(b, npad, D)
        dw_qkv = gemm(_T(dqkv_op), _prep(xp.view(b * npad, D), mode))                        # [3 inner, D]
        ops.axpby_(dw_qkv[:inner], dw_qkv[:inner].clone(), d ** -0.5, 0.0)
        dx = dxp[:, pad:]
        return dx, dw_qkv, dw_out, db_out, dw_conv, None, None, None, None


def nystrom_attention(x, w_qkv, w_out, b_out, w_conv, heads, num_landmarks, pinv_iterations, precision="fp32"):
    return NystromAttentionFn.apply(x, w_qkv, w_out, b_out, w_conv, heads, num_landmarks, pinv_iterations, precision)

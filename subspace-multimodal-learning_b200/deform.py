"""Deformable cross-attention (1-D / 2-D) forward/backward on the C-ABI kernels.

Implements /root/reference/models/DeformableAttention1D.py:156-240 and
models/DeformableAttention2D.py:223-325 on token-major tensors; the bias MLP (CPB) is never materialised.
"""
import ctypes

import torch

from . import ops
from ._lib import c_f32, c_i32, c_u64, check, lib, ptr, stream
from .ops import F32, Split, gemm, to_bf16


import os as _os
# "bf16" mode, 2-D, one head per offset group: SML_DEFORM_TC5=1 runs the forward as the first tcgen05 / TMEM kernel
# (csrc/deform_tc5.cu, one key per MMA round trip: measured 43 ms against 16.4 ms for the warp-level kernel at B=8/16k, so off)
TC5_FWD = _os.environ.get("SML_DEFORM_TC5", "0") != "0"
# "bf16" mode, 1-D: the bias MLP as its exact piecewise-linear table (csrc/cpb_table.cu) inside the fused attention kernels
PWL_1D = _os.environ.get("SML_DEFORM_PWL", "1") != "0"


def cpb1d_table(cpb, OUT):
    """device-built table of the 1-D bias MLP: (brk [slots], coef [slots + 1, OUT, 2], nseg int32 [1])"""
    slots = lib.sml_cpb1d_table_slots()
    brk = torch.empty(slots, device=cpb.device, dtype=F32)
    coef = torch.empty(slots + 1, OUT, 2, device=cpb.device, dtype=F32)
    nseg = torch.empty(1, device=cpb.device, dtype=torch.int32)
    check(lib.sml_cpb1d_table_build(ptr(cpb), c_i32(OUT), ptr(brk), ptr(coef), ptr(nseg), stream()), "cpb1d_table_build")
    return brk, coef, nseg


def key_ranges(acfg, vs):
    """(min, max) key position of every 32-key tile: scratch of the 1-D table kernels (one small launch per forward)"""
    kr = torch.empty(lib.sml_deform_pwl_krange_floats(acfg), device=vs.device, dtype=F32)
    check(lib.sml_deform_pwl_key_ranges(acfg, ptr(vs), ptr(kr), stream()), "deform_pwl_key_ranges")
    return kr


def _prof_begin():
    if ops.PROFILE is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


def _prof_end(e0, eng, flops, tag):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        ops.PROFILE.append((eng, flops, e0, e1, tag))


def _cfg(*vals):
    return (c_i32 * len(vals))(*[int(v) for v in vals])


def _op(t, mode):
    """GEMM operand for the engine of ``mode``: fp32 tensor (SIMT engine, "fp32") or bf16 hi + lo copies (tcgen05 engine,
    3-pass split).  Both tensor-core modes run the projections fp32-grade: q feeds the offset network, whose outputs decide
    the integer sampling taps (bit-exact contract), and the GEMMs are < 10 % of the step either way; "bf16" only changes the
    arithmetic inside the fused attention kernel."""
    return t if mode == "fp32" else to_bf16(t.contiguous(), split=True)


def _v(op, fn):
    return op.map(fn) if isinstance(op, Split) else fn(op)


def _block_diag(w2d, G):
    """[Cout, Cin/G] grouped 1x1 weight -> dense block-diagonal [Cout, Cin].  In the tensor-core modes a grouped conv with
    16 input channels per group is run as ONE dense GEMM against this matrix: the per-group form is a batch of K = 16
    products (one k-block per 128 x 64 tile: ~10 us of pipeline latency per tile, 594 us per call at 16 k tokens, measured)
    and its weight gradient a single CTA per group streaming all tokens (248 us x 8 slides); the dense form does 8x the
    FLOPs on zeros at full tensor-core rate and is bound by the same output bytes."""
    Cout, ci = w2d.shape
    co = Cout // G
    wbd = torch.zeros(G, co, G, ci, device=w2d.device, dtype=w2d.dtype)
    idx = torch.arange(G, device=w2d.device)
    wbd[idx, :, idx, :] = w2d.view(G, co, ci)
    return wbd.view(Cout, G * ci)


def gconv(x, w2d, G, out=None, mode="fp32"):
    """grouped 1x1 conv on token-major x [b,T,Cin]: w2d [Cout, Cin/G] -> [b,T,Cout]."""
    b, T, Cin = x.shape
    Cout = w2d.shape[0]
    if out is None:
        out = torch.empty(b, T, Cout, device=x.device, dtype=F32)
    if G > 1 and mode != "fp32":
        w2d, G = _block_diag(w2d, G), 1
    xo, wo = _op(x, mode), _op(w2d, mode)
    if G == 1:
        gemm(xo, _v(wo, lambda t: t.t()), out=out)
    else:
        ci, co = Cin // G, Cout // G
        gemm(_v(xo, lambda t: t.view(b, T, G, ci).permute(0, 2, 1, 3)), _v(wo, lambda t: t.view(1, G, co, ci).transpose(-1, -2)),
             out=out.view(b, T, G, co).permute(0, 2, 1, 3))
    return out


def gconv_bwd(x, w2d, G, dy, need_dx=True, dx_out=None, dx_beta=0.0, mode="fp32"):
    """returns (dx [b,T,Cin] or None, dw [Cout, Cin/G])."""
    b, T, Cin = x.shape
    Cout = w2d.shape[0]
    dx = None
    dense = G > 1 and mode != "fp32"                     # tensor-core modes: dense block-diagonal form (see _block_diag)
    Gd = G
    if dense:
        w2d, G = _block_diag(w2d, G), 1
    xo, dyo = _op(x, mode), _op(dy, mode)
    wo = _op(w2d, mode) if need_dx else None
    if G == 1:
        dw = gemm(_v(dyo, lambda t: t.reshape(b * T, Cout).t()), _v(xo, lambda t: t.reshape(b * T, Cin)))
        if need_dx:
            dx = dx_out if dx_out is not None else torch.empty_like(x)
            gemm(dyo, wo, out=dx, beta=dx_beta)
        if dense:                                        # keep the diagonal blocks of the dense [Cout, Cin] gradient
            ci, co = Cin // Gd, Cout // Gd
            idx = torch.arange(Gd, device=x.device)
            dw = dw.view(Gd, co, Gd, ci)[idx, :, idx, :].reshape(Cout, ci)
    else:
        ci, co = Cin // G, Cout // G
        # dw_g = sum over (b, t) of dy_g^T x_g: the batch is folded into the contraction ([b, T, C] is [b T, C] in memory)
        xg = _v(xo, lambda t: t.view(b * T, G, ci).permute(1, 0, 2))              # [G, b T, ci]
        dyg = _v(dyo, lambda t: t.view(b * T, G, co).permute(1, 0, 2))            # [G, b T, co]
        dw = gemm(_v(dyg, lambda t: t.transpose(-1, -2)), xg).view(Cout, ci)
        if need_dx:
            dx = dx_out if dx_out is not None else torch.empty_like(x)
            gemm(_v(dyo, lambda t: t.view(b, T, G, co).permute(0, 2, 1, 3)), _v(wo, lambda t: t.view(1, G, co, ci)),
                 out=dx.view(b, T, G, ci).permute(0, 2, 1, 3), beta=dx_beta)
    return dx, dw


class DeformAttnFn(torch.autograd.Function):
    """(x1t, x2t) [b,T,C] token-major -> (y [b,T,C], vgrid).  cfg: dict(dim2d, side, heads, groups, ds, ksize, scale,
    gq, gkv, p_drop, seed, mode, want_taps)."""

    @staticmethod
    def forward(ctx, x1t, x2t, wq, wk, wv, wout, bout, wc, bc, w2, cw1, cb1, cw2, cb2, cw3, cb3, cfg):
        ops.require_cuda(x1t, x2t)
        x1t, x2t = x1t.contiguous(), x2t.contiguous()
        b, T, C = x1t.shape
        heads, G, ds, ks = cfg["heads"], cfg["groups"], cfg["ds"], cfg["ksize"]
        inner = heads * 64
        gd = inner // G
        two_d = cfg["dim2d"]
        dev = x1t.device
        gq = G if cfg["gq"] else 1
        gkv = G if cfg["gkv"] else 1
        wq2, wk2, wv2 = (w.detach().reshape(w.shape[0], -1).contiguous() for w in (wq, wk, wv))
        wo2 = wout.detach().reshape(wout.shape[0], -1).contiguous()
        prec = cfg.get("precision", "bf16x3")
        assert prec in ("fp32", "bf16x3", "bf16"), prec
        q = gconv(x1t, wq2, gq, mode=prec)                                       # 1D.py:175 / 2D.py:246
        pad = (ks - ds) // 2
        if two_d:
            side = cfg["side"]
            assert side * side == T
            ocfg = _cfg(b, G, gd, inner, side, side, ks, ks, ds, ds, pad, pad, 2)
            hk = (side + 2 * pad - ks) // ds + 1
            nkv, OC = hk * hk, 2
            vshape = (b * G, 2, hk, hk)
            gcfg = _cfg(b, G, C // G, side, side, nkv, 2, 0)
        else:
            ocfg = _cfg(b, G, gd, inner, 1, T, 1, ks, 1, ds, 0, pad, 1)
            nkv, OC = (T + 2 * pad - ks) // ds + 1, 1
            vshape = (b * G, nkv)
            gcfg = (_cfg(b, G, C // G, T, 1, nkv, 1, 0) if cfg["mode"] == 0 else _cfg(b, G, C // G, T, 1, nkv, 1, 1))
        wc2 = wc.detach().reshape(gd, -1).contiguous()
        w22 = w2.detach().reshape(OC, gd).contiguous()
        off = torch.empty(b * G, nkv, OC, device=dev, dtype=F32)
        vgrid = torch.empty(vshape, device=dev, dtype=F32)
        vs = torch.empty(b * G, nkv, OC, device=dev, dtype=F32)
        check(lib.sml_deform_offsets_fwd(ocfg, c_f32(cfg["scale"]), ptr(q), ptr(wc2), ptr(bc.detach()), ptr(w22), ptr(off),
                                         ptr(vgrid), ptr(vs), stream()), "deform_offsets_fwd")
        kv = torch.empty(b, nkv, C, device=dev, dtype=F32)
        taps = torch.empty(b * G, nkv, 2, device=dev, dtype=torch.int32) if cfg.get("want_taps") else None
        check(lib.sml_deform_gather_fwd(gcfg, ptr(x2t), ptr(vs), ptr(kv), ptr(taps), stream()), "deform_gather_fwd")
        k = gconv(kv, wk2, gkv, mode=prec)                                       # 1D.py:199 / 2D.py:279
        v = gconv(kv, wv2, gkv, mode=prec)
        qs = torch.empty_like(q)
        ops.axpby_(qs, q, 64 ** -0.5, 0.0)                                       # q * scale after the offsets (1D.py:203)
        cpb = torch.cat([t.detach().reshape(-1) for t in (cw1, cb1, cw2, cb2, cw3, cb3)]).contiguous()
        acfg = _cfg(b, heads, T, nkv, G, OC, cfg["side"] if two_d else 1, 64)
        out = torch.empty(b, T, inner, device=dev, dtype=F32)
        lse = torch.empty(b, heads, T, device=dev, dtype=F32)
        pairs = float(b) * heads * T * nkv
        pwl = None
        prof = _prof_begin()
        if prec == "fp32":          # SIMT fp32 kernel
            check(lib.sml_deform_attn_fwd(acfg, c_f32(cfg["p_drop"]), c_u64(cfg["seed"]), ptr(qs), ptr(k), ptr(v), ptr(vs), ptr(cpb),
                                          ptr(out), ptr(lse), stream()), "deform_attn_fwd")
        elif prec == "bf16" and PWL_1D and not two_d:                  # 1-D: bias from the piecewise-linear table
            brk, coef, nseg = cpb1d_table(cpb, heads // G)
            kh, vh = ops.to_bf16(k).hi, ops.to_bf16(v).hi              # the mode's operand rounding, once per call
            kr = key_ranges(acfg, vs)
            pwl = (brk, coef, nseg, kh, vh, kr)
            check(lib.sml_deform_attn_fwd_pwl(acfg, c_f32(cfg["p_drop"]), c_u64(cfg["seed"]), ptr(qs), ptr(kh), ptr(vh), ptr(vs),
                                              ptr(kr), ptr(brk), ptr(coef), ptr(nseg), ptr(out), ptr(lse), stream()), "deform_attn_fwd_pwl")
        elif prec == "bf16" and TC5_FWD and two_d and heads == G:     # tcgen05 / TMEM kernel
            check(lib.sml_deform_attn_fwd_tc5(acfg, c_f32(cfg["p_drop"]), c_u64(cfg["seed"]), ptr(qs), ptr(k), ptr(v), ptr(vs), ptr(cpb),
                                              ptr(out), ptr(lse), stream()), "deform_attn_fwd_tc5")
        else:                       # register-fragment MMA kernel (bf16 hi/lo 3-pass or plain bf16 operands)
            check(lib.sml_deform_attn_fwd_tc(acfg, c_f32(cfg["p_drop"]), c_u64(cfg["seed"]), ptr(qs), ptr(k), ptr(v), ptr(vs), ptr(cpb),
                                             ptr(out), ptr(lse), c_i32(1 if prec == "bf16x3" else 0), stream()), "deform_attn_fwd_tc")
        # algorithmic work per (query, key) pair: bias MLP 2 (CD 32 + 32 32 + 32) + q.k and p.v 4 d  
        # (1-D table path: the MLP is not evaluated per pair; count q.k, p.v and ~8 table operations)
        _prof_end(prof, "deform_fwd", pairs * ((8 if pwl is not None else 2 * (OC * 32 + 32 * 32 + 32)) + 4 * 64),
                  (b, heads, T, nkv, OC, prec))
        y = torch.empty(b, T, C, device=dev, dtype=F32)
        gemm(_op(out.view(b * T, inner), prec), _v(_op(wo2, prec), lambda t: t.t()), out=y.view(b * T, C),
             bias=bout.detach())                                                               # 1D.py:233 / 2D.py:314
        ctx.save_for_backward(x1t, x2t, wq2, wk2, wv2, wo2, wc2, bc, w22, q, qs, off, vs, kv, k, v, cpb, out, lse)
        ctx.cfg = dict(cfg, ocfg=ocfg, gcfg=gcfg, acfg=acfg, nkv=nkv, OC=OC, gd=gd, gq=gq, gkv=gkv, inner=inner,
                       shapes=[t.shape for t in (wq, wk, wv, wout, wc, w2, cw1, cb1, cw2, cb2, cw3, cb3)])
        ctx.taps = taps
        ctx.pwl = pwl
        ctx.mark_non_differentiable(*([] if taps is None else []))
        return y, vgrid

    @staticmethod
    def backward(ctx, dy, dvgrid):
        cfg = ctx.cfg
        (x1t, x2t, wq2, wk2, wv2, wo2, wc2, bc, w22, q, qs, off, vs, kv, k, v, cpb, out, lse) = ctx.saved_tensors
        b, T, C = x1t.shape
        G, heads, inner, nkv, OC, gd = cfg["groups"], cfg["heads"], cfg["inner"], cfg["nkv"], cfg["OC"], cfg["gd"]
        dev = x1t.device
        dy = dy.contiguous()
        dy2 = dy.view(b * T, C)
        prec = cfg.get("precision", "bf16x3")
        db_out = ops.colsum(dy2)
        dy2o = _op(dy2, prec)
        dw_out = gemm(_v(dy2o, lambda t: t.t()), _op(out.view(b * T, inner), prec))
        dout = gemm(dy2o, _op(wo2, prec)).view(b, T, inner)
        dqs = torch.empty_like(q)
        dk = torch.zeros_like(k)
        dv = torch.zeros_like(v)
        dvs_b = torch.zeros_like(vs)
        nblk = lib.sml_deform_attn_bwd_blocks(cfg["acfg"])
        npar = lib.sml_deform_cpb_nparams(c_i32(OC), c_i32(heads // G))
        pwl = ctx.pwl
        part = torch.empty(nblk, npar, device=dev, dtype=F32) if pwl is None else None
        prof = _prof_begin()
        if pwl is not None:         # 1-D, bias from the table: per-segment sums, mapped to the MLP parameters afterwards
            brk, coef, nseg, kh, vh, kr = pwl
            gseg = torch.zeros(heads // G, lib.sml_cpb1d_seg_rows(), 2, device=dev, dtype=F32)
            dcpb = torch.empty(npar, device=dev, dtype=F32)
            check(lib.sml_deform_attn_bwd_pwl(cfg["acfg"], c_f32(cfg["p_drop"]), c_u64(cfg["seed"]), ptr(qs), ptr(kh), ptr(vh), ptr(vs),
                                              ptr(kr), ptr(brk), ptr(coef), ptr(nseg), ptr(out), ptr(lse), ptr(dout), ptr(dqs), ptr(dk),
                                              ptr(dv), ptr(dvs_b), ptr(gseg), stream()), "deform_attn_bwd_pwl")
            check(lib.sml_cpb1d_table_bwd(ptr(cpb), c_i32(heads // G), ptr(brk), ptr(nseg), ptr(gseg), ptr(dcpb), stream()),
                  "cpb1d_table_bwd")
        elif prec == "fp32":
            check(lib.sml_deform_attn_bwd(cfg["acfg"], c_f32(cfg["p_drop"]), c_u64(cfg["seed"]), ptr(qs), ptr(k), ptr(v), ptr(vs),
                                          ptr(cpb), ptr(out), ptr(lse), ptr(dout), ptr(dqs), ptr(dk), ptr(dv), ptr(dvs_b), ptr(part),
                                          stream()), "deform_attn_bwd")
        else:
            check(lib.sml_deform_attn_bwd_tc(cfg["acfg"], c_f32(cfg["p_drop"]), c_u64(cfg["seed"]), ptr(qs), ptr(k), ptr(v), ptr(vs),
                                             ptr(cpb), ptr(out), ptr(lse), ptr(dout), ptr(dqs), ptr(dk), ptr(dv), ptr(dvs_b),
                                             ptr(part), c_i32(1 if prec == "bf16x3" else 0), stream()), "deform_attn_bwd_tc")
        # backward = 2x the forward contractions (the forward recomputation inside the kernel is not counted)
        # (1-D table path: five 64-deep products (S, dP, dQ, dK, dV) + ~24 table / segment-sum operations per pair)
        _prof_end(prof, "deform_bwd", float(b) * heads * T * nkv * (10 * 64 + 24 if pwl is not None else
                                                                     2 * (2 * (OC * 32 + 32 * 32 + 32) + 4 * 64)),
                  (b, heads, T, nkv, OC, prec))
        if pwl is None:
            dcpb = ops.colsum(part)
        # k, v projections
        dkv, dwk = gconv_bwd(kv, wk2, cfg["gkv"], dk, mode=prec)
        _, dwv = gconv_bwd(kv, wv2, cfg["gkv"], dv, dx_out=dkv, dx_beta=1.0, mode=prec)
        # sampling
        dx2t = torch.zeros_like(x2t)
        dvs = torch.empty_like(vs)
        check(lib.sml_deform_gather_bwd(cfg["gcfg"], ptr(x2t), ptr(vs), ptr(dkv), ptr(dx2t), ptr(dvs_b), ptr(dvs), stream()),
              "deform_gather_bwd")
        # offsets network: dq = scale * dqs (+ offset path, accumulated in place)
        dq = torch.empty_like(q)
        ops.axpby_(dq, dqs, 64 ** -0.5, 0.0)
        nb = lib.sml_deform_offsets_bwd_blocks(cfg["ocfg"])
        KK = wc2.shape[1]
        nacc = gd * KK + gd + OC * gd
        dpre = torch.empty(b * G, nkv, gd, device=dev, dtype=F32)
        opart = torch.empty(nb, nacc, device=dev, dtype=F32)
        dvg = dvgrid.contiguous() if dvgrid is not None else None
        check(lib.sml_deform_offsets_bwd(cfg["ocfg"], c_f32(cfg["scale"]), ptr(q), ptr(wc2), ptr(bc.detach()), ptr(w22), ptr(off),
                                         ptr(dvs), ptr(dvg), ptr(dpre), ptr(opart), ptr(dq), stream()), "deform_offsets_bwd")
        osum = ops.colsum(opart)
        dwc, dbc, dw2 = osum[:gd * KK], osum[gd * KK:gd * KK + gd], osum[gd * KK + gd:]
        dx1t, dwq = gconv_bwd(x1t, wq2, cfg["gq"], dq, mode=prec)
        sh = cfg["shapes"]
        sizes = [int(torch.Size(s).numel()) for s in sh[6:]]
        dc = torch.split(dcpb, sizes)
        return (dx1t, dx2t, dwq.reshape(sh[0]), dwk.reshape(sh[1]), dwv.reshape(sh[2]), dw_out.reshape(sh[3]), db_out,
                dwc.reshape(sh[4]), dbc.clone(), dw2.reshape(sh[5]), *[g.reshape(s) for g, s in zip(dc, sh[6:])], None)


def gather_with_taps(x2t, vs, G, H, W, mode=0):
    """Forward bilinear gather alone (test / inspection entry): x2t [b, H*W, C], vs [(b G), nkv, CD]
    -> (kv [b, nkv, C], taps int32 [(b G), nkv, 2] = (x0, y0) top-left tap of every sample)."""
    ops.require_cuda(x2t, vs)
    x2t, vs = x2t.contiguous(), vs.contiguous()
    b, T, C = x2t.shape
    nkv, CD = vs.shape[1], vs.shape[2]
    kv = torch.empty(b, nkv, C, device=x2t.device, dtype=F32)
    taps = torch.empty(b * G, nkv, 2, device=x2t.device, dtype=torch.int32)
    check(lib.sml_deform_gather_fwd(_cfg(b, G, C // G, H, W, nkv, CD, mode), ptr(x2t), ptr(vs), ptr(kv), ptr(taps), stream()),
          "deform_gather_fwd")
    return kv, taps

"""Nystrom attention forward/backward on the C-ABI kernels (torch.autograd.Function).

Implements /root/reference/models/NystromAttention.py:74-157 (+ pinv :20-35) with mask=None.
Differences in *schedule* only (results match the reference within the tolerances in tests/):
  * q-scale folded into the q rows of W_qkv (exact for dim_head = 64: scale = 2^-3);
  * (A1 @ Z) @ (A3 @ V) is evaluated as A1 @ (Z @ (A3 @ V))  (4x fewer FLOPs);
  * the similarity softmaxes over the landmark axis are fused into the GEMM epilogue on the tensor-core engine;
    identity terms of the pinv iteration are GEMM-epilogue diagonals;
  * every producer (GEMM epilogue, softmax, LayerNorm, conv, landmark backward) also emits the bf16 (hi[, lo]) copy its
    consumer contraction needs, so there are no stand-alone conversion passes over the big tensors.

Precision modes (``precision`` argument):
  "fp32"   — SIMT fp32 FFMA engine everywhere (device-side exact-fp32 path);
  "bf16x3" — tcgen05 engine with the 3-pass bf16 split on every contraction (fp32-grade, ~2^-16);
  "bf16"   — tcgen05 single-pass bf16 operands, fp32 accumulate; the pinv chain stays bf16x3.
"""
import torch

from . import ops
from ._lib import c_i32, check, lib, ptr, stream
from .ops import F32, TB, gemm, mm

PRECISIONS = ("fp32", "bf16x3", "bf16")
# Run the Moore-Penrose GEMM chains as single persistent cooperative launches (sml_gemm_tc_program).  Measured on B200 at
# batch 64 x 256^3: no gain over back-to-back launches (each step is bound by L2->SMEM operand traffic, 64 MB/step, not by
# launch latency), so it is off by default; tests/test_gpu_nystrom.py keeps the path parity-checked.
FUSE_PINV_CHAIN = False
# bf16 mode, d = 64, m = 256: run the backward of the token-by-landmark path (dA1, softmax backward, dq, dk~) as the fused
# tcgen05 kernel sml_nystrom_a1_bwd instead of two GEMMs + a softmax-backward pass + two more GEMMs over [n', m] tensors
FUSE_A1_BWD = True
import os as _os
FUSE_A1_BWD = _os.environ.get("SML_FUSE_A1_BWD", "1") != "0"
FUSE_PINV_CHAIN = _os.environ.get("SML_FUSE_PINV_CHAIN", "0") == "1"
# bf16 mode, d = 64, m = 256: A3 V = softmax(ql k^T) v as a flash attention over key chunks (csrc/nystrom_a3.cu) - the [m, n']
# logits / probabilities and their gradients never touch HBM; the backward recomputes P from the saved row log-sum-exp
FUSE_A3 = _os.environ.get("SML_FUSE_A3", "1") != "0"
# "bf16" mode: run the Moore-Penrose chain with hi/lo split operands (True, fp32-grade) or in plain bf16 (False)
PINV_HP = _os.environ.get("SML_PINV_HP", "1") != "0"
# bf16 mode: keep q, k, v as bf16 only (landmark / residual-conv kernels read the bf16 copy).  Measured neutral on the 16 k
# step (the projection is not bound by its fp32 output and the conv kernels lose what it gains), so off by default; it saves
# 0.8 GB of activations per layer at B = 8 and stays parity-tested.
QKV_BF16_ONLY = _os.environ.get("SML_QKV_BF16_ONLY", "0") != "0"


def landmarks_fwd(qkv, b, npad, h, d, m, qkv_bf16=None):
    """qkv: the fp32 tensor, or None to read its bf16 copy ``qkv_bf16`` (bf16 mode keeps no fp32 qkv)."""
    ref = qkv if qkv is not None else qkv_bf16
    ql = torch.empty(b, h, m, d, device=ref.device, dtype=F32)
    kl = torch.empty_like(ql)
    check(lib.sml_nystrom_landmarks_fwd(ptr(qkv), ptr(ql), ptr(kl), c_i32(b), c_i32(npad), c_i32(h), c_i32(d), c_i32(m),
                                        None if qkv is not None else ptr(qkv_bf16), stream()), "landmarks_fwd")
    return ql, kl


def _prof(eng, flops, tag, fn):
    """run fn() between CUDA events when bench.py's instrumented pass is on (ops.PROFILE is a list)"""
    if ops.PROFILE is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    ops.PROFILE.append((eng, flops, e0, e1, tag))
    return r


def _a3_cfg(b, h, m, d, npad, nchunks):
    """cfg11 of sml_nystrom_a3_*: K / V as views into the packed [b, n', 3, h, d] tensor (element strides)"""
    bs = npad * 3 * h * d
    lo = bs & 0xffffffff
    return (c_i32 * 11)(b, h, m, d, npad, 3 * h * d, lo - 2 ** 32 if lo >= 2 ** 31 else lo, bs >> 32, h * d, 2 * h * d, nchunks)


def _a3_chunks(b, h, npad, tile, ctas_per_wave, rows_per_cta=1):
    """number of key chunks: enough CTAs for a few full waves, every chunk non-empty"""
    ntot = npad // tile
    want = max(1, (4 * ctas_per_wave) // max(b * h * rows_per_cta, 1))
    n = max(1, min(want, ntot))
    tpc = -(-ntot // n)
    return -(-ntot // tpc)


def landmark_ids(npad, m, device):
    ids = torch.empty(npad, device=device, dtype=torch.int32)
    check(lib.sml_nystrom_landmark_ids(ptr(ids), c_i32(npad), c_i32(m), stream()), "landmark_ids")
    return ids


def _new_like(t, mode):
    """bundle for a pinv-chain matrix batch: fp32 + (hi, lo) in tensor-core modes (the chain is always bf16x3)."""
    return TB.new(t.shape, t.device, mode, hp=PINV_HP)


def _pinv_step(pa2, z, mode):
    """one Moore-Penrose iteration on bundles: returns (xz, t2, t3).  With t1 = 7I - xz folded in,
    t2 = 15I - xz (7I - xz) = xz xz - 7 xz + 15 I is ONE contraction whose epilogue adds the -7 xz and 15 I terms."""
    op_only = mode != "fp32"     # t2, t3 are only ever GEMM operands: in tensor-core modes emit just their bf16 hi/lo
    xz = mm(pa2, z, _new_like(z.f, mode))                                        # x z
    t2 = mm(xz, xz, TB.new(z.f.shape, z.f.device, mode, hp=PINV_HP, f32=not op_only), diag=15.0,
            rowtab=xz.f, rowdiv=1, rowtab_scale=-7.0)                            # 15I - xz (7I - xz)
    t3 = mm(xz, t2, TB.new(z.f.shape, z.f.device, mode, hp=PINV_HP, f32=not op_only), alpha=-1.0, diag=13.0)   # 13I - xz t2
    return xz, t2, t3


def pinv_fwd(a2, iters, mode, a2_bundle=None, keep_steps=False):
    """Moore-Penrose iteration (NystromAttention.py:20-35).  Returns z (fp32) and what backward needs.  ``keep_steps`` (tensor-
    core modes) also keeps the bf16 operands xz, t2, t3 of every iteration (50 MB per iteration at 64 x 256^2) so that the
    backward does not recompute them: 18 of its 66 dependent, latency-bound GEMM launches."""
    bh, m = a2.numel() // (a2.shape[-1] ** 2), a2.shape[-1]
    dev = a2.device
    scal = torch.empty(2, device=dev, dtype=F32)
    arg = torch.empty(2, device=dev, dtype=torch.int32)
    ws = torch.empty(2, device=dev, dtype=torch.int64)
    check(lib.sml_pinv_scale(ptr(a2), ptr(scal), ptr(arg), ptr(ws), c_i32(bh), c_i32(m), stream()), "pinv_scale")
    z0 = torch.empty_like(a2)
    check(lib.sml_pinv_init(ptr(a2), ptr(scal), ptr(z0), c_i32(bh), c_i32(m), stream()), "pinv_init")
    pa2 = a2_bundle if (a2_bundle is not None and (mode == "fp32" or a2_bundle.lo is not None)) else TB.of(a2, mode, hp=PINV_HP)
    z = TB.of(z0, mode, hp=PINV_HP)
    zs = [z.f]
    keep_steps = keep_steps and mode != "fp32"
    steps = []
    zbs = [TB(None, z.hi, z.lo)]
    with ops.gemm_program(enabled=(mode != "fp32" and FUSE_PINV_CHAIN)):         # 4 * iters dependent GEMMs, one launch
        for _ in range(iters):
            xz, t2, t3 = _pinv_step(pa2, z, mode)
            if keep_steps:
                steps.append((TB(None, xz.hi, xz.lo), TB(None, t2.hi, t2.lo), TB(None, t3.hi, t3.lo)))
            z = mm(z, t3, _new_like(z.f, mode), alpha=0.25)                      # z <- 1/4 z t3
            zs.append(z.f)
            zbs.append(TB(None, z.hi, z.lo))
    if keep_steps:
        return z, (zs[:-1], scal, arg, steps, zbs[:-1])
    return z, (zs[:-1], scal, arg)


def pinv_bwd(a2, saved, dz, mode):
    """d a2 given d z_final; recomputes xz, t1..t3 per iteration from the saved z_k."""
    zs, scal, arg = saved[:3]
    steps = saved[3] if len(saved) > 3 else None                                 # (xz, t2, t3) operands kept by the forward
    zbs_kept = saved[4] if len(saved) > 4 else None                              # bf16 copies of the z_k, idem
    pa2_kept = saved[5] if len(saved) > 5 else None
    bh, m = a2.numel() // (a2.shape[-1] ** 2), a2.shape[-1]
    pa2 = pa2_kept if pa2_kept is not None else TB.of(a2, mode, hp=PINV_HP)
    da2 = torch.zeros_like(a2)
    g = TB.of(dz.contiguous(), mode, hp=PINV_HP)
    if zbs_kept is not None:
        zbs = [TB(zk, zb.hi, zb.lo) for zk, zb in zip(zs, zbs_kept)]
    else:
        zbs = [TB.of(zk, mode, hp=PINV_HP) for zk in zs]                            # conversions first: only GEMMs inside the program
    prog = ops.gemm_program(enabled=(mode != "fp32" and FUSE_PINV_CHAIN))
    prog.__enter__()
    for it, (zk, z) in enumerate(zip(reversed(zs), reversed(zbs))):
        xz, t2, t3 = steps[len(zs) - 1 - it] if steps else _pinv_step(pa2, z, mode)
        dzk = mm(g, t3.T(), _new_like(zk, mode), alpha=0.25)                     # 1/4 G t3^T
        dt3 = mm(z.T(), g, _new_like(zk, mode), alpha=0.25)                      # 1/4 z^T G
        dxz = TB(torch.empty_like(zk))
        mm(dt3, t2.T(), dxz, alpha=-1.0)                                         # t3 = 13I - xz t2:  dxz  = -dt3 t2^T
        dt2 = mm(xz.T(), dt3, _new_like(zk, mode), alpha=-1.0)                   #                    dt2  = -xz^T dt3
        mm(dt2, xz.T(), dxz, beta=1.0, rowtab=dt2.f, rowdiv=1, rowtab_scale=-7.0)   # t2 = xz xz - 7 xz + 15 I:  dxz += dt2 xz^T - 7 dt2
        dxz_b = _new_like(zk, mode)
        dxz_b.f = dxz.f
        mm(xz.T(), dt2, dxz_b, beta=1.0)                                         #                    dxz += xz^T dt2  (emits hi/lo)
        mm(dxz_b, z.T(), TB(da2), beta=1.0)                                      # dx += dxz z^T
        dzk2 = _new_like(zk, mode)
        dzk2.f = dzk.f
        mm(pa2.T(), dxz_b, dzk2, beta=1.0)                                       # dz += x^T dxz  (re-emits hi/lo of the sum)
        g = dzk2
    prog.__exit__(None, None, None)                                              # 11 * iters dependent GEMMs, <= 2 launches
    tmp = torch.empty(1, device=a2.device, dtype=F32)
    check(lib.sml_pinv_init_bwd(ptr(a2), ptr(scal), ptr(arg), ptr(g.f), ptr(da2), ptr(tmp), c_i32(bh), c_i32(m), stream()),
          "pinv_init_bwd")
    return da2


class NystromAttentionFn(torch.autograd.Function):
    """y = NystromAttention(x) without the trailing dropout.  x [b,n,D] fp32 CUDA.

    ``pre_norm`` = (ln_weight, ln_bias, eps) fuses the TransLayer LayerNorm (mil.py:186): the normalised rows are
    written straight into the front-padded buffer (and its bf16 copy)."""

    @staticmethod
    def forward(ctx, x, w_qkv, w_out, b_out, w_conv, ln_w, ln_b, heads, m, iters, mode, ln_eps, skip=None, inference=False):
        """skip = (p, seed, offset): return x + dropout(attention(LN(x)), p) (TransLayer, mil.py:186) - the skip gradient is then
        added inside the LayerNorm-backward kernel instead of by autograd's accumulation pass."""
        assert mode in PRECISIONS
        ops.require_cuda(x, w_qkv, w_out, b_out, w_conv)
        b, n, D = x.shape
        inner = w_qkv.shape[0] // 3
        h, d = heads, inner // heads
        pad = (m - n % m) % m
        npad = n + pad
        dev = x.device
        x = x.contiguous()
        tc = mode != "fp32"
        hv = lambda t, i: t[:, :, i].permute(0, 2, 1, 3)                             # [b,h,npad,d] view of a qkv chunk
        # ---- (LayerNorm +) front pad (:79-82)
        # tensor-core modes with the fused LayerNorm only ever read the bf16 copies of xp: no fp32 plane then
        xp = TB.new((b, npad, D), dev, mode, f32=not (tc and ln_w is not None))
        if pad:
            for t in (xp.f, xp.hi, xp.lo):
                if t is not None:
                    t[:, :pad].zero_()
        ln_saved = None
        if ln_w is not None:
            mean, rstd = ops.layernorm_fwd(x.view(b * n, D), ln_w.detach(), ln_b.detach(), xp.f, n, npad, pad, ln_eps,
                                           y_hi=xp.hi, y_lo=xp.lo)
            ln_saved = (mean, rstd)
        else:
            xp.f[:, pad:] = x
            if tc:
                s = ops.to_bf16(xp.f, split=(mode == "bf16x3"))
                xp.hi, xp.lo = s.hi, s.lo
        xp2 = xp.map(lambda t: t.view(b * npad, D))
        weff = w_qkv.detach().clone()
        ops.axpby_(weff[:inner], w_qkv.detach()[:inner].contiguous(), d ** -0.5, 0.0)    # fold q * scale (:98)
        weff_b = TB.of(weff, mode)
        # ---- qkv projection (:89), fp32 + bf16 copy from the same epilogue
        # bf16 mode (33-tap conv: the vector kernels read either type): q, k, v live as bf16 only - the projection writes a
        # third of the bytes and the landmark / residual-conv kernels read half
        ksize = w_conv.shape[2]
        qkv = TB.new((b, npad, 3, h, d), dev, mode, f32=not (QKV_BF16_ONLY and mode == "bf16" and ksize == 33 and d % 4 == 0))
        mm(xp2, weff_b.T(), qkv.map(lambda t: t.view(b * npad, 3 * inner)))
        ql, kl = landmarks_fwd(qkv.f, b, npad, h, d, m, qkv.hi)                      # :102-118
        ql_b, kl_b = TB.of(ql, mode), TB.of(kl, mode)
        q_b, k_b, v_b = (qkv.map(lambda t, i=i: hv(t, i)) for i in range(3))
        # ---- similarities + softmaxes (:122-137)
        fuse = tc and m <= 256
        lowp = mode == "bf16" and fuse          # keep the big probability tensors (A1, A3) in bf16 only
        # inference: the token-by-landmark product is fused too (no A1): see below, after W = Z (A3 V) is known
        fuse_a1_inf = FUSE_A3 and inference and mode == "bf16" and d == 64 and m % 64 == 0 and npad % 64 == 0 and qkv.hi is not None
        a1 = TB(None, None, None) if fuse_a1_inf else TB.new((b, h, npad, m), dev, mode, f32=not lowp)
        a2 = TB.new((b, h, m, m), dev, mode, hp=PINV_HP)
        # (the forward kernel takes any multiple of 64 landmarks, the backward 256: 512-landmark bags fuse when no gradient is needed)
        fuse_a3 = (FUSE_A3 and mode == "bf16" and d == 64 and ksize == 33 and npad % 64 == 0 and qkv.hi is not None and
                   (m == 256 or (m % 64 == 0 and inference)))
        if not fuse_a1_inf:
            mm(q_b, kl_b.T(), a1, softmax=fuse)
        mm(ql_b, kl_b.T(), a2, softmax=fuse)
        lse3 = None
        if fuse_a3:
            # sim3 -> softmax over tokens -> . v in one flash pass over key chunks: no [m, n'] tensor (csrc/nystrom_a3.cu)
            a3 = TB(None, None, None)
            a3v = TB.new((b, h, m, d), dev, mode)
            lse3 = torch.empty(b, h, m, device=dev, dtype=F32)
            nch = _a3_chunks(b, h, npad, 64, 3 * 148, rows_per_cta=m // 64)
            opart = torch.empty(b * h, nch, m, d, device=dev, dtype=F32)
            ml = torch.empty(b * h, nch, m, 2, device=dev, dtype=F32)
            _prof("a3_fwd", 4.0 * d * b * h * m * npad, (b, h, m, npad),
                  lambda: check(lib.sml_nystrom_a3_fwd(_a3_cfg(b, h, m, d, npad, nch), ptr(ql_b.hi), ptr(qkv.hi), ptr(opart), ptr(ml),
                                                       ptr(a3v.f), ptr(a3v.hi), ptr(lse3), stream()), "nystrom_a3_fwd"))
            del opart, ml
        else:
            a3 = TB.new((b, h, m, npad), dev, mode, f32=not lowp)
            s3 = a3.f if a3.f is not None else torch.empty(b, h, m, npad, device=dev, dtype=F32)   # logits scratch
            mm(ql_b, k_b.T(), TB(s3))
        if not fuse:
            if not fuse_a1_inf:
                ops.softmax_(a1.f.view(-1, m), a1.hi, a1.lo)
            ops.softmax_(a2.f.view(-1, m), a2.hi, a2.lo)
        if not fuse_a3:
            ops.softmax_(s3.view(-1, npad), a3.hi, a3.lo, write_f32=not lowp)            # softmax over tokens
            del s3
        z, pinv_saved = pinv_fwd(a2.f, iters, mode, a2, keep_steps=any(ctx.needs_input_grad) and not inference)     # :138
        if not fuse_a3:
            a3v = mm(a3, v_b, TB.new((b, h, m, d), dev, mode))                       # [b,h,m,d]
        z_op = z if mode != "bf16" else TB(z.f, z.hi, None)
        wm = mm(z_op, a3v, TB.new((b, h, m, d), dev, mode))                          # Z (A3 V)
        merged = TB.new((b, npad, inner), dev, mode)
        if fuse_a1_inf:
            check(lib.sml_nystrom_a1_fwd_nograd(_a3_cfg(b, h, m, d, npad, 1), ptr(qkv.hi), ptr(kl_b.hi), ptr(wm.hi), ptr(merged.f), stream()),
                  "nystrom_a1_fwd_nograd")
        else:
            mm(a1, wm, TB(merged.f.view(b, npad, h, d).permute(0, 2, 1, 3)))         # A1 (Z A3 V)  :140
        ksize = w_conv.shape[2]
        wc = w_conv.detach().reshape(h, ksize).contiguous()
        fused_cvt = ksize == 33
        check(lib.sml_nystrom_resconv_fwd(ptr(qkv.f), ptr(wc), ptr(merged.f), c_i32(b), c_i32(npad), c_i32(h), c_i32(d),
                                          c_i32(ksize), ptr(merged.hi) if fused_cvt else None,
                                          ptr(merged.lo) if fused_cvt else None, None if qkv.f is not None else ptr(qkv.hi),
                                          stream()), "resconv_fwd")   # :144-145
        if tc and not fused_cvt:
            s = ops.to_bf16(merged.f, split=(mode == "bf16x3"))
            merged.hi, merged.lo = s.hi, s.lo
        y = torch.empty(b, n, D, device=dev, dtype=F32)
        wout_b = TB.of(w_out.detach().contiguous(), mode)
        mm(merged.map(lambda t: t[:, pad:]), wout_b.T(), TB(y), bias=b_out.detach())  # :150-151
        # bf16 operand copies the backward needs again (each re-conversion is a ~13 us latency-bound launch: 14 per layer)
        extras = []
        if len(pinv_saved) > 3:
            for bnd in (ql_b, kl_b, wm, a3v, z, a2, weff_b, wout_b, *pinv_saved[4]):
                extras += [bnd.hi, bnd.lo]
        ctx.save_for_backward(x, ln_w, ln_b, xp.f, xp.hi, xp.lo, weff, w_out, qkv.f, qkv.hi, qkv.lo, ql, kl, a1.f, a1.hi, a1.lo,
                              a2.f, lse3 if fuse_a3 else a3.f, a3.hi, a3.lo, z.f, a3v.f, wm.f, merged.f, merged.hi, merged.lo, wc,
                              *(ln_saved or ()), *pinv_saved[0], pinv_saved[1], pinv_saved[2],
                              *([t for st in pinv_saved[3] for bnd in st for t in (bnd.hi, bnd.lo)] if len(pinv_saved) > 3 else []),
                              *extras)
        ctx.n_extras = len(extras)
        ctx.cfg = (b, n, D, h, d, m, iters, pad, npad, inner, ksize, mode, ln_w is not None)
        ctx.skip = skip
        ctx.fuse_a3 = fuse_a3
        if skip is not None:
            assert ln_w is not None, "the fused skip connection needs the fused LayerNorm"
            return ops.dropout_add(x, y, skip[0], skip[1], skip[2])
        return y

    @staticmethod
    def backward(ctx, dy):
        b, n, D, h, d, m, iters, pad, npad, inner, ksize, mode, has_ln = ctx.cfg
        sv = list(ctx.saved_tensors)
        (x, ln_w, ln_b, xp_f, xp_hi, xp_lo, weff, w_out, qkv_f, qkv_hi, qkv_lo, ql, kl, a1_f, a1_hi, a1_lo, a2_f, a3_f, a3_hi,
         a3_lo, z_f, a3v_f, wm_f, mg_f, mg_hi, mg_lo, wc) = sv[:27]
        rest = sv[27:]
        if has_ln:
            mean, rstd = rest[0], rest[1]
            rest = rest[2:]
        zs, scal, arg = list(rest[:iters]), rest[iters], rest[iters + 1]
        ex = rest[len(rest) - ctx.n_extras:] if ctx.n_extras else []
        kept = rest[iters + 2:len(rest) - ctx.n_extras]
        steps = None
        if len(kept) == 6 * iters:                                                   # (xz, t2, t3) x (hi, lo) per iteration
            steps = [tuple(TB(None, kept[6 * i + 2 * j], kept[6 * i + 2 * j + 1]) for j in range(3)) for i in range(iters)]
        exb = [(ex[2 * i], ex[2 * i + 1]) for i in range(len(ex) // 2)]                # (hi, lo) pairs in the order saved

        def kept_or_of(i, f32, single=False):
            """bundle of fp32 tensor ``f32`` with the bf16 copies the forward kept (else converted now)"""
            if not exb:
                return TB.of(f32, mode)
            hi, lo = exb[i]
            return TB(f32, hi, None if (single or mode == "bf16") else lo)
        dev = dy.device
        tc = mode != "fp32"
        hv = lambda t, i: t[:, :, i].permute(0, 2, 1, 3)
        xp = TB(xp_f, xp_hi, xp_lo); qkv = TB(qkv_f, qkv_hi, qkv_lo); a1 = TB(a1_f, a1_hi, a1_lo); a3 = TB(a3_f, a3_hi, a3_lo)
        merged = TB(mg_f, mg_hi, mg_lo)
        dy = dy.contiguous()
        dy_skip = None
        if ctx.skip is not None:
            dy_skip = dy                                                             # gradient of the skip connection
            if ctx.skip[0] > 0:
                dy = ops.dropout_add(None, dy, ctx.skip[0], ctx.skip[1], ctx.skip[2])    # same mask as the forward
        if pad and tc and D % 4 == 0:
            # tensor-core modes: the front-padded gradient only exists as the bf16 GEMM operand (zero pad rows written by the
            # conversion kernel itself): no padded fp32 copy (zero fill + copy + re-read of 0.27 GB per layer at B = 8)
            hi = torch.empty(b * npad, D, device=dev, dtype=torch.bfloat16)
            lo = torch.empty_like(hi) if mode == "bf16x3" else None
            check(lib.sml_split_bf16_pad(ptr(dy), ptr(hi), ptr(lo), ops.c_i64(b), ops.c_i64(n), ops.c_i64(npad), ops.c_i64(pad),
                                         c_i32(D), stream()), "split_bf16_pad")
            dy_b = TB(None, hi, lo)
            db_out = ops.colsum(dy.view(b * n, D))
        else:
            if pad:
                dyp = torch.zeros(b, npad, D, device=dev, dtype=F32)
                dyp[:, pad:] = dy
            else:
                dyp = dy
            dyp2 = dyp.view(b * npad, D)
            dy_b = TB.of(dyp2, mode)
            db_out = ops.colsum(dyp2)
        # ---- to_out
        dw_out = mm(dy_b.T(), merged.map(lambda t: t.view(b * npad, inner))).f                  # [D, inner]
        dmerged = TB.new((b, npad, inner), dev, mode)
        mm(dy_b, kept_or_of(7, w_out.detach().contiguous()), dmerged.map(lambda t: t.view(b * npad, inner)))
        dout_b = dmerged.map(lambda t: t.view(b, npad, h, d).permute(0, 2, 1, 3))
        q_b, k_b, v_b = (qkv.map(lambda t, i=i: hv(t, i)) for i in range(3))
        ql_b, kl_b = kept_or_of(0, ql), kept_or_of(1, kl)
        dqkv = TB.new((b, npad, 3, h, d), dev, mode)
        dq, dk, dv = (hv(dqkv.f, i) for i in range(3))
        # ---- out = A1 Wm ; Wm = Z A3V ; A3V = A3 V
        wm_b = kept_or_of(2, wm_f)
        fuse_a1 = FUSE_A1_BWD and mode == "bf16" and a1.f is None and m == 256 and d == 64 and fused_cvt_ok(ksize)
        da1 = None if fuse_a1 else mm(dout_b, wm_b.T(), TB(torch.empty(b, h, npad, m, device=dev, dtype=F32)))
        dwm = mm(a1.T(), dout_b, TB.new((b, h, m, d), dev, mode))                               # [b,h,m,d]
        dz = mm(dwm, kept_or_of(3, a3v_f).T()).f                                                # [b,h,m,m]
        da3v = mm(kept_or_of(4, z_f).T(), dwm, TB.new((b, h, m, d), dev, mode))                 # [b,h,m,d]
        fuse_a3 = ctx.fuse_a3
        if fuse_a3:
            # dS3 recomputed tile by tile from the saved log-sum-exp: dql (+=), dk (bf16 slice, landmark term added below) and
            # dv (fp32 slice) in one kernel, nothing of size [m, n'] in HBM
            lse3 = a3_f
            dql = TB(torch.zeros(b, h, m, d, device=dev, dtype=F32))
            a3cfg = _a3_cfg(b, h, m, d, npad, _a3_chunks(b, h, npad, 32, 148))
            _prof("a3_bwd", 10.0 * d * b * h * m * npad, (b, h, m, npad),
                  lambda: check(lib.sml_nystrom_a3_bwd(a3cfg, ptr(ql_b.hi), ptr(qkv.hi), ptr(da3v.f), ptr(a3v_f), ptr(lse3), None,
                                                       c_i32(npad // m), ptr(dql.f), ptr(dqkv.hi), ptr(dqkv.f), stream()), "nystrom_a3_bwd"))
        else:
            da3 = mm(da3v, v_b.T(), TB(torch.empty(b, h, m, npad, device=dev, dtype=F32)))
            mm(a3.T(), da3v, TB(dv))                                                            # dv = A3^T dA3V
        dwc_part = torch.zeros(lib.sml_nystrom_resconv_bwd_blocks(c_i32(b), c_i32(npad), c_i32(h * d)), h * ksize, device=dev,
                               dtype=F32)      # upper bound on the partial rows; unused rows stay zero
        fused_cvt = ksize == 33
        check(lib.sml_nystrom_resconv_bwd(ptr(qkv.f), ptr(wc), ptr(dmerged.f), ptr(dqkv.f), ptr(dwc_part), c_i32(b), c_i32(npad),
                                          c_i32(h), c_i32(d), c_i32(ksize), ptr(dqkv.hi) if fused_cvt else None,
                                          ptr(dqkv.lo) if fused_cvt else None, None if qkv.f is not None else ptr(qkv.hi),
                                          stream()), "resconv_bwd")   # dv += conv^T
        dw_conv = ops.colsum(dwc_part).view(h, 1, ksize, 1)
        # ---- pinv + softmaxes (dS emitted with its bf16 copy by the softmax-backward kernel)
        if steps and exb:
            pa2 = TB(a2_f, *exb[5])
            zbs_kept = [TB(None, hi, lo) for hi, lo in exb[8:8 + iters]]
            da2 = pinv_bwd(a2_f, (zs, scal, arg, steps, zbs_kept, pa2), dz, mode)
        else:
            da2 = pinv_bwd(a2_f, (zs, scal, arg, steps) if steps else (zs, scal, arg), dz, mode)
        lowp = a1.f is None                     # bf16-only probabilities: dS also only needs its bf16 copy
        p1 = a1.f if a1.f is not None else a1.hi
        if not fuse_a3:
            ds3 = TB.new((b, h, m, npad), dev, mode, f32=False); ds3.f = da3.f
            p3 = a3.f if a3.f is not None else a3.hi
        if not fuse_a1:
            ds1 = TB.new((b, h, npad, m), dev, mode, f32=False); ds1.f = da1.f
            ops.softmax_bwd_(p1.view(-1, m), ds1.f.view(-1, m), ds1.hi, ds1.lo, write_f32=not lowp)
        ops.softmax_bwd_(a2_f.view(-1, m), da2.view(-1, m))
        if not fuse_a3:
            ops.softmax_bwd_(p3.view(-1, npad), ds3.f.view(-1, npad), ds3.hi, ds3.lo, write_f32=not lowp)
        if lowp:
            if not fuse_a3:
                ds3.f = None
            if not fuse_a1:
                ds1.f = None
        ds2 = TB.of(da2, mode)
        # ---- similarities; the landmark-mean backward (dq[t] += dql[t // l] / l, same for k) rides in the epilogue of
        #      the dq / dk contractions, which also emit the bf16 copy of dqkv for the projection backward
        l = npad // m
        dkl = mm(ds2.T(), ql_b)                                                                 # dkl = dS2^T ql (+ dS1^T q below)
        if fuse_a3:
            mm(ds2, kl_b, dql, beta=1.0)                                                        # dql = dS3 k (above) + dS2 kl
        else:
            dql = mm(ds2, kl_b)
            mm(ds3, k_b, dql, beta=1.0)
        # bf16 mode: only the bf16 copies of dq / dk are ever read (projection backward), so their fp32 halves are not written
        skip_f32 = mode == "bf16" and fused_cvt
        dq_b = TB(None if skip_f32 else dq, *(hv(t, 0) if t is not None else None for t in (dqkv.hi, dqkv.lo)))
        dk_b = TB(None if skip_f32 else dk, *(hv(t, 1) if t is not None else None for t in (dqkv.hi, dqkv.lo)))
        if fuse_a1:
            # dA1 = dout W^T, softmax backward, dq = dS1 kl + bcast(dql)/l and dkl += dS1^T q in one tcgen05 kernel
            check(lib.sml_nystrom_a1_bwd(ptr(dmerged.hi), ptr(wm_b.hi), ptr(a1.hi), ptr(kl_b.hi), ptr(qkv.hi), ptr(dql.f),
                                         None if skip_f32 else ptr(dqkv.f), ptr(dqkv.hi), ptr(dkl.f), c_i32(b), c_i32(npad), c_i32(h), c_i32(d), c_i32(m), stream()),
                  "nystrom_a1_bwd")
        else:
            mm(ds1.T(), q_b, dkl, beta=1.0)                                                     # dkl += dS1^T q
            mm(ds1, kl_b, dq_b, rowtab=dql.f, rowdiv=l, rowtab_scale=1.0 / l)                   # dq = dS1 kl + bcast(dql)/l
        if fuse_a3:
            check(lib.sml_nystrom_a3_dk_landmarks(a3cfg, ptr(dkl.f), c_i32(l), ptr(dqkv.hi), stream()), "nystrom_a3_dk_landmarks")
        else:
            mm(ds3.T(), ql_b, dk_b, rowtab=dkl.f, rowdiv=l, rowtab_scale=1.0 / l)               # dk = dS3^T ql + bcast(dkl)/l
        if tc and not fused_cvt:
            s = ops.to_bf16(dqkv.f, split=(mode == "bf16x3"))
            dqkv.hi, dqkv.lo = s.hi, s.lo
        # ---- qkv projection
        dqkv2 = dqkv.map(lambda t: t.view(b * npad, 3 * inner))
        dxp = mm(dqkv2, kept_or_of(6, weff)).f.view(b, npad, D)
        dw_qkv = mm(dqkv2.T(), xp.map(lambda t: t.view(b * npad, D))).f                         # [3 inner, D]
        ops.axpby_(dw_qkv[:inner], dw_qkv[:inner].clone(), d ** -0.5, 0.0)
        dln_w = dln_b = None
        if has_ln:
            dx2, dln_w, dln_b = ops.layernorm_bwd(dxp.view(b * npad, D), x.view(b * n, D), ln_w.detach(), mean, rstd, n, npad, pad,
                                                  res=None if dy_skip is None else dy_skip.view(b * n, D))
            dx = dx2.view(b, n, D)
        else:
            dx = dxp[:, pad:]
        return dx, dw_qkv, dw_out, db_out, dw_conv, dln_w, dln_b, None, None, None, None, None, None, None


@torch.no_grad()
def nystrom_attn_map(x, w_qkv, heads, m, iters):
    """``return_attn=True`` of NystromAttention.forward (NystromAttention.py:153-155): the dense approximate attention map
    A1 . A2^+ . A3 [b, h, n', n'] over the front-padded tokens.  Inspection only (no caller in the reference uses it): evaluated
    on the exact-fp32 engine, detached from autograd; n'^2 floats per head, so meant for small bags."""
    ops.require_cuda(x, w_qkv)
    b, n, D = x.shape
    inner = w_qkv.shape[0] // 3
    h, d = heads, inner // heads
    pad = (m - n % m) % m
    npad = n + pad
    dev = x.device
    xp = torch.zeros(b, npad, D, device=dev, dtype=F32)
    xp[:, pad:] = x
    weff = w_qkv.detach().clone()
    ops.axpby_(weff[:inner], w_qkv.detach()[:inner].contiguous(), d ** -0.5, 0.0)
    qkv = torch.empty(b, npad, 3, h, d, device=dev, dtype=F32)
    gemm(xp.view(b * npad, D), weff.t(), out=qkv.view(b * npad, 3 * inner))
    ql, kl = landmarks_fwd(qkv, b, npad, h, d, m)
    q, k = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(2))
    a1 = gemm(q, kl.transpose(-1, -2)); ops.softmax_(a1.view(-1, m))
    a2 = gemm(ql, kl.transpose(-1, -2)); ops.softmax_(a2.view(-1, m))
    a3 = gemm(ql, k.transpose(-1, -2)); ops.softmax_(a3.view(-1, npad))
    z, _ = pinv_fwd(a2, iters, "fp32")
    return gemm(gemm(a1, z.f), a3)


def fused_cvt_ok(ksize):
    return ksize == 33


def nystrom_attention(x, w_qkv, w_out, b_out, w_conv, heads, num_landmarks, pinv_iterations, precision="fp32",
                      pre_norm=None, skip=None):
    ln_w, ln_b, eps = pre_norm if pre_norm is not None else (None, None, 1e-5)
    # inference = called under torch.no_grad() (ctx.needs_input_grad does not tell: it stays True for parameters)
    return NystromAttentionFn.apply(x, w_qkv, w_out, b_out, w_conv, ln_w, ln_b, heads, num_landmarks, pinv_iterations,
                                    precision, eps, skip, not torch.is_grad_enabled())

"""Thin Python wrappers over the C-ABI primitives: tensors in, tensors out, current stream, no fallback."""
import ctypes

import torch

from ._lib import GemmDesc, c_f32, c_i32, c_i64, c_u64, c_vp, check, lib, ptr, require_cuda, stream

F32, BF16 = torch.float32, torch.bfloat16

class GemmProgram:
    """Record the tensor-core GEMMs issued inside the ``with`` block and run them as ONE persistent cooperative launch
    (sml_gemm_tc_program): for chains of small dependent products (the Moore-Penrose iteration) whose individual
    launches are dominated by fixed per-launch latency.  Only GEMMs may depend on each other inside the block; every
    tensor a recorded step touches is kept alive until the launch, so the caching allocator cannot recycle it early."""

    def __init__(self, enabled=True):
        self.enabled = enabled
        self.descs, self.keep, self.flops = [], [], 0.0

    def __enter__(self):
        global _PROGRAM
        if self.enabled:
            assert _PROGRAM is None, "GemmProgram blocks do not nest"
            _PROGRAM = self
        return self

    def __exit__(self, et, ev, tb):
        global _PROGRAM
        if not self.enabled:
            return False
        _PROGRAM = None
        if et is None and self.descs:
            self.launch()
        return False

    def launch(self):
        n = len(self.descs)
        arr = (GemmDesc * n)(*self.descs)
        dev = self.keep[0].device
        counter = torch.empty(1, device=dev, dtype=torch.int32)
        if PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(lib.sml_gemm_tc_program(arr, c_i32(n), ptr(counter), stream()), "gemm_tc_program")
        if PROFILE is not None:
            e1.record()
            PROFILE.append(("tc", self.flops, e0, e1, ("program", n)))
        self.descs, self.keep = [], []


_PROGRAM = None


def gemm_program(enabled=True):
    return GemmProgram(enabled)


# bench instrumentation: when a list, every gemm() appends (engine, algorithmic flops, start event, end event)
PROFILE = None


class Split:
    """bf16 operand for the tensor-core engine: ``hi`` (+ optional ``lo`` residual for the bf16x3 split).
    ``map(f)`` applies the same view function to both parts."""
    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo=None):
        self.hi, self.lo = hi, lo

    def map(self, f):
        return Split(f(self.hi), None if self.lo is None else f(self.lo))

    @property
    def shape(self):
        return self.hi.shape

    def transpose(self, a, b):
        return self.map(lambda t: t.transpose(a, b))


class TB:
    """Tensor bundle: fp32 tensor ``f`` (may be None) plus optional bf16 copies ``hi`` / ``lo`` that producers fill in
    the same pass (GEMM epilogues, softmax, LayerNorm, ...), so no separate conversion pass is needed before a
    tensor-core contraction.  ``map(fn)`` applies one view function to every part."""
    __slots__ = ("f", "hi", "lo")

    def __init__(self, f=None, hi=None, lo=None):
        self.f, self.hi, self.lo = f, hi, lo

    @staticmethod
    def new(shape, device, mode, hp=False, f32=True):
        """mode "fp32": fp32 only; "bf16": + hi (and lo when ``hp``); "bf16x3": + hi, lo."""
        f = torch.empty(shape, device=device, dtype=F32) if (f32 or mode == "fp32") else None
        hi = lo = None
        if mode != "fp32":
            hi = torch.empty(shape, device=device, dtype=BF16)
            if mode == "bf16x3" or hp:
                lo = torch.empty(shape, device=device, dtype=BF16)
        return TB(f, hi, lo)

    @staticmethod
    def of(x, mode, hp=False):
        """wrap an existing contiguous fp32 tensor, converting with one split pass if the mode needs bf16."""
        if mode == "fp32":
            return TB(x)
        s = to_bf16(x.contiguous(), split=(mode == "bf16x3" or hp))
        return TB(x, s.hi, s.lo)

    def map(self, fn):
        return TB(None if self.f is None else fn(self.f), None if self.hi is None else fn(self.hi),
                  None if self.lo is None else fn(self.lo))

    def T(self):
        return self.map(lambda t: t.transpose(-1, -2))

    def operand(self):
        return self.f if self.hi is None else Split(self.hi, self.lo)

    def ptrs(self):
        return ptr(self.hi), ptr(self.lo)


def mm(A, B, out=None, **kw):
    """gemm on bundles: operands pick the engine (bf16 parts present -> tcgen05); ``out`` bundle receives fp32 and/or
    bf16 hi/lo straight from the epilogue."""
    a, b = A.operand(), B.operand()
    if isinstance(a, Split) != isinstance(b, Split):
        raise RuntimeError("mm: operands prepared for different engines")
    if out is None:
        return TB(gemm(a, b, **kw))
    if isinstance(a, Split):
        gemm(a, b, out=out.f, out_bf16=out.hi, out_bf16_lo=out.lo, want_f32=False, **kw)
    else:
        gemm(a, b, out=out.f, **kw)
    return out


def to_bf16(x, split=False):
    """fp32 -> Split(bf16 hi[, lo]) with one pass of the split kernel (x must be contiguous)."""
    require_cuda(x)
    assert x.dtype == F32 and x.is_contiguous()
    hi = torch.empty_like(x, dtype=BF16)
    lo = torch.empty_like(x, dtype=BF16) if split else None
    check(lib.sml_split_bf16(ptr(x), ptr(hi), ptr(lo), c_i64(x.numel()), stream()), "split_bf16")
    return Split(hi, lo)


def _batch_strides(t, nb):
    """t has shape [*batch, r, c] with len(batch) <= 2; returns (s_b1, s_b2) honouring broadcast (size-1) dims."""
    bs = t.shape[:-2]
    st = t.stride()[:-2]
    pads = 2 - len(bs)
    bs = (1,) * pads + tuple(bs)
    st = (0,) * pads + tuple(st)
    out = []
    for want, have, s in zip(nb, bs, st):
        assert have == want or have == 1, "batch dims %s do not broadcast to %s" % (bs, nb)
        out.append(0 if have == 1 else s)
    return out


def gemm(A, B, out=None, *, alpha=1.0, beta=0.0, diag=0.0, bias=None, relu=False, softmax=False,
         out_bf16=None, out_bf16_lo=None, want_f32=True, rowtab=None, rowdiv=1, rowtab_scale=1.0):
    """out[..., M, N] = epi(alpha * A[..., M, K] @ B[..., K, N]).

    fp32 tensors -> SIMT fp32 engine (sml_gemm_f32); ``Split`` operands -> tcgen05 engine (sml_gemm_tc).
    Any strides / transposed views are passed through as strides (no copies).  ``softmax`` (row softmax
    over N fused in the epilogue) is only available on the tensor-core engine with N <= 256.
    """
    tc = isinstance(A, Split)
    assert tc == isinstance(B, Split), "both operands must use the same engine"
    a, b = (A.hi, B.hi) if tc else (A, B)
    require_cuda(a, b, out)
    M, K = a.shape[-2:]
    K2, N = b.shape[-2:]
    assert K == K2, "inner dims differ: %s @ %s" % (tuple(a.shape), tuple(b.shape))
    ba, bb = tuple(a.shape[:-2]), tuple(b.shape[:-2])
    if ba == bb:
        nbs = ba
    else:                                       # broadcast of the (at most two) batch dims, without torch's generic machinery
        n_ = max(len(ba), len(bb))
        pa, pb = (1,) * (n_ - len(ba)) + ba, (1,) * (n_ - len(bb)) + bb
        assert all(x == y or x == 1 or y == 1 for x, y in zip(pa, pb)), "batch dims %s / %s do not broadcast" % (ba, bb)
        nbs = tuple(max(x, y) for x, y in zip(pa, pb))
    assert len(nbs) <= 2
    nb = (1,) * (2 - len(nbs)) + tuple(nbs)
    if out is None and (want_f32 or not tc):
        out = torch.empty(tuple(nbs) + (M, N), device=a.device, dtype=F32)
        assert beta == 0.0
    d = GemmDesc()
    d.M, d.N, d.K, d.nb1, d.nb2 = M, N, K, nb[0], nb[1]
    d.A = a.data_ptr(); d.sa_m, d.sa_k = a.stride(-2), a.stride(-1)
    d.sa_b1, d.sa_b2 = _batch_strides(a, nb)
    d.B = b.data_ptr(); d.sb_k, d.sb_n = b.stride(-2), b.stride(-1)
    d.sb_b1, d.sb_b2 = _batch_strides(b, nb)
    ref = out if out is not None else out_bf16
    assert ref is not None and tuple(ref.shape[-2:]) == (M, N)
    d.C = out.data_ptr() if out is not None else None
    d.sc_m, d.sc_n = ref.stride(-2), ref.stride(-1)
    d.sc_b1, d.sc_b2 = _batch_strides(ref, nb)
    d.alpha, d.beta, d.diag = alpha, beta, diag
    d.bias = bias.data_ptr() if bias is not None else None
    d.relu = (1 if relu else 0) | (2 if softmax else 0)
    if rowtab is not None:        # + rowtab_scale * rowtab[..., m // rowdiv, n]
        assert rowtab.dtype == F32 and rowtab.stride(-1) == 1 and rowtab.shape[-1] == N
        d.rowtab = rowtab.data_ptr(); d.rt_m = rowtab.stride(-2)
        d.rt_b1, d.rt_b2 = _batch_strides(rowtab, nb)
        d.rowdiv, d.rowtab_scale = int(rowdiv), float(rowtab_scale)
    if tc:
        assert a.dtype == BF16 and b.dtype == BF16
        if (A.lo is None) != (B.lo is None):     # mixed: give the single-precision side a zero residual? no: drop to single pass
            A, B = Split(A.hi), Split(B.hi)
        d.A_lo = A.lo.data_ptr() if A.lo is not None else None
        d.B_lo = B.lo.data_ptr() if B.lo is not None else None
        if A.lo is not None:
            assert A.lo.stride() == a.stride() and B.lo.stride() == b.stride()
        for ob in (out_bf16, out_bf16_lo):
            if ob is not None:
                assert ob.dtype == BF16 and ob.stride() == ref.stride()
        d.Cb_hi = out_bf16.data_ptr() if out_bf16 is not None else None
        d.Cb_lo = out_bf16_lo.data_ptr() if out_bf16_lo is not None else None
        fn, eng = lib.sml_gemm_tc, "tc"
        if _PROGRAM is not None and A.lo is not None and not softmax:
            _PROGRAM.descs.append(d)
            _PROGRAM.keep.extend(t for t in (a, b, A.lo, B.lo, out, out_bf16, out_bf16_lo, bias, rowtab) if t is not None)
            _PROGRAM.flops += 2.0 * M * N * K * nb[0] * nb[1]
            return out
    else:
        assert a.dtype == F32 and b.dtype == F32 and not softmax and out_bf16 is None
        fn, eng = lib.sml_gemm_f32, "simt"
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(ctypes.byref(d), stream()), "gemm_" + eng)
        e1.record()
        tag = (M, N, K, nb[0] * nb[1], int(d.sa_k != 1), int(d.sb_k != 1), int(bool(d.A_lo)), d.relu, int(bool(d.C)), int(bool(d.Cb_hi)), int(bool(d.Cb_lo)), float(beta))
        PROFILE.append((eng, 2.0 * M * N * K * nb[0] * nb[1], e0, e1, tag))
    else:
        check(fn(ctypes.byref(d), stream()), "gemm_" + eng)
    return out


def layernorm_fwd(x2d, w, b, y, n, npad, pad, eps=1e-5, y_hi=None, y_lo=None):
    R, D = x2d.shape
    mean = torch.empty(R, device=x2d.device, dtype=F32)
    rstd = torch.empty(R, device=x2d.device, dtype=F32)
    check(lib.sml_layernorm_fwd(ptr(x2d), ptr(w), ptr(b), ptr(y), ptr(mean), ptr(rstd), c_i64(R), c_i32(D), c_f32(eps),
                                c_i64(n), c_i64(npad), c_i64(pad), ptr(y_hi), ptr(y_lo), stream()), "layernorm_fwd")
    return mean, rstd


def layernorm_bwd(dy, x2d, w, mean, rstd, n, npad, pad, res=None):
    R, D = x2d.shape
    nb = lib.sml_layernorm_bwd_blocks(c_i64(R))
    dx = torch.empty_like(x2d)
    dwp = torch.empty(nb, D, device=x2d.device, dtype=F32)
    dbp = torch.empty(nb, D, device=x2d.device, dtype=F32)
    check(lib.sml_layernorm_bwd(ptr(dy), ptr(x2d), ptr(w), ptr(mean), ptr(rstd), ptr(dx), ptr(dwp), ptr(dbp), c_i64(R),
                                c_i32(D), c_i64(n), c_i64(npad), c_i64(pad), ptr(res), stream()), "layernorm_bwd")
    return dx, colsum(dwp), colsum(dbp)


def colsum(x2d, out=None, beta=0.0):
    R, C = x2d.shape
    if out is None:
        out = torch.empty(C, device=x2d.device, dtype=F32)
    part = torch.empty(lib.sml_colsum_parts(c_i64(R)), C, device=x2d.device, dtype=F32)
    check(lib.sml_colsum(ptr(x2d), ptr(out), ptr(part), c_i64(R), c_i64(C), c_f32(beta), stream()), "colsum")
    return out


def softmax_(x2d, hi=None, lo=None, write_f32=True):
    R, C = x2d.shape
    assert x2d.is_contiguous()
    check(lib.sml_softmax_fwd(ptr(x2d), c_i64(R), c_i64(C), ptr(hi), ptr(lo), c_i32(1 if write_f32 else 0), stream()), "softmax_fwd")
    return x2d


def softmax_bwd_(A2d, dA2d, hi=None, lo=None, write_f32=True):
    """A2d: fp32 or bf16 probabilities; dS overwrites dA2d (write_f32) and/or goes to the bf16 copies."""
    R, C = A2d.shape
    assert A2d.is_contiguous() and dA2d.is_contiguous()
    a32, a16 = (A2d, None) if A2d.dtype == F32 else (None, A2d)
    check(lib.sml_softmax_bwd(ptr(a32), ptr(a16), ptr(dA2d), c_i64(R), c_i64(C), ptr(hi), ptr(lo), c_i32(1 if write_f32 else 0),
                              stream()), "softmax_bwd")
    return dA2d


def dropout_add(x, u, p, seed, offset=0, out=None):
    """out = x + dropout(u, p) (x may be None).  Same (seed, offset) regenerates the mask (used for backward)."""
    assert u.is_contiguous() and (x is None or x.is_contiguous())
    if out is None:
        out = torch.empty_like(u)
    check(lib.sml_dropout_add(ptr(x), ptr(u), ptr(out), c_i64(u.numel()), c_f32(p), c_u64(seed), c_u64(offset), stream()),
          "dropout_add")
    return out


def axpby_(y, x, a, b, hi=None, lo=None):
    assert y.is_contiguous() and x.is_contiguous() and y.numel() == x.numel()
    check(lib.sml_axpby(ptr(y), ptr(x), c_f32(a), c_f32(b), c_i64(y.numel()), ptr(hi), ptr(lo), stream()), "axpby")
    return y


def diag_axpby(x, a, dg, out=None, hi=None, lo=None):
    """out = a*x + dg*I for a batch of square matrices."""
    m = x.shape[-1]
    assert x.is_contiguous() and x.shape[-2] == m
    if out is None:
        out = torch.empty_like(x)
    check(lib.sml_diag_axpby(ptr(out), ptr(x), c_f32(a), c_f32(dg), c_i64(x.numel() // (m * m)), c_i32(m), ptr(hi), ptr(lo), stream()),
          "diag_axpby")
    return out

// deform.cu — deformable cross-attention 1-D / 2-D (models/DeformableAttention1D.py:106-240,
// models/DeformableAttention2D.py:161-325): offset network, bilinear gather / scatter, and a fused
// flash-style attention whose logits carry the continuous-position-bias MLP evaluated on the fly
// (the reference materialises [B*g, n, n_kv, 32] three times).
//
// Layout: everything is token-major — x, q, kv, k, v, out are [b, tokens, channels] with the channel axis
// contiguous (so a warp reads 128 B of one token), the transpose of the reference's [b, C, n] views.
// A 1-D sequence is the 2-D case with H = 1 (offsets) / W = 1 (sampling), so one kernel serves both.
#include "common.cuh"
#include "deform_attn.cuh"

namespace sml {

struct OffCfg {
  int B, G, gd, inner;          // batch, offset groups, channels per group (inner / G), q row width
  int H, W;                     // query lattice (1-D: H = 1, W = n)
  int KH, KW, SH, SW, PH, PW;   // depthwise conv geometry
  int HK, WK, OC;               // output lattice and offset channels (2-D: 2 = (x, y); 1-D: 1)
  float scale, den0, den1;      // offset scale; normalisation denominators max(size-1, 1) for channel 0 / 1
};

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  return cdf + x * 0.39894228040143267794f * __expf(-0.5f * x * x);
}

// depthwise-conv pre-activation of channel c at output location (oy, ox)
__device__ __forceinline__ float off_preact(const OffCfg& cf, const float* __restrict__ qb, const float* __restrict__ wc,
                                            float bias, int c, int oy, int ox) {
  float acc = bias;
  for (int ky = 0; ky < cf.KH; ++ky) {
    const int iy = oy * cf.SH - cf.PH + ky;
    if (iy < 0 || iy >= cf.H) continue;
    for (int kx = 0; kx < cf.KW; ++kx) {
      const int ix = ox * cf.SW - cf.PW + kx;
      if (ix < 0 || ix >= cf.W) continue;
      acc = fmaf(__ldg(wc + c * cf.KH * cf.KW + ky * cf.KW + kx), qb[(int64_t)(iy * cf.W + ix) * cf.inner + c], acc);
    }
  }
  return acc;
}

// one warp per (bg, p): lanes span the gd channels.  Writes off (tanh*scale), vgrid (reference layout
// [(b g), OC, HK, WK]) and vs (normalised, [(b g), HK*WK, OC]).
__global__ void __launch_bounds__(256) offsets_fwd_kernel(const OffCfg cf, const float* __restrict__ q,
                                                          const float* __restrict__ wc, const float* __restrict__ bc,
                                                          const float* __restrict__ w2, float* __restrict__ off,
                                                          float* __restrict__ vgrid, float* __restrict__ vs) {
  const int lane = threadIdx.x & 31;
  const int nkv = cf.HK * cf.WK;
  const int64_t loc = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (loc >= (int64_t)cf.B * cf.G * nkv) return;
  const int p = (int)(loc % nkv), bg = (int)(loc / nkv);
  const int b = bg / cf.G, g = bg % cf.G;
  const int oy = p / cf.WK, ox = p % cf.WK;
  const float* qb = q + (int64_t)b * cf.H * cf.W * cf.inner + g * cf.gd;
  float s0 = 0.f, s1 = 0.f;
  for (int c = lane; c < cf.gd; c += 32) {
    const float a = gelu_exact(off_preact(cf, qb, wc, __ldg(bc + c), c, oy, ox));
    s0 = fmaf(__ldg(w2 + c), a, s0);
    if (cf.OC == 2) s1 = fmaf(__ldg(w2 + cf.gd + c), a, s1);
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (lane == 0) {
    const float o0 = tanhf(s0) * cf.scale;
    const float v0 = (float)ox + o0;                               // grid channel 0 = x (column index)
    off[loc * cf.OC] = o0;
    vgrid[((int64_t)bg * cf.OC) * nkv + p] = v0;
    vs[loc * cf.OC] = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, v0), cf.den0), -1.f);
    if (cf.OC == 2) {
      const float o1 = tanhf(s1) * cf.scale;
      const float v1 = (float)oy + o1;                             // channel 1 = y (row index)
      off[loc * 2 + 1] = o1;
      vgrid[((int64_t)bg * 2 + 1) * nkv + p] = v1;
      vs[loc * 2 + 1] = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, v1), cf.den1), -1.f);
    }
  }
}

// backward, stage 1: per (bg, p) turn (d vs, d vgrid) into d pre-activation per channel (stored for stage 2)
// and accumulate the parameter gradients in shared memory -> one partial row per block.
// part layout per block: [gd*KK (dwc)] [gd (dbc)] [OC*gd (dw2)]
__global__ void __launch_bounds__(256) offsets_bwd_kernel(const OffCfg cf, const float* __restrict__ q,
                                                          const float* __restrict__ wc, const float* __restrict__ bc,
                                                          const float* __restrict__ w2, const float* __restrict__ off,
                                                          const float* __restrict__ dvs, const float* __restrict__ dvgrid,
                                                          float* __restrict__ dpre, float* __restrict__ part,
                                                          int64_t locs_per_block) {
  extern __shared__ float acc[];
  const int KK = cf.KH * cf.KW;
  const int nacc = cf.gd * KK + cf.gd + cf.OC * cf.gd;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  float* a_wc = acc; float* a_bc = acc + cf.gd * KK; float* a_w2 = a_bc + cf.gd;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nkv = cf.HK * cf.WK;
  const int64_t total = (int64_t)cf.B * cf.G * nkv;
  const int64_t l0 = (int64_t)blockIdx.x * locs_per_block, l1 = min(total, l0 + locs_per_block);
  for (int64_t loc = l0 + wid; loc < l1; loc += nw) {
    const int p = (int)(loc % nkv), bg = (int)(loc / nkv);
    const int b = bg / cf.G, g = bg % cf.G;
    const int oy = p / cf.WK, ox = p % cf.WK;
    const float* qb = q + (int64_t)b * cf.H * cf.W * cf.inner + g * cf.gd;
    float dt[2];
    for (int oc = 0; oc < cf.OC; ++oc) {
      float gv = dvs[loc * cf.OC + oc] * 2.f / (oc == 0 ? cf.den0 : cf.den1);
      if (dvgrid) gv += dvgrid[((int64_t)bg * cf.OC + oc) * nkv + p];
      const float t = off[loc * cf.OC + oc] / cf.scale;            // tanh value
      dt[oc] = gv * cf.scale * (1.f - t * t);
    }
    for (int c = lane; c < cf.gd; c += 32) {
      const float pre = off_preact(cf, qb, wc, __ldg(bc + c), c, oy, ox);
      const float a = gelu_exact(pre);
      float da = dt[0] * __ldg(w2 + c);
      atomicAdd(&a_w2[c], dt[0] * a);
      if (cf.OC == 2) { da = fmaf(dt[1], __ldg(w2 + cf.gd + c), da); atomicAdd(&a_w2[cf.gd + c], dt[1] * a); }
      const float dp = da * gelu_grad(pre);
      dpre[loc * cf.gd + c] = dp;
      atomicAdd(&a_bc[c], dp);
      for (int ky = 0; ky < cf.KH; ++ky) {
        const int iy = oy * cf.SH - cf.PH + ky;
        if (iy < 0 || iy >= cf.H) continue;
        for (int kx = 0; kx < cf.KW; ++kx) {
          const int ix = ox * cf.SW - cf.PW + kx;
          if (ix < 0 || ix >= cf.W) continue;
          atomicAdd(&a_wc[c * KK + ky * cf.KW + kx], dp * qb[(int64_t)(iy * cf.W + ix) * cf.inner + c]);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) part[(int64_t)blockIdx.x * nacc + i] = acc[i];
}

// Same stage for the two geometries of the reference path (1-D: 1 x 6 taps, 128 channels per group; 2-D: 6 x 6 taps, 64
// channels per group): a lane owns CPL channels and keeps their parameter-gradient sums (KH KW + 1 + OC per channel) in
// registers over all the locations of its warp; shared memory sees one add per sum and warp at the end.  The generic kernel
// above issues KH KW + 2 shared-memory fp32 atomics (compare-and-swap loops, 8 warps on the same addresses) per (location,
// channel): 639 us at B = 8 / 16 k tokens.
template <int KH, int KW, int CPL>
__global__ void __launch_bounds__(256) offsets_bwd_reg_kernel(const OffCfg cf, const float* __restrict__ q,
                                                              const float* __restrict__ wc, const float* __restrict__ bc,
                                                              const float* __restrict__ w2, const float* __restrict__ off,
                                                              const float* __restrict__ dvs, const float* __restrict__ dvgrid,
                                                              float* __restrict__ dpre, float* __restrict__ part,
                                                              int64_t locs_per_block) {
  extern __shared__ float acc[];
  constexpr int KK = KH * KW;
  const int nacc = cf.gd * KK + cf.gd + cf.OC * cf.gd;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  float* a_wc = acc; float* a_bc = acc + cf.gd * KK; float* a_w2 = a_bc + cf.gd;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nkv = cf.HK * cf.WK;
  const int64_t total = (int64_t)cf.B * cf.G * nkv;
  const int64_t l0 = (int64_t)blockIdx.x * locs_per_block, l1 = min(total, l0 + locs_per_block);
  float r_wc[CPL][KK], r_bc[CPL], r_w2[CPL][2], wreg[CPL][KK], breg[CPL], w2r[CPL][2];
#pragma unroll
  for (int u = 0; u < CPL; ++u) {
    const int c = lane + 32 * u;
    r_bc[u] = 0.f; r_w2[u][0] = r_w2[u][1] = 0.f;
    breg[u] = __ldg(bc + c);
    w2r[u][0] = __ldg(w2 + c); w2r[u][1] = cf.OC == 2 ? __ldg(w2 + cf.gd + c) : 0.f;
#pragma unroll
    for (int k = 0; k < KK; ++k) { r_wc[u][k] = 0.f; wreg[u][k] = __ldg(wc + c * KK + k); }
  }
  for (int64_t loc = l0 + wid; loc < l1; loc += nw) {
    const int p = (int)(loc % nkv), bg = (int)(loc / nkv);
    const int b = bg / cf.G, g = bg % cf.G;
    const int oy = p / cf.WK, ox = p % cf.WK;
    const float* qb = q + (int64_t)b * cf.H * cf.W * cf.inner + g * cf.gd;
    float dt[2] = {0.f, 0.f};
    for (int oc = 0; oc < cf.OC; ++oc) {
      float gv = dvs[loc * cf.OC + oc] * 2.f / (oc == 0 ? cf.den0 : cf.den1);
      if (dvgrid) gv += dvgrid[((int64_t)bg * cf.OC + oc) * nkv + p];
      const float t = off[loc * cf.OC + oc] / cf.scale;            // tanh value
      dt[oc] = gv * cf.scale * (1.f - t * t);
    }
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
      const int c = lane + 32 * u;
      float qv[KK];
      float pre = breg[u];
#pragma unroll
      for (int ky = 0; ky < KH; ++ky) {
        const int iy = oy * cf.SH - cf.PH + ky;
#pragma unroll
        for (int kx = 0; kx < KW; ++kx) {
          const int ix = ox * cf.SW - cf.PW + kx;
          const bool ok = iy >= 0 && iy < cf.H && ix >= 0 && ix < cf.W;
          qv[ky * KW + kx] = ok ? qb[(int64_t)(iy * cf.W + ix) * cf.inner + c] : 0.f;
          pre = fmaf(wreg[u][ky * KW + kx], qv[ky * KW + kx], pre);    // same tap order as off_preact (zero taps add 0)
        }
      }
      const float a = gelu_exact(pre);
      float da = dt[0] * w2r[u][0];
      r_w2[u][0] = fmaf(dt[0], a, r_w2[u][0]);
      if (cf.OC == 2) { da = fmaf(dt[1], w2r[u][1], da); r_w2[u][1] = fmaf(dt[1], a, r_w2[u][1]); }
      const float dp = da * gelu_grad(pre);
      dpre[loc * cf.gd + c] = dp;
      r_bc[u] += dp;
#pragma unroll
      for (int k = 0; k < KK; ++k) r_wc[u][k] = fmaf(dp, qv[k], r_wc[u][k]);
    }
  }
#pragma unroll
  for (int u = 0; u < CPL; ++u) {
    const int c = lane + 32 * u;
    atomicAdd(&a_bc[c], r_bc[u]);
    atomicAdd(&a_w2[c], r_w2[u][0]);
    if (cf.OC == 2) atomicAdd(&a_w2[cf.gd + c], r_w2[u][1]);
#pragma unroll
    for (int k = 0; k < KK; ++k) atomicAdd(&a_wc[c * KK + k], r_wc[u][k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) part[(int64_t)blockIdx.x * nacc + i] = acc[i];
}

// backward, stage 2 (atomic-free gather form): dq[b, pix, g*gd + c] += sum over windows containing pix
// grid (B ceil(H W / 8) pixel groups, ceil(G gd / 256)), one thread per channel of eight pixels: coalesced dpre / dq rows, no 64-bit index
// arithmetic per element (the flat grid-stride form spent ~200 integer instructions per element on div / mod: 872 us at
// B = 8 / 16 k tokens against ~90 us of memory traffic)
constexpr int ODQ_PIX = 8;      // pixels per thread: eight independent read-modify-write chains in flight
__global__ void __launch_bounds__(256) offsets_dq_kernel(const OffCfg cf, const float* __restrict__ wc,
                                                         const float* __restrict__ dpre, float* __restrict__ dq) {
  const int nkv = cf.HK * cf.WK;
  const int call = blockIdx.y * 256 + threadIdx.x;
  if (call >= cf.G * cf.gd) return;
  const int g = call / cf.gd, c = call - g * cf.gd;
  const int hw = cf.H * cf.W;
  const int per_b = (hw + ODQ_PIX - 1) / ODQ_PIX;
  const int b = blockIdx.x / per_b, pix0 = (blockIdx.x - b * per_b) * ODQ_PIX;
  const float* dp = dpre + ((int64_t)(b * cf.G + g)) * nkv * cf.gd + c;
  const float* wp = wc + c * cf.KH * cf.KW;
  float* dqb = dq + (int64_t)b * hw * cf.inner + call;
  float s[ODQ_PIX], old[ODQ_PIX];
#pragma unroll
  for (int u = 0; u < ODQ_PIX; ++u) old[u] = pix0 + u < hw ? dqb[(int64_t)(pix0 + u) * cf.inner] : 0.f;
#pragma unroll
  for (int u = 0; u < ODQ_PIX; ++u) {
    const int pix = pix0 + u;
    s[u] = 0.f;
    if (pix >= hw) continue;
    const int iy = pix / cf.W, ix = pix - iy * cf.W;
    // windows (oy, ox) with oy*SH - PH <= iy < oy*SH - PH + KH
    const int oy_hi = min(cf.HK - 1, (iy + cf.PH) / cf.SH);
    const int ox_hi = min(cf.WK - 1, (ix + cf.PW) / cf.SW);
    for (int oy = oy_hi; oy >= 0; --oy) {
      const int ky = iy - (oy * cf.SH - cf.PH);
      if (ky >= cf.KH) break;
      for (int ox = ox_hi; ox >= 0; --ox) {
        const int kx = ix - (ox * cf.SW - cf.PW);
        if (kx >= cf.KW) break;
        s[u] = fmaf(dp[(int64_t)(oy * cf.WK + ox) * cf.gd], __ldg(wp + ky * cf.KW + kx), s[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < ODQ_PIX; ++u)
    if (pix0 + u < hw) dqb[(int64_t)(pix0 + u) * cf.inner] = old[u] + s[u];
}

// ------------------------------------------------------------------------------------------ bilinear gather
struct GatherCfg {
  int B, G, cg, C;      // batch, groups, channels per group, row width of x2 / kv (= G*cg)
  int H, W;             // sampled image (1-D reference pad order: H = n, W = 1; see F5)
  int nkv, CD;          // samples per (b,g); coordinate dims
  int mode;             // CD == 1: 0 = reference pad order (x = vs, y = 0), 1 = upstream (x = 0, y = vs)
};

struct Taps { int x0, y0; float wx1, wy1; };
__device__ __forceinline__ Taps make_taps(const GatherCfg& cf, const float* __restrict__ vs, int64_t loc) {
  float gx, gy;
  if (cf.CD == 2) { gx = vs[loc * 2]; gy = vs[loc * 2 + 1]; }
  else if (cf.mode == 0) { gx = vs[loc]; gy = 0.f; }
  else { gx = 0.f; gy = vs[loc]; }
  // grid_sampler unnormalise, align_corners=False: ((coord + 1) * size - 1) / 2, in the reference's op order
  const float ix = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)cf.W), -1.f), 2.f);
  const float iy = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)cf.H), -1.f), 2.f);
  const float fx = floorf(ix), fy = floorf(iy);
  Taps t;
  t.x0 = (int)fx; t.y0 = (int)fy; t.wx1 = ix - fx; t.wy1 = iy - fy;
  return t;
}

// thread per (bg, p, c); taps [(b g), nkv, 2] int32 (x0, y0) optional
__global__ void __launch_bounds__(256) gather_fwd_kernel(const GatherCfg cf, const float* __restrict__ x2,
                                                         const float* __restrict__ vs, float* __restrict__ kv,
                                                         int32_t* __restrict__ taps) {
  const int64_t total = (int64_t)cf.B * cf.G * cf.nkv * cf.cg;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % cf.cg);
  const int64_t loc = i / cf.cg;
  const int p = (int)(loc % cf.nkv), bg = (int)(loc / cf.nkv);
  const int b = bg / cf.G, g = bg % cf.G;
  const Taps t = make_taps(cf, vs, loc);
  if (taps && c == 0) { taps[loc * 2] = t.x0; taps[loc * 2 + 1] = t.y0; }
  const float* xb = x2 + (int64_t)b * cf.H * cf.W * cf.C + g * cf.cg + c;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
    if (xx < 0 || xx >= cf.W || yy < 0 || yy >= cf.H) continue;
    const float w = ((k & 1) ? t.wx1 : 1.f - t.wx1) * ((k >> 1) ? t.wy1 : 1.f - t.wy1);
    acc = fmaf(w, xb[(int64_t)(yy * cf.W + xx) * cf.C], acc);
  }
  kv[((int64_t)b * cf.nkv + p) * cf.C + g * cf.cg + c] = acc;
}

// thread per (bg, p, c): scatter d kv into d x2 and reduce the coordinate gradient over the cg channels.
// dvs_out[loc, cd] = dvs_in[loc, cd] (bias path, may be NULL) + sampling path.
__global__ void __launch_bounds__(256) gather_bwd_kernel(const GatherCfg cf, const float* __restrict__ x2,
                                                         const float* __restrict__ vs, const float* __restrict__ dkv,
                                                         float* __restrict__ dx2, const float* __restrict__ dvs_in,
                                                         float* __restrict__ dvs_out, const bool scatter) {
  const int64_t total = (int64_t)cf.B * cf.G * cf.nkv * cf.cg;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < total;
  const int c = live ? (int)(i % cf.cg) : 0;
  const int64_t loc = live ? i / cf.cg : 0;
  float gix = 0.f, giy = 0.f;
  if (live) {
    const int p = (int)(loc % cf.nkv), bg = (int)(loc / cf.nkv);
    const int b = bg / cf.G, g = bg % cf.G;
    const Taps t = make_taps(cf, vs, loc);
    const float gout = dkv[((int64_t)b * cf.nkv + p) * cf.C + g * cf.cg + c];
    const int64_t base = (int64_t)b * cf.H * cf.W * cf.C + g * cf.cg + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
      if (xx < 0 || xx >= cf.W || yy < 0 || yy >= cf.H) continue;
      const float wx = (k & 1) ? t.wx1 : 1.f - t.wx1, wy = (k >> 1) ? t.wy1 : 1.f - t.wy1;
      const int64_t o = base + (int64_t)(yy * cf.W + xx) * cf.C;
      if (scatter) atomicAdd(dx2 + o, wx * wy * gout);
      const float v = x2[o];
      gix += ((k & 1) ? 1.f : -1.f) * wy * v * gout;
      giy += ((k >> 1) ? 1.f : -1.f) * wx * v * gout;
    }
  }
  // reduce over the cg (power of two <= 32) channel lanes of this sample
  for (int o = cf.cg >> 1; o > 0; o >>= 1) {
    gix += __shfl_xor_sync(0xffffffffu, gix, o);
    giy += __shfl_xor_sync(0xffffffffu, giy, o);
  }
  if (live && c == 0) {
    const float dgx = gix * 0.5f * (float)cf.W, dgy = giy * 0.5f * (float)cf.H;
    if (cf.CD == 2) {
      dvs_out[loc * 2] = (dvs_in ? dvs_in[loc * 2] : 0.f) + dgx;
      dvs_out[loc * 2 + 1] = (dvs_in ? dvs_in[loc * 2 + 1] : 0.f) + dgy;
    } else {
      dvs_out[loc] = (dvs_in ? dvs_in[loc] : 0.f) + (cf.mode == 0 ? dgx : dgy);
    }
  }
}

// Vectorised forms (channels per group a multiple of 4, the only case on the reference path: 16 in 2-D, 32 in 1-D): one thread
// per (sample, 4 channels).  Forward: the four tap rows are fetched with one 128-bit load each, all four in flight before the
// first FMA (out-of-range taps read a clamped address with weight 0), one 128-bit store.  Same FMA order as the scalar kernel.
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(256) gather_fwd_vec_kernel(const GatherCfg cf, const float* __restrict__ x2,
                                                             const float* __restrict__ vs, float* __restrict__ kv,
                                                             int32_t* __restrict__ taps) {
  const int qpg = cf.cg >> 2;
  const int64_t total = (int64_t)cf.B * cf.G * cf.nkv * qpg;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % qpg);
  const int64_t loc = i / qpg;
  const int p = (int)(loc % cf.nkv), bg = (int)(loc / cf.nkv);
  const int b = bg / cf.G, g = bg % cf.G;
  const Taps t = make_taps(cf, vs, loc);
  if (taps && c4 == 0) *reinterpret_cast<int2*>(taps + loc * 2) = make_int2(t.x0, t.y0);
  const float* xb = x2 + (int64_t)b * cf.H * cf.W * cf.C + g * cf.cg + 4 * c4;
  float w[4];
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
    const bool ok = xx >= 0 && xx < cf.W && yy >= 0 && yy < cf.H;
    w[k] = ok ? ((k & 1) ? t.wx1 : 1.f - t.wx1) * ((k >> 1) ? t.wy1 : 1.f - t.wy1) : 0.f;
    const int xc = min(max(xx, 0), cf.W - 1), yc = min(max(yy, 0), cf.H - 1);
    v[k] = ldg4(xb + (int64_t)(yc * cf.W + xc) * cf.C);
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc.x = fmaf(w[k], v[k].x, acc.x); acc.y = fmaf(w[k], v[k].y, acc.y);
    acc.z = fmaf(w[k], v[k].z, acc.z); acc.w = fmaf(w[k], v[k].w, acc.w);
  }
  *reinterpret_cast<float4*>(kv + ((int64_t)b * cf.nkv + p) * cf.C + g * cf.cg + 4 * c4) = acc;
}

// backward, coordinate gradients (and, when `scatter`, the fp32-atomic d x2 fallback): thread per (sample, 4 channels), the
// partial sums of a sample's cg / 4 threads are combined with shuffles.
__global__ void __launch_bounds__(256) gather_bwd_vec_kernel(const GatherCfg cf, const float* __restrict__ x2,
                                                             const float* __restrict__ vs, const float* __restrict__ dkv,
                                                             float* __restrict__ dx2, const float* __restrict__ dvs_in,
                                                             float* __restrict__ dvs_out, const bool scatter) {
  const int qpg = cf.cg >> 2;
  const int64_t total = (int64_t)cf.B * cf.G * cf.nkv * qpg;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < total;
  const int c4 = live ? (int)(i % qpg) : 0;
  const int64_t loc = live ? i / qpg : 0;
  float gix = 0.f, giy = 0.f;
  if (live) {
    const int p = (int)(loc % cf.nkv), bg = (int)(loc / cf.nkv);
    const int b = bg / cf.G, g = bg % cf.G;
    const Taps t = make_taps(cf, vs, loc);
    const float4 go = ldg4(dkv + ((int64_t)b * cf.nkv + p) * cf.C + g * cf.cg + 4 * c4);
    const int64_t base = (int64_t)b * cf.H * cf.W * cf.C + g * cf.cg + 4 * c4;
    float4 v[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
      ok[k] = xx >= 0 && xx < cf.W && yy >= 0 && yy < cf.H;
      const int xc = min(max(xx, 0), cf.W - 1), yc = min(max(yy, 0), cf.H - 1);
      v[k] = ldg4(x2 + base + (int64_t)(yc * cf.W + xc) * cf.C);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!ok[k]) continue;
      const float wx = (k & 1) ? t.wx1 : 1.f - t.wx1, wy = (k >> 1) ? t.wy1 : 1.f - t.wy1;
      if (scatter) {
        float* o = dx2 + base + (int64_t)((t.y0 + (k >> 1)) * cf.W + t.x0 + (k & 1)) * cf.C;
        const float ww = wx * wy;
        atomicAdd(o, ww * go.x); atomicAdd(o + 1, ww * go.y); atomicAdd(o + 2, ww * go.z); atomicAdd(o + 3, ww * go.w);
      }
      const float d = fmaf(v[k].x, go.x, fmaf(v[k].y, go.y, fmaf(v[k].z, go.z, v[k].w * go.w)));
      gix += ((k & 1) ? 1.f : -1.f) * wy * d;
      giy += ((k >> 1) ? 1.f : -1.f) * wx * d;
    }
  }
  for (int o = qpg >> 1; o > 0; o >>= 1) {
    gix += __shfl_xor_sync(0xffffffffu, gix, o);
    giy += __shfl_xor_sync(0xffffffffu, giy, o);
  }
  if (live && c4 == 0) {
    const float dgx = gix * 0.5f * (float)cf.W, dgy = giy * 0.5f * (float)cf.H;
    if (cf.CD == 2) {
      dvs_out[loc * 2] = (dvs_in ? dvs_in[loc * 2] : 0.f) + dgx;
      dvs_out[loc * 2 + 1] = (dvs_in ? dvs_in[loc * 2 + 1] : 0.f) + dgy;
    } else {
      dvs_out[loc] = (dvs_in ? dvs_in[loc] : 0.f) + (cf.mode == 0 ? dgx : dgy);
    }
  }
}

// Atomic-free, deterministic scatter of d kv into d x2: one block per (b, g) owns every destination pixel of its channel
// group.  The 4 nkv (sample, tap) entries are sorted by destination in shared memory (bitonic network on the composite key
// dest * NP + entry, so the order inside a destination is canonical), then every run of equal destinations is reduced by one
// warp (lanes = the group's channels: coalesced d kv reads, one coalesced store per destination).  dx2 must be zeroed by the
// caller (pixels no sample touches).
constexpr int GS_THREADS = 512;
// NP = padded entry count (power of two), EPS = entries per sample: 4 bilinear taps in 2-D; 2 in 1-D, where the sampled image is
// one pixel wide (F5) and only the tap column xx == 0 can be in range.  skey / wgt live in dynamic shared memory (8 NP bytes:
// 128 KB for the 1-D path at 16 k tokens with n_kv = 4096 -> NP = 8192 ... 2-D lattices up to n_kv = 4096 -> NP = 16384).
template <int NP, int EPS>
__global__ void __launch_bounds__(GS_THREADS) gather_bwd_sorted_kernel(const GatherCfg cf, const float* __restrict__ vs,
                                                                       const float* __restrict__ dkv, float* __restrict__ dx2) {
  extern __shared__ __align__(16) unsigned char gs_smem[];
  uint32_t* skey = reinterpret_cast<uint32_t*>(gs_smem);
  float* wgt = reinterpret_cast<float*>(gs_smem) + NP;
  const int bg = blockIdx.x, b = bg / cf.G, g = bg % cf.G;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t INVALID = 0xffffffffu;
  constexpr int ESH = EPS == 4 ? 2 : 1;
  for (int e = tid; e < NP; e += GS_THREADS) {
    uint32_t key = INVALID;
    float w = 0.f;
    if (e < EPS * cf.nkv) {
      const int p = e >> ESH;
      const Taps t = make_taps(cf, vs, (int64_t)bg * cf.nkv + p);
      // EPS == 2: tap row ky = e & 1, tap column = the one that lands on xx == 0 (x0 == 0 -> kx = 0, x0 == -1 -> kx = 1)
      const int kx = EPS == 4 ? (e & 1) : -t.x0, ky = EPS == 4 ? ((e >> 1) & 1) : (e & 1);
      const int xx = t.x0 + kx, yy = t.y0 + ky;
      if ((kx == 0 || kx == 1) && xx >= 0 && xx < cf.W && yy >= 0 && yy < cf.H) {
        key = (uint32_t)(yy * cf.W + xx) * (uint32_t)NP + (uint32_t)e;
        w = (kx ? t.wx1 : 1.f - t.wx1) * (ky ? t.wy1 : 1.f - t.wy1);
      }
    }
    skey[e] = key; wgt[e] = w;
  }
  __syncthreads();
  for (int k = 2; k <= NP; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < NP; i += GS_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint32_t a = skey[i], c = skey[ixj];
          if ((a > c) == ((i & k) == 0)) { skey[i] = c; skey[ixj] = a; }
        }
      }
      __syncthreads();
    }
  // runs of equal destination -> one warp each (the warp that owns the run's first position)
  constexpr int PER_WARP = NP / (GS_THREADS / 32);
  for (int pos = warp * PER_WARP; pos < (warp + 1) * PER_WARP; ++pos) {
    const uint32_t ck = skey[pos];
    if (ck == INVALID) break;
    const uint32_t dest = ck / NP;
    if (pos > 0 && skey[pos - 1] / NP == dest) continue;
    float acc = 0.f;
    for (int qn = pos; qn < NP; ++qn) {
      const uint32_t kq = skey[qn];
      if (kq == INVALID || kq / NP != dest) break;
      const int e = (int)(kq % NP);
      if (lane < cf.cg) acc = fmaf(wgt[e], dkv[((int64_t)b * cf.nkv + (e >> ESH)) * cf.C + g * cf.cg + lane], acc);
    }
    if (lane < cf.cg) dx2[((int64_t)b * cf.H * cf.W + dest) * cf.C + g * cf.cg + lane] = acc;
  }
}

// Same contract, with the bitonic network (78 block-wide passes at 4096 entries) replaced by a stable least-significant-digit
// radix sort on the destination: two 8-bit passes (destinations < 65536), each = per-warp digit counts (match.any groups, no
// atomics), one exclusive scan of the [256 digits x 16 warps] table, and a ranked scatter.  Entries start in ascending entry
// order and the sort is stable, so the order inside a destination is canonical (deterministic sums) without sorting on it.
// Shared memory: two key buffers + weights (12 NP bytes) + the 16 KB count table: 64 KB at NP = 4096 (three blocks per SM).
constexpr int GR_THREADS = 512, GR_WARPS = GR_THREADS / 32, GR_RADIX = 256;
constexpr int GR_LONG = 128;          // runs longer than this are reduced by the whole block
constexpr int GR_MAXLONG = 128;       // NP / GR_LONG such runs at most (16384 / 128)
template <int NP, int EPS>
__global__ void __launch_bounds__(GR_THREADS) gather_bwd_radix_kernel(const GatherCfg cf, const float* __restrict__ vs,
                                                                      const float* __restrict__ dkv, float* __restrict__ dx2) {
  extern __shared__ __align__(16) unsigned char gs_smem[];
  uint32_t* keyA = reinterpret_cast<uint32_t*>(gs_smem);              // (dest << 16) | entry
  uint32_t* keyB = keyA + NP;
  float* wgt = reinterpret_cast<float*>(keyB + NP);
  uint32_t* hist = reinterpret_cast<uint32_t*>(wgt + NP);             // [digit][warp]
  __shared__ uint32_t wsum[GR_WARPS];
  __shared__ int nlong, longpos[GR_MAXLONG];
  const int bg = blockIdx.x, b = bg / cf.G, g = bg % cf.G;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t INVALID = 0xffffffffu;
  constexpr int ESH = EPS == 4 ? 2 : 1;
  static_assert(NP % GR_THREADS == 0 && NP <= 65536, "entry ids are 16 bits");
  if (tid == 0) nlong = 0;
  for (int e = tid; e < NP; e += GR_THREADS) {
    uint32_t key = INVALID;
    float w = 0.f;
    if (e < EPS * cf.nkv) {
      const int p = e >> ESH;
      const Taps t = make_taps(cf, vs, (int64_t)bg * cf.nkv + p);
      const int kx = EPS == 4 ? (e & 1) : -t.x0, ky = EPS == 4 ? ((e >> 1) & 1) : (e & 1);
      const int xx = t.x0 + kx, yy = t.y0 + ky;
      if ((kx == 0 || kx == 1) && xx >= 0 && xx < cf.W && yy >= 0 && yy < cf.H) {
        key = ((uint32_t)(yy * cf.W + xx) << 16) | (uint32_t)e;
        w = (kx ? t.wx1 : 1.f - t.wx1) * (ky ? t.wy1 : 1.f - t.wy1);
      }
    }
    keyA[e] = key; wgt[e] = w;
  }
  // warp w owns the contiguous entries [w E, (w + 1) E) of the source buffer and walks them in order, 32 at a time
  constexpr int E = NP / GR_WARPS;
  uint32_t* src = keyA;
  uint32_t* dst = keyB;
#pragma unroll 1
  for (int shift = 16; shift < 32; shift += 8) {
    for (int i = tid; i < GR_RADIX * GR_WARPS; i += GR_THREADS) hist[i] = 0u;
    __syncthreads();
    for (int r = 0; r < E; r += 32) {
      const uint32_t d = (src[warp * E + r + lane] >> shift) & 0xffu;
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      if (lane == __ffs(peers) - 1) hist[d * GR_WARPS + warp] += __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    {                                                               // exclusive scan of hist (digit-major): 8 values per thread
      constexpr int PER = GR_RADIX * GR_WARPS / GR_THREADS;
      uint32_t v[PER], sum = 0u;
#pragma unroll
      for (int i = 0; i < PER; ++i) { v[i] = hist[tid * PER + i]; sum += v[i]; }
      uint32_t inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      uint32_t base = 0u;
      for (int w2 = 0; w2 < warp; ++w2) base += wsum[w2];
      uint32_t run = base + inc - sum;
#pragma unroll
      for (int i = 0; i < PER; ++i) { hist[tid * PER + i] = run; run += v[i]; }
    }
    __syncthreads();
    for (int r = 0; r < E; r += 32) {
      const uint32_t k = src[warp * E + r + lane];
      const uint32_t d = (k >> shift) & 0xffu;
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      const uint32_t pos = hist[d * GR_WARPS + warp] + __popc(peers & ((1u << lane) - 1u));
      dst[pos] = k;
      __syncwarp();
      if (lane == __ffs(peers) - 1) hist[d * GR_WARPS + warp] += __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    uint32_t* tmp = src; src = dst; dst = tmp;
  }
  const uint32_t* skey = src;                                         // sorted by destination, entry order inside
  // runs of equal destination: one group of cg / 4 lanes (a float4 of channels per lane) per sorted position, 32 / (cg / 4)
  // positions in flight per warp; the group whose position starts a run reduces the run (1-2 entries for bounded offsets;
  // the degenerate 1-D reference pad order puts every sample on one destination and is walked by a single group)
  constexpr int PER_WARP = NP / GR_WARPS;
  if ((cf.cg & 3) == 0) {
    const int qpg = cf.cg >> 2, gpw = 32 / qpg;
    const int sub = lane / qpg, c4 = (lane % qpg) * 4;
    // four positions per group and step: the d kv rows of their first entries are requested together (the loop is bound by
    // the latency of that load, not by bandwidth)
    constexpr int UN = 4;
    const float* dkvb = dkv + (int64_t)b * cf.nkv * cf.C + g * cf.cg + c4;
    for (int pos0 = warp * PER_WARP; pos0 < (warp + 1) * PER_WARP; pos0 += gpw * UN) {
      uint32_t ck[UN];
      bool start[UN];
      float4 gv[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int pos = pos0 + u * gpw + sub;
        ck[u] = pos < (warp + 1) * PER_WARP ? skey[pos] : INVALID;
        start[u] = ck[u] != INVALID && (pos == 0 || (skey[pos - 1] >> 16) != (ck[u] >> 16));
      }
      if (__all_sync(0xffffffffu, ck[0] == INVALID)) break;
#pragma unroll
      for (int u = 0; u < UN; ++u)
        gv[u] = start[u] ? ldg4(dkvb + (int64_t)((int)(ck[u] & 0xffffu) >> ESH) * cf.C) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (!start[u]) continue;
        const uint32_t dest = ck[u] >> 16;
        const float w0 = wgt[ck[u] & 0xffffu];
        float4 acc = make_float4(w0 * gv[u].x, w0 * gv[u].y, w0 * gv[u].z, w0 * gv[u].w);     // = fma(w, g, 0)
        const int first = pos0 + u * gpw + sub;
        bool longrun = false;
        for (int qn = first + 1; qn < NP; ++qn) {
          const uint32_t kq = skey[qn];
          if (kq == INVALID || (kq >> 16) != dest) break;
          if (qn - first >= GR_LONG) { longrun = true; break; }
          const int e = (int)(kq & 0xffffu);
          const float w = wgt[e];
          const float4 g2 = ldg4(dkvb + (int64_t)(e >> ESH) * cf.C);
          acc.x = fmaf(w, g2.x, acc.x); acc.y = fmaf(w, g2.y, acc.y); acc.z = fmaf(w, g2.z, acc.z); acc.w = fmaf(w, g2.w, acc.w);
        }
        if (!longrun) *reinterpret_cast<float4*>(dx2 + ((int64_t)b * cf.H * cf.W + dest) * cf.C + g * cf.cg + c4) = acc;
        else if (c4 == 0) { const int slot = atomicAdd(&nlong, 1); if (slot < GR_MAXLONG) longpos[slot] = first; }
      }
    }
    // Long runs (the degenerate 1-D reference pad order puts every sample of a group on the one or two middle tokens, SURVEY
    // F5): reduced by the whole block - warp w sums a fixed slice of the run, its lane groups fixed strided subsets of the
    // slice, partials are combined in a fixed order, so the result stays deterministic.
    __syncthreads();
    const int nl = min(nlong, GR_MAXLONG);
    for (int li = 0; li < nl; ++li) {
      int first = longpos[0];                                   // process in ascending position order (list order is arbitrary)
      {
        int best = -1;
        for (int q = 0; q < nl; ++q) { const int v = longpos[q]; int rank = 0; for (int r = 0; r < nl; ++r) rank += longpos[r] < v; if (rank == li) best = v; }
        first = best;
      }
      const uint32_t dest = skey[first] >> 16;
      int lo = first, hi = NP;                                   // end of the run: first position with another destination
      while (lo < hi) { const int mid = (lo + hi) >> 1; if ((skey[mid] >> 16) <= dest && skey[mid] != INVALID) lo = mid + 1; else hi = mid; }
      const int len = lo - first, per = (len + GR_WARPS - 1) / GR_WARPS;
      const int s0 = first + warp * per, s1 = min(s0 + per, first + len);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int qn = s0 + sub; qn < s1; qn += gpw) {
        const int e = (int)(skey[qn] & 0xffffu);
        const float w = wgt[e];
        const float4 g2 = ldg4(dkvb + (int64_t)(e >> ESH) * cf.C);
        acc.x = fmaf(w, g2.x, acc.x); acc.y = fmaf(w, g2.y, acc.y); acc.z = fmaf(w, g2.z, acc.z); acc.w = fmaf(w, g2.w, acc.w);
      }
      for (int o = qpg; o < 32; o <<= 1) {                       // lane groups of the warp, fixed tree
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
      }
      float4* part = reinterpret_cast<float4*>(hist);            // the count table is free now: [warp][cg / 4] float4
      __syncthreads();
      if (sub == 0) part[warp * qpg + lane] = acc;
      __syncthreads();
      if (warp == 0 && sub == 0) {
        float4 t = part[lane];
        for (int w2 = 1; w2 < GR_WARPS; ++w2) { const float4 v = part[w2 * qpg + lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
        *reinterpret_cast<float4*>(dx2 + ((int64_t)b * cf.H * cf.W + dest) * cf.C + g * cf.cg + c4) = t;
      }
    }
    return;
  }
  for (int pos = warp * PER_WARP; pos < (warp + 1) * PER_WARP; ++pos) {
    const uint32_t ck = skey[pos];
    if (ck == INVALID) break;
    const uint32_t dest = ck >> 16;
    if (pos > 0 && (skey[pos - 1] >> 16) == dest) continue;
    float acc = 0.f;
    for (int qn = pos; qn < NP; ++qn) {
      const uint32_t kq = skey[qn];
      if (kq == INVALID || (kq >> 16) != dest) break;
      const int e = (int)(kq & 0xffffu);
      if (lane < cf.cg) acc = fmaf(wgt[e], dkv[((int64_t)b * cf.nkv + (e >> ESH)) * cf.C + g * cf.cg + lane], acc);
    }
    if (lane < cf.cg) dx2[((int64_t)b * cf.H * cf.W + dest) * cf.C + g * cf.cg + lane] = acc;
  }
}
template <int NP, int EPS>
static void launch_radix(const GatherCfg& cf, const float* vs, const float* dkv, float* dx2, cudaStream_t st) {
  const int smem = NP * 12 + GR_RADIX * GR_WARPS * 4;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gather_bwd_radix_kernel<NP, EPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  gather_bwd_radix_kernel<NP, EPS><<<(unsigned)(cf.B * cf.G), GR_THREADS, smem, st>>>(cf, vs, dkv, dx2);
}

template <int NP, int EPS>
static void launch_sorted(const GatherCfg& cf, const float* vs, const float* dkv, float* dx2, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gather_bwd_sorted_kernel<NP, EPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, NP * 8);
    attr = true;
  }
  gather_bwd_sorted_kernel<NP, EPS><<<(unsigned)(cf.B * cf.G), GS_THREADS, NP * 8, st>>>(cf, vs, dkv, dx2);
}

// ------------------------------------------------------------------------------------------ fused attention
constexpr int ATQ = 64, ATK = 64, ATHREADS = 128;

struct AttnSmem {
  float Qt[AD][ATQ + 4];     // [dim][query]
  float Kt[AD][ATK + 4];     // [dim][key]
  float V[ATK][AD + 4];      // [key][dim]
  float S[ATQ][ATK + 1];     // logits -> probabilities
  float Pt[ATK][ATQ + 4];    // dropped / scaled probabilities, [key][query] for the PV product
  float kpos[ATK][2];
  float m[ATQ], l[ATQ], alpha[ATQ];
  float W1[CPB_H * 2], b1[CPB_H], W2[CPB_H * CPB_H], b2[CPB_H], w3[CPB_H];
  float b3;
};

// bias for two keys at once (register blocking: one LDS.128 of W2 feeds 8 FMAs)
template <int CD>
__device__ __forceinline__ void cpb_pair(const AttnSmem& sm, float qx, float qy, const float* kp0, const float* kp1,
                                         float& bias0, float& bias1) {
  float h0[CPB_H], h1[CPB_H];
  const float ax = slogf(qx - kp0[0]), bx = slogf(qx - kp1[0]);
  float ay = 0.f, by = 0.f;
  if (CD == 2) { ay = slogf(qy - kp0[1]); by = slogf(qy - kp1[1]); }
#pragma unroll
  for (int k = 0; k < CPB_H; ++k) {
    float u = fmaf(sm.W1[k * CD], ax, sm.b1[k]), v = fmaf(sm.W1[k * CD], bx, sm.b1[k]);
    if (CD == 2) { u = fmaf(sm.W1[k * CD + 1], ay, u); v = fmaf(sm.W1[k * CD + 1], by, v); }
    h0[k] = fmaxf(u, 0.f); h1[k] = fmaxf(v, 0.f);
  }
  float o0 = sm.b3, o1 = sm.b3;
#pragma unroll 4
  for (int c = 0; c < CPB_H; ++c) {
    float a0 = sm.b2[c], a1 = a0;
    const float4* wr = reinterpret_cast<const float4*>(&sm.W2[c * CPB_H]);
#pragma unroll
    for (int k4 = 0; k4 < CPB_H / 4; ++k4) {
      const float4 w = wr[k4];
      a0 = fmaf(w.x, h0[4 * k4], a0); a1 = fmaf(w.x, h1[4 * k4], a1);
      a0 = fmaf(w.y, h0[4 * k4 + 1], a0); a1 = fmaf(w.y, h1[4 * k4 + 1], a1);
      a0 = fmaf(w.z, h0[4 * k4 + 2], a0); a1 = fmaf(w.z, h1[4 * k4 + 2], a1);
      a0 = fmaf(w.w, h0[4 * k4 + 3], a0); a1 = fmaf(w.w, h1[4 * k4 + 3], a1);
    }
    o0 = fmaf(sm.w3[c], fmaxf(a0, 0.f), o0);
    o1 = fmaf(sm.w3[c], fmaxf(a1, 0.f), o1);
  }
  bias0 = o0; bias1 = o1;
}

__device__ __forceinline__ void load_cpb(AttnSmem& sm, const float* __restrict__ cpb, int CD, int OUT, int o) {
  const float* W1 = cpb; const float* b1 = W1 + CPB_H * CD; const float* W2 = b1 + CPB_H;
  const float* b2 = W2 + CPB_H * CPB_H; const float* W3 = b2 + CPB_H; const float* b3 = W3 + OUT * CPB_H;
  for (int i = threadIdx.x; i < CPB_H * CD; i += blockDim.x) sm.W1[i] = W1[i];
  for (int i = threadIdx.x; i < CPB_H; i += blockDim.x) { sm.b1[i] = b1[i]; sm.b2[i] = b2[i]; sm.w3[i] = W3[o * CPB_H + i]; }
  for (int i = threadIdx.x; i < CPB_H * CPB_H; i += blockDim.x) sm.W2[i] = W2[i];
  if (threadIdx.x == 0) sm.b3 = b3[o];
}

// grid (ceil(n/ATQ), heads, B).  q is pre-scaled by d^-1/2.  out [b, n, inner] (+ lse [b, heads, n]).
template <int CD>
__global__ void __launch_bounds__(ATHREADS) deform_attn_fwd_kernel(const AttnCfg cf, const float* __restrict__ q,
                                                                   const float* __restrict__ k, const float* __restrict__ v,
                                                                   const float* __restrict__ vs, const float* __restrict__ cpb,
                                                                   float* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * ATQ;
  const int g = head / cf.OUT, o = head % cf.OUT;
  const int bg = b * cf.G + g;
  load_cpb(sm, cpb, CD, cf.OUT, o);
  // Q tile, transposed
  for (int e = tid; e < ATQ * AD; e += ATHREADS) {
    const int r = e / AD, d = e % AD;
    sm.Qt[d][r] = (q0 + r < cf.n) ? q[((int64_t)b * cf.n + q0 + r) * cf.inner + head * AD + d] : 0.f;
  }
  if (tid < ATQ) { sm.m[tid] = -INFINITY; sm.l[tid] = 0.f; }
  const int ty = tid / 8, tx = tid % 8;        // 16 x 8 thread grid: rows ty*4.., cols / dims tx*8..
  float oacc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) oacc[i][j] = 0.f;
  const int qi = tid % ATQ, cb = tid / ATQ;    // CPB phase: my query row and my 32-key half
  float qx, qy;
  query_pos(cf, q0 + qi, qx, qy);
  const float keep_scale = cf.p_drop > 0.f ? 1.f / (1.f - cf.p_drop) : 1.f;
  const uint32_t thresh = (uint32_t)(cf.p_drop * 4294967296.0);

  for (int k0 = 0; k0 < cf.nkv; k0 += ATK) {
    __syncthreads();
    for (int e = tid; e < ATK * AD; e += ATHREADS) {
      const int r = e / AD, d = e % AD;
      const bool ok = k0 + r < cf.nkv;
      const int64_t off = ((int64_t)b * cf.nkv + k0 + r) * cf.inner + head * AD + d;
      sm.Kt[d][r] = ok ? k[off] : 0.f;
      sm.V[r][d] = ok ? v[off] : 0.f;
    }
    if (tid < ATK) {
      const bool ok = k0 + tid < cf.nkv;
      sm.kpos[tid][0] = ok ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD] : 0.f;
      sm.kpos[tid][1] = (ok && CD == 2) ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD + 1] : 0.f;
    }
    __syncthreads();
    // ---- S = Q K^T
    {
      float s[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s[i][j] = 0.f;
#pragma unroll 8
      for (int d = 0; d < AD; ++d) {
        const float4 a = *reinterpret_cast<const float4*>(&sm.Qt[d][ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&sm.Kt[d][tx * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&sm.Kt[d][tx * 8 + 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) sm.S[ty * 4 + i][tx * 8 + j] = s[i][j];
    }
    __syncthreads();
    // ---- + continuous position bias (the dominant cost: ~1.1 kFMA per pair)
    for (int j = cb * 32; j < cb * 32 + 32; j += 2) {
      float b0, b1;
      cpb_pair<CD>(sm, qx, qy, sm.kpos[j], sm.kpos[j + 1], b0, b1);
      sm.S[qi][j] = (k0 + j < cf.nkv) ? sm.S[qi][j] + b0 : -INFINITY;
      sm.S[qi][j + 1] = (k0 + j + 1 < cf.nkv) ? sm.S[qi][j + 1] + b1 : -INFINITY;
    }
    __syncthreads();
    // ---- online softmax: warp w owns rows w*16 .. w*16+15
    {
      const int lane = tid & 31, w = tid >> 5;
      for (int r = w * 16; r < w * 16 + 16; ++r) {
        const float s0 = sm.S[r][lane], s1 = sm.S[r][lane + 32];
        const float mt = warp_max(fmaxf(s0, s1));
        const float mo = sm.m[r];
        const float mn = fmaxf(mo, mt);
        const float p0 = __expf(s0 - mn), p1 = __expf(s1 - mn);
        const float ps = warp_sum(p0 + p1);
        float d0 = p0, d1 = p1;
        if (cf.p_drop > 0.f) {
          const unsigned long long base = (((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + r)) * cf.nkv) + k0;
          d0 = pair_hash(salted(cf.seed, cf.salt), base + lane) >= thresh ? p0 * keep_scale : 0.f;
          d1 = pair_hash(salted(cf.seed, cf.salt), base + lane + 32) >= thresh ? p1 * keep_scale : 0.f;
        }
        sm.Pt[lane][r] = d0;
        sm.Pt[lane + 32][r] = d1;
        if (lane == 0) {
          const float al = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
          sm.alpha[r] = al;
          sm.l[r] = sm.l[r] * al + ps;
          sm.m[r] = mn;
        }
      }
    }
    __syncthreads();
    // ---- O = alpha * O + P V
    {
      const float4 al = *reinterpret_cast<const float4*>(&sm.alpha[ty * 4]);
      const float alv[4] = {al.x, al.y, al.z, al.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) oacc[i][j] *= alv[i];
#pragma unroll 8
      for (int j = 0; j < ATK; ++j) {
        const float4 p = *reinterpret_cast<const float4*>(&sm.Pt[j][ty * 4]);
        const float4 v0 = *reinterpret_cast<const float4*>(&sm.V[j][tx * 8]);
        const float4 v1 = *reinterpret_cast<const float4*>(&sm.V[j][tx * 8 + 4]);
        const float pv[4] = {p.x, p.y, p.z, p.w};
        const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) oacc[i][jj] = fmaf(pv[i], vv[jj], oacc[i][jj]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (q0 + r >= cf.n) continue;
    const float inv = 1.f / sm.l[r];
    float* orow = out + ((int64_t)b * cf.n + q0 + r) * cf.inner + head * AD + tx * 8;
    *reinterpret_cast<float4*>(orow) = make_float4(oacc[i][0] * inv, oacc[i][1] * inv, oacc[i][2] * inv, oacc[i][3] * inv);
    *reinterpret_cast<float4*>(orow + 4) = make_float4(oacc[i][4] * inv, oacc[i][5] * inv, oacc[i][6] * inv, oacc[i][7] * inv);
  }
  if (tid < ATQ && q0 + tid < cf.n) lse[((int64_t)b * cf.heads + head) * cf.n + q0 + tid] = sm.m[tid] + logf(sm.l[tid]);
}


// ------------------------------------------------------------------------------------------ attention backward
// One block per (query tile, head, batch); loops over key tiles.  Recomputes logits + bias, forms
// dS = P * (dP - delta), accumulates dQ in registers, dK / dV / d key-coordinates with global atomics (each
// element receives one add per query tile), and the CPB parameter gradients in registers (dW2 as a per-warp
// 32x32 outer-product accumulation over staged (h1, h2, dS) rows) -> one partial row per block.
constexpr int BTK = 32;
struct AttnBwdSmem {
  float Q[ATQ][AD + 4], dO[ATQ][AD + 4];
  float K[BTK][AD + 4], V[BTK][AD + 4];
  float S[ATQ][BTK + 1], dP[ATQ][BTK + 1];
  float bufA[4][32][33], bufB[4][32][33];
  float bufS[4][32], bufP[4][32][2];
  float kpos[BTK][2], dkpos[BTK][2];
  float lse[ATQ], delta[ATQ];
  float W1[CPB_H * 2], b1[CPB_H], W2[CPB_H * CPB_H], W2t[CPB_H * CPB_H], b2[CPB_H], w3[CPB_H];
  float b3;
  float red[CPB_H * 2 + CPB_H + CPB_H * CPB_H + CPB_H + CPB_H + 1];
};

template <int CD>
__global__ void __launch_bounds__(ATHREADS) deform_attn_bwd_kernel(
    const AttnCfg cf, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const float* __restrict__ vs, const float* __restrict__ cpb, const float* __restrict__ out,
    const float* __restrict__ lse, const float* __restrict__ dout, float* __restrict__ dq, float* __restrict__ dk,
    float* __restrict__ dv, float* __restrict__ dvs, float* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnBwdSmem& sm = *reinterpret_cast<AttnBwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * ATQ;
  const int g = head / cf.OUT, o = head % cf.OUT;
  const int bg = b * cf.G + g;
  {
    const float* W1 = cpb; const float* b1 = W1 + CPB_H * CD; const float* W2 = b1 + CPB_H;
    const float* b2 = W2 + CPB_H * CPB_H; const float* W3 = b2 + CPB_H; const float* b3 = W3 + cf.OUT * CPB_H;
    for (int i = tid; i < CPB_H * CD; i += ATHREADS) sm.W1[i] = W1[i];
    for (int i = tid; i < CPB_H; i += ATHREADS) { sm.b1[i] = b1[i]; sm.b2[i] = b2[i]; sm.w3[i] = W3[o * CPB_H + i]; }
    for (int i = tid; i < CPB_H * CPB_H; i += ATHREADS) { sm.W2[i] = W2[i]; sm.W2t[(i % CPB_H) * CPB_H + i / CPB_H] = W2[i]; }
    if (tid == 0) sm.b3 = b3[o];
  }
  for (int e = tid; e < ATQ * AD; e += ATHREADS) {
    const int r = e / AD, d = e % AD;
    const bool ok = q0 + r < cf.n;
    const int64_t off = ((int64_t)b * cf.n + q0 + r) * cf.inner + head * AD + d;
    sm.Q[r][d] = ok ? q[off] : 0.f;
    sm.dO[r][d] = ok ? dout[off] : 0.f;
  }
  __syncthreads();
  if (tid < ATQ) {
    const bool ok = q0 + tid < cf.n;
    float dl = 0.f;
    if (ok) {
      const float* orow = out + ((int64_t)b * cf.n + q0 + tid) * cf.inner + head * AD;
      for (int d = 0; d < AD; ++d) dl = fmaf(sm.dO[tid][d], orow[d], dl);
    }
    sm.delta[tid] = dl;
    sm.lse[tid] = ok ? lse[((int64_t)b * cf.heads + head) * cf.n + q0 + tid] : INFINITY;
  }
  const int ty = tid / 8, tx = tid % 8;
  float dqacc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dqacc[i][j] = 0.f;
  // CPB gradient accumulators (per lane; each warp holds a full copy, reduced at the end)
  const int c0 = (lane / 4) * 4, kk0 = (lane % 4) * 8;
  float gW2[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) gW2[i][j] = 0.f;
  float gb2 = 0.f, gw3 = 0.f, gb1 = 0.f, gW1x = 0.f, gW1y = 0.f, gb3 = 0.f;
  const int qi = tid % ATQ, jb = (tid / ATQ) * 16;
  float qx, qy;
  query_pos(cf, q0 + qi, qx, qy);
  const float keep_scale = cf.p_drop > 0.f ? 1.f / (1.f - cf.p_drop) : 1.f;
  const uint32_t thresh = (uint32_t)(cf.p_drop * 4294967296.0);
  float (*bufA)[33] = sm.bufA[wid];
  float (*bufB)[33] = sm.bufB[wid];

  for (int k0 = 0; k0 < cf.nkv; k0 += BTK) {
    __syncthreads();
    for (int e = tid; e < BTK * AD; e += ATHREADS) {
      const int r = e / AD, d = e % AD;
      const bool ok = k0 + r < cf.nkv;
      const int64_t off = ((int64_t)b * cf.nkv + k0 + r) * cf.inner + head * AD + d;
      sm.K[r][d] = ok ? k[off] : 0.f;
      sm.V[r][d] = ok ? v[off] : 0.f;
    }
    if (tid < BTK) {
      const bool ok = k0 + tid < cf.nkv;
      sm.kpos[tid][0] = ok ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD] : 0.f;
      sm.kpos[tid][1] = (ok && CD == 2) ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD + 1] : 0.f;
      sm.dkpos[tid][0] = 0.f; sm.dkpos[tid][1] = 0.f;
    }
    __syncthreads();
    // ---- S = Q K^T and dP = dO V^T: rows ty*4.., cols tx*4..
    {
      float s[4][4], dp[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[i][j] = 0.f; dp[i][j] = 0.f; }
#pragma unroll 4
      for (int d = 0; d < AD; d += 4) {
        float4 a[4], da[4], bk[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i] = *reinterpret_cast<const float4*>(&sm.Q[ty * 4 + i][d]);
          da[i] = *reinterpret_cast<const float4*>(&sm.dO[ty * 4 + i][d]);
          bk[i] = *reinterpret_cast<const float4*>(&sm.K[tx * 4 + i][d]);
          bv[i] = *reinterpret_cast<const float4*>(&sm.V[tx * 4 + i][d]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s[i][j] = fmaf(a[i].x, bk[j].x, fmaf(a[i].y, bk[j].y, fmaf(a[i].z, bk[j].z, fmaf(a[i].w, bk[j].w, s[i][j]))));
            dp[i][j] = fmaf(da[i].x, bv[j].x, fmaf(da[i].y, bv[j].y, fmaf(da[i].z, bv[j].z, fmaf(da[i].w, bv[j].w, dp[i][j]))));
          }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { sm.S[ty * 4 + i][tx * 4 + j] = s[i][j]; sm.dP[ty * 4 + i][tx * 4 + j] = dp[i][j]; }
    }
    __syncthreads();
    // ---- CPB forward + backward, one pair per thread per round; each warp reduces its own 32 pairs
    for (int r = 0; r < 16; ++r) {
      const int j = jb + r;
      const bool jok = k0 + j < cf.nkv;
      float h1[CPB_H], h2[CPB_H];
      const float dx = qx - sm.kpos[j][0];
      const float dy = (CD == 2) ? qy - sm.kpos[j][1] : 0.f;
      const float px = slogf(dx), py = (CD == 2) ? slogf(dy) : 0.f;
#pragma unroll
      for (int kk = 0; kk < CPB_H; ++kk) {
        float u = fmaf(sm.W1[kk * CD], px, sm.b1[kk]);
        if (CD == 2) u = fmaf(sm.W1[kk * CD + 1], py, u);
        h1[kk] = fmaxf(u, 0.f);
      }
      float bias = sm.b3;
#pragma unroll 4
      for (int c = 0; c < CPB_H; ++c) {
        float a0 = sm.b2[c];
        const float4* wr = reinterpret_cast<const float4*>(&sm.W2[c * CPB_H]);
#pragma unroll
        for (int k4 = 0; k4 < CPB_H / 4; ++k4) {
          const float4 w = wr[k4];
          a0 = fmaf(w.x, h1[4 * k4], a0); a0 = fmaf(w.y, h1[4 * k4 + 1], a0);
          a0 = fmaf(w.z, h1[4 * k4 + 2], a0); a0 = fmaf(w.w, h1[4 * k4 + 3], a0);
        }
        h2[c] = fmaxf(a0, 0.f);
        bias = fmaf(sm.w3[c], h2[c], bias);
      }
      const float sv = sm.S[qi][j] + bias;
      const float P = jok ? __expf(sv - sm.lse[qi]) : 0.f;
      float keep = 1.f;
      if (cf.p_drop > 0.f) {
        const unsigned long long idx = (((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + qi)) * cf.nkv) + k0 + j;
        keep = pair_hash(salted(cf.seed, cf.salt), idx) >= thresh ? keep_scale : 0.f;
      }
      const float dS = P * (sm.dP[qi][j] * keep - sm.delta[qi]);
      sm.S[qi][j] = dS;
      sm.dP[qi][j] = P * keep;
      gb3 += dS;
      // stage h1, h2, dS for the warp-cooperative accumulations
      __syncwarp();
#pragma unroll
      for (int c = 0; c < CPB_H; ++c) { bufA[lane][c] = h1[c]; bufB[lane][c] = h2[c]; }
      sm.bufS[wid][lane] = dS;
      __syncwarp();
      {
        const float4 w3q = *reinterpret_cast<const float4*>(&sm.w3[c0]);
        const float w3l = sm.w3[lane];
        for (int p = 0; p < 32; ++p) {
          const float ds = sm.bufS[wid][p];
          const float hq0 = bufB[p][c0], hq1 = bufB[p][c0 + 1], hq2 = bufB[p][c0 + 2], hq3 = bufB[p][c0 + 3];
          const float d0 = hq0 > 0.f ? ds * w3q.x : 0.f, d1 = hq1 > 0.f ? ds * w3q.y : 0.f;
          const float d2 = hq2 > 0.f ? ds * w3q.z : 0.f, d3 = hq3 > 0.f ? ds * w3q.w : 0.f;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const float hv = bufA[p][kk0 + kk];
            gW2[0][kk] = fmaf(d0, hv, gW2[0][kk]); gW2[1][kk] = fmaf(d1, hv, gW2[1][kk]);
            gW2[2][kk] = fmaf(d2, hv, gW2[2][kk]); gW2[3][kk] = fmaf(d3, hv, gW2[3][kk]);
          }
          const float hl = bufB[p][lane];
          gw3 = fmaf(ds, hl, gw3);
          gb2 += hl > 0.f ? ds * w3l : 0.f;
        }
      }
      // dh1 = W2^T dh2 (thread-local), then through the first ReLU
      float dpx = 0.f, dpy = 0.f;
      {
#pragma unroll
        for (int c = 0; c < CPB_H; ++c) h2[c] = h2[c] > 0.f ? dS * sm.w3[c] : 0.f;    // h2 <- dh2
        __syncwarp();
#pragma unroll 4
        for (int kk = 0; kk < CPB_H; ++kk) {
          float a0 = 0.f;
          const float4* wr = reinterpret_cast<const float4*>(&sm.W2t[kk * CPB_H]);
#pragma unroll
          for (int c4 = 0; c4 < CPB_H / 4; ++c4) {
            const float4 w = wr[c4];
            a0 = fmaf(w.x, h2[4 * c4], a0); a0 = fmaf(w.y, h2[4 * c4 + 1], a0);
            a0 = fmaf(w.z, h2[4 * c4 + 2], a0); a0 = fmaf(w.w, h2[4 * c4 + 3], a0);
          }
          const float dpre = h1[kk] > 0.f ? a0 : 0.f;
          bufA[lane][kk] = dpre;
          dpx = fmaf(sm.W1[kk * CD], dpre, dpx);
          if (CD == 2) dpy = fmaf(sm.W1[kk * CD + 1], dpre, dpy);
        }
        sm.bufP[wid][lane][0] = px; sm.bufP[wid][lane][1] = py;
        __syncwarp();
        for (int p = 0; p < 32; ++p) {
          const float dv1 = bufA[p][lane];
          gb1 += dv1;
          gW1x = fmaf(dv1, sm.bufP[wid][p][0], gW1x);
          if (CD == 2) gW1y = fmaf(dv1, sm.bufP[wid][p][1], gW1y);
        }
      }
      // d slog(x)/dx = 1/(|x|+1); key position enters with a minus sign
      atomicAdd(&sm.dkpos[j][0], -dpx / (fabsf(dx) + 1.f));
      if (CD == 2) atomicAdd(&sm.dkpos[j][1], -dpy / (fabsf(dy) + 1.f));
    }
    __syncthreads();
    // ---- dQ += dS K ; dK += dS^T Q ; dV += Pd^T dO
    {
#pragma unroll 4
      for (int j = 0; j < BTK; ++j) {
        const float4 k0v = *reinterpret_cast<const float4*>(&sm.K[j][tx * 8]);
        const float4 k1v = *reinterpret_cast<const float4*>(&sm.K[j][tx * 8 + 4]);
        const float kv8[8] = {k0v.x, k0v.y, k0v.z, k0v.w, k1v.x, k1v.y, k1v.z, k1v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = sm.S[ty * 4 + i][j];
#pragma unroll
          for (int dd = 0; dd < 8; ++dd) dqacc[i][dd] = fmaf(a, kv8[dd], dqacc[i][dd]);
        }
      }
      float dkacc[2][8], dvacc[2][8];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int dd = 0; dd < 8; ++dd) { dkacc[i][dd] = 0.f; dvacc[i][dd] = 0.f; }
      const int j0 = ty * 2;
#pragma unroll 4
      for (int i = 0; i < ATQ; ++i) {
        const float4 q0v = *reinterpret_cast<const float4*>(&sm.Q[i][tx * 8]);
        const float4 q1v = *reinterpret_cast<const float4*>(&sm.Q[i][tx * 8 + 4]);
        const float4 g0v = *reinterpret_cast<const float4*>(&sm.dO[i][tx * 8]);
        const float4 g1v = *reinterpret_cast<const float4*>(&sm.dO[i][tx * 8 + 4]);
        const float q8[8] = {q0v.x, q0v.y, q0v.z, q0v.w, q1v.x, q1v.y, q1v.z, q1v.w};
        const float g8[8] = {g0v.x, g0v.y, g0v.z, g0v.w, g1v.x, g1v.y, g1v.z, g1v.w};
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const float ds = sm.S[i][j0 + jj], pd = sm.dP[i][j0 + jj];
#pragma unroll
          for (int dd = 0; dd < 8; ++dd) { dkacc[jj][dd] = fmaf(ds, q8[dd], dkacc[jj][dd]); dvacc[jj][dd] = fmaf(pd, g8[dd], dvacc[jj][dd]); }
        }
      }
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        if (k0 + j0 + jj >= cf.nkv) continue;
        const int64_t off = ((int64_t)b * cf.nkv + k0 + j0 + jj) * cf.inner + head * AD + tx * 8;
#pragma unroll
        for (int dd = 0; dd < 8; ++dd) { atomicAdd(dk + off + dd, dkacc[jj][dd]); atomicAdd(dv + off + dd, dvacc[jj][dd]); }
      }
      if (tid < BTK && k0 + tid < cf.nkv) {
        atomicAdd(dvs + ((int64_t)bg * cf.nkv + k0 + tid) * CD, sm.dkpos[tid][0]);
        if (CD == 2) atomicAdd(dvs + ((int64_t)bg * cf.nkv + k0 + tid) * CD + 1, sm.dkpos[tid][1]);
      }
    }
  }
  // ---- dQ
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (q0 + r >= cf.n) continue;
    float* row = dq + ((int64_t)b * cf.n + q0 + r) * cf.inner + head * AD + tx * 8;
    *reinterpret_cast<float4*>(row) = make_float4(dqacc[i][0], dqacc[i][1], dqacc[i][2], dqacc[i][3]);
    *reinterpret_cast<float4*>(row + 4) = make_float4(dqacc[i][4], dqacc[i][5], dqacc[i][6], dqacc[i][7]);
  }
  // ---- CPB parameter partials: [W1 32*CD][b1 32][W2 1024][b2 32][W3 OUT*32][b3 OUT]
  constexpr int NRED = CPB_H * 2 + CPB_H + CPB_H * CPB_H + CPB_H + CPB_H + 1;
  __syncthreads();
  for (int i = tid; i < NRED; i += ATHREADS) sm.red[i] = 0.f;
  __syncthreads();
  {
    float* rW1 = sm.red; float* rb1 = rW1 + CPB_H * 2; float* rW2 = rb1 + CPB_H; float* rb2 = rW2 + CPB_H * CPB_H;
    float* rw3 = rb2 + CPB_H; float* rb3 = rw3 + CPB_H;
    atomicAdd(&rW1[lane * 2], gW1x); atomicAdd(&rW1[lane * 2 + 1], gW1y);
    atomicAdd(&rb1[lane], gb1); atomicAdd(&rb2[lane], gb2); atomicAdd(&rw3[lane], gw3);
    gb3 = warp_sum(gb3);
    if (lane == 0) atomicAdd(rb3, gb3);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) atomicAdd(&rW2[(c0 + i) * CPB_H + kk0 + kk], gW2[i][kk]);
  }
  __syncthreads();
  {
    const int np = CPB_H * CD + CPB_H + CPB_H * CPB_H + CPB_H + cf.OUT * CPB_H + cf.OUT;
    float* prow = part + ((int64_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * np;
    const float* rW1 = sm.red; const float* rb1 = rW1 + CPB_H * 2; const float* rW2 = rb1 + CPB_H;
    const float* rb2 = rW2 + CPB_H * CPB_H; const float* rw3 = rb2 + CPB_H; const float* rb3 = rw3 + CPB_H;
    for (int i = tid; i < np; i += ATHREADS) {
      int e = i;
      float val;
      if (e < CPB_H * CD) val = rW1[(e / CD) * 2 + (e % CD)];
      else if ((e -= CPB_H * CD) < CPB_H) val = rb1[e];
      else if ((e -= CPB_H) < CPB_H * CPB_H) val = rW2[e];
      else if ((e -= CPB_H * CPB_H) < CPB_H) val = rb2[e];
      else if ((e -= CPB_H) < cf.OUT * CPB_H) val = (e / CPB_H == o) ? rw3[e % CPB_H] : 0.f;
      else { e -= cf.OUT * CPB_H; val = (e == o) ? rb3[0] : 0.f; }
      prow[i] = val;
    }
  }
}

}  // namespace sml

using namespace sml;

static OffCfg make_offcfg(const int32_t* c, float scale) {
  OffCfg cf;
  cf.B = c[0]; cf.G = c[1]; cf.gd = c[2]; cf.inner = c[3]; cf.H = c[4]; cf.W = c[5];
  cf.KH = c[6]; cf.KW = c[7]; cf.SH = c[8]; cf.SW = c[9]; cf.PH = c[10]; cf.PW = c[11];
  cf.HK = (cf.H + 2 * cf.PH - cf.KH) / cf.SH + 1;
  cf.WK = (cf.W + 2 * cf.PW - cf.KW) / cf.SW + 1;
  cf.OC = c[12];
  cf.scale = scale;
  if (cf.OC == 2) { cf.den0 = (float)(cf.HK - 1 > 1 ? cf.HK - 1 : 1); cf.den1 = (float)(cf.WK - 1 > 1 ? cf.WK - 1 : 1); }
  else { cf.den0 = (float)(cf.WK - 1 > 1 ? cf.WK - 1 : 1); cf.den1 = 1.f; }
  return cf;
}

extern "C" int sml_deform_offsets_fwd(const int32_t* cfg13, float scale, const float* q, const float* wc, const float* bc,
                                      const float* w2, float* off, float* vgrid, float* vs, void* stream) {
  const OffCfg cf = make_offcfg(cfg13, scale);
  SML_CHECK(cf.HK > 0 && cf.WK > 0 && (cf.OC == 1 || cf.OC == 2), "sml_deform_offsets_fwd: bad geometry");
  const int64_t locs = (int64_t)cf.B * cf.G * cf.HK * cf.WK;
  offsets_fwd_kernel<<<(unsigned)((locs + 7) / 8), 256, 0, as_stream(stream)>>>(cf, q, wc, bc, w2, off, vgrid, vs);
  SML_LAUNCH_OK("sml_deform_offsets_fwd");
  return 0;
}
extern "C" int sml_deform_offsets_bwd_blocks(const int32_t* cfg13) {
  const OffCfg cf = make_offcfg(cfg13, 1.f);
  const int64_t locs = (int64_t)cf.B * cf.G * cf.HK * cf.WK;
  int64_t nb = (locs + 31) / 32;
  if (nb > 2 * kNumSMs) nb = 2 * kNumSMs;
  return (int)(nb < 1 ? 1 : nb);
}
extern "C" int sml_deform_offsets_bwd(const int32_t* cfg13, float scale, const float* q, const float* wc, const float* bc,
                                      const float* w2, const float* off, const float* dvs, const float* dvgrid,
                                      float* dpre, float* part, float* dq, void* stream) {
  const OffCfg cf = make_offcfg(cfg13, scale);
  const int nb = sml_deform_offsets_bwd_blocks(cfg13);
  const int64_t locs = (int64_t)cf.B * cf.G * cf.HK * cf.WK;
  const int64_t lpb = (locs + nb - 1) / nb;
  const int nacc = cf.gd * cf.KH * cf.KW + cf.gd + cf.OC * cf.gd;
  SML_CHECK(nacc * 4 <= 48 * 1024, "sml_deform_offsets_bwd: accumulator block too large");
  cudaStream_t s = as_stream(stream);
  if (cf.KH == 1 && cf.KW == 6 && cf.gd == 128)
    offsets_bwd_reg_kernel<1, 6, 4><<<nb, 256, nacc * sizeof(float), s>>>(cf, q, wc, bc, w2, off, dvs, dvgrid, dpre, part, lpb);
  else if (cf.KH == 6 && cf.KW == 6 && cf.gd == 64)
    offsets_bwd_reg_kernel<6, 6, 2><<<nb, 256, nacc * sizeof(float), s>>>(cf, q, wc, bc, w2, off, dvs, dvgrid, dpre, part, lpb);
  else
    offsets_bwd_kernel<<<nb, 256, nacc * sizeof(float), s>>>(cf, q, wc, bc, w2, off, dvs, dvgrid, dpre, part, lpb);
  SML_LAUNCH_OK("sml_deform_offsets_bwd");
  offsets_dq_kernel<<<dim3((unsigned)(cf.B * ((cf.H * cf.W + ODQ_PIX - 1) / ODQ_PIX)), (unsigned)((cf.G * cf.gd + 255) / 256)), 256, 0, s>>>(cf, wc, dpre, dq);
  SML_LAUNCH_OK("sml_deform_offsets_bwd(dq)");
  return 0;
}

static GatherCfg make_gcfg(const int32_t* c) {
  GatherCfg cf;
  cf.B = c[0]; cf.G = c[1]; cf.cg = c[2]; cf.C = c[1] * c[2]; cf.H = c[3]; cf.W = c[4]; cf.nkv = c[5]; cf.CD = c[6]; cf.mode = c[7];
  return cf;
}
extern "C" int sml_deform_gather_fwd(const int32_t* cfg8, const float* x2, const float* vs, float* kv, int32_t* taps,
                                     void* stream) {
  const GatherCfg cf = make_gcfg(cfg8);
  if (cf.cg % 4 == 0) {
    const int64_t total = (int64_t)cf.B * cf.G * cf.nkv * (cf.cg / 4);
    gather_fwd_vec_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(cf, x2, vs, kv, taps);
  } else {
    const int64_t total = (int64_t)cf.B * cf.G * cf.nkv * cf.cg;
    gather_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(cf, x2, vs, kv, taps);
  }
  SML_LAUNCH_OK("sml_deform_gather_fwd");
  return 0;
}
extern "C" int sml_deform_gather_bwd(const int32_t* cfg8, const float* x2, const float* vs, const float* dkv, float* dx2,
                                     const float* dvs_in, float* dvs_out, void* stream) {
  const GatherCfg cf = make_gcfg(cfg8);
  SML_CHECK(cf.cg <= 32 && (cf.cg & (cf.cg - 1)) == 0, "sml_deform_gather_bwd: channels per group must be a power of two <= 32");
  // d x2: sorted, atomic-free, deterministic scatter whenever the (b, g) entry list fits shared memory (<= 16384 entries:
  // 2-D lattices up to n_kv = 4096, 1-D bags up to n_kv = 8192) and the composite key fits 32 bits; fp32 atomics beyond that
  const int eps = cf.CD == 1 ? 2 : 4;
  const int entries = eps * cf.nkv;
  int np = 1024;
  while (np < entries) np <<= 1;
  const bool sorted = np <= 16384 && (int64_t)cf.H * cf.W * np <= (1ll << 32);
  cudaStream_t st = as_stream(stream);
  if (cf.cg % 4 == 0) {
    const int64_t total = (int64_t)cf.B * cf.G * cf.nkv * (cf.cg / 4);
    gather_bwd_vec_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(cf, x2, vs, dkv, dx2, dvs_in, dvs_out, !sorted);
  } else {
    const int64_t total = (int64_t)cf.B * cf.G * cf.nkv * cf.cg;
    gather_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(cf, x2, vs, dkv, dx2, dvs_in, dvs_out, !sorted);
  }
  SML_LAUNCH_OK("sml_deform_gather_bwd");
  if (sorted && (int64_t)cf.H * cf.W < 65535) {
    // destinations and entry ids fit 16 bits each: radix-sorted scatter
    if (eps == 4) {
      if (np == 1024) launch_radix<1024, 4>(cf, vs, dkv, dx2, st);
      else if (np == 2048) launch_radix<2048, 4>(cf, vs, dkv, dx2, st);
      else if (np == 4096) launch_radix<4096, 4>(cf, vs, dkv, dx2, st);
      else if (np == 8192) launch_radix<8192, 4>(cf, vs, dkv, dx2, st);
      else launch_radix<16384, 4>(cf, vs, dkv, dx2, st);
    } else {
      if (np == 1024) launch_radix<1024, 2>(cf, vs, dkv, dx2, st);
      else if (np == 2048) launch_radix<2048, 2>(cf, vs, dkv, dx2, st);
      else if (np == 4096) launch_radix<4096, 2>(cf, vs, dkv, dx2, st);
      else if (np == 8192) launch_radix<8192, 2>(cf, vs, dkv, dx2, st);
      else launch_radix<16384, 2>(cf, vs, dkv, dx2, st);
    }
    SML_LAUNCH_OK("sml_deform_gather_bwd(radix-sorted scatter)");
  } else if (sorted) {
    if (eps == 4) {
      if (np == 1024) launch_sorted<1024, 4>(cf, vs, dkv, dx2, st);
      else if (np == 2048) launch_sorted<2048, 4>(cf, vs, dkv, dx2, st);
      else if (np == 4096) launch_sorted<4096, 4>(cf, vs, dkv, dx2, st);
      else if (np == 8192) launch_sorted<8192, 4>(cf, vs, dkv, dx2, st);
      else launch_sorted<16384, 4>(cf, vs, dkv, dx2, st);
    } else {
      if (np == 1024) launch_sorted<1024, 2>(cf, vs, dkv, dx2, st);
      else if (np == 2048) launch_sorted<2048, 2>(cf, vs, dkv, dx2, st);
      else if (np == 4096) launch_sorted<4096, 2>(cf, vs, dkv, dx2, st);
      else if (np == 8192) launch_sorted<8192, 2>(cf, vs, dkv, dx2, st);
      else launch_sorted<16384, 2>(cf, vs, dkv, dx2, st);
    }
    SML_LAUNCH_OK("sml_deform_gather_bwd(sorted scatter)");
  }
  return 0;
}

extern "C" int sml_deform_attn_fwd(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const float* k,
                                   const float* v, const float* vs, const float* cpb, float* out, float* lse,
                                   void* stream) {
  AttnCfg cf;
  if (make_acfg(cf, cfg8, p_drop, seed)) return 1;
  dim3 grid((cf.n + ATQ - 1) / ATQ, cf.heads, cf.B);
  const size_t smem = sizeof(AttnSmem);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(deform_attn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(deform_attn_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  if (cf.CD == 2) deform_attn_fwd_kernel<2><<<grid, ATHREADS, smem, as_stream(stream)>>>(cf, q, k, v, vs, cpb, out, lse);
  else deform_attn_fwd_kernel<1><<<grid, ATHREADS, smem, as_stream(stream)>>>(cf, q, k, v, vs, cpb, out, lse);
  SML_LAUNCH_OK("sml_deform_attn_fwd");
  return 0;
}

extern "C" int sml_deform_attn_bwd_blocks(const int32_t* cfg8) {
  return cfg8[0] * cfg8[1] * ((cfg8[2] + ATQ - 1) / ATQ);
}
extern "C" int sml_deform_cpb_nparams(int32_t CD, int32_t OUT) { return CPB_H * CD + CPB_H + CPB_H * CPB_H + CPB_H + OUT * CPB_H + OUT; }
/* dk, dv, dvs must be zero-initialised by the caller (they are accumulated with atomics) */
extern "C" int sml_deform_attn_bwd(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const float* k,
                                   const float* v, const float* vs, const float* cpb, const float* out, const float* lse,
                                   const float* dout, float* dq, float* dk, float* dv, float* dvs, float* part,
                                   void* stream) {
  AttnCfg cf;
  if (make_acfg(cf, cfg8, p_drop, seed)) return 1;
  dim3 grid((cf.n + ATQ - 1) / ATQ, cf.heads, cf.B);
  const size_t smem = sizeof(AttnBwdSmem);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(deform_attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(deform_attn_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  if (cf.CD == 2) deform_attn_bwd_kernel<2><<<grid, ATHREADS, smem, as_stream(stream)>>>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs, part);
  else deform_attn_bwd_kernel<1><<<grid, ATHREADS, smem, as_stream(stream)>>>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs, part);
  SML_LAUNCH_OK("sml_deform_attn_bwd");
  return 0;
}

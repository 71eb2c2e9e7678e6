// rowwise.cu — HBM-bound row kernels: LayerNorm fwd/bwd, softmax fwd/bwd, dropout+add, axpby, column sums.
#include "common.cuh"

namespace sml {

// ------------------------------------------------------------------ LayerNorm
// one warp per row; D % 4 == 0 (float4 I/O), D <= 1024
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, float* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd,
                                                            int64_t R, int D, float eps, int64_t n, int64_t npad,
                                                            int64_t pad, __nv_bfloat16* __restrict__ yhi,
                                                            __nv_bfloat16* __restrict__ ylo) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int64_t yrow = (row / n) * npad + pad + (row % n);   // destination row in the front-padded layout
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  const int nv = D >> 2;
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + i * 32;
    if (c < nv) { v[i] = xr[c]; s += v[i].x + v[i].y + v[i].z + v[i].w; }
  }
  const float mu = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float a = v[i].x - mu, bb = v[i].y - mu, cc = v[i].z - mu, dd = v[i].w - mu;
      q += a * a + bb * bb + cc * cc + dd * dd;
    }
  }
  const float rs = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  float4* yr = y ? reinterpret_cast<float4*>(y + yrow * D) : nullptr;      // fp32 output optional (tensor-core modes)
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float4 ww = __ldg(w4 + c), bb = __ldg(b4 + c);
      float4 o;
      o.x = (v[i].x - mu) * rs * ww.x + bb.x;
      o.y = (v[i].y - mu) * rs * ww.y + bb.y;
      o.z = (v[i].z - mu) * rs * ww.z + bb.z;
      o.w = (v[i].w - mu) * rs * ww.w + bb.w;
      if (yr) yr[c] = o;
      if (yhi) {
        const int64_t e = yrow * D + 4 * c;
        store_split(yhi, ylo, e, o.x); store_split(yhi, ylo, e + 1, o.y);
        store_split(yhi, ylo, e + 2, o.z); store_split(yhi, ylo, e + 3, o.w);
      }
    }
  }
}

// Each block owns a contiguous slab of rows; a warp handles rows, accumulating dw/db partials in registers,
// then the block reduces them through shared memory into dw_part[blk], db_part[blk].
constexpr int LNB_WARPS = 8;
template <int VPL>
__global__ void __launch_bounds__(LNB_WARPS * 32) layernorm_bwd_kernel(
    const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
    const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx,
    float* __restrict__ dw_part, float* __restrict__ db_part, int64_t R, int D, int64_t rows_per_block, int64_t n,
    int64_t npad, int64_t pad, const float* __restrict__ res) {
  extern __shared__ float sm[];   // [LNB_WARPS][2][D]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nv = D >> 2;
  float4 aw[VPL], ab[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) { aw[i] = make_float4(0, 0, 0, 0); ab[i] = make_float4(0, 0, 0, 0); }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(R, r0 + rows_per_block);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  for (int64_t row = r0 + wid; row < r1; row += LNB_WARPS) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    const float4* gr = reinterpret_cast<const float4*>(dy + ((row / n) * npad + pad + (row % n)) * D);
    const float mu = mean[row], rs = rstd[row];
    float4 xh[VPL], g[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        const float4 xv = xr[c], gv = gr[c], ww = __ldg(w4 + c);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        ab[i].x += gv.x; ab[i].y += gv.y; ab[i].z += gv.z; ab[i].w += gv.w;
        aw[i].x += gv.x * xh[i].x; aw[i].y += gv.y * xh[i].y; aw[i].z += gv.z * xh[i].z; aw[i].w += gv.w * xh[i].w;
        g[i] = make_float4(gv.x * ww.x, gv.y * ww.y, gv.z * ww.z, gv.w * ww.w);
        s1 += g[i].x + g[i].y + g[i].z + g[i].w;
        s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
      }
    }
    s1 = warp_sum(s1) / D; s2 = warp_sum(s2) / D;
    float4* dr = reinterpret_cast<float4*>(dx + row * D);
    const float4* rr = res ? reinterpret_cast<const float4*>(res + row * D) : nullptr;   // + gradient of the skip connection
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        float4 o;
        o.x = rs * (g[i].x - s1 - xh[i].x * s2);
        o.y = rs * (g[i].y - s1 - xh[i].y * s2);
        o.z = rs * (g[i].z - s1 - xh[i].z * s2);
        o.w = rs * (g[i].w - s1 - xh[i].w * s2);
        if (rr) { const float4 rv = rr[c]; o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w; }
        dr[c] = o;
      }
    }
  }
  float4* sw = reinterpret_cast<float4*>(sm) + (size_t)wid * 2 * nv;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + i * 32;
    if (c < nv) { sw[c] = aw[i]; sw[nv + c] = ab[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < LNB_WARPS; ++wv) s += sm[(size_t)wv * 2 * D + c];
    if (c < D) dw_part[(size_t)blockIdx.x * D + c] = s;
    else db_part[(size_t)blockIdx.x * D + (c - D)] = s;
  }
}

// ------------------------------------------------------------------ softmax
// short rows (C <= 1024): one warp per row, row kept in registers
template <int VPT>
__global__ void __launch_bounds__(256) softmax_fwd_warp_kernel(float* __restrict__ x, int64_t R, int C,
                                                               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                               int write_f32) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  float* xr = x + row * C;
  float v[VPT];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    v[i] = (c < C) ? xr[c] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) { v[i] = __expf(v[i] - mx); s += v[i]; }
  const float inv = 1.f / warp_sum(s);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    if (c < C) { if (write_f32) xr[c] = v[i] * inv; store_split(hi, lo, row * C + c, v[i] * inv); }
  }
}
// long rows: one block per row, three passes through L2-resident data
__global__ void __launch_bounds__(1024) softmax_fwd_block_kernel(float* __restrict__ x, int64_t C,
                                                                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                                 int write_f32) {
  __shared__ float red[32];
  float* xr = x + (int64_t)blockIdx.x * C;
  float mx = -INFINITY;
  for (int64_t c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, xr[c]);
  mx = block_max(mx, red);
  float s = 0.f;
  for (int64_t c = threadIdx.x; c < C; c += blockDim.x) s += __expf(xr[c] - mx);
  s = block_sum(s, red);
  const float inv = 1.f / s;
  for (int64_t c = threadIdx.x; c < C; c += blockDim.x) {     // the row (<= a few hundred KB) is L2-resident by now
    const float pv = __expf(xr[c] - mx) * inv;
    if (write_f32) xr[c] = pv;
    store_split(hi, lo, (int64_t)blockIdx.x * C + c, pv);
  }
}
template <int VPT>
__global__ void __launch_bounds__(256) softmax_bwd_warp_kernel(const float* __restrict__ A, const __nv_bfloat16* __restrict__ Ab,
                                                               float* __restrict__ dA, int64_t R, int C,
                                                               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                               int write_f32) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  float* gr = dA + row * C;
  float a[VPT], g[VPT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    a[i] = (c < C) ? (A ? A[row * C + c] : __bfloat162float(Ab[row * C + c])) : 0.f;
    g[i] = (c < C) ? gr[c] : 0.f;
    s += a[i] * g[i];
  }
  s = warp_sum(s);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    if (c < C) { const float dsv = a[i] * (g[i] - s); if (write_f32) gr[c] = dsv; store_split(hi, lo, row * C + c, dsv); }
  }
}
__global__ void __launch_bounds__(1024) softmax_bwd_block_kernel(const float* __restrict__ A, const __nv_bfloat16* __restrict__ Ab,
                                                                 float* __restrict__ dA, int64_t C, __nv_bfloat16* __restrict__ hi,
                                                                 __nv_bfloat16* __restrict__ lo, int write_f32) {
  __shared__ float red[32];
  const int64_t base = (int64_t)blockIdx.x * C;
  float* gr = dA + base;
  float s = 0.f;
  for (int64_t c = threadIdx.x; c < C; c += blockDim.x) s += (A ? A[base + c] : __bfloat162float(Ab[base + c])) * gr[c];
  s = block_sum(s, red);
  for (int64_t c = threadIdx.x; c < C; c += blockDim.x) {
    const float dsv = (A ? A[base + c] : __bfloat162float(Ab[base + c])) * (gr[c] - s);
    if (write_f32) gr[c] = dsv;
    store_split(hi, lo, base + c, dsv);
  }
}


// ---- long rows, register-resident: one 1024-thread block per row, NV float4 per thread (C <= 4096*NV, C % 4 == 0).
// One global read, exp evaluated once, 128-bit loads / 64-bit bf16 stores; replaces the 3-pass kernel whose time went
// into per-element 64-bit index arithmetic (ncu: 45 % integer instructions, 29 % of DRAM peak).
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t e, float4 v) {
  if (hi) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&h0); pk.y = *reinterpret_cast<const uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(hi + e) = pk;
    if (lo) {
      const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
      const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
      pk.x = *reinterpret_cast<const uint32_t*>(&l0); pk.y = *reinterpret_cast<const uint32_t*>(&l1);
      *reinterpret_cast<uint2*>(lo + e) = pk;
    }
  }
}
template <int NV>
__global__ void __launch_bounds__(1024) softmax_fwd_reg_kernel(float* __restrict__ x, int C4, __nv_bfloat16* __restrict__ hi,
                                                               __nv_bfloat16* __restrict__ lo, int write_f32) {
  __shared__ float red[32];
  const int64_t base = (int64_t)blockIdx.x * C4 * 4;
  float4* xr = reinterpret_cast<float4*>(x + base);
  float4 v[NV];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 1024;
    if (c < C4) { v[i] = xr[c]; mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w))); }
  }
  mx = block_max(mx, red);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 1024;
    if (c < C4) {
      v[i].x = __expf(v[i].x - mx); v[i].y = __expf(v[i].y - mx); v[i].z = __expf(v[i].z - mx); v[i].w = __expf(v[i].w - mx);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  s = block_sum(s, red);
  const float inv = 1.f / s;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 1024;
    if (c < C4) {
      const float4 p = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
      if (write_f32) xr[c] = p;
      store_bf16x4(hi, lo, base + 4 * (int64_t)c, p);
    }
  }
}
template <int NV>
__global__ void __launch_bounds__(1024) softmax_bwd_reg_kernel(const float* __restrict__ A, const __nv_bfloat16* __restrict__ Ab,
                                                               float* __restrict__ dA, int C4, __nv_bfloat16* __restrict__ hi,
                                                               __nv_bfloat16* __restrict__ lo, int write_f32) {
  __shared__ float red[32];
  const int64_t base = (int64_t)blockIdx.x * C4 * 4;
  float4* gr = reinterpret_cast<float4*>(dA + base);
  float4 a[NV], g[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 1024;
    if (c < C4) {
      g[i] = gr[c];
      if (A) a[i] = reinterpret_cast<const float4*>(A + base)[c];
      else {
        const uint2 pk = reinterpret_cast<const uint2*>(Ab + base)[c];
        const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
        const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
        a[i] = make_float4(f0.x, f0.y, f1.x, f1.y);
      }
      s += (a[i].x * g[i].x + a[i].y * g[i].y) + (a[i].z * g[i].z + a[i].w * g[i].w);
    }
  }
  s = block_sum(s, red);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 1024;
    if (c < C4) {
      const float4 d4 = make_float4(a[i].x * (g[i].x - s), a[i].y * (g[i].y - s), a[i].z * (g[i].z - s), a[i].w * (g[i].w - s));
      if (write_f32) gr[c] = d4;
      store_bf16x4(hi, lo, base + 4 * (int64_t)c, d4);
    }
  }
}

// ------------------------------------------------------------------ dropout + add (counter-based RNG)
__device__ __forceinline__ uint32_t mix32(uint64_t z) {   // splitmix64 finaliser -> 32 bits
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 16);
}
__global__ void __launch_bounds__(256) dropout_add_kernel(const float* __restrict__ x, const float* __restrict__ u,
                                                          float* __restrict__ y, int64_t n, float p, uint64_t seed,
                                                          uint64_t offset, const unsigned long long* __restrict__ salt) {
  seed = salted(seed, salt);
  const float scale = 1.f / (1.f - p);
  const uint32_t thresh = (uint32_t)(p * 4294967296.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float uv = u[i];
    if (p > 0.f) {
      const uint32_t r = mix32(seed * 0xD1342543DE82EF95ull + offset + (uint64_t)i);
      uv = (r >= thresh) ? uv * scale : 0.f;
    }
    y[i] = (x ? x[i] : 0.f) + uv;
  }
}
// AlphaDropout (models/model.py:152-170, torch.nn.AlphaDropout): kept elements a x + b, dropped ones a alpha' + b with
// alpha' = -selu_scale * selu_alpha, a = ((1 - p)(1 + p alpha'^2))^-1/2, b = -a alpha' p.  bwd = 1: y = dx given x = dy.
__global__ void __launch_bounds__(256) alpha_dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p,
                                                            float a, float dropped, float b, int bwd, uint64_t seed,
                                                            uint64_t offset, const unsigned long long* __restrict__ salt) {
  seed = salted(seed, salt);
  const uint32_t thresh = (uint32_t)(p * 4294967296.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool keep = mix32(seed * 0xD1342543DE82EF95ull + offset + (uint64_t)i) >= thresh;
    y[i] = bwd ? (keep ? a * x[i] : 0.f) : (keep ? fmaf(a, x[i], b) : dropped);
  }
}
__global__ void __launch_bounds__(256) axpby_kernel(float* __restrict__ y, const float* __restrict__ x, float a, float b,
                                                    int64_t n, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = a * x[i] + (b != 0.f ? b * y[i] : 0.f);
    y[i] = v;
    store_split(hi, lo, i, v);
  }
}
__global__ void __launch_bounds__(256) diag_axpby_kernel(float* __restrict__ out, const float* __restrict__ x, float a,
                                                         float dg, int64_t total, int m, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i % ((int64_t)m * m);
    const int r = (int)(e / m), c = (int)(e % m);
    const float v = a * x[i] + (r == c ? dg : 0.f);
    out[i] = v;
    store_split(hi, lo, i, v);
  }
}
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                       float* __restrict__ dx, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}
// y[b,t,:] = (x ? x[b,t,:] : 0) + a * vec[b,:]
__global__ void __launch_bounds__(256) add_rowvec_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                         const float* __restrict__ vec, float a, int64_t total, int64_t TC,
                                                         int C) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / TC;
    const int c = (int)(i % C);
    y[i] = (x ? x[i] : 0.f) + a * vec[b * C + c];
  }
}
// unary ops: 0 tanh, 1 sigmoid, 2 relu, 3 elu(alpha=1)
__device__ __forceinline__ float unary_f(int op, float x) {
  switch (op) {
    case 0: return tanhf(x);
    case 1: return 1.f / (1.f + expf(-x));
    case 2: return fmaxf(x, 0.f);
    default: return x > 0.f ? x : expm1f(x);
  }
}
__device__ __forceinline__ float unary_g(int op, float x, float y) {
  switch (op) {
    case 0: return 1.f - y * y;
    case 1: return y * (1.f - y);
    case 2: return x > 0.f ? 1.f : 0.f;
    default: return x > 0.f ? 1.f : y + 1.f;
  }
}
__global__ void __launch_bounds__(256) unary_fwd_kernel(int op, const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = unary_f(op, x[i]);
}
__global__ void __launch_bounds__(256) unary_bwd_kernel(int op, const float* __restrict__ x, const float* __restrict__ y,
                                                        const float* __restrict__ dy, float* __restrict__ dx, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = dy[i] * unary_g(op, x[i], y[i]);
}
// out[c] = sum_r x[r,c] (+ beta*out).  grid.x covers columns in chunks of 32, 32x8 threads, rows strided by 8*gridDim.y
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ part, int64_t R,
                                                     int64_t C) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < C)
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < R; r += (int64_t)gridDim.y * 8) s += x[r * C + c];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][tx];
    part[(int64_t)blockIdx.y * C + c] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int nparts, int64_t C,
                                    float beta) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(int64_t)p * C + c];
  out[c] = s + (beta != 0.f ? beta * out[c] : 0.f);
}

}  // namespace sml

using namespace sml;

extern "C" int sml_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd,
                                 int64_t R, int32_t D, float eps, int64_t n, int64_t npad, int64_t pad, void* y_hi,
                                 void* y_lo, void* stream) {
  SML_CHECK(n > 0 && npad >= n + pad && R % n == 0, "sml_layernorm_fwd: bad row mapping n=%lld npad=%lld pad=%lld",
            (long long)n, (long long)npad, (long long)pad);
  SML_CHECK(D % 4 == 0 && D <= 1024, "sml_layernorm_fwd: D=%d must be a multiple of 4 and <= 1024", D);
  if (R == 0) return 0;
  const int wpb = 8;
  layernorm_fwd_kernel<<<(unsigned)((R + wpb - 1) / wpb), wpb * 32, 0, as_stream(stream)>>>(x, w, b, y, mean, rstd, R, D, eps, n, npad, pad, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo);
  SML_LAUNCH_OK("sml_layernorm_fwd");
  return 0;
}
extern "C" int sml_layernorm_bwd_blocks(int64_t R) {
  int64_t nb = (R + 63) / 64;
  if (nb > 4 * kNumSMs) nb = 4 * kNumSMs;
  if (nb < 1) nb = 1;
  return (int)nb;
}
extern "C" int sml_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd,
                                 float* dx, float* dw_part, float* db_part, int64_t R, int32_t D, int64_t n, int64_t npad,
                                 int64_t pad, const float* res, void* stream) {
  SML_CHECK(D % 4 == 0 && D <= 1024, "sml_layernorm_bwd: D=%d must be a multiple of 4 and <= 1024", D);
  const int nb = sml_layernorm_bwd_blocks(R);
  const int64_t rpb = (R + nb - 1) / nb;
  const size_t smem = (size_t)LNB_WARPS * 2 * D * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(layernorm_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_set = true;
  }
  const int nv = D / 4;
  cudaStream_t st = as_stream(stream);
  if (nv <= 32) layernorm_bwd_kernel<1><<<nb, LNB_WARPS * 32, smem, st>>>(dy, x, w, mean, rstd, dx, dw_part, db_part, R, D, rpb, n, npad, pad, res);
  else if (nv <= 64) layernorm_bwd_kernel<2><<<nb, LNB_WARPS * 32, smem, st>>>(dy, x, w, mean, rstd, dx, dw_part, db_part, R, D, rpb, n, npad, pad, res);
  else if (nv <= 128) layernorm_bwd_kernel<4><<<nb, LNB_WARPS * 32, smem, st>>>(dy, x, w, mean, rstd, dx, dw_part, db_part, R, D, rpb, n, npad, pad, res);
  else layernorm_bwd_kernel<8><<<nb, LNB_WARPS * 32, smem, st>>>(dy, x, w, mean, rstd, dx, dw_part, db_part, R, D, rpb, n, npad, pad, res);
  SML_LAUNCH_OK("sml_layernorm_bwd");
  return 0;
}
extern "C" int sml_softmax_fwd(float* x, int64_t R, int64_t C, void* hi_, void* lo_, int32_t write_f32, void* stream) {
  __nv_bfloat16* hi = (__nv_bfloat16*)hi_; __nv_bfloat16* lo = (__nv_bfloat16*)lo_;
  if (R == 0 || C == 0) return 0;
  cudaStream_t s = as_stream(stream);
  const unsigned gw = (unsigned)((R + 7) / 8);
  if (C <= 128) softmax_fwd_warp_kernel<4><<<gw, 256, 0, s>>>(x, R, (int)C, hi, lo, write_f32);
  else if (C <= 256) softmax_fwd_warp_kernel<8><<<gw, 256, 0, s>>>(x, R, (int)C, hi, lo, write_f32);
  else if (C <= 512) softmax_fwd_warp_kernel<16><<<gw, 256, 0, s>>>(x, R, (int)C, hi, lo, write_f32);
  else {
    SML_CHECK(R <= 2147483647LL, "sml_softmax_fwd: too many rows");
    const bool al = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (!hi || (reinterpret_cast<uintptr_t>(hi) & 7) == 0) &&
                    (!lo || (reinterpret_cast<uintptr_t>(lo) & 7) == 0);
    const int C4 = (int)(C / 4);
    if (al && C4 <= 1024) softmax_fwd_reg_kernel<1><<<(unsigned)R, 1024, 0, s>>>(x, C4, hi, lo, write_f32);
    else if (al && C4 <= 2048) softmax_fwd_reg_kernel<2><<<(unsigned)R, 1024, 0, s>>>(x, C4, hi, lo, write_f32);
    else if (al && C4 <= 4096) softmax_fwd_reg_kernel<4><<<(unsigned)R, 1024, 0, s>>>(x, C4, hi, lo, write_f32);
    else if (al && C4 <= 5120) softmax_fwd_reg_kernel<5><<<(unsigned)R, 1024, 0, s>>>(x, C4, hi, lo, write_f32);
    else if (al && C4 <= 8192) softmax_fwd_reg_kernel<8><<<(unsigned)R, 1024, 0, s>>>(x, C4, hi, lo, write_f32);
    else softmax_fwd_block_kernel<<<(unsigned)R, C >= 8192 ? 1024 : 256, 0, s>>>(x, C, hi, lo, write_f32);
  }
  SML_LAUNCH_OK("sml_softmax_fwd");
  return 0;
}
extern "C" int sml_softmax_bwd(const float* A, const void* A_bf16, float* dA, int64_t R, int64_t C, void* hi_, void* lo_,
                               int32_t write_f32, void* stream) {
  __nv_bfloat16* hi = (__nv_bfloat16*)hi_; __nv_bfloat16* lo = (__nv_bfloat16*)lo_;
  const __nv_bfloat16* Ab = (const __nv_bfloat16*)A_bf16;
  SML_CHECK(A || Ab, "sml_softmax_bwd: need the probabilities in fp32 or bf16");
  if (R == 0 || C == 0) return 0;
  cudaStream_t s = as_stream(stream);
  const unsigned gw = (unsigned)((R + 7) / 8);
  if (C <= 128) softmax_bwd_warp_kernel<4><<<gw, 256, 0, s>>>(A, Ab, dA, R, (int)C, hi, lo, write_f32);
  else if (C <= 256) softmax_bwd_warp_kernel<8><<<gw, 256, 0, s>>>(A, Ab, dA, R, (int)C, hi, lo, write_f32);
  else if (C <= 512) softmax_bwd_warp_kernel<16><<<gw, 256, 0, s>>>(A, Ab, dA, R, (int)C, hi, lo, write_f32);
  else {
    const bool al = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(dA) & 15) == 0) && (!A || (reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                    (!Ab || (reinterpret_cast<uintptr_t>(Ab) & 7) == 0) && (!hi || (reinterpret_cast<uintptr_t>(hi) & 7) == 0) &&
                    (!lo || (reinterpret_cast<uintptr_t>(lo) & 7) == 0);
    const int C4 = (int)(C / 4);
    if (al && C4 <= 1024) softmax_bwd_reg_kernel<1><<<(unsigned)R, 1024, 0, s>>>(A, Ab, dA, C4, hi, lo, write_f32);
    else if (al && C4 <= 2048) softmax_bwd_reg_kernel<2><<<(unsigned)R, 1024, 0, s>>>(A, Ab, dA, C4, hi, lo, write_f32);
    else if (al && C4 <= 4096) softmax_bwd_reg_kernel<4><<<(unsigned)R, 1024, 0, s>>>(A, Ab, dA, C4, hi, lo, write_f32);
    else if (al && C4 <= 5120) softmax_bwd_reg_kernel<5><<<(unsigned)R, 1024, 0, s>>>(A, Ab, dA, C4, hi, lo, write_f32);
    else if (al && C4 <= 8192) softmax_bwd_reg_kernel<8><<<(unsigned)R, 1024, 0, s>>>(A, Ab, dA, C4, hi, lo, write_f32);
    else softmax_bwd_block_kernel<<<(unsigned)R, C >= 8192 ? 1024 : 256, 0, s>>>(A, Ab, dA, C, hi, lo, write_f32);
  }
  SML_LAUNCH_OK("sml_softmax_bwd");
  return 0;
}
static inline unsigned ew_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
extern "C" int sml_dropout_add(const float* x, const float* u, float* y, int64_t n, float p, uint64_t seed,
                               uint64_t offset, void* stream) {
  SML_CHECK(p >= 0.f && p < 1.f, "sml_dropout_add: p=%f out of range", p);
  if (n == 0) return 0;
  dropout_add_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, u, y, n, p, seed, offset, seed_salt());
  SML_LAUNCH_OK("sml_dropout_add");
  return 0;
}
extern "C" int sml_alpha_dropout(const float* x, float* y, int64_t n, float p, int32_t bwd, uint64_t seed, uint64_t offset,
                                 void* stream) {
  SML_CHECK(p > 0.f && p < 1.f, "sml_alpha_dropout: p=%f out of range", p);
  if (n == 0) return 0;
  const double ap = -1.0507009873554804934193349852946 * 1.6732632423543772848170429916717;
  const double a = 1.0 / sqrt((1.0 - p) * (1.0 + p * ap * ap)), b = -a * ap * p;
  alpha_dropout_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, y, n, p, (float)a, (float)(a * ap + b), (float)b, bwd, seed, offset, seed_salt());
  SML_LAUNCH_OK("sml_alpha_dropout");
  return 0;
}
extern "C" int sml_axpby(float* y, const float* x, float a, float b, int64_t n, void* hi, void* lo, void* stream) {
  if (n == 0) return 0;
  axpby_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(y, x, a, b, n, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  SML_LAUNCH_OK("sml_axpby");
  return 0;
}
extern "C" int sml_colsum_parts(int64_t R) {
  int64_t p = (R + 255) / 256;
  return (int)(p < 1 ? 1 : (p > 64 ? 64 : p));
}
extern "C" int sml_colsum(const float* x, float* out, float* part, int64_t R, int64_t C, float beta, void* stream) {
  const int np = sml_colsum_parts(R);
  dim3 grid((unsigned)((C + 31) / 32), np);
  colsum_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, part, R, C);
  SML_LAUNCH_OK("sml_colsum");
  colsum_final_kernel<<<(unsigned)((C + 255) / 256), 256, 0, as_stream(stream)>>>(part, out, np, C, beta);
  SML_LAUNCH_OK("sml_colsum_final");
  return 0;
}

extern "C" int sml_diag_axpby(float* out, const float* x, float a, float dg, int64_t nmat, int32_t m, void* hi, void* lo,
                              void* stream) {
  const int64_t total = nmat * m * m;
  if (total == 0) return 0;
  diag_axpby_kernel<<<ew_grid(total), 256, 0, as_stream(stream)>>>(out, x, a, dg, total, m, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  SML_LAUNCH_OK("sml_diag_axpby");
  return 0;
}
extern "C" int sml_relu_bwd(const float* y, const float* dy, float* dx, int64_t n, void* stream) {
  if (n == 0) return 0;
  relu_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(y, dy, dx, n);
  SML_LAUNCH_OK("sml_relu_bwd");
  return 0;
}

extern "C" int sml_add_rowvec(float* y, const float* x, const float* vec, float a, int64_t B, int64_t T, int32_t C, void* stream) {
  const int64_t total = B * T * C;
  if (total == 0) return 0;
  add_rowvec_kernel<<<ew_grid(total), 256, 0, as_stream(stream)>>>(y, x, vec, a, total, T * C, C);
  SML_LAUNCH_OK("sml_add_rowvec");
  return 0;
}
extern "C" int sml_unary_fwd(int32_t op, const float* x, float* y, int64_t n, void* stream) {
  SML_CHECK(op >= 0 && op <= 3, "sml_unary_fwd: unknown op %d", op);
  if (n == 0) return 0;
  unary_fwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(op, x, y, n);
  SML_LAUNCH_OK("sml_unary_fwd");
  return 0;
}
extern "C" int sml_unary_bwd(int32_t op, const float* x, const float* y, const float* dy, float* dx, int64_t n, void* stream) {
  SML_CHECK(op >= 0 && op <= 3, "sml_unary_bwd: unknown op %d", op);
  if (n == 0) return 0;
  unary_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(op, x, y, dy, dx, n);
  SML_LAUNCH_OK("sml_unary_bwd");
  return 0;
}

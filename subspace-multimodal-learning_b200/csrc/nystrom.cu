// nystrom.cu — the non-GEMM pieces of Nystrom attention (models/NystromAttention.py:20-35, :74-157)
// and the PPEG stencil of TransMIL (models/mil.py:192-206).  All HBM-bound; layouts are token-major so
// that the channel axis (h*d or C) is the coalesced one.
#include "common.cuh"

namespace sml {

// ------------------------------------------------------------------ landmarks
// grid (m, B); block = h*d threads (<= 1024).  qkv [B, npad, 3, h, d]
// source-type helpers: the qkv tensor is read as fp32, or as its bf16 copy in bf16 mode (no fp32 plane is written then)
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}

template <typename TS>
__global__ void landmarks_fwd_kernel(const TS* __restrict__ qkv, float* __restrict__ ql, float* __restrict__ kl,
                                     int npad, int h, int d, int m) {
  const int j = blockIdx.x, b = blockIdx.y, c = threadIdx.x;   // c = head*d + dd
  const int hd = h * d, l = npad / m;
  const TS* base = qkv + ((int64_t)b * npad + (int64_t)j * l) * 3 * hd;
  // 5 independent load streams per tensor (l = 65 at 16 k patches): the segment sum is latency-bound otherwise
  float aq[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, ak[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  int t = 0;
  for (; t + 5 <= l; t += 5) {
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      aq[u] += ld1(base + (int64_t)(t + u) * 3 * hd + c);
      ak[u] += ld1(base + (int64_t)(t + u) * 3 * hd + hd + c);
    }
  }
  for (; t < l; ++t) {
    aq[0] += ld1(base + (int64_t)t * 3 * hd + c);
    ak[0] += ld1(base + (int64_t)t * 3 * hd + hd + c);
  }
  const float sq = ((aq[0] + aq[1]) + (aq[2] + aq[3])) + aq[4], sk = ((ak[0] + ak[1]) + (ak[2] + ak[3])) + ak[4];
  const int head = c / d, dd = c % d;
  const int64_t o = (((int64_t)b * h + head) * m + j) * d + dd;
  ql[o] = sq / (float)l;
  kl[o] = sk / (float)l;
}
__global__ void landmarks_bwd_kernel(float* __restrict__ dqkv, const float* __restrict__ dql,
                                     const float* __restrict__ dkl, int npad, int h, int d, int m,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int j = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
  const int hd = h * d, l = npad / m;
  const int head = c / d, dd = c % d;
  const int64_t o = (((int64_t)b * h + head) * m + j) * d + dd;
  const float gq = dql[o] / (float)l, gk = dkl[o] / (float)l;
  float* base = dqkv + ((int64_t)b * npad + (int64_t)j * l) * 3 * hd;
  const int64_t boff = ((int64_t)b * npad + (int64_t)j * l) * 3 * hd;
  for (int t = 0; t < l; ++t) {
    const float vq = base[(int64_t)t * 3 * hd + c] + gq, vk = base[(int64_t)t * 3 * hd + hd + c] + gk;
    base[(int64_t)t * 3 * hd + c] = vq;
    base[(int64_t)t * 3 * hd + hd + c] = vk;
    store_split(hi, lo, boff + (int64_t)t * 3 * hd + c, vq);
    store_split(hi, lo, boff + (int64_t)t * 3 * hd + hd + c, vk);
  }
}
__global__ void landmark_ids_kernel(int32_t* ids, int npad, int l) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < npad) ids[t] = t / l;
}

// ------------------------------------------------------------------ pinv scale (global max of row / column sums)
// grid BH blocks, 256 threads.  packed[0] / packed[1]: (float bits << 32) | (0xFFFFFFFF - flat index)
__global__ void __launch_bounds__(256) pinv_scale_kernel(const float* __restrict__ x, unsigned long long* packed, int m) {
  __shared__ unsigned long long best[2];
  const int bh = blockIdx.x;
  const float* X = x + (int64_t)bh * m * m;
  if (threadIdx.x < 2) best[threadIdx.x] = 0ull;
  __syncthreads();
  // column sums: thread per column, coalesced across threads
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < m; ++i) s += fabsf(X[(int64_t)i * m + j]);
    const unsigned long long p = ((unsigned long long)__float_as_uint(s) << 32) | (0xFFFFFFFFu - (unsigned)(bh * m + j));
    atomicMax(&best[1], p);
  }
  // row sums: warp per row
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = wid; i < m; i += nw) {
    float s = 0.f;
    for (int j = lane; j < m; j += 32) s += fabsf(X[(int64_t)i * m + j]);
    s = warp_sum(s);
    if (lane == 0) {
      const unsigned long long p = ((unsigned long long)__float_as_uint(s) << 32) | (0xFFFFFFFFu - (unsigned)(bh * m + i));
      atomicMax(&best[0], p);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) atomicMax(&packed[threadIdx.x], best[threadIdx.x]);
}
__global__ void pinv_scale_final_kernel(const unsigned long long* packed, float* scal, int32_t* arg) {
  const int t = threadIdx.x;
  if (t < 2) {
    scal[t] = __uint_as_float((unsigned)(packed[t] >> 32));
    arg[t] = (int32_t)(0xFFFFFFFFu - (unsigned)(packed[t] & 0xFFFFFFFFull));
  }
}
// z0 = x^T / (scal0*scal1): 32x32 smem-tiled transpose. grid (m/32, m/32, BH), block (32,8)
__global__ void pinv_init_kernel(const float* __restrict__ x, const float* __restrict__ scal, float* __restrict__ z0, int m) {
  __shared__ float tile[32][33];
  const float inv = 1.f / (scal[0] * scal[1]);
  const int64_t off = (int64_t)blockIdx.z * m * m;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = by + r, j = bx + threadIdx.x;
    tile[r][threadIdx.x] = (i < m && j < m) ? x[off + (int64_t)i * m + j] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = bx + r, j = by + threadIdx.x;     // z0[i,j] = x[j,i]
    if (i < m && j < m) z0[off + (int64_t)i * m + j] = tile[threadIdx.x][r] * inv;
  }
}
// dx[j,i] += dz0[i,j]/s ; acc += sum dz0[i,j]*x[j,i]
__global__ void pinv_init_bwd_kernel(const float* __restrict__ x, const float* __restrict__ scal,
                                     const float* __restrict__ dz0, float* __restrict__ dx, float* __restrict__ acc, int m) {
  __shared__ float tile[32][33];
  __shared__ float red[32];
  const float inv = 1.f / (scal[0] * scal[1]);
  const int64_t off = (int64_t)blockIdx.z * m * m;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = by + r, j = bx + threadIdx.x;
    tile[r][threadIdx.x] = (i < m && j < m) ? dz0[off + (int64_t)i * m + j] : 0.f;
  }
  __syncthreads();
  float part = 0.f;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int jj = bx + r, ii = by + threadIdx.x;   // x[jj, ii] pairs with dz0[ii, jj]
    if (ii < m && jj < m) {
      const float g = tile[threadIdx.x][r];
      const int64_t o = off + (int64_t)jj * m + ii;
      part += g * x[o];
      dx[o] += g * inv;
    }
  }
  // block reduce (256 threads laid out 32x8)
  const int tid = threadIdx.y * 32 + threadIdx.x;
  part = warp_sum(part);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(acc, s);
  }
}
// scalar path: s = c*r, ds = -acc / s^2 ; dc = ds*r -> row arg[0] ; dr = ds*c -> column arg[1]
__global__ void pinv_init_bwd_scalar_kernel(const float* __restrict__ x, const float* __restrict__ scal,
                                            const int32_t* __restrict__ arg, const float* __restrict__ acc,
                                            float* __restrict__ dx, int m) {
  const float c = scal[0], r = scal[1];
  const float s = c * r;
  const float ds = -acc[0] / (s * s);
  const float dc = ds * r, dr = ds * c;
  const int64_t rowbase = (int64_t)arg[0] * m;                          // flat (bh*m + i) * m
  const int64_t colbh = (int64_t)(arg[1] / m) * m * m, colj = arg[1] % m;
  for (int t = threadIdx.x; t < m; t += blockDim.x) {
    const float xv = x[rowbase + t];
    atomicAdd(&dx[rowbase + t], dc * (xv > 0.f ? 1.f : (xv < 0.f ? -1.f : 0.f)));
    const int64_t o = colbh + (int64_t)t * m + colj;
    const float xc = x[o];
    atomicAdd(&dx[o], dr * (xc > 0.f ? 1.f : (xc < 0.f ? -1.f : 0.f)));
  }
}

// ------------------------------------------------------------------ residual depthwise conv along tokens
// src rows have stride src_ld, channel c; dst rows stride dst_ld.  dst[t,c] += sum_k w[c/d][k'] src[t + k - half, c]
// with k' = k (fwd) or K-1-k (FLIP: the transposed conv of the backward).  Each thread owns TT consecutive tokens of
// one channel: the K taps live in registers and every (tap, output) pair is resolved at compile time.
constexpr int RC_TT = 8;
// SLD / DLD = the row pitches at compile time (0 = run-time)
template <bool FLIP, int K, typename TS, int SLD = 0, int DLD = 0>
__global__ void __launch_bounds__(128) resconv_kernel(const TS* __restrict__ src, int64_t src_ld_rt,
                                                      const float* __restrict__ w, float* __restrict__ dst,
                                                      int64_t dst_ld_rt, int npad, int hd, int d,
                                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  // thread = 4 adjacent channels (one 128-bit load per token) x RC_TT tokens; requires d % 4 == 0 and 16-byte rows
  constexpr int HALF = K / 2;
  const int64_t src_ld = SLD ? SLD : src_ld_rt, dst_ld = DLD ? DLD : dst_ld_rt;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * RC_TT;
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c >= hd) return;
  float wr[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wr[k] = __ldg(w + (c / d) * K + (FLIP ? (K - 1 - k) : k));
  const TS* s = src + (int64_t)b * npad * src_ld + c;
  float4 acc[RC_TT];
#pragma unroll
  for (int i = 0; i < RC_TT; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int u = -HALF; u < RC_TT + HALF; ++u) {       // source token t0+u feeds output i through tap k = u - i + HALF
    const int ts = t0 + u;
    const float4 v = (ts >= 0 && ts < npad) ? ld4(s + (int64_t)ts * src_ld) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < RC_TT; ++i) {
      const int k = u - i + HALF;
      if (k >= 0 && k < K) {
        acc[i].x = fmaf(wr[k], v.x, acc[i].x); acc[i].y = fmaf(wr[k], v.y, acc[i].y);
        acc[i].z = fmaf(wr[k], v.z, acc[i].z); acc[i].w = fmaf(wr[k], v.w, acc[i].w);
      }
    }
  }
  const int64_t ooff = (int64_t)b * npad * dst_ld + c;
#pragma unroll
  for (int i = 0; i < RC_TT; ++i)
    if (t0 + i < npad) {
      float4* o = reinterpret_cast<float4*>(dst + ooff + (int64_t)(t0 + i) * dst_ld);
      float4 v = *o;
      v.x += acc[i].x; v.y += acc[i].y; v.z += acc[i].z; v.w += acc[i].w;
      *o = v;
      if (hi) {
        const int64_t e = ooff + (int64_t)(t0 + i) * dst_ld;
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
}
// dw_part[blk, head, k] = sum over the block's token strip and the head's d channels of dout[t,c] * v[t+k-half, c].
// Thread = one channel; it walks its strip in chunks of 16 tokens with a (16 + K - 1)-wide register window of v.
constexpr int RCB_TOK = 256;
template <int K, typename TS>
__global__ void __launch_bounds__(128) resconv_dw_kernel(const TS* __restrict__ v, int64_t v_ld,
                                                         const float* __restrict__ dout, int64_t o_ld,
                                                         float* __restrict__ dw_part, int npad, int hd, int d) {
  // thread = 4 adjacent channels of one head (128-bit loads); the 4 channel products fold into one scalar per tap
  constexpr int HALF = K / 2, CH = 8;
  extern __shared__ float sdw[];   // [h*K]
  const int h = hd / d;
  for (int i = threadIdx.x; i < h * K; i += blockDim.x) sdw[i] = 0.f;
  __syncthreads();
  const int b = blockIdx.z, t0 = blockIdx.x * RCB_TOK;
  const int t1 = min(npad, t0 + RCB_TOK);
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c < hd) {
    const TS* vp = v + (int64_t)b * npad * v_ld + c;
    const float* gp = dout + (int64_t)b * npad * o_ld + c;
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    for (int tc = t0; tc < t1; tc += CH) {
      float4 g[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i)
        g[i] = (tc + i < t1) ? *reinterpret_cast<const float4*>(gp + (int64_t)(tc + i) * o_ld) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = -HALF; u < CH + HALF; ++u) {          // v token tc+u pairs with output i through tap k = u - i + HALF
        const int ts = tc + u;
        const float4 x = (ts >= 0 && ts < npad) ? ld4(vp + (int64_t)ts * v_ld) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int k = u - i + HALF;
          if (k >= 0 && k < K) acc[k] = fmaf(g[i].x, x.x, fmaf(g[i].y, x.y, fmaf(g[i].z, x.z, fmaf(g[i].w, x.w, acc[k]))));
        }
      }
    }
    // reduce over the d/4 threads of a head: whole warps when d/4 >= 32, else per-thread shared atomics
    const bool whole_warp_one_head = ((d / 4) & 31) == 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (whole_warp_one_head) {
        const float sum = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sdw[(c / d) * K + k], sum);
      } else {
        atomicAdd(&sdw[(c / d) * K + k], acc[k]);
      }
    }
  }
  __syncthreads();
  float* o = dw_part + (((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * h * K;
  for (int i = threadIdx.x; i < h * K; i += blockDim.x) o[i] = sdw[i];
}
// generic-K fallbacks (any odd kernel size; the reference path uses 33)
template <bool FLIP>
__global__ void __launch_bounds__(256) resconv_generic_kernel(const float* __restrict__ src, int64_t src_ld,
                                                              const float* __restrict__ w, float* __restrict__ dst,
                                                              int64_t dst_ld, int npad, int hd, int d, int K) {
  const int b = blockIdx.z, t = blockIdx.x, half = K / 2;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= hd) return;
  const float* s = src + (int64_t)b * npad * src_ld + c;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const int ts = t + k - half;
    if (ts >= 0 && ts < npad) acc = fmaf(__ldg(w + (c / d) * K + (FLIP ? K - 1 - k : k)), s[(int64_t)ts * src_ld], acc);
  }
  dst[((int64_t)b * npad + t) * dst_ld + c] += acc;
}
__global__ void __launch_bounds__(256) resconv_dw_generic_kernel(const float* __restrict__ v, int64_t v_ld,
                                                                 const float* __restrict__ dout, int64_t o_ld,
                                                                 float* __restrict__ dw_part, int npad, int hd, int d, int K) {
  extern __shared__ float sdw[];
  const int h = hd / d, half = K / 2;
  for (int i = threadIdx.x; i < h * K; i += blockDim.x) sdw[i] = 0.f;
  __syncthreads();
  const int b = blockIdx.z, t0 = blockIdx.x * RCB_TOK, t1 = min(npad, t0 + RCB_TOK);
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c < hd) {
    const float* vp = v + (int64_t)b * npad * v_ld + c;
    const float* gp = dout + (int64_t)b * npad * o_ld + c;
    for (int k = 0; k < K; ++k) {
      float s = 0.f;
      for (int t = t0; t < t1; ++t) {
        const int ts = t + k - half;
        if (ts >= 0 && ts < npad) s = fmaf(gp[(int64_t)t * o_ld], vp[(int64_t)ts * v_ld], s);
      }
      atomicAdd(&sdw[(c / d) * K + k], s);
    }
  }
  __syncthreads();
  float* o = dw_part + (((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * h * K;
  for (int i = threadIdx.x; i < h * K; i += blockDim.x) o[i] = sdw[i];
}

// ------------------------------------------------------------------ PPEG 7x7 folded stencil, token-major
// x,y [B, 1+H*W, C]; w7 [49, C] (tap-major); bias [C].  Block = 16x8 spatial tile x 32 channels staged in shared
// memory with a 3-pixel halo; thread = (channel, tile row) sliding along the row with its 49 taps in registers.
constexpr int PT_W = 16, PT_H = 8, PT_C = 32;
template <bool FLIP>
__global__ void __launch_bounds__(256) ppeg_kernel(const float* __restrict__ x, const float* __restrict__ w7,
                                                   const float* __restrict__ bias, float* __restrict__ y, int H, int W,
                                                   int C, int tiles_x) {
  __shared__ float tile[PT_H + 6][PT_W + 6][PT_C + 1];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * PT_C;
  const int ty0 = (blockIdx.x / tiles_x) * PT_H, tx0 = (blockIdx.x % tiles_x) * PT_W;
  const int64_t base = (int64_t)b * (1 + H * W) * C;
  const int lc = threadIdx.x & 31, lr = threadIdx.x >> 5;     // channel lane, 8 row groups
  const int c = c0 + lc;
  if (blockIdx.x == 0 && lr == 0 && c < C) y[base + c] = x[base + c];      // cls row passes through (mil.py:201-205)
  for (int ry = lr; ry < PT_H + 6; ry += 8) {                 // one warp per halo row: row checks and base address hoisted
    const int gy = ty0 + ry - 3;
    const bool rok = gy >= 0 && gy < H && c < C;
    const float* rp = x + base + (int64_t)(1 + gy * W + tx0 - 3) * C + c;
#pragma unroll
    for (int rx = 0; rx < PT_W + 6; ++rx) {
      const int gx = tx0 + rx - 3;
      tile[ry][rx][lc] = (rok && gx >= 0 && gx < W) ? rp[(int64_t)rx * C] : 0.f;
    }
  }
  __syncthreads();
  if (c >= C) return;
  float wr[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) wr[t] = __ldg(w7 + (FLIP ? 48 - t : t) * C + c);
  const float bv = bias ? __ldg(bias + c) : 0.f;
  const int oy = ty0 + lr;
  if (oy >= H) return;
  float acc[PT_W];
#pragma unroll
  for (int ox = 0; ox < PT_W; ++ox) acc[ox] = bv;
#pragma unroll
  for (int di = 0; di < 7; ++di) {
    float rowv[PT_W + 6];                                   // one halo row of the tile, reused by 7 taps x 16 outputs
#pragma unroll
    for (int u = 0; u < PT_W + 6; ++u) rowv[u] = tile[lr + di][u][lc];
#pragma unroll
    for (int ox = 0; ox < PT_W; ++ox)
#pragma unroll
      for (int dj = 0; dj < 7; ++dj) acc[ox] = fmaf(wr[di * 7 + dj], rowv[ox + dj], acc[ox]);
  }
#pragma unroll
  for (int ox = 0; ox < PT_W; ++ox)
    if (tx0 + ox < W) y[base + (int64_t)(1 + oy * W + tx0 + ox) * C + c] = acc[ox];
}
// dw_part[grp, tap, c], db_part[grp, c]: grid (C/32, groups, B); a block walks its share of spatial tiles
constexpr int PPB_GROUPS = 32;
__global__ void __launch_bounds__(256) ppeg_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                      float* __restrict__ dw_part, float* __restrict__ db_part, int H,
                                                      int W, int C, int tiles_x, int ntiles) {
  __shared__ float tile[PT_H + 6][PT_W + 6][PT_C + 1];
  __shared__ float red[50][PT_C + 1];
  const int b = blockIdx.z, c0 = blockIdx.x * PT_C;
  const int lc = threadIdx.x & 31, lr = threadIdx.x >> 5;
  const int c = c0 + lc;
  const int64_t base = (int64_t)b * (1 + H * W) * C;
  float acc[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) acc[t] = 0.f;
  float ab = 0.f;
  for (int tl = blockIdx.y; tl < ntiles; tl += gridDim.y) {
    const int ty0 = (tl / tiles_x) * PT_H, tx0 = (tl % tiles_x) * PT_W;
    __syncthreads();
    for (int ry = lr; ry < PT_H + 6; ry += 8) {
      const int gy = ty0 + ry - 3;
      const bool rok = gy >= 0 && gy < H && c < C;
      const float* rp = x + base + (int64_t)(1 + gy * W + tx0 - 3) * C + c;
#pragma unroll
      for (int rx = 0; rx < PT_W + 6; ++rx) {
        const int gx = tx0 + rx - 3;
        tile[ry][rx][lc] = (rok && gx >= 0 && gx < W) ? rp[(int64_t)rx * C] : 0.f;
      }
    }
    __syncthreads();
    const int oy = ty0 + lr;
    if (c < C && oy < H) {
      float g[PT_W];
#pragma unroll
      for (int ox = 0; ox < PT_W; ++ox) {
        g[ox] = (tx0 + ox < W) ? dy[base + (int64_t)(1 + oy * W + tx0 + ox) * C + c] : 0.f;
        ab += g[ox];
      }
#pragma unroll
      for (int di = 0; di < 7; ++di) {
        float rowv[PT_W + 6];
#pragma unroll
        for (int u = 0; u < PT_W + 6; ++u) rowv[u] = tile[lr + di][u][lc];
#pragma unroll
        for (int dj = 0; dj < 7; ++dj)
#pragma unroll
          for (int ox = 0; ox < PT_W; ++ox) acc[di * 7 + dj] = fmaf(g[ox], rowv[ox + dj], acc[di * 7 + dj]);
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 50 * (PT_C + 1); e += blockDim.x) (&red[0][0])[e] = 0.f;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < 49; ++t) atomicAdd(&red[t][lc], acc[t]);
  atomicAdd(&red[49][lc], ab);
  __syncthreads();
  const int64_t grp = (int64_t)blockIdx.z * gridDim.y + blockIdx.y;
  for (int e = threadIdx.x; e < 50 * PT_C; e += blockDim.x) {
    const int t = e / PT_C, l = e % PT_C;
    if (c0 + l >= C) continue;
    const float s = red[t][l];
    if (t < 49) dw_part[(grp * 49 + t) * C + c0 + l] = s;
    else db_part[grp * C + c0 + l] = s;
  }
}


// ---- register-tiled PPEG stencil (round 2).  The shared-memory version above reads 0.78 B of shared memory per FMA (one
// 22-float halo row per 112 FMAs) and is co-limited by the LSU pipe and its load -> barrier -> compute phases (350 us forward
// at B = 8 / 128 x 128 x 512, 35 % of the issue peak).  Here a warp is 32 consecutive channels of ONE pixel (so every load is a
// 128-byte line and every bounds predicate is warp-uniform) and a thread computes an 8 x 4 output patch: one 14-float input
// row feeds 4 x 8 x 7 = 224 FMAs (0.25 B per FMA, from L1), no shared memory, no barrier.  Block = 8 warps = 16 x 16 pixels.
// Measured: the transposed stencil (dx) 549 -> 390 us, the weight gradient 568 -> 526 us, the forward 350 -> 367 us (so the
// forward stays on the kernel above); ncu: issue 57 %, FMA pipe 38 %, 2.5 instructions per FMA - not memory-bound.
constexpr int PR_W = 8, PR_H = 4, PRB_W = 16, PRB_H = 16;
// CT = channel count at compile time (512 on the reference path; 0 = run-time): with a constant row pitch every load of a
// 14-pixel input row is base + immediate, which removes the 64-bit address arithmetic that made up a third of the instructions
template <bool FLIP, int CT>
__global__ void __launch_bounds__(256, 2) ppeg_rt_kernel(const float* __restrict__ x, const float* __restrict__ w7,
                                                         const float* __restrict__ bias, float* __restrict__ y, int H, int W,
                                                         int Crt, int tiles_x) {
  const int C = CT ? CT : Crt;
  const int b = blockIdx.z;
  const int lc = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + lc;
  const int ox0 = (blockIdx.x % tiles_x) * PRB_W + (wp & 1) * PR_W;
  const int oy0 = (blockIdx.x / tiles_x) * PRB_H + (wp >> 1) * PR_H;
  const int64_t base = (int64_t)b * (1 + H * W) * C;
  if (c >= C) return;
  if (blockIdx.x == 0 && wp == 0) y[base + c] = x[base + c];      // cls row passes through (mil.py:201-205)
  if (ox0 >= W || oy0 >= H) return;
  float wr[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) wr[t] = __ldg(w7 + (FLIP ? 48 - t : t) * C + c);
  const float bv = bias ? __ldg(bias + c) : 0.f;
  float acc[PR_H][PR_W];
#pragma unroll
  for (int r = 0; r < PR_H; ++r)
#pragma unroll
    for (int u = 0; u < PR_W; ++u) acc[r][u] = bv;
  const bool interior = ox0 >= 3 && ox0 + PR_W + 3 <= W;          // warp-uniform: no column predicate needed
  const float* xb = x + base + (int64_t)C + c + (int64_t)(ox0 - 3) * C;      // input column ox0 - 3 of lattice row 0
#pragma unroll
  for (int iy = 0; iy < PR_H + 6; ++iy) {
    const int gy = oy0 - 3 + iy;
    if (gy < 0 || gy >= H) continue;                              // warp-uniform
    const float* rp = xb + (int64_t)gy * W * C;
    float v[PR_W + 6];
    if (interior) {
#pragma unroll
      for (int u = 0; u < PR_W + 6; ++u) v[u] = __ldg(rp + u * C);
    } else {
#pragma unroll
      for (int u = 0; u < PR_W + 6; ++u) {
        const int gx = ox0 - 3 + u;
        v[u] = (gx >= 0 && gx < W) ? __ldg(rp + u * C) : 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < PR_H; ++r) {
      const int di = iy - r;                                      // compile-time after unrolling
      if (di < 0 || di > 6) continue;
#pragma unroll
      for (int u = 0; u < PR_W; ++u)
#pragma unroll
        for (int dj = 0; dj < 7; ++dj) acc[r][u] = fmaf(wr[di * 7 + dj], v[u + dj], acc[r][u]);
    }
  }
  float* yb = y + base + (int64_t)C + c + ((int64_t)oy0 * W + ox0) * C;
#pragma unroll
  for (int r = 0; r < PR_H; ++r) {
    if (oy0 + r >= H) break;
#pragma unroll
    for (int u = 0; u < PR_W; ++u)
      if (ox0 + u < W) yb[(int64_t)r * W * C + u * C] = acc[r][u];
  }
}
// weight / bias gradient, same tiling: grid (C/32, groups, B); a block walks its share of the 16 x 16 tiles, every thread keeps
// the 49 tap sums of its channel; the 8 warps are reduced through shared memory at the end
template <int CT>
__global__ void __launch_bounds__(256, 2) ppeg_dw_rt_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            float* __restrict__ dw_part, float* __restrict__ db_part, int H,
                                                            int W, int Crt, int tiles_x, int ntiles) {
  const int C = CT ? CT : Crt;
  __shared__ float red[50][33];
  const int b = blockIdx.z, c0 = blockIdx.x * 32;
  const int lc = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c = c0 + lc;
  const int64_t base = (int64_t)b * (1 + H * W) * C;
  float acc[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) acc[t] = 0.f;
  float ab = 0.f;
  if (c < C) {
    const float* xb = x + base + (int64_t)C + c;
    const float* gb = dy + base + (int64_t)C + c;
    for (int tl = blockIdx.y; tl < ntiles; tl += gridDim.y) {
      const int ox0 = (tl % tiles_x) * PRB_W + (wp & 1) * PR_W;
      const int oy0 = (tl / tiles_x) * PRB_H + (wp >> 1) * PR_H;
      if (ox0 >= W || oy0 >= H) continue;
      float g[PR_H][PR_W];
      const float* gp = gb + ((int64_t)oy0 * W + ox0) * C;
#pragma unroll
      for (int r = 0; r < PR_H; ++r)
#pragma unroll
        for (int u = 0; u < PR_W; ++u) {
          g[r][u] = (oy0 + r < H && ox0 + u < W) ? __ldg(gp + (int64_t)r * W * C + u * C) : 0.f;
          ab += g[r][u];
        }
      const bool interior = ox0 >= 3 && ox0 + PR_W + 3 <= W;
#pragma unroll
      for (int iy = 0; iy < PR_H + 6; ++iy) {
        const int gy = oy0 - 3 + iy;
        if (gy < 0 || gy >= H) continue;
        const float* rp = xb + ((int64_t)gy * W + ox0 - 3) * C;
        float v[PR_W + 6];
        if (interior) {
#pragma unroll
          for (int u = 0; u < PR_W + 6; ++u) v[u] = __ldg(rp + u * C);
        } else {
#pragma unroll
          for (int u = 0; u < PR_W + 6; ++u) {
            const int gx = ox0 - 3 + u;
            v[u] = (gx >= 0 && gx < W) ? __ldg(rp + u * C) : 0.f;
          }
        }
#pragma unroll
        for (int r = 0; r < PR_H; ++r) {
          const int di = iy - r;
          if (di < 0 || di > 6) continue;
#pragma unroll
          for (int dj = 0; dj < 7; ++dj)
#pragma unroll
            for (int u = 0; u < PR_W; ++u) acc[di * 7 + dj] = fmaf(g[r][u], v[u + dj], acc[di * 7 + dj]);
        }
      }
    }
  }
  for (int e = threadIdx.x; e < 50 * 33; e += blockDim.x) (&red[0][0])[e] = 0.f;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < 49; ++t) atomicAdd(&red[t][lc], acc[t]);
  atomicAdd(&red[49][lc], ab);
  __syncthreads();
  const int64_t grp = (int64_t)blockIdx.z * gridDim.y + blockIdx.y;
  for (int e = threadIdx.x; e < 50 * 32; e += blockDim.x) {
    const int t = e / 32, l = e % 32;
    if (c0 + l >= C) continue;
    const float sacc = red[t][l];
    if (t < 49) dw_part[(grp * 49 + t) * C + c0 + l] = sacc;
    else db_part[grp * C + c0 + l] = sacc;
  }
}

}  // namespace sml

using namespace sml;

extern "C" int sml_nystrom_landmarks_fwd(const float* qkv, float* ql, float* kl, int32_t B, int32_t npad, int32_t h,
                                         int32_t d, int32_t m, const void* qkv_bf16, void* stream) {
  SML_CHECK(h * d <= 1024 && npad % m == 0, "sml_nystrom_landmarks_fwd: need h*d<=1024 and npad %% m == 0 (npad=%d m=%d)", npad, m);
  SML_CHECK(qkv || qkv_bf16, "sml_nystrom_landmarks_fwd: no source");
  if (qkv) landmarks_fwd_kernel<float><<<dim3(m, B), h * d, 0, as_stream(stream)>>>(qkv, ql, kl, npad, h, d, m);
  else landmarks_fwd_kernel<__nv_bfloat16><<<dim3(m, B), h * d, 0, as_stream(stream)>>>((const __nv_bfloat16*)qkv_bf16, ql, kl, npad, h, d, m);
  SML_LAUNCH_OK("sml_nystrom_landmarks_fwd");
  return 0;
}
extern "C" int sml_nystrom_landmarks_bwd(float* dqkv, const float* dql, const float* dkl, int32_t B, int32_t npad,
                                         int32_t h, int32_t d, int32_t m, void* hi, void* lo, void* stream) {
  SML_CHECK(h * d <= 1024 && npad % m == 0, "sml_nystrom_landmarks_bwd: bad sizes");
  landmarks_bwd_kernel<<<dim3(m, B), h * d, 0, as_stream(stream)>>>(dqkv, dql, dkl, npad, h, d, m, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  SML_LAUNCH_OK("sml_nystrom_landmarks_bwd");
  return 0;
}
extern "C" int sml_nystrom_landmark_ids(int32_t* ids, int32_t npad, int32_t m, void* stream) {
  SML_CHECK(m > 0 && npad % m == 0, "sml_nystrom_landmark_ids: npad %% m != 0");
  landmark_ids_kernel<<<(npad + 255) / 256, 256, 0, as_stream(stream)>>>(ids, npad, npad / m);
  SML_LAUNCH_OK("sml_nystrom_landmark_ids");
  return 0;
}
extern "C" int sml_pinv_scale(const float* x, float* scal, int32_t* arg, void* packed_ws, int32_t BH, int32_t m,
                              void* stream) {
  cudaStream_t s = as_stream(stream);
  SML_CHECK(cudaMemsetAsync(packed_ws, 0, 16, s) == cudaSuccess, "sml_pinv_scale: memset failed");
  pinv_scale_kernel<<<BH, 256, 0, s>>>(x, (unsigned long long*)packed_ws, m);
  SML_LAUNCH_OK("sml_pinv_scale");
  pinv_scale_final_kernel<<<1, 32, 0, s>>>((const unsigned long long*)packed_ws, scal, arg);
  SML_LAUNCH_OK("sml_pinv_scale_final");
  return 0;
}
extern "C" int sml_pinv_init(const float* x, const float* scal, float* z0, int32_t BH, int32_t m, void* stream) {
  pinv_init_kernel<<<dim3((m + 31) / 32, (m + 31) / 32, BH), dim3(32, 8), 0, as_stream(stream)>>>(x, scal, z0, m);
  SML_LAUNCH_OK("sml_pinv_init");
  return 0;
}
extern "C" int sml_pinv_init_bwd(const float* x, const float* scal, const int32_t* arg, const float* dz0, float* dx,
                                 float* tmp, int32_t BH, int32_t m, void* stream) {
  cudaStream_t s = as_stream(stream);
  SML_CHECK(cudaMemsetAsync(tmp, 0, 4, s) == cudaSuccess, "sml_pinv_init_bwd: memset failed");
  pinv_init_bwd_kernel<<<dim3((m + 31) / 32, (m + 31) / 32, BH), dim3(32, 8), 0, s>>>(x, scal, dz0, dx, tmp, m);
  SML_LAUNCH_OK("sml_pinv_init_bwd");
  pinv_init_bwd_scalar_kernel<<<1, 256, 0, s>>>(x, scal, arg, tmp, dx, m);
  SML_LAUNCH_OK("sml_pinv_init_bwd_scalar");
  return 0;
}
static inline dim3 rc_grid(int npad, int hd, int B, int tok) { return dim3((npad + tok - 1) / tok, (hd + 255) / 256, B); }
static inline dim3 rc_grid4(int npad, int hd, int B, int tok) { return dim3((npad + tok - 1) / tok, (hd / 4 + 127) / 128, B); }
extern "C" int sml_nystrom_resconv_fwd(const float* qkv, const float* w, float* out, int32_t B, int32_t npad, int32_t h,
                                       int32_t d, int32_t ksize, void* out_hi, void* out_lo, const void* qkv_bf16, void* stream) {
  const int hd = h * d;
  SML_CHECK(ksize % 2 == 1, "sml_nystrom_resconv_fwd: kernel size must be odd");
  SML_CHECK(qkv || (qkv_bf16 && ksize == 33 && d % 4 == 0), "sml_nystrom_resconv_fwd: the bf16 source needs the 33-tap vector kernel");
  if (!qkv) {
    const __nv_bfloat16* qb = (const __nv_bfloat16*)qkv_bf16;
    resconv_kernel<false, 33, __nv_bfloat16><<<rc_grid4(npad, hd, B, RC_TT), 128, 0, as_stream(stream)>>>(qb + 2 * hd, 3 * hd, w, out, hd, npad, hd, d, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  } else if (ksize == 33 && d % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    if (hd == 512) resconv_kernel<false, 33, float, 1536, 512><<<rc_grid4(npad, hd, B, RC_TT), 128, 0, as_stream(stream)>>>(qkv + 2 * hd, 3 * hd, w, out, hd, npad, hd, d, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
    else resconv_kernel<false, 33, float><<<rc_grid4(npad, hd, B, RC_TT), 128, 0, as_stream(stream)>>>(qkv + 2 * hd, 3 * hd, w, out, hd, npad, hd, d, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  else
    resconv_generic_kernel<false><<<rc_grid(npad, hd, B, 1), 256, 0, as_stream(stream)>>>(qkv + 2 * hd, 3 * hd, w, out, hd, npad, hd, d, ksize);
  SML_LAUNCH_OK("sml_nystrom_resconv_fwd");
  return 0;
}
extern "C" int sml_nystrom_resconv_bwd_blocks(int32_t B, int32_t npad, int32_t hd) {
  return B * ((hd + 255) / 256) * ((npad + RCB_TOK - 1) / RCB_TOK);   // upper bound for both the vector and the generic grid
}
extern "C" int sml_nystrom_resconv_bwd(const float* qkv, const float* w, const float* dout, float* dqkv, float* dw_part,
                                       int32_t B, int32_t npad, int32_t h, int32_t d, int32_t ksize, void* dqkv_hi, void* dqkv_lo,
                                       const void* qkv_bf16, void* stream) {
  const int hd = h * d;
  cudaStream_t s = as_stream(stream);
  SML_CHECK(qkv || (qkv_bf16 && ksize == 33 && d % 4 == 0), "sml_nystrom_resconv_bwd: the bf16 source needs the 33-tap vector kernel");
  if (ksize == 33 && d % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dout) & 15) == 0) {
    resconv_kernel<true, 33, float><<<rc_grid4(npad, hd, B, RC_TT), 128, 0, s>>>(dout, hd, w, dqkv + 2 * hd, 3 * hd, npad, hd, d,
        dqkv_hi ? (__nv_bfloat16*)dqkv_hi + 2 * hd : nullptr, dqkv_lo ? (__nv_bfloat16*)dqkv_lo + 2 * hd : nullptr);
    SML_LAUNCH_OK("sml_nystrom_resconv_bwd(dv)");
    if (qkv) resconv_dw_kernel<33, float><<<rc_grid4(npad, hd, B, RCB_TOK), 128, h * ksize * sizeof(float), s>>>(qkv + 2 * hd, 3 * hd, dout, hd, dw_part, npad, hd, d);
    else resconv_dw_kernel<33, __nv_bfloat16><<<rc_grid4(npad, hd, B, RCB_TOK), 128, h * ksize * sizeof(float), s>>>((const __nv_bfloat16*)qkv_bf16 + 2 * hd, 3 * hd, dout, hd, dw_part, npad, hd, d);
  } else {
    resconv_generic_kernel<true><<<rc_grid(npad, hd, B, 1), 256, 0, s>>>(dout, hd, w, dqkv + 2 * hd, 3 * hd, npad, hd, d, ksize);
    SML_LAUNCH_OK("sml_nystrom_resconv_bwd(dv)");
    resconv_dw_generic_kernel<<<rc_grid(npad, hd, B, RCB_TOK), 256, h * ksize * sizeof(float), s>>>(qkv + 2 * hd, 3 * hd, dout, hd, dw_part, npad, hd, d, ksize);
  }
  SML_LAUNCH_OK("sml_nystrom_resconv_bwd(dw)");
  return 0;
}
// PPEG parameter folding (models/mil.py:192-206): identity + 7x7 + 5x5 + 3x3 depthwise kernels -> one tap-major [49,C] stencil
// and one bias; `unfold` is its backward (slices of d w7 and three copies of d bias).  One thread per (tap, channel).
__global__ void ppeg_fold_kernel(const float* __restrict__ w7, const float* __restrict__ w5, const float* __restrict__ w3,
                                 const float* __restrict__ b7, const float* __restrict__ b5, const float* __restrict__ b3,
                                 float* __restrict__ wf, float* __restrict__ bf, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 49 * C) return;
  const int tap = i / C, c = i % C, ky = tap / 7, kx = tap % 7;
  float v = w7[c * 49 + tap];
  if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) v += w5[c * 25 + (ky - 1) * 5 + (kx - 1)];
  if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) v += w3[c * 9 + (ky - 2) * 3 + (kx - 2)];
  if (tap == 24) v += 1.f;                                  // the residual identity
  wf[i] = v;
  if (tap == 0) bf[c] = b7[c] + b5[c] + b3[c];
}
__global__ void ppeg_unfold_kernel(const float* __restrict__ dwf, float* __restrict__ dw7, float* __restrict__ dw5,
                                   float* __restrict__ dw3, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 49 * C) return;
  const int tap = i / C, c = i % C, ky = tap / 7, kx = tap % 7;
  const float g = dwf[i];
  dw7[c * 49 + tap] = g;
  if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) dw5[c * 25 + (ky - 1) * 5 + (kx - 1)] = g;
  if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) dw3[c * 9 + (ky - 2) * 3 + (kx - 2)] = g;
}
extern "C" int sml_ppeg_fold(const float* w7, const float* w5, const float* w3, const float* b7, const float* b5, const float* b3,
                             float* wf, float* bf, int32_t C, void* stream) {
  ppeg_fold_kernel<<<(49 * C + 255) / 256, 256, 0, as_stream(stream)>>>(w7, w5, w3, b7, b5, b3, wf, bf, C);
  SML_LAUNCH_OK("sml_ppeg_fold");
  return 0;
}
extern "C" int sml_ppeg_unfold(const float* dwf, float* dw7, float* dw5, float* dw3, int32_t C, void* stream) {
  ppeg_unfold_kernel<<<(49 * C + 255) / 256, 256, 0, as_stream(stream)>>>(dwf, dw7, dw5, dw3, C);
  SML_LAUNCH_OK("sml_ppeg_unfold");
  return 0;
}
extern "C" int sml_ppeg_fwd(const float* x, const float* w7, const float* bias, float* y, int32_t B, int32_t H, int32_t W,
                            int32_t C, void* stream) {
  if (C == 512) {                 // constant row pitch: the register-tiled kernel
    const int tx = (W + PRB_W - 1) / PRB_W, tyc = (H + PRB_H - 1) / PRB_H;
    dim3 grid(tx * tyc, (C + 31) / 32, B);
    ppeg_rt_kernel<false, 512><<<grid, 256, 0, as_stream(stream)>>>(x, w7, bias, y, H, W, C, tx);
  } else {                        // other widths: the shared-memory kernel
    const int tx = (W + PT_W - 1) / PT_W, tyc = (H + PT_H - 1) / PT_H;
    dim3 grid(tx * tyc, (C + PT_C - 1) / PT_C, B);
    ppeg_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(x, w7, bias, y, H, W, C, tx);
  }
  SML_LAUNCH_OK("sml_ppeg_fwd");
  return 0;
}
extern "C" int sml_ppeg_bwd_blocks(int32_t B, int32_t H, int32_t W) {
  const int nt = ((W + PRB_W - 1) / PRB_W) * ((H + PRB_H - 1) / PRB_H);
  return B * (nt < PPB_GROUPS ? nt : PPB_GROUPS);
}
extern "C" int sml_ppeg_bwd(const float* x, const float* w7, const float* dy, float* dx, float* dw_part, float* db_part,
                            int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  cudaStream_t s = as_stream(stream);
  const int tx = (W + PRB_W - 1) / PRB_W, tyc = (H + PRB_H - 1) / PRB_H, nt = tx * tyc;
  dim3 grid(nt, (C + 31) / 32, B);                                              // backward: register-tiled kernels (390 / 526 us against 549 / 568 us)
  if (C == 512) ppeg_rt_kernel<true, 512><<<grid, 256, 0, s>>>(dy, w7, nullptr, dx, H, W, C, tx);   // transposed stencil = flipped taps
  else ppeg_rt_kernel<true, 0><<<grid, 256, 0, s>>>(dy, w7, nullptr, dx, H, W, C, tx);
  SML_LAUNCH_OK("sml_ppeg_bwd(dx)");
  dim3 grid2((C + 31) / 32, nt < PPB_GROUPS ? nt : PPB_GROUPS, B);
  if (C == 512) ppeg_dw_rt_kernel<512><<<grid2, 256, 0, s>>>(x, dy, dw_part, db_part, H, W, C, tx, nt);
  else ppeg_dw_rt_kernel<0><<<grid2, 256, 0, s>>>(x, dy, dw_part, db_part, H, W, C, tx, nt);
  SML_LAUNCH_OK("sml_ppeg_bwd(dw)");
  return 0;
}

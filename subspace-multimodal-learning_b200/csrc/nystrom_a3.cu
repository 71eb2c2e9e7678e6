// nystrom_a3.cu — the landmark-by-token product of the Nystrom attention without its probability matrix:
//   A3V = softmax_rows(ql k^T) v      (models/NystromAttention.py:124, 130-137, 140: sim3, attn3, attn3 @ v)
// ql [b, h, m = 256, 64] are the landmark queries, k / v the n' front-padded tokens (n' = 16 640 at 16 k-patch bags).  The
// un-fused path writes the [m, n'] logits in fp32 (1.0 GB per layer at B = 8), re-reads them for the softmax over n' and again
// for the product with v, and the backward repeats that with dA3 and dS3: ~3 GB of HBM traffic for 0.3 GB of operands.  Here
// the forward is a flash attention over key chunks (partial (O, max, sum) per chunk + a small combine) and the backward
// recomputes P from the saved log-sum-exp.  bf16 operands, fp32 accumulation, base-2 exponentials ("bf16" mode only; the
// fp32-grade modes keep the materialised path).
//
// Backward layout: one CTA = ALL 256 landmark rows of one (b, h) x one chunk of keys, 16 warps x 16 rows.  The reduction of
// dK / dV over the query rows is then complete inside the CTA (P / dS tiles go through shared memory, every warp owns a
// 16-key x 16-dim slice of dK or dV) and the results are written once, with the landmark-mean term of dk (dkl[t / l] / l,
// NystromAttention.py:102-118 backward) and the bf16 conversion in the store: no atomics, no zero fill.  dql (256 x 64 per
// (b, h)) is reduced over the chunks with one vector reduction per element and CTA.
#include "common.cuh"

namespace sml {
namespace a3 {

constexpr int D = 64;            // head dim
constexpr int M = 256;           // landmarks
constexpr int LDQ = 72;          // bf16 row stride of the padded Q / dO tiles (144 B: conflict-free ldmatrix)
constexpr int PLD = 40;          // bf16 row stride of the P / dS tiles
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo_f(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(d), "l"(src));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

// token-major bf16 K / V planes inside the packed qkv tensor [b, n', 3, h, 64]: element (b, t, head, c) of K at
// kv + b * bs + t * rs + k_off + head * 64 + c  (V: v_off)
struct Cfg {
  int B, H, npad, m;                 // m = landmark rows per (b, h): a multiple of 64 (forward), 256 (backward)
  int64_t rs, bs;
  int k_off, v_off;
  int tiles_per_chunk, nchunks;      // in tiles of the kernel's key tile (forward 64 keys, backward 32)
};

// one stage of the key ring: NK keys x 64 dims of K and V, rows of 128 bytes with the 16-byte chunk c of row r at c ^ (r & 7)
template <int NK>
struct Stage { __nv_bfloat16 K[NK][D], V[NK][D]; };
__device__ __forceinline__ const __nv_bfloat16* sw(const __nv_bfloat16 (*T)[D], int row, int col) {
  return &T[row][((((col >> 3) ^ row) & 7) << 3) | (col & 7)];
}
template <int NK, int NT>
__device__ __forceinline__ void stage_load(Stage<NK>& st, const __nv_bfloat16* __restrict__ kb, const __nv_bfloat16* __restrict__ vb,
                                           int64_t rs, int k0, int tid) {
  for (int e = tid; e < NK * 8; e += NT) {
    const int r = e >> 3, c8 = (e & 7) * 8;
    const int cs = (((c8 >> 3) ^ r) & 7) << 3;
    const int64_t off = (int64_t)(k0 + r) * rs + c8;
    cp16(&st.K[r][cs], kb + off);
    cp16(&st.V[r][cs], vb + off);
  }
}

// ------------------------------------------------------------------ forward: 64 landmark rows x one key chunk per CTA
constexpr int FK = 64, FT = 128;
struct FwdSmem {
  __nv_bfloat16 Q[64][LDQ];
  Stage<FK> st[3];
};
// opart [B H, nchunks, 256, 64] un-normalised partial outputs, ml [B H, nchunks, 256, 2] = (base-2 running max, sum)
__global__ void __launch_bounds__(FT, 3) a3_fwd_kernel(const Cfg cf, const __nv_bfloat16* __restrict__ ql, const __nv_bfloat16* __restrict__ kv,
                                                        float* __restrict__ opart, float2* __restrict__ ml) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int chunk = blockIdx.x, q0 = blockIdx.y * 64, bh = blockIdx.z, b = bh / cf.H, head = bh % cf.H;
  const __nv_bfloat16* kb = kv + (int64_t)b * cf.bs + cf.k_off + head * D;
  const __nv_bfloat16* vb = kv + (int64_t)b * cf.bs + cf.v_off + head * D;
  const int ntot = cf.npad / FK;
  const int t0 = chunk * cf.tiles_per_chunk, t1 = min(t0 + cf.tiles_per_chunk, ntot);
  const int ntiles = t1 - t0;
  {
    const __nv_bfloat16* qb = ql + ((int64_t)bh * cf.m + q0) * D;
    for (int e = tid; e < 64 * 8; e += FT) cp16(&sm.Q[e >> 3][(e & 7) * 8], qb + (e >> 3) * D + (e & 7) * 8);
  }
  if (ntiles > 0) stage_load<FK, FT>(sm.st[0], kb, vb, cf.rs, t0 * FK, tid);
  cp_commit();
  if (ntiles > 1) stage_load<FK, FT>(sm.st[1], kb, vb, cf.rs, (t0 + 1) * FK, tid);
  cp_commit();
  float oacc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[j][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;          // running max of log2(e) s, running sum

  for (int it = 0; it < ntiles; ++it) {
    cp_wait<1>();
    __syncthreads();
    if (it + 2 < ntiles) stage_load<FK, FT>(sm.st[(it + 2) % 3], kb, vb, cf.rs, (t0 + it + 2) * FK, tid);
    cp_commit();
    const Stage<FK>& st = sm.st[it % 3];
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4];
      ldsm_x4(ah[0], ah[1], ah[2], ah[3], &sm.Q[w * 16 + (lane & 15)][16 * ks + 8 * (lane >> 4)]);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        uint32_t b00, b01, b10, b11;
        ldsm_x4(b00, b01, b10, b11, sw(st.K, 8 * j + (lane & 7) + 8 * (lane >> 4), 16 * ks + 8 * ((lane >> 3) & 1)));
        mma16816(s[j], ah, b00, b01);
        mma16816(s[j + 1], ah, b10, b11);
      }
    }
    float mt0 = -INFINITY, mt1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mt0 = fmaxf(mt0, fmaxf(s[j][0], s[j][1])); mt1 = fmaxf(mt1, fmaxf(s[j][2], s[j][3]));
    }
    mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 1)); mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 1));
    mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 2)); mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 2));
    const float mn0 = fmaxf(m0, mt0 * LOG2E), mn1 = fmaxf(m1, mt1 * LOG2E);
    const float al0 = ex2(m0 - mn0), al1 = ex2(m1 - mn1);               // first tile: 2^(-inf) = 0
    m0 = mn0; m1 = mn1;
    l0 *= al0; l1 *= al1;
#pragma unroll
    for (int j = 0; j < 8; ++j) { oacc[j][0] *= al0; oacc[j][1] *= al0; oacc[j][2] *= al1; oacc[j][3] *= al1; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = ex2(fmaf(s[j][0], LOG2E, -mn0)); s[j][1] = ex2(fmaf(s[j][1], LOG2E, -mn0));
      s[j][2] = ex2(fmaf(s[j][2], LOG2E, -mn1)); s[j][3] = ex2(fmaf(s[j][3], LOG2E, -mn1));
      l0 += s[j][0] + s[j][1]; l1 += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4];
      ah[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
      ah[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
      ah[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      ah[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        uint32_t b00, b01, b10, b11;       // V^T fragments of dims 8 j .. 8 j + 15: keys 16 ks .. + 15 (rows) x dims (columns), transposed
        ldsm_x4_t(b00, b01, b10, b11, sw(st.V, 16 * ks + (lane & 15), 8 * j + 8 * (lane >> 4)));
        mma16816(oacc[j], ah, b00, b01);
        mma16816(oacc[j + 1], ah, b10, b11);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const int r0 = q0 + w * 16 + g, r1 = r0 + 8;
  const int64_t pb = ((int64_t)bh * cf.nchunks + chunk) * cf.m;
  float* o0 = opart + (pb + r0) * D + 2 * t;
  float* o1 = opart + (pb + r1) * D + 2 * t;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<float2*>(o0 + 8 * j) = make_float2(oacc[j][0], oacc[j][1]);
    *reinterpret_cast<float2*>(o1 + 8 * j) = make_float2(oacc[j][2], oacc[j][3]);
  }
  if (t == 0) { ml[pb + r0] = make_float2(m0, l0); ml[pb + r1] = make_float2(m1, l1); }
}

// ------------------------------------------------------------------ token-by-landmark side, inference only
// out[t] = softmax_j(q_t . kl_j) W_j  (NystromAttention.py:122, 130, 140 with W = Z (A3 V) folded: A1 @ W) for 64 token rows per CTA
// and ALL m landmarks (m / 64 key tiles through the same ring): the [n', m] probabilities are not written.  Only usable when no
// gradient is needed (the training backward, sml_nystrom_a1_bwd, reads the saved bf16 A1).  q = strided bf16 view of the packed
// qkv tensor (offset 0), kl / wm contiguous bf16 [B H, m, 64]; out = fp32 [B, n', H 64] (head slice of the merged tensor).
__global__ void __launch_bounds__(FT, 3) a1_fwd_kernel(const Cfg cf, const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ kl,
                                                        const __nv_bfloat16* __restrict__ wm, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 64, bh = blockIdx.y, b = bh / cf.H, head = bh % cf.H;
  const __nv_bfloat16* kb = kl + (int64_t)bh * cf.m * D;
  const __nv_bfloat16* vb = wm + (int64_t)bh * cf.m * D;
  const int ntiles = cf.m / FK;
  {
    const __nv_bfloat16* qb = qkv + (int64_t)b * cf.bs + (int64_t)q0 * cf.rs + head * D;
    for (int e = tid; e < 64 * 8; e += FT) cp16(&sm.Q[e >> 3][(e & 7) * 8], qb + (int64_t)(e >> 3) * cf.rs + (e & 7) * 8);
  }
  stage_load<FK, FT>(sm.st[0], kb, vb, D, 0, tid);
  cp_commit();
  if (ntiles > 1) stage_load<FK, FT>(sm.st[1], kb, vb, D, FK, tid);
  cp_commit();
  float oacc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[j][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  for (int it = 0; it < ntiles; ++it) {
    cp_wait<1>();
    __syncthreads();
    if (it + 2 < ntiles) stage_load<FK, FT>(sm.st[(it + 2) % 3], kb, vb, D, (it + 2) * FK, tid);
    cp_commit();
    const Stage<FK>& st = sm.st[it % 3];
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4];
      ldsm_x4(ah[0], ah[1], ah[2], ah[3], &sm.Q[w * 16 + (lane & 15)][16 * ks + 8 * (lane >> 4)]);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        uint32_t b00, b01, b10, b11;
        ldsm_x4(b00, b01, b10, b11, sw(st.K, 8 * j + (lane & 7) + 8 * (lane >> 4), 16 * ks + 8 * ((lane >> 3) & 1)));
        mma16816(s[j], ah, b00, b01);
        mma16816(s[j + 1], ah, b10, b11);
      }
    }
    float mt0 = -INFINITY, mt1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mt0 = fmaxf(mt0, fmaxf(s[j][0], s[j][1])); mt1 = fmaxf(mt1, fmaxf(s[j][2], s[j][3]));
    }
    mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 1)); mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 1));
    mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 2)); mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 2));
    const float mn0 = fmaxf(m0, mt0 * LOG2E), mn1 = fmaxf(m1, mt1 * LOG2E);
    const float al0 = ex2(m0 - mn0), al1 = ex2(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= al0; l1 *= al1;
#pragma unroll
    for (int j = 0; j < 8; ++j) { oacc[j][0] *= al0; oacc[j][1] *= al0; oacc[j][2] *= al1; oacc[j][3] *= al1; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = ex2(fmaf(s[j][0], LOG2E, -mn0)); s[j][1] = ex2(fmaf(s[j][1], LOG2E, -mn0));
      s[j][2] = ex2(fmaf(s[j][2], LOG2E, -mn1)); s[j][3] = ex2(fmaf(s[j][3], LOG2E, -mn1));
      l0 += s[j][0] + s[j][1]; l1 += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4];
      ah[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
      ah[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
      ah[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      ah[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        uint32_t b00, b01, b10, b11;
        ldsm_x4_t(b00, b01, b10, b11, sw(st.V, 16 * ks + (lane & 15), 8 * j + 8 * (lane >> 4)));
        mma16816(oacc[j], ah, b00, b01);
        mma16816(oacc[j + 1], ah, b10, b11);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = q0 + w * 16 + g, r1 = r0 + 8;
  float* o0 = out + ((int64_t)b * cf.npad + r0) * (cf.H * D) + head * D + 2 * t;
  float* o1 = out + ((int64_t)b * cf.npad + r1) * (cf.H * D) + head * D + 2 * t;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<float2*>(o0 + 8 * j) = make_float2(oacc[j][0] * i0, oacc[j][1] * i0);
    *reinterpret_cast<float2*>(o1 + 8 * j) = make_float2(oacc[j][2] * i1, oacc[j][3] * i1);
  }
}

// a3v [B H, 256, 64] (fp32 + bf16 copy), lse [B H, 256] (natural log) from the chunk partials; one thread per (row, 4 dims)
__global__ void __launch_bounds__(256) a3_combine_kernel(const float* __restrict__ opart, const float2* __restrict__ ml, int nchunks,
                                                         int64_t rows, int M, float* __restrict__ a3v, __nv_bfloat16* __restrict__ a3v_hi,
                                                         float* __restrict__ lse) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 16) return;
  const int64_t row = i >> 4;                 // bh * m + r
  const int c4 = (int)(i & 15) * 4;
  const int64_t bh = row / M, r = row % M;
  float mx = -INFINITY;
  for (int c = 0; c < nchunks; ++c) mx = fmaxf(mx, ml[(bh * nchunks + c) * M + r].x);
  float L = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < nchunks; ++c) {
    const float2 v = ml[(bh * nchunks + c) * M + r];
    const float sc = ex2(v.x - mx);
    L = fmaf(v.y, sc, L);
    const float4 o = *reinterpret_cast<const float4*>(opart + ((bh * nchunks + c) * M + r) * D + c4);
    acc.x = fmaf(o.x, sc, acc.x); acc.y = fmaf(o.y, sc, acc.y); acc.z = fmaf(o.z, sc, acc.z); acc.w = fmaf(o.w, sc, acc.w);
  }
  const float inv = 1.f / L;
  acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
  *reinterpret_cast<float4*>(a3v + row * D + c4) = acc;
  if (a3v_hi) *reinterpret_cast<uint2*>(a3v_hi + row * D + c4) = make_uint2(pack_bf16(acc.x, acc.y), pack_bf16(acc.z, acc.w));
  if (c4 == 0) lse[row] = (mx + log2f(L)) * LN2;
}

// ------------------------------------------------------------------ backward: all 256 landmark rows x one key chunk per CTA
constexpr int BK = 32, BT = 512;
struct BwdSmem {
  __nv_bfloat16 Q[M][LDQ], dO[M][LDQ];
  Stage<BK> st[3];
  __nv_bfloat16 P[2][M][PLD], dS[2][M][PLD];
};
// ql bf16, da3v / a3v fp32 [B H, 256, 64], lse [B H, 256]; dql [B H, 256, 64] fp32 accumulated (vector reductions);
// dk_hi = bf16 K slice of the packed gradient (same addressing as K), dv = fp32 V slice of the packed fp32 gradient
// (addressing with the same element strides); dkl [B H, 256, 64] fp32, lsz = tokens per landmark: dk[t] = dS^T ql + dkl[t / lsz] / lsz
__global__ void __launch_bounds__(BT, 1) a3_bwd_kernel(const Cfg cf, const __nv_bfloat16* __restrict__ ql, const __nv_bfloat16* __restrict__ kv,
                                                        const float* __restrict__ da3v, const float* __restrict__ a3v,
                                                        const float* __restrict__ lse, const float* __restrict__ dkl, int lsz,
                                                        float* __restrict__ dql, __nv_bfloat16* __restrict__ dkv_hi, float* __restrict__ dkv_f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int chunk = blockIdx.x, bh = blockIdx.y, b = bh / cf.H, head = bh % cf.H;
  const __nv_bfloat16* kb = kv + (int64_t)b * cf.bs + cf.k_off + head * D;
  const __nv_bfloat16* vb = kv + (int64_t)b * cf.bs + cf.v_off + head * D;
  const int ntot = cf.npad / BK;
  const int t0 = chunk * cf.tiles_per_chunk, t1 = min(t0 + cf.tiles_per_chunk, ntot);
  const int ntiles = t1 - t0;
  if (ntiles <= 0) return;
  stage_load<BK, BT>(sm.st[0], kb, vb, cf.rs, t0 * BK, tid);
  cp_commit();
  if (ntiles > 1) stage_load<BK, BT>(sm.st[1], kb, vb, cf.rs, (t0 + 1) * BK, tid);
  cp_commit();
  {
    const __nv_bfloat16* qb = ql + (int64_t)bh * M * D;
    for (int e = tid; e < M * 8; e += BT) cp16(&sm.Q[e >> 3][(e & 7) * 8], qb + (e >> 3) * D + (e & 7) * 8);
    cp_commit();
    const float* gb = da3v + (int64_t)bh * M * D;
    for (int e = tid; e < M * 16; e += BT) {
      const int r = e >> 4, c4 = (e & 15) * 4;
      const float4 y = *reinterpret_cast<const float4*>(gb + r * D + c4);
      *reinterpret_cast<uint2*>(&sm.dO[r][c4]) = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
    }
  }
  const int r0 = w * 16 + g, r1 = r0 + 8;
  float delta0, delta1, lse0, lse1;            // delta = bf16(dO) . O per row; lse in base-2 units
  {
    float d[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int64_t ro = ((int64_t)bh * M + (rr ? r1 : r0)) * D + 16 * t;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(a3v + ro + 4 * i), c = *reinterpret_cast<const float4*>(da3v + ro + 4 * i);
        const uint32_t c01 = pack_bf16(c.x, c.y), c23 = pack_bf16(c.z, c.w);
        acc = fmaf(a.x, bf16_lo_f(c01), fmaf(a.y, bf16_hi_f(c01), fmaf(a.z, bf16_lo_f(c23), fmaf(a.w, bf16_hi_f(c23), acc))));
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      d[rr] = acc;
    }
    delta0 = d[0]; delta1 = d[1];
    lse0 = lse[(int64_t)bh * M + r0] * LOG2E; lse1 = lse[(int64_t)bh * M + r1] * LOG2E;
  }
  float dqacc[8][4];
#pragma unroll
  for (int jd = 0; jd < 8; ++jd)
#pragma unroll
    for (int e = 0; e < 4; ++e) dqacc[jd][e] = 0.f;
  // my slice of dK (warps 0 .. 7) or dV (warps 8 .. 15): keys 16 mt .. + 15 of the tile, dims 16 dh .. + 15
  const int mt = w & 1, dh = (w >> 1) & 3, which = w >> 3;
  const float inv_l = 1.f / (float)lsz;
  cp_wait<0>();                                 // tiles 0, 1 and the Q tile
  __syncthreads();

  for (int it = 0; it < ntiles; ++it) {
    const int k0 = (t0 + it) * BK;
    const Stage<BK>& st = sm.st[it % 3];
    __nv_bfloat16 (*Pt)[PLD] = sm.P[it & 1];
    __nv_bfloat16 (*dSt)[PLD] = sm.dS[it & 1];
    // ---- S = ql k^T and dP = dO v^T of my 16 rows x the 32 keys of the tile
    float s[4][4], dp[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[j][e] = 0.f; dp[j][e] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4], gh[4];
      ldsm_x4(ah[0], ah[1], ah[2], ah[3], &sm.Q[w * 16 + (lane & 15)][16 * ks + 8 * (lane >> 4)]);
      ldsm_x4(gh[0], gh[1], gh[2], gh[3], &sm.dO[w * 16 + (lane & 15)][16 * ks + 8 * (lane >> 4)]);
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const int krow = 8 * j + (lane & 7) + 8 * (lane >> 4), kcol = 16 * ks + 8 * ((lane >> 3) & 1);
        uint32_t k00, k01, k10, k11, v00, v01, v10, v11;
        ldsm_x4(k00, k01, k10, k11, sw(st.K, krow, kcol));
        ldsm_x4(v00, v01, v10, v11, sw(st.V, krow, kcol));
        mma16816(s[j], ah, k00, k01);
        mma16816(dp[j], gh, v00, v01);
        mma16816(s[j + 1], ah, k10, k11);
        mma16816(dp[j + 1], gh, v10, v11);
      }
    }
    // ---- P = 2^(log2(e) S - lse2), dS = P (dP - delta): bf16 tiles for the transposed products
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float p0 = ex2(fmaf(s[j][0], LOG2E, -lse0)), p1 = ex2(fmaf(s[j][1], LOG2E, -lse0));
      const float p2 = ex2(fmaf(s[j][2], LOG2E, -lse1)), p3 = ex2(fmaf(s[j][3], LOG2E, -lse1));
      *reinterpret_cast<uint32_t*>(&Pt[r0][8 * j + 2 * t]) = pack_bf16(p0, p1);
      *reinterpret_cast<uint32_t*>(&Pt[r1][8 * j + 2 * t]) = pack_bf16(p2, p3);
      *reinterpret_cast<uint32_t*>(&dSt[r0][8 * j + 2 * t]) = pack_bf16(p0 * (dp[j][0] - delta0), p1 * (dp[j][1] - delta0));
      *reinterpret_cast<uint32_t*>(&dSt[r1][8 * j + 2 * t]) = pack_bf16(p2 * (dp[j][2] - delta1), p3 * (dp[j][3] - delta1));
    }
    cp_wait<0>();                                       // my copies of tile it + 1 have landed
    __syncthreads();                                    // P / dS of this tile and tile it + 1 visible; everyone is past tile it - 1
    if (it + 2 < ntiles) stage_load<BK, BT>(sm.st[(it + 2) % 3], kb, vb, cf.rs, k0 + 2 * BK, tid);
    cp_commit();
    // ---- dQ += dS K
#pragma unroll
    for (int ks = 0; ks < BK / 16; ++ks) {
      uint32_t ah[4];
      ldsm_x4(ah[0], ah[1], ah[2], ah[3], &dSt[w * 16 + (lane & 15)][16 * ks + 8 * (lane >> 4)]);
#pragma unroll
      for (int jd = 0; jd < 8; jd += 2) {
        uint32_t b00, b01, b10, b11;
        ldsm_x4_t(b00, b01, b10, b11, sw(st.K, 16 * ks + (lane & 15), 8 * jd + 8 * (lane >> 4)));
        mma16816(dqacc[jd], ah, b00, b01);
        mma16816(dqacc[jd + 1], ah, b10, b11);
      }
    }
    // ---- my 16-key x 16-dim slice of dK = dS^T ql (warps 0-7) or dV = P^T dO (warps 8-15): contraction over the 256 rows
    {
      // two accumulator sets (even / odd 16-row steps): four independent MMA chains of 8 instead of two chains of 16 - with
      // two chains this loop was bound by the MMA latency (40 % of the kernel's stall samples)
      float acc2[2][2][4];
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int jd = 0; jd < 2; ++jd)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc2[p][jd][e] = 0.f;
      __nv_bfloat16 (*T)[PLD] = which ? Pt : dSt;
      const __nv_bfloat16 (*Bq)[LDQ] = which ? sm.dO : sm.Q;
      const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
      for (int ks = 0; ks < M / 16; ++ks) {
        uint32_t a[4], b00, b01, b10, b11;
        ldsm_x4_t(a[0], a[1], a[2], a[3], &T[16 * ks + rr + 8 * (mi >> 1)][16 * mt + 8 * (mi & 1)]);
        ldsm_x4_t(b00, b01, b10, b11, &Bq[16 * ks + (lane & 15)][16 * dh + 8 * (lane >> 4)]);
        mma16816(acc2[ks & 1][0], a, b00, b01);
        mma16816(acc2[ks & 1][1], a, b10, b11);
      }
      float acc[2][4];
#pragma unroll
      for (int jd = 0; jd < 2; ++jd)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[jd][e] = acc2[0][jd][e] + acc2[1][jd][e];
      const int key0 = k0 + 16 * mt + g, key1 = key0 + 8;
      const int64_t e0 = (int64_t)b * cf.bs + (int64_t)key0 * cf.rs + head * D + 16 * dh + 2 * t;
      const int64_t e1 = e0 + 8 * cf.rs;
      if (which) {
#pragma unroll
        for (int jd = 0; jd < 2; ++jd) {
          *reinterpret_cast<float2*>(dkv_f + e0 + cf.v_off + 8 * jd) = make_float2(acc[jd][0], acc[jd][1]);
          *reinterpret_cast<float2*>(dkv_f + e1 + cf.v_off + 8 * jd) = make_float2(acc[jd][2], acc[jd][3]);
        }
      } else {
        float2 a0[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, a1[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if (dkl) {
          const float* l0p = dkl + ((int64_t)bh * M + key0 / lsz) * D + 16 * dh + 2 * t;
          const float* l1p = dkl + ((int64_t)bh * M + key1 / lsz) * D + 16 * dh + 2 * t;
#pragma unroll
          for (int jd = 0; jd < 2; ++jd) { a0[jd] = *reinterpret_cast<const float2*>(l0p + 8 * jd); a1[jd] = *reinterpret_cast<const float2*>(l1p + 8 * jd); }
        }
#pragma unroll
        for (int jd = 0; jd < 2; ++jd) {
          *reinterpret_cast<uint32_t*>(dkv_hi + e0 + cf.k_off + 8 * jd) = pack_bf16(fmaf(a0[jd].x, inv_l, acc[jd][0]), fmaf(a0[jd].y, inv_l, acc[jd][1]));
          *reinterpret_cast<uint32_t*>(dkv_hi + e1 + cf.k_off + 8 * jd) = pack_bf16(fmaf(a1[jd].x, inv_l, acc[jd][2]), fmaf(a1[jd].y, inv_l, acc[jd][3]));
        }
      }
    }
  }
  // ---- dql += my rows (one chunk of keys): vector reductions
  {
    float* q0p = dql + ((int64_t)bh * M + r0) * D + 2 * t;
    float* q1p = dql + ((int64_t)bh * M + r1) * D + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(reinterpret_cast<float2*>(q0p + 8 * j), make_float2(dqacc[j][0], dqacc[j][1]));
      atomicAdd(reinterpret_cast<float2*>(q1p + 8 * j), make_float2(dqacc[j][2], dqacc[j][3]));
    }
  }
}

// dk[b, t, head, :] += dkl[b, head, t / lsz, :] / lsz on the bf16 K slice of the packed gradient (the landmark-mean term, when it
// is only known after the kernel above has run: dkl also collects dS1^T q from the A1 backward).  One thread per 8 dims.
__global__ void __launch_bounds__(256) a3_bcast_add_kernel(const Cfg cf, __nv_bfloat16* __restrict__ dkv_hi, const float* __restrict__ dkl, int lsz) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)cf.B * cf.npad * cf.H * 8;
  if (i >= total) return;
  const int c8 = (int)(i & 7) * 8;
  const int head = (int)((i >> 3) % cf.H);
  const int64_t bt = (i >> 3) / cf.H;
  const int tkn = (int)(bt % cf.npad), b = (int)(bt / cf.npad);
  __nv_bfloat16* p = dkv_hi + (int64_t)b * cf.bs + (int64_t)tkn * cf.rs + cf.k_off + head * D + c8;
  const float* a = dkl + (((int64_t)b * cf.H + head) * M + tkn / lsz) * D + c8;
  const float inv_l = 1.f / (float)lsz;
  uint4 v = *reinterpret_cast<uint4*>(p);
  const float4 a0 = *reinterpret_cast<const float4*>(a), a1 = *reinterpret_cast<const float4*>(a + 4);
  v.x = pack_bf16(fmaf(a0.x, inv_l, bf16_lo_f(v.x)), fmaf(a0.y, inv_l, bf16_hi_f(v.x)));
  v.y = pack_bf16(fmaf(a0.z, inv_l, bf16_lo_f(v.y)), fmaf(a0.w, inv_l, bf16_hi_f(v.y)));
  v.z = pack_bf16(fmaf(a1.x, inv_l, bf16_lo_f(v.z)), fmaf(a1.y, inv_l, bf16_hi_f(v.z)));
  v.w = pack_bf16(fmaf(a1.z, inv_l, bf16_lo_f(v.w)), fmaf(a1.w, inv_l, bf16_hi_f(v.w)));
  *reinterpret_cast<uint4*>(p) = v;
}

}  // namespace a3
}  // namespace sml

using namespace sml;

static int a3_cfg(a3::Cfg& cf, const int32_t* c, int tile, const char* who) {
  // c = (B, H, m, d, npad, row stride, batch stride (elements, as two int32: lo, hi), k_off, v_off, nchunks)
  cf.B = c[0]; cf.H = c[1]; cf.m = c[2];
  SML_CHECK(c[3] == a3::D && c[2] > 0 && c[2] % 64 == 0 && (tile == a3::FK || c[2] == a3::M),
            "%s: needs head dim 64 and a multiple of 64 landmarks (256 for the backward) (got %d, %d)", who, c[2], c[3]);
  cf.npad = c[4];
  SML_CHECK(cf.npad % 64 == 0 && cf.npad > 0, "%s: padded length must be a multiple of 64 (got %d)", who, cf.npad);
  cf.rs = c[5];
  cf.bs = (int64_t)(uint32_t)c[6] | ((int64_t)c[7] << 32);
  cf.k_off = c[8]; cf.v_off = c[9];
  cf.nchunks = c[10];
  SML_CHECK(cf.nchunks >= 1, "%s: nchunks", who);
  const int ntot = cf.npad / tile;
  cf.tiles_per_chunk = (ntot + cf.nchunks - 1) / cf.nchunks;
  SML_CHECK(cf.rs % 8 == 0 && cf.bs % 8 == 0 && cf.k_off % 8 == 0 && cf.v_off % 8 == 0, "%s: 16-byte aligned strides", who);
  return 0;
}

/* A3V = softmax(ql k^T) v per (b, h) without materialising the [256, n'] probabilities (models/NystromAttention.py:124,
 * 130-137, 140).  cfg11 = (B, H, 256, 64, n', row stride, batch stride lo, hi, k offset, v offset, nchunks): K / V are bf16 views
 * into the packed qkv tensor (element strides).  opart [B H nchunks 256 64] fp32 and ml [B H nchunks 256 2] fp32 are scratch;
 * a3v [B, H, 256, 64] fp32 (+ optional bf16 copy a3v_hi), lse [B, H, 256] natural log-sum-exp of the rows (saved for backward). */
extern "C" int sml_nystrom_a3_fwd(const int32_t* cfg11, const void* ql_bf16, const void* qkv_bf16, float* opart, float* ml, float* a3v,
                                  void* a3v_hi, float* lse, void* stream) {
  a3::Cfg cf;
  if (a3_cfg(cf, cfg11, a3::FK, "sml_nystrom_a3_fwd")) return 1;
  SML_CHECK(ql_bf16 && qkv_bf16 && opart && ml && a3v && lse, "sml_nystrom_a3_fwd: null pointer");
  const size_t smem = sizeof(a3::FwdSmem);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(a3::a3_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  dim3 grid(cf.nchunks, cf.m / 64, cf.B * cf.H);
  a3::a3_fwd_kernel<<<grid, a3::FT, smem, as_stream(stream)>>>(cf, (const __nv_bfloat16*)ql_bf16, (const __nv_bfloat16*)qkv_bf16, opart,
                                                               reinterpret_cast<float2*>(ml));
  SML_LAUNCH_OK("sml_nystrom_a3_fwd");
  const int64_t rows = (int64_t)cf.B * cf.H * cf.m;
  a3::a3_combine_kernel<<<(unsigned)((rows * 16 + 255) / 256), 256, 0, as_stream(stream)>>>(opart, reinterpret_cast<const float2*>(ml), cf.nchunks,
                                                                                            rows, cf.m, a3v, (__nv_bfloat16*)a3v_hi, lse);
  SML_LAUNCH_OK("sml_nystrom_a3_fwd(combine)");
  return 0;
}

/* Backward of sml_nystrom_a3_fwd given da3v: dql += dS3 k (accumulated into the existing [B, H, 256, 64] fp32 tensor),
 * dk = dS3^T ql -> the bf16 K slice of the packed gradient dqkv_bf16, with the landmark-mean term dkl[t / lsz] / lsz
 * (NystromAttention.py:102-118 backward) added when dkl is given (else: sml_nystrom_a3_dk_landmarks afterwards),
 * dv = P^T da3v -> the fp32 V slice of dqkv_f32 (same element strides as qkv).  P is recomputed from lse; nothing of size
 * [256, n'] touches global memory. */
extern "C" int sml_nystrom_a3_bwd(const int32_t* cfg11, const void* ql_bf16, const void* qkv_bf16, const float* da3v, const float* a3v,
                                  const float* lse, const float* dkl, int32_t lsz, float* dql, void* dqkv_bf16, float* dqkv_f32,
                                  void* stream) {
  a3::Cfg cf;
  if (a3_cfg(cf, cfg11, a3::BK, "sml_nystrom_a3_bwd")) return 1;
  SML_CHECK(ql_bf16 && qkv_bf16 && da3v && a3v && lse && dql && dqkv_bf16 && dqkv_f32 && lsz > 0, "sml_nystrom_a3_bwd: null pointer");
  const size_t smem = sizeof(a3::BwdSmem);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(a3::a3_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  dim3 grid(cf.nchunks, cf.B * cf.H);
  a3::a3_bwd_kernel<<<grid, a3::BT, smem, as_stream(stream)>>>(cf, (const __nv_bfloat16*)ql_bf16, (const __nv_bfloat16*)qkv_bf16, da3v, a3v, lse,
                                                               dkl, lsz, dql, (__nv_bfloat16*)dqkv_bf16, dqkv_f32);
  SML_LAUNCH_OK("sml_nystrom_a3_bwd");
  return 0;
}

/* dk += dkl[t / lsz] / lsz on the bf16 K slice (for callers whose dkl is complete only after sml_nystrom_a3_bwd has run) */
extern "C" int sml_nystrom_a3_dk_landmarks(const int32_t* cfg11, const float* dkl, int32_t lsz, void* dqkv_bf16, void* stream) {
  a3::Cfg cf;
  if (a3_cfg(cf, cfg11, a3::BK, "sml_nystrom_a3_dk_landmarks")) return 1;
  SML_CHECK(dkl && dqkv_bf16 && lsz > 0, "sml_nystrom_a3_dk_landmarks: null pointer");
  const int64_t total = (int64_t)cf.B * cf.npad * cf.H * 8;
  a3::a3_bcast_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(cf, (__nv_bfloat16*)dqkv_bf16, dkl, lsz);
  SML_LAUNCH_OK("sml_nystrom_a3_dk_landmarks");
  return 0;
}

/* Inference only: out[b, t, head 64 ..] = softmax_j(q_t . kl_j) wm_j without the [n', m] probabilities (models/NystromAttention.py:
 * 122, 130, 140 with wm = Z (A3 V)).  cfg11 as above (q = the packed tensor at offset 0); kl, wm bf16 [B, H, m, 64]; out fp32
 * [B, n', H 64]. */
extern "C" int sml_nystrom_a1_fwd_nograd(const int32_t* cfg11, const void* qkv_bf16, const void* kl_bf16, const void* wm_bf16, float* out,
                                         void* stream) {
  a3::Cfg cf;
  if (a3_cfg(cf, cfg11, a3::FK, "sml_nystrom_a1_fwd_nograd")) return 1;
  SML_CHECK(qkv_bf16 && kl_bf16 && wm_bf16 && out, "sml_nystrom_a1_fwd_nograd: null pointer");
  const size_t smem = sizeof(a3::FwdSmem);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(a3::a1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  dim3 grid(cf.npad / 64, cf.B * cf.H);
  a3::a1_fwd_kernel<<<grid, a3::FT, smem, as_stream(stream)>>>(cf, (const __nv_bfloat16*)qkv_bf16, (const __nv_bfloat16*)kl_bf16,
                                                               (const __nv_bfloat16*)wm_bf16, out);
  SML_LAUNCH_OK("sml_nystrom_a1_fwd_nograd");
  return 0;
}

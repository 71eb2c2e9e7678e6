// deform_mma.cu — tensor-core version of the fused deformable attention (models/DeformableAttention1D.py:60-102,
// 211-231; models/DeformableAttention2D.py:120-157, 292-312).
//
// The continuous-position-bias MLP is evaluated per (query, key) pair:  Linear(CD,32) ReLU Linear(32,32) ReLU Linear(32,1).
// Its 32x32 layer is 2 kFLOP of the 2.2 kFLOP per pair, so it is run as a [16 pairs x 32] x [32 x 32] warp-level MMA:
// every thread builds its slice of the layer-1 activations directly in the A-fragment registers (no shared-memory round
// trip), the accumulators come back in registers, and ReLU + layer 3 is a thread-local dot product finished by a
// 4-lane reduction.  Q K^T and P V run on the same register-fragment pipeline (flash-attention style, online softmax).
// Precision: SPLIT = true runs every product as bf16 hi/lo 3-pass (a_lo b_hi + a_hi b_lo + a_hi b_hi, fp32 accumulate;
// ~2^-16 relative, meets the fp32 tolerance), SPLIT = false is plain bf16 operands.
//
// Per pair the CUDA cores still do ~110 (bf16) / ~160 (split) operations (signed-log distance, layer 1, packing, ReLU,
// layer 3), which bounds the kernel; the MMA work hides under it.
#include "common.cuh"
#include "deform_attn.cuh"

namespace sml {

constexpr int MQ = 64;        // queries per CTA (4 warps x 16 rows)
constexpr int MTHREADS = 128;
constexpr int MLD = 72;       // bf16 row stride of the 64-wide operand tiles (144 B: conflict-free fragment loads)

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {       // low 16 bits <- lo
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo_f(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }
template <bool SPLIT>
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);
  if (SPLIT) lo = pack_bf16(a - bf16_lo_f(hi), b - bf16_hi_f(hi));
  else lo = 0u;
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// d = a b + c with the addend in its own registers (no accumulator-initialising moves)
__device__ __forceinline__ void mma16816_c(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c0, float c1,
                                           float c2, float c3) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};\n"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}
// d += a b with hi/lo operands (small terms first)
template <bool SPLIT>
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                     uint32_t bl0, uint32_t bl1) {
  if (SPLIT) {
    mma16816(d, al, bh0, bh1);
    mma16816(d, ah, bl0, bl1);
  }
  mma16816(d, ah, bh0, bh1);
}
// three-term bf16 split (24 mantissa bits: exact for fp32 inputs) and the 6-pass product that keeps every term above 2^-24
__device__ __forceinline__ void split3(float a, float b, uint32_t& h, uint32_t& m, uint32_t& l) {
  h = pack_bf16(a, b);
  const float ra = a - bf16_lo_f(h), rb = b - bf16_hi_f(h);
  m = pack_bf16(ra, rb);
  l = pack_bf16(ra - bf16_lo_f(m), rb - bf16_hi_f(m));
}
__device__ __forceinline__ void mma6(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&am)[4], const uint32_t (&al)[4],
                                     uint32_t bh0, uint32_t bh1, uint32_t bm0, uint32_t bm1, uint32_t bl0, uint32_t bl1) {
  mma16816(d, al, bh0, bh1);
  mma16816(d, ah, bl0, bl1);
  mma16816(d, am, bm0, bm1);
  mma16816(d, am, bh0, bh1);
  mma16816(d, ah, bm0, bm1);
  mma16816(d, ah, bh0, bh1);
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ uint32_t movm_trans(uint32_t a) {
  uint32_t d;
  asm("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(d) : "r"(a));
  return d;
}
// sign(x) log(|x| + 1) (1D.py:93-94, 2D.py:150-151); MUFU.LG2 based (abs. error < 1e-6 on |x| <= 4)
__device__ __forceinline__ float slog_fast(float x) { return copysignf(__logf(fabsf(x) + 1.f), x); }

// fp32 rows [rows x 64] of a token-major tensor -> bf16 hi (/lo) tiles [64][MLD]; rows >= limit are zero-filled
template <bool SPLIT, int ROWS>
__device__ __forceinline__ void load_tile(__nv_bfloat16 (*hi)[MLD], __nv_bfloat16 (*lo)[MLD], const float* __restrict__ src,
                                          int64_t row_stride, int row0, int limit) {
  for (int e = threadIdx.x; e < ROWS * 16; e += MTHREADS) {
    const int r = e >> 4, c4 = (e & 15) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < limit) x = *reinterpret_cast<const float4*>(src + (int64_t)(row0 + r) * row_stride + c4);
    uint32_t h0, l0, h1, l1;
    split2<SPLIT>(x.x, x.y, h0, l0);
    split2<SPLIT>(x.z, x.w, h1, l1);
    *reinterpret_cast<uint2*>(&hi[r][c4]) = make_uint2(h0, h1);
    if (SPLIT) *reinterpret_cast<uint2*>(&lo[r][c4]) = make_uint2(l0, l1);
  }
}

// same, three planes (hi / mid / lo): operands of the fp32-exact products (P V and dO V^T, see the SPLIT note on delta)
template <int ROWS>
__device__ __forceinline__ void load_tile3(__nv_bfloat16 (*p0)[MLD], __nv_bfloat16 (*p1)[MLD], __nv_bfloat16 (*p2)[MLD],
                                           const float* __restrict__ src, int64_t row_stride, int row0, int limit) {
  for (int e = threadIdx.x; e < ROWS * 16; e += MTHREADS) {
    const int r = e >> 4, c4 = (e & 15) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < limit) x = *reinterpret_cast<const float4*>(src + (int64_t)(row0 + r) * row_stride + c4);
    uint32_t h0, m0, l0, h1, m1, l1;
    split3(x.x, x.y, h0, m0, l0);
    split3(x.z, x.w, h1, m1, l1);
    *reinterpret_cast<uint2*>(&p0[r][c4]) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(&p1[r][c4]) = make_uint2(m0, m1);
    *reinterpret_cast<uint2*>(&p2[r][c4]) = make_uint2(l0, l1);
  }
}

// Per-thread constants of the CPB MLP in MMA-fragment order.  Thread (g = lane / 4, t = lane % 4):
//   layer-1 hidden units k(s, e) = 16 s + 2 t + (e & 1) + 8 (e >> 1)   (the A-fragment columns of k-step s)
//   layer-2 outputs      c(j, e) = 8 j + 2 t + e                        (the C-fragment columns of n-tile j)
template <int CD, bool SPLIT>
struct CpbRegs {
  float w1x[2][4], w1y[2][4], b1[2][4];
  float b2[4][2], w3[4][2], b3;
  uint32_t wh[4][2][2], wl[4][2][2];      // B fragments of W2^T: [n-tile j][k-step s][reg]

  __device__ __forceinline__ void load(const float* __restrict__ cpb, int OUT, int o, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const float* W1 = cpb; const float* B1 = W1 + CPB_H * CD; const float* W2 = B1 + CPB_H;
    const float* B2 = W2 + CPB_H * CPB_H; const float* W3 = B2 + CPB_H; const float* B3 = W3 + OUT * CPB_H;
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = 16 * s + 2 * t + (e & 1) + 8 * (e >> 1);
        w1x[s][e] = W1[k * CD];
        w1y[s][e] = CD == 2 ? W1[k * CD + 1] : 0.f;
        b1[s][e] = B1[k];
      }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) { b2[j][e] = B2[8 * j + 2 * t + e]; w3[j][e] = W3[o * CPB_H + 8 * j + 2 * t + e]; }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float* row = W2 + (8 * j + g) * CPB_H + 16 * s + 2 * t;      // B[k][n] = W2[c = n][k]
        split2<SPLIT>(row[0], row[1], wh[j][s][0], wl[j][s][0]);
        split2<SPLIT>(row[8], row[9], wh[j][s][1], wl[j][s][1]);
      }
    }
    b3 = B3[o];
  }
};

// layer 1 for the two pairs (rows g and g + 8 of the 16-pair tile): A fragments of relu(W1 pos + b1)
template <int CD, bool SPLIT>
__device__ __forceinline__ void cpb_layer1(const CpbRegs<CD, SPLIT>& c, float ax0, float ay0, float ax1, float ay1,
                                           uint32_t (&ah)[2][4], uint32_t (&al)[2][4], float (*hraw)[2][4] = nullptr) {
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    float h0[4], h1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float u0 = fmaf(c.w1x[s][e], ax0, c.b1[s][e]), u1 = fmaf(c.w1x[s][e], ax1, c.b1[s][e]);
      if (CD == 2) { u0 = fmaf(c.w1y[s][e], ay0, u0); u1 = fmaf(c.w1y[s][e], ay1, u1); }
      h0[e] = fmaxf(u0, 0.f); h1[e] = fmaxf(u1, 0.f);
      if (hraw) { hraw[s][0][e] = h0[e]; hraw[s][1][e] = h1[e]; }
    }
    split2<SPLIT>(h0[0], h0[1], ah[s][0], al[s][0]);
    split2<SPLIT>(h1[0], h1[1], ah[s][1], al[s][1]);
    split2<SPLIT>(h0[2], h0[3], ah[s][2], al[s][2]);
    split2<SPLIT>(h1[2], h1[3], ah[s][3], al[s][3]);
  }
}
// layer 2 pre-activations: acc[j] = C fragment (rows g / g+8, columns c(j, 0..1)) of b2 + h1 W2^T
template <int CD, bool SPLIT>
__device__ __forceinline__ void cpb_layer2(const CpbRegs<CD, SPLIT>& c, const uint32_t (&ah)[2][4], const uint32_t (&al)[2][4],
                                           float (&acc)[4][4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc[j][0] = c.b2[j][0]; acc[j][1] = c.b2[j][1]; acc[j][2] = c.b2[j][0]; acc[j][3] = c.b2[j][1];
#pragma unroll
    for (int s = 0; s < 2; ++s) mma3<SPLIT>(acc[j], ah[s], al[s], c.wh[j][s][0], c.wh[j][s][1], c.wl[j][s][0], c.wl[j][s][1]);
  }
}

template <bool SPLIT>
struct FwdSmem {
  __nv_bfloat16 Q[SPLIT ? 2 : 1][MQ][MLD];
  __nv_bfloat16 K[SPLIT ? 2 : 1][64][MLD];
  __nv_bfloat16 V[SPLIT ? 3 : 2][64][MLD];     // P V always runs one precision level above the rest (see the backward's delta)
  float kpos[64][2];
};

// grid (ceil(n / 64), heads, B), 128 threads.  q pre-scaled by d^-1/2.  out [b, n, inner], lse [b, heads, n].
template <int CD, bool SPLIT>
__global__ void __launch_bounds__(MTHREADS, SPLIT ? 2 : 3) deform_attn_fwd_mma_kernel(const AttnCfg cf, const float* __restrict__ q,
                                                                       const float* __restrict__ k, const float* __restrict__ v,
                                                                       const float* __restrict__ vs, const float* __restrict__ cpb,
                                                                       float* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem<SPLIT>& sm = *reinterpret_cast<FwdSmem<SPLIT>*>(smem_raw);
  constexpr int LO = SPLIT ? 1 : 0;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * MQ;
  const int grp = head / cf.OUT, o = head % cf.OUT;
  const int bg = b * cf.G + grp;
  CpbRegs<CD, SPLIT> cr;
  cr.load(cpb, cf.OUT, o, lane);
  load_tile<SPLIT, MQ>(sm.Q[0], sm.Q[LO], q + (int64_t)b * cf.n * cf.inner + head * AD, cf.inner, q0, cf.n);
  const int r0 = w * 16 + g, r1 = r0 + 8;               // my two query rows inside the CTA tile
  float qx0, qy0, qx1, qy1;
  query_pos(cf, q0 + r0, qx0, qy0);
  query_pos(cf, q0 + r1, qx1, qy1);
  float oacc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[j][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;       // l: per-thread partial row sums
  const float keep_scale = cf.p_drop > 0.f ? 1.f / (1.f - cf.p_drop) : 1.f;
  const uint32_t thresh = (uint32_t)(cf.p_drop * 4294967296.0);
  const float* kb = k + (int64_t)b * cf.nkv * cf.inner + head * AD;
  const float* vb = v + (int64_t)b * cf.nkv * cf.inner + head * AD;

  for (int k0 = 0; k0 < cf.nkv; k0 += 64) {
    __syncthreads();
    load_tile<SPLIT, 64>(sm.K[0], sm.K[LO], kb, cf.inner, k0, cf.nkv);
    if (SPLIT) load_tile3<64>(sm.V[0], sm.V[LO], sm.V[2 * LO], vb, cf.inner, k0, cf.nkv);
    else load_tile<true, 64>(sm.V[0], sm.V[1], vb, cf.inner, k0, cf.nkv);
    if (tid < 64) {
      const bool ok = k0 + tid < cf.nkv;
      sm.kpos[tid][0] = ok ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD] : 0.f;
      sm.kpos[tid][1] = (ok && CD == 2) ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD + 1] : 0.f;
    }
    __syncthreads();
    // ---- S = Q K^T : C fragments s[j] = rows (r0, r1) x keys 8 j + 2 t + {0, 1}
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4], al[4] = {0u, 0u, 0u, 0u};
      ah[0] = *reinterpret_cast<const uint32_t*>(&sm.Q[0][r0][16 * ks + 2 * t]);
      ah[1] = *reinterpret_cast<const uint32_t*>(&sm.Q[0][r1][16 * ks + 2 * t]);
      ah[2] = *reinterpret_cast<const uint32_t*>(&sm.Q[0][r0][16 * ks + 2 * t + 8]);
      ah[3] = *reinterpret_cast<const uint32_t*>(&sm.Q[0][r1][16 * ks + 2 * t + 8]);
      if (SPLIT) {
        al[0] = *reinterpret_cast<const uint32_t*>(&sm.Q[LO][r0][16 * ks + 2 * t]);
        al[1] = *reinterpret_cast<const uint32_t*>(&sm.Q[LO][r1][16 * ks + 2 * t]);
        al[2] = *reinterpret_cast<const uint32_t*>(&sm.Q[LO][r0][16 * ks + 2 * t + 8]);
        al[3] = *reinterpret_cast<const uint32_t*>(&sm.Q[LO][r1][16 * ks + 2 * t + 8]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&sm.K[0][8 * j + g][16 * ks + 2 * t]);
        const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&sm.K[0][8 * j + g][16 * ks + 2 * t + 8]);
        uint32_t bl0 = 0u, bl1 = 0u;
        if (SPLIT) {
          bl0 = *reinterpret_cast<const uint32_t*>(&sm.K[LO][8 * j + g][16 * ks + 2 * t]);
          bl1 = *reinterpret_cast<const uint32_t*>(&sm.K[LO][8 * j + g][16 * ks + 2 * t + 8]);
        }
        mma3<SPLIT>(s[j], ah, al, bh0, bh1, bl0, bl1);
      }
    }
    // ---- + continuous position bias: for key 8 j + u the 16 queries of this warp form one 16-pair MMA tile
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      float bias[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float2 kp = *reinterpret_cast<const float2*>(sm.kpos[8 * j + u]);
        const float ax0 = slog_fast(qx0 - kp.x), ax1 = slog_fast(qx1 - kp.x);
        const float ay0 = CD == 2 ? slog_fast(qy0 - kp.y) : 0.f, ay1 = CD == 2 ? slog_fast(qy1 - kp.y) : 0.f;
        uint32_t ah[2][4], al[2][4];
        cpb_layer1<CD, SPLIT>(cr, ax0, ay0, ax1, ay1, ah, al);
        float acc[4][4];
        cpb_layer2<CD, SPLIT>(cr, ah, al, acc);
        float p0 = 0.f, p1 = 0.f;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          p0 = fmaf(cr.w3[jj][0], fmaxf(acc[jj][0], 0.f), p0); p0 = fmaf(cr.w3[jj][1], fmaxf(acc[jj][1], 0.f), p0);
          p1 = fmaf(cr.w3[jj][0], fmaxf(acc[jj][2], 0.f), p1); p1 = fmaf(cr.w3[jj][1], fmaxf(acc[jj][3], 0.f), p1);
        }
        p0 += __shfl_xor_sync(0xffffffffu, p0, 1); p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
        p0 += __shfl_xor_sync(0xffffffffu, p0, 2); p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
        if ((u >> 1) == t) { bias[u & 1] = p0 + cr.b3; bias[2 + (u & 1)] = p1 + cr.b3; }
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)          // static register indices (the j loop stays rolled)
        if (jj == j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) s[jj][e] += bias[e];
        }
    }
    // ---- online softmax on the fragments (rows r0 / r1 live in the 4 lanes of a quad)
    float mt0 = -INFINITY, mt1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (k0 + 8 * j + 2 * t + e >= cf.nkv) { s[j][e] = -INFINITY; s[j][2 + e] = -INFINITY; }
        mt0 = fmaxf(mt0, s[j][e]); mt1 = fmaxf(mt1, s[j][2 + e]);
      }
    mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 1)); mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 1));
    mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 2)); mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 2));
    const float mn0 = fmaxf(m0, mt0), mn1 = fmaxf(m1, mt1);
    const float al0 = m0 == -INFINITY ? 0.f : __expf(m0 - mn0), al1 = m1 == -INFINITY ? 0.f : __expf(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= al0; l1 *= al1;
#pragma unroll
    for (int j = 0; j < 8; ++j) { oacc[j][0] *= al0; oacc[j][1] *= al0; oacc[j][2] *= al1; oacc[j][3] *= al1; }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float p0 = __expf(s[j][e] - mn0), p1 = __expf(s[j][2 + e] - mn1);
        l0 += p0; l1 += p1;
        if (cf.p_drop > 0.f) {
          const unsigned long long key = (unsigned long long)(k0 + 8 * j + 2 * t + e);
          const unsigned long long base = (unsigned long long)(b * cf.heads + head) * cf.n;
          p0 = pair_hash(salted(cf.seed, cf.salt), (base + (q0 + r0)) * cf.nkv + key) >= thresh ? p0 * keep_scale : 0.f;
          p1 = pair_hash(salted(cf.seed, cf.salt), (base + (q0 + r1)) * cf.nkv + key) >= thresh ? p1 * keep_scale : 0.f;
        }
        s[j][e] = p0; s[j][2 + e] = p1;
      }
    // ---- O += P V : A fragments straight from the probability fragments, B fragments by transposing ldmatrix
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4], am[4], al[4];
      if (SPLIT) {
        split3(s[2 * ks][0], s[2 * ks][1], ah[0], am[0], al[0]);
        split3(s[2 * ks][2], s[2 * ks][3], ah[1], am[1], al[1]);
        split3(s[2 * ks + 1][0], s[2 * ks + 1][1], ah[2], am[2], al[2]);
        split3(s[2 * ks + 1][2], s[2 * ks + 1][3], ah[3], am[3], al[3]);
      } else {
        split2<true>(s[2 * ks][0], s[2 * ks][1], ah[0], al[0]);
        split2<true>(s[2 * ks][2], s[2 * ks][3], ah[1], al[1]);
        split2<true>(s[2 * ks + 1][0], s[2 * ks + 1][1], ah[2], al[2]);
        split2<true>(s[2 * ks + 1][2], s[2 * ks + 1][3], ah[3], al[3]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t bh0, bh1;
        ldmatrix_x2_trans(bh0, bh1, &sm.V[0][16 * ks + (lane & 15)][8 * j]);
        if (SPLIT) {
          uint32_t bm0, bm1, bl0, bl1;
          ldmatrix_x2_trans(bm0, bm1, &sm.V[LO][16 * ks + (lane & 15)][8 * j]);
          ldmatrix_x2_trans(bl0, bl1, &sm.V[2 * LO][16 * ks + (lane & 15)][8 * j]);
          mma6(oacc[j], ah, am, al, bh0, bh1, bm0, bm1, bl0, bl1);
        } else {
          uint32_t bl0, bl1;
          ldmatrix_x2_trans(bl0, bl1, &sm.V[1][16 * ks + (lane & 15)][8 * j]);
          mma3<true>(oacc[j], ah, al, bh0, bh1, bl0, bl1);
        }
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  if (q0 + r0 < cf.n) {
    float* orow = out + ((int64_t)b * cf.n + q0 + r0) * cf.inner + head * AD + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(orow + 8 * j) = make_float2(oacc[j][0] * i0, oacc[j][1] * i0);
    if (t == 0) lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r0] = m0 + logf(l0);
  }
  if (q0 + r1 < cf.n) {
    float* orow = out + ((int64_t)b * cf.n + q0 + r1) * cf.inner + head * AD + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(orow + 8 * j) = make_float2(oacc[j][2] * i1, oacc[j][3] * i1);
    if (t == 0) lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r1] = m1 + logf(l1);
  }
}


// ------------------------------------------------------------------------------------------ backward
// One CTA per (64-query tile, head, batch), 4 warps x 16 query rows, key tiles of 32.  Per key tile:
//   phase 1  S = Q K^T and dP = dO V^T (MMA) -> fp32 staging tiles in shared memory (warp-private rows)
//   phase 2  per key: the warp's 16 (query, key) pairs form one MMA tile: CPB forward (layers 1-3), P = exp(S + bias - lse),
//            dS = P (dP keep - delta); CPB backward: dh2 -> dh1 = dh2 W2 (MMA) -> layer-1 / coordinate gradients, and
//            dW2 += dh2^T h1 as an MMA over the pair axis (operands transposed in registers with movmatrix).
//            P keep and dS are written as bf16 (hi/lo) tiles for phase 3.
//   phase 3  dQ += dS K (registers), dK += dS^T Q and dV += (P keep)^T dO per 16-key x 32-dim warp slice -> global atomics.
// CPB parameter gradients stay in registers for the whole kernel and leave as one partial row per CTA.
constexpr int BK = 32;        // keys per tile
constexpr int PLD = 40;       // bf16 row stride of the P / dS tiles (80 B rows: conflict-free ldmatrix)
constexpr int NRED = CPB_H * 2 + CPB_H + CPB_H * CPB_H + CPB_H + CPB_H + 1;

template <bool SPLIT>
struct BwdSmem {
  // dO and V carry one more bf16 plane than the other operands: dP = dO V^T (and O = P V in the forward) run one precision
  // level above the rest (SPLIT: 3-term / 6-pass = fp32-exact; plain bf16 mode: hi/lo 3-pass) so that delta = dO . O equals
  // sum_j P dP to rounding and sum_j dS_ij cancels as it does in exact arithmetic; every CPB parameter gradient is a sum of
  // dS against a slowly varying factor and inherits that cancellation.
  __nv_bfloat16 Q[SPLIT ? 2 : 1][MQ][MLD], dO[SPLIT ? 3 : 2][MQ][MLD];
  __nv_bfloat16 K[SPLIT ? 2 : 1][BK][MLD], V[SPLIT ? 3 : 2][BK][MLD];
  __nv_bfloat16 P[SPLIT ? 2 : 1][MQ][PLD], dS[SPLIT ? 2 : 1][MQ][PLD];     // `red` aliases P at the end of the kernel
  float dQ[32][MTHREADS];            // thread-private dQ accumulators (C-fragment order), live across key tiles
  float stg[4][2][16][9];            // per warp: S and dP of the current 8-key group (fp32)
  float kgs[4][8][16];               // per warp: key-coordinate gradient partials of the current 8-key group [g][2 u + coord]
  float c1[4][24];                   // layer-1 constants by t: w1x[2][4], w1y[2][4], b1[2][4]
  float c2[4][16];                   // b2[4][2], w3[4][2] by t
  float kpos[BK][2], dkpos[BK][2];
};
// volatile shared-memory vector load: re-issued at every use instead of being held in (spilled) registers
__device__ __forceinline__ float4 lds128v(const float* p) {
  float4 r;
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
  return r;
}

__device__ __forceinline__ void store_bf16_split(__nv_bfloat16* hi, __nv_bfloat16* lo, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  *hi = h;
  if (lo) *lo = __float2bfloat16_rn(v - __bfloat162float(h));
}

template <int CD, bool SPLIT>
__global__ void __launch_bounds__(MTHREADS, 2) deform_attn_bwd_mma_kernel(
    const AttnCfg cf, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const float* __restrict__ vs, const float* __restrict__ cpb, const float* __restrict__ out,
    const float* __restrict__ lse, const float* __restrict__ dout, float* __restrict__ dq, float* __restrict__ dk,
    float* __restrict__ dv, float* __restrict__ dvs, float* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem<SPLIT>& sm = *reinterpret_cast<BwdSmem<SPLIT>*>(smem_raw);
  constexpr int LO = SPLIT ? 1 : 0;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * MQ;
  const int grp = head / cf.OUT, o = head % cf.OUT;
  const int bg = b * cf.G + grp;
  // ---- constants: W2 fragments in registers (both orientations), the small vectors in shared memory (by t)
  uint32_t wh[4][2][2], wl[4][2][2];        // layer 2:  B[k][n = c] = W2[c][k]      [n-tile j][k-step s][reg]
  uint32_t wth[4][2][2], wtl[4][2][2];      // dh1 = dh2 W2:  B[k' = c][n = k] = W2[c][k]   [n-tile jn][k-step s'][reg]
  uint32_t w1bh[2][2], w1bl[2][2];          // d pos = dpre W1:  B[k][n < CD] = W1[k][n]     [k-step s][reg]
  float b3;
  {
    const float* W1 = cpb; const float* B1 = W1 + CPB_H * CD; const float* W2 = B1 + CPB_H;
    const float* B2 = W2 + CPB_H * CPB_H; const float* W3 = B2 + CPB_H; const float* B3 = W3 + cf.OUT * CPB_H;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float* row = W2 + (8 * j + g) * CPB_H + 16 * s + 2 * t;
        split2<SPLIT>(row[0], row[1], wh[j][s][0], wl[j][s][0]);
        split2<SPLIT>(row[8], row[9], wh[j][s][1], wl[j][s][1]);
        const float* col = W2 + (16 * s + 2 * t) * CPB_H + 8 * j + g;
        split2<SPLIT>(col[0], col[CPB_H], wth[j][s][0], wtl[j][s][0]);
        split2<SPLIT>(col[8 * CPB_H], col[9 * CPB_H], wth[j][s][1], wtl[j][s][1]);
      }
    b3 = B3[o];
#pragma unroll
    for (int s = 0; s < 2; ++s) {            // B[k][n] = W1[k][n] for n < CD (else 0): operand of d pos = dpre W1
      const int k0_ = 16 * s + 2 * t;
      const bool on = g < CD;
      const float v0 = on ? W1[k0_ * CD + g] : 0.f, v1 = on ? W1[(k0_ + 1) * CD + g] : 0.f;
      const float v2 = on ? W1[(k0_ + 8) * CD + g] : 0.f, v3 = on ? W1[(k0_ + 9) * CD + g] : 0.f;
      split2<SPLIT>(v0, v1, w1bh[s][0], w1bl[s][0]);
      split2<SPLIT>(v2, v3, w1bh[s][1], w1bl[s][1]);
    }
    if (tid < 4 * 24) {
      const int tt = tid / 24, i = tid % 24, which = i / 8, s = (i % 8) / 4, e = i % 4;
      const int kk = 16 * s + 2 * tt + (e & 1) + 8 * (e >> 1);
      sm.c1[tt][i] = which == 0 ? W1[kk * CD] : (which == 1 ? (CD == 2 ? W1[kk * CD + 1] : 0.f) : B1[kk]);
    }
    if (tid < 4 * 16) {
      const int tt = tid / 16, i = tid % 16, which = i / 8, j = (i % 8) / 2, e = i % 2;
      const int c = 8 * j + 2 * tt + e;
      sm.c2[tt][i] = which == 0 ? B2[c] : W3[o * CPB_H + c];
    }
  }
  const int64_t qoff = (int64_t)b * cf.n * cf.inner + head * AD;
  load_tile<SPLIT, MQ>(sm.Q[0], sm.Q[LO], q + qoff, cf.inner, q0, cf.n);
  if (SPLIT) load_tile3<MQ>(sm.dO[0], sm.dO[LO], sm.dO[2 * LO], dout + qoff, cf.inner, q0, cf.n);
  else load_tile<true, MQ>(sm.dO[0], sm.dO[1], dout + qoff, cf.inner, q0, cf.n);
#pragma unroll
  for (int i = 0; i < 32; ++i) sm.dQ[i][tid] = 0.f;
  const int r0 = w * 16 + g, r1 = r0 + 8;
  const bool ok0 = q0 + r0 < cf.n, ok1 = q0 + r1 < cf.n;
  float qx0, qy0, qx1, qy1;
  query_pos(cf, q0 + r0, qx0, qy0);
  query_pos(cf, q0 + r1, qx1, qy1);
  // delta = rowsum(dO * O), lse (the four lanes of a quad share a row)
  float delta0 = 0.f, delta1 = 0.f, lse0 = INFINITY, lse1 = INFINITY;
  {
    if (ok0) {
      const float* orow = out + qoff + (int64_t)(q0 + r0) * cf.inner + 16 * t;
      const float* grow = dout + qoff + (int64_t)(q0 + r0) * cf.inner + 16 * t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(orow + 4 * i), c = *reinterpret_cast<const float4*>(grow + 4 * i);
        delta0 = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, fmaf(a.w, c.w, delta0))));
      }
      lse0 = lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r0];
    }
    if (ok1) {
      const float* orow = out + qoff + (int64_t)(q0 + r1) * cf.inner + 16 * t;
      const float* grow = dout + qoff + (int64_t)(q0 + r1) * cf.inner + 16 * t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(orow + 4 * i), c = *reinterpret_cast<const float4*>(grow + 4 * i);
        delta1 = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, fmaf(a.w, c.w, delta1))));
      }
      lse1 = lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r1];
    }
    delta0 += __shfl_xor_sync(0xffffffffu, delta0, 1); delta1 += __shfl_xor_sync(0xffffffffu, delta1, 1);
    delta0 += __shfl_xor_sync(0xffffffffu, delta0, 2); delta1 += __shfl_xor_sync(0xffffffffu, delta1, 2);
  }
  // CPB parameter gradient accumulators (fragment order)
  float gW2[2][4][4];                       // C fragments of dW2[c][k]: [m-tile of c][n-tile of k]
  float gL1[2][4];                          // C fragments of [dW1x | dW1y | db1 | 0...][k][n]: [m-tile of k]
  float gb2[4][2], gw3[4][2], gb3 = 0.f;    // by output c(j, e)
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      gL1[a][c] = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) gW2[a][c][e] = 0.f;
    }
#pragma unroll
  for (int j = 0; j < 4; ++j) { gb2[j][0] = 0.f; gb2[j][1] = 0.f; gw3[j][0] = 0.f; gw3[j][1] = 0.f; }
  const float keep_scale = cf.p_drop > 0.f ? 1.f / (1.f - cf.p_drop) : 1.f;
  const uint32_t thresh = (uint32_t)(cf.p_drop * 4294967296.0);
  const unsigned long long hbase = (unsigned long long)(b * cf.heads + head) * cf.n;
  const float* kb = k + (int64_t)b * cf.nkv * cf.inner + head * AD;
  const float* vb = v + (int64_t)b * cf.nkv * cf.inner + head * AD;
  const float* c1 = sm.c1[t];
  const float* c2 = sm.c2[t];
  float (*stS)[9] = sm.stg[w][0];
  float (*stD)[9] = sm.stg[w][1];

  for (int k0 = 0; k0 < cf.nkv; k0 += BK) {
    __syncthreads();
    load_tile<SPLIT, BK>(sm.K[0], sm.K[LO], kb, cf.inner, k0, cf.nkv);
    if (SPLIT) load_tile3<BK>(sm.V[0], sm.V[LO], sm.V[2 * LO], vb, cf.inner, k0, cf.nkv);
    else load_tile<true, BK>(sm.V[0], sm.V[1], vb, cf.inner, k0, cf.nkv);
    if (tid < BK) {
      const bool ok = k0 + tid < cf.nkv;
      sm.kpos[tid][0] = ok ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD] : 0.f;
      sm.kpos[tid][1] = (ok && CD == 2) ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD + 1] : 0.f;
      sm.dkpos[tid][0] = 0.f; sm.dkpos[tid][1] = 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < BK / 8; ++j) {
      // ---- phase 1: S = Q K^T and dP = dO V^T of my 16 rows x this 8-key group -> warp-private staging
      {
        float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t ah[4], al[4] = {0u, 0u, 0u, 0u}, gh[4];
          const int c0 = 16 * ks + 2 * t;
          ah[0] = *reinterpret_cast<const uint32_t*>(&sm.Q[0][r0][c0]);     ah[1] = *reinterpret_cast<const uint32_t*>(&sm.Q[0][r1][c0]);
          ah[2] = *reinterpret_cast<const uint32_t*>(&sm.Q[0][r0][c0 + 8]); ah[3] = *reinterpret_cast<const uint32_t*>(&sm.Q[0][r1][c0 + 8]);
          gh[0] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r0][c0]);     gh[1] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r1][c0]);
          gh[2] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r0][c0 + 8]); gh[3] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r1][c0 + 8]);
          const uint32_t kh0 = *reinterpret_cast<const uint32_t*>(&sm.K[0][8 * j + g][c0]);
          const uint32_t kh1 = *reinterpret_cast<const uint32_t*>(&sm.K[0][8 * j + g][c0 + 8]);
          const uint32_t vh0 = *reinterpret_cast<const uint32_t*>(&sm.V[0][8 * j + g][c0]);
          const uint32_t vh1 = *reinterpret_cast<const uint32_t*>(&sm.V[0][8 * j + g][c0 + 8]);
          if (SPLIT) {
            uint32_t gm[4], gl[4];
            al[0] = *reinterpret_cast<const uint32_t*>(&sm.Q[LO][r0][c0]);     al[1] = *reinterpret_cast<const uint32_t*>(&sm.Q[LO][r1][c0]);
            al[2] = *reinterpret_cast<const uint32_t*>(&sm.Q[LO][r0][c0 + 8]); al[3] = *reinterpret_cast<const uint32_t*>(&sm.Q[LO][r1][c0 + 8]);
            gm[0] = *reinterpret_cast<const uint32_t*>(&sm.dO[LO][r0][c0]);     gm[1] = *reinterpret_cast<const uint32_t*>(&sm.dO[LO][r1][c0]);
            gm[2] = *reinterpret_cast<const uint32_t*>(&sm.dO[LO][r0][c0 + 8]); gm[3] = *reinterpret_cast<const uint32_t*>(&sm.dO[LO][r1][c0 + 8]);
            gl[0] = *reinterpret_cast<const uint32_t*>(&sm.dO[2 * LO][r0][c0]);     gl[1] = *reinterpret_cast<const uint32_t*>(&sm.dO[2 * LO][r1][c0]);
            gl[2] = *reinterpret_cast<const uint32_t*>(&sm.dO[2 * LO][r0][c0 + 8]); gl[3] = *reinterpret_cast<const uint32_t*>(&sm.dO[2 * LO][r1][c0 + 8]);
            const uint32_t kl0 = *reinterpret_cast<const uint32_t*>(&sm.K[LO][8 * j + g][c0]);
            const uint32_t kl1 = *reinterpret_cast<const uint32_t*>(&sm.K[LO][8 * j + g][c0 + 8]);
            const uint32_t vm0 = *reinterpret_cast<const uint32_t*>(&sm.V[LO][8 * j + g][c0]);
            const uint32_t vm1 = *reinterpret_cast<const uint32_t*>(&sm.V[LO][8 * j + g][c0 + 8]);
            const uint32_t vl0 = *reinterpret_cast<const uint32_t*>(&sm.V[2 * LO][8 * j + g][c0]);
            const uint32_t vl1 = *reinterpret_cast<const uint32_t*>(&sm.V[2 * LO][8 * j + g][c0 + 8]);
            mma3<true>(s, ah, al, kh0, kh1, kl0, kl1);
            mma6(dp, gh, gm, gl, vh0, vh1, vm0, vm1, vl0, vl1);
          } else {
            uint32_t gl[4];
            gl[0] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r0][c0]);     gl[1] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r1][c0]);
            gl[2] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r0][c0 + 8]); gl[3] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r1][c0 + 8]);
            const uint32_t vl0 = *reinterpret_cast<const uint32_t*>(&sm.V[1][8 * j + g][c0]);
            const uint32_t vl1 = *reinterpret_cast<const uint32_t*>(&sm.V[1][8 * j + g][c0 + 8]);
            mma16816(s, ah, kh0, kh1);
            mma3<true>(dp, gh, gl, vh0, vh1, vl0, vl1);
          }
        }
        __syncwarp();
        stS[g][2 * t] = s[0]; stS[g][2 * t + 1] = s[1]; stS[g + 8][2 * t] = s[2]; stS[g + 8][2 * t + 1] = s[3];
        stD[g][2 * t] = dp[0]; stD[g][2 * t + 1] = dp[1]; stD[g + 8][2 * t] = dp[2]; stD[g + 8][2 * t + 1] = dp[3];
        __syncwarp();
      }
      // ---- phase 2: CPB forward + backward, one key (16 pairs) per MMA tile
#pragma unroll 1
      for (int u = 0; u < 8; ++u) {
        const int key = 8 * j + u;
        const float2 kp = *reinterpret_cast<const float2*>(sm.kpos[key]);
        const float dx0 = qx0 - kp.x, dx1 = qx1 - kp.x, dy0 = qy0 - kp.y, dy1 = qy1 - kp.y;
        const float ax0 = slog_fast(dx0), ax1 = slog_fast(dx1);
        const float ay0 = CD == 2 ? slog_fast(dy0) : 0.f, ay1 = CD == 2 ? slog_fast(dy1) : 0.f;
        // layer 1 -> A fragments (hidden unit of (s, e): 16 s + 2 t + (e & 1) + 8 (e >> 1))
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const float4 wx = lds128v(c1 + 4 * s), wy = lds128v(c1 + 8 + 4 * s), bb = lds128v(c1 + 16 + 4 * s);
          const float wxa[4] = {wx.x, wx.y, wx.z, wx.w}, wya[4] = {wy.x, wy.y, wy.z, wy.w}, bba[4] = {bb.x, bb.y, bb.z, bb.w};
          float h0[4], h1[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float u0 = fmaf(wxa[e], ax0, bba[e]), u1 = fmaf(wxa[e], ax1, bba[e]);
            if (CD == 2) { u0 = fmaf(wya[e], ay0, u0); u1 = fmaf(wya[e], ay1, u1); }
            h0[e] = fmaxf(u0, 0.f); h1[e] = fmaxf(u1, 0.f);
          }
          split2<SPLIT>(h0[0], h0[1], ah[s][0], al[s][0]);
          split2<SPLIT>(h1[0], h1[1], ah[s][1], al[s][1]);
          split2<SPLIT>(h0[2], h0[3], ah[s][2], al[s][2]);
          split2<SPLIT>(h1[2], h1[3], ah[s][3], al[s][3]);
        }
        // layer 2 pre-activations
        float acc[4][4];
        {
          const float4 ba = lds128v(c2), bbv = lds128v(c2 + 4);
          const float b2v[8] = {ba.x, ba.y, ba.z, ba.w, bbv.x, bbv.y, bbv.z, bbv.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            acc[jj][0] = b2v[2 * jj]; acc[jj][1] = b2v[2 * jj + 1]; acc[jj][2] = b2v[2 * jj]; acc[jj][3] = b2v[2 * jj + 1];
#pragma unroll
            for (int s = 0; s < 2; ++s) mma3<SPLIT>(acc[jj], ah[s], al[s], wh[jj][s][0], wh[jj][s][1], wl[jj][s][0], wl[jj][s][1]);
          }
        }
        const float4 w3a = lds128v(c2 + 8), w3b = lds128v(c2 + 12);
        const float w3v[8] = {w3a.x, w3a.y, w3a.z, w3a.w, w3b.x, w3b.y, w3b.z, w3b.w};
        float p0 = 0.f, p1 = 0.f;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          p0 = fmaf(w3v[2 * jj], fmaxf(acc[jj][0], 0.f), p0); p0 = fmaf(w3v[2 * jj + 1], fmaxf(acc[jj][1], 0.f), p0);
          p1 = fmaf(w3v[2 * jj], fmaxf(acc[jj][2], 0.f), p1); p1 = fmaf(w3v[2 * jj + 1], fmaxf(acc[jj][3], 0.f), p1);
        }
        p0 += __shfl_xor_sync(0xffffffffu, p0, 1); p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
        p0 += __shfl_xor_sync(0xffffffffu, p0, 2); p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
        // probabilities and logit gradients of the two pairs (computed by all four lanes of the quad)
        const bool kok = k0 + key < cf.nkv;
        const float P0 = kok ? __expf(stS[g][u] + p0 + b3 - lse0) : 0.f;
        const float P1 = kok ? __expf(stS[g + 8][u] + p1 + b3 - lse1) : 0.f;
        float keep0 = 1.f, keep1 = 1.f;
        if (cf.p_drop > 0.f) {
          keep0 = pair_hash(salted(cf.seed, cf.salt), (hbase + (q0 + r0)) * cf.nkv + k0 + key) >= thresh ? keep_scale : 0.f;
          keep1 = pair_hash(salted(cf.seed, cf.salt), (hbase + (q0 + r1)) * cf.nkv + k0 + key) >= thresh ? keep_scale : 0.f;
        }
        const float dS0 = P0 * (stD[g][u] * keep0 - delta0);
        const float dS1 = P1 * (stD[g + 8][u] * keep1 - delta1);
        if (t < 2) {                                     // lane t = 0 stores row r0, lane t = 1 row r1
          const int rs = t ? r1 : r0;
          store_bf16_split(&sm.P[0][rs][key], SPLIT ? &sm.P[LO][rs][key] : nullptr, t ? P1 * keep1 : P0 * keep0);
          store_bf16_split(&sm.dS[0][rs][key], SPLIT ? &sm.dS[LO][rs][key] : nullptr, t ? dS1 : dS0);
        }
        if (t == 0) gb3 += dS0 + dS1;
        // layer 3 / second ReLU backward: dh2 (C-fragment layout) and its A fragments for dh1 = dh2 W2
        uint32_t dah[2][4], dal[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          float d0[2][2], d1[2][2];                         // [n-tile 2 s + i][col e], rows r0 / r1
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int jj = 2 * s + i;
              const float a0 = acc[jj][e], a1 = acc[jj][2 + e];
              gw3[jj][e] = fmaf(dS0, fmaxf(a0, 0.f), fmaf(dS1, fmaxf(a1, 0.f), gw3[jj][e]));
              d0[i][e] = a0 > 0.f ? dS0 * w3v[2 * jj + e] : 0.f;
              d1[i][e] = a1 > 0.f ? dS1 * w3v[2 * jj + e] : 0.f;
              gb2[jj][e] += d0[i][e] + d1[i][e];
            }
          split2<SPLIT>(d0[0][0], d0[0][1], dah[s][0], dal[s][0]);
          split2<SPLIT>(d1[0][0], d1[0][1], dah[s][1], dal[s][1]);
          split2<SPLIT>(d0[1][0], d0[1][1], dah[s][2], dal[s][2]);
          split2<SPLIT>(d1[1][0], d1[1][1], dah[s][3], dal[s][3]);
        }
        // dW2[c][k] += sum_pairs dh2[p][c] h1[p][k]: A = dh2^T, B = h1, both transposed in registers
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t tah[4], tal[4] = {0u, 0u, 0u, 0u};
          tah[0] = movm_trans(dah[mt][0]); tah[1] = movm_trans(dah[mt][2]); tah[2] = movm_trans(dah[mt][1]); tah[3] = movm_trans(dah[mt][3]);
          if (SPLIT) { tal[0] = movm_trans(dal[mt][0]); tal[1] = movm_trans(dal[mt][2]); tal[2] = movm_trans(dal[mt][1]); tal[3] = movm_trans(dal[mt][3]); }
#pragma unroll
          for (int jn = 0; jn < 4; ++jn) {
            const int s = jn >> 1, rb = (jn & 1) * 2;
            const uint32_t bh0 = movm_trans(ah[s][rb]), bh1 = movm_trans(ah[s][rb + 1]);
            uint32_t bl0 = 0u, bl1 = 0u;
            if (SPLIT) { bl0 = movm_trans(al[s][rb]); bl1 = movm_trans(al[s][rb + 1]); }
            mma3<SPLIT>(gW2[mt][jn], tah, tal, bh0, bh1, bl0, bl1);
          }
        }
        // dh1 = dh2 W2
        float dh1[4][4];
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          dh1[jn][0] = 0.f; dh1[jn][1] = 0.f; dh1[jn][2] = 0.f; dh1[jn][3] = 0.f;
#pragma unroll
          for (int s = 0; s < 2; ++s) mma3<SPLIT>(dh1[jn], dah[s], dal[s], wth[jn][s][0], wth[jn][s][1], wtl[jn][s][0], wtl[jn][s][1]);
        }
        // first ReLU backward: dpre = 1[h1 > 0] dh1 (C-fragment layout) packed as bf16 A fragments
        uint32_t peh[4][2], pel[4][2];                     // [n-tile jn of k][row r0 / r1]
#pragma unroll
        for (int jn = 0; jn < 4; ++jn)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t mk = ah[jn >> 1][(jn & 1) * 2 + r];            // h1 (bf16 pair) of hidden units 8 jn + 2 t + {0, 1}
            const float v0 = (mk & 0xffffu) != 0u ? dh1[jn][2 * r] : 0.f;
            const float v1 = (mk >> 16) != 0u ? dh1[jn][2 * r + 1] : 0.f;
            split2<SPLIT>(v0, v1, peh[jn][r], pel[jn][r]);
          }
        // d bias / d (signed-log distance) = dpre W1: MMA over the hidden axis; lanes t == 0 hold (x, y) of rows r0 / r1
        float dpo[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const uint32_t pah[4] = {peh[2 * s][0], peh[2 * s][1], peh[2 * s + 1][0], peh[2 * s + 1][1]};
          const uint32_t pal[4] = {pel[2 * s][0], pel[2 * s][1], pel[2 * s + 1][0], pel[2 * s + 1][1]};
          mma3<SPLIT>(dpo, pah, pal, w1bh[s][0], w1bh[s][1], w1bl[s][0], w1bl[s][1]);
        }
        const float dax0 = dpo[0], day0 = dpo[1], dax1 = dpo[2], day1 = dpo[3];
        // layer-1 parameter gradients [dW1x | dW1y | db1] = dpre^T [ax ay 1]: MMA over the pair axis (operands transposed in
        // registers); the [16 pairs x 8] right operand lives in C-fragment layout: lanes t == 0 hold (ax, ay), t == 1 (1, 0)
        {
          uint32_t xh0, xl0, xh1, xl1;
          split2<SPLIT>(ax0, ay0, xh0, xl0);
          split2<SPLIT>(ax1, ay1, xh1, xl1);
          if (t != 0) { xh0 = t == 1 ? 0x00003f80u : 0u; xh1 = xh0; xl0 = 0u; xl1 = 0u; }
          const uint32_t bxh0 = movm_trans(xh0), bxh1 = movm_trans(xh1);
          uint32_t bxl0 = 0u, bxl1 = 0u;
          if (SPLIT) { bxl0 = movm_trans(xl0); bxl1 = movm_trans(xl1); }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            uint32_t tah[4], tal[4] = {0u, 0u, 0u, 0u};
            tah[0] = movm_trans(peh[2 * mt][0]); tah[1] = movm_trans(peh[2 * mt + 1][0]);
            tah[2] = movm_trans(peh[2 * mt][1]); tah[3] = movm_trans(peh[2 * mt + 1][1]);
            if (SPLIT) {
              tal[0] = movm_trans(pel[2 * mt][0]); tal[1] = movm_trans(pel[2 * mt + 1][0]);
              tal[2] = movm_trans(pel[2 * mt][1]); tal[3] = movm_trans(pel[2 * mt + 1][1]);
            }
            mma3<SPLIT>(gL1[mt], tah, tal, bxh0, bxh1, bxl0, bxl1);
          }
        }
        // d slog(x)/dx = 1 / (|x| + 1); the key coordinate enters with a minus sign
        // only the lanes t == 0 hold (dax, day): park this key's per-lane partials (2 query rows each) in the warp's staging
        // rows; they are summed over the 8 lanes once per 8-key group
        if (t == 0) {
          const float kgx = -(__fdividef(dax0, fabsf(dx0) + 1.f) + __fdividef(dax1, fabsf(dx1) + 1.f));
          const float kgy = CD == 2 ? -(__fdividef(day0, fabsf(dy0) + 1.f) + __fdividef(day1, fabsf(dy1) + 1.f)) : 0.f;
          *reinterpret_cast<float2*>(&sm.kgs[w][g][2 * u]) = make_float2(kgx, kgy);
        }
      }
      __syncwarp();
      if (lane < 16) {                                     // lane = 2 u + coord
        float r = 0.f;
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) r += sm.kgs[w][gg][lane];
        if (CD == 2 || (lane & 1) == 0) atomicAdd(&sm.dkpos[8 * j + (lane >> 1)][lane & 1], r);
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- phase 3: dQ += dS K (thread-private accumulators parked in shared memory between key tiles)
    {
      float dqacc[8][4];
#pragma unroll
      for (int jd = 0; jd < 8; ++jd)
#pragma unroll
        for (int e = 0; e < 4; ++e) dqacc[jd][e] = sm.dQ[jd * 4 + e][tid];
#pragma unroll
      for (int ks = 0; ks < BK / 16; ++ks) {
        uint32_t ah[4], al[4] = {0u, 0u, 0u, 0u};
        const int c0 = 16 * ks + 2 * t;
        ah[0] = *reinterpret_cast<const uint32_t*>(&sm.dS[0][r0][c0]);     ah[1] = *reinterpret_cast<const uint32_t*>(&sm.dS[0][r1][c0]);
        ah[2] = *reinterpret_cast<const uint32_t*>(&sm.dS[0][r0][c0 + 8]); ah[3] = *reinterpret_cast<const uint32_t*>(&sm.dS[0][r1][c0 + 8]);
        if (SPLIT) {
          al[0] = *reinterpret_cast<const uint32_t*>(&sm.dS[LO][r0][c0]);     al[1] = *reinterpret_cast<const uint32_t*>(&sm.dS[LO][r1][c0]);
          al[2] = *reinterpret_cast<const uint32_t*>(&sm.dS[LO][r0][c0 + 8]); al[3] = *reinterpret_cast<const uint32_t*>(&sm.dS[LO][r1][c0 + 8]);
        }
#pragma unroll
        for (int jd = 0; jd < 8; ++jd) {
          uint32_t bh0, bh1, bl0 = 0u, bl1 = 0u;
          ldmatrix_x2_trans(bh0, bh1, &sm.K[0][16 * ks + (lane & 15)][8 * jd]);
          if (SPLIT) ldmatrix_x2_trans(bl0, bl1, &sm.K[LO][16 * ks + (lane & 15)][8 * jd]);
          mma3<SPLIT>(dqacc[jd], ah, al, bh0, bh1, bl0, bl1);
        }
      }
#pragma unroll
      for (int jd = 0; jd < 8; ++jd)
#pragma unroll
        for (int e = 0; e < 4; ++e) sm.dQ[jd * 4 + e][tid] = dqacc[jd][e];
    }
    // ---- dK += dS^T Q and dV += (P keep)^T dO: warp slice = 16 keys x 32 dims, contraction over the CTA's 64 queries
    {
      const int mt = w & 1, dh = w >> 1;
      float dkacc[4][4], dvacc[4][4];
#pragma unroll
      for (int jd = 0; jd < 4; ++jd)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dkacc[jd][e] = 0.f; dvacc[jd][e] = 0.f; }
      const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int qrow = 16 * ks + rr + 8 * (mi >> 1), kcol = 16 * mt + 8 * (mi & 1);
        uint32_t sh[4], sl[4] = {0u, 0u, 0u, 0u}, ph[4], pl[4] = {0u, 0u, 0u, 0u};
        ldmatrix_x4_trans(sh[0], sh[1], sh[2], sh[3], &sm.dS[0][qrow][kcol]);
        ldmatrix_x4_trans(ph[0], ph[1], ph[2], ph[3], &sm.P[0][qrow][kcol]);
        if (SPLIT) {
          ldmatrix_x4_trans(sl[0], sl[1], sl[2], sl[3], &sm.dS[LO][qrow][kcol]);
          ldmatrix_x4_trans(pl[0], pl[1], pl[2], pl[3], &sm.P[LO][qrow][kcol]);
        }
#pragma unroll
        for (int jd = 0; jd < 4; ++jd) {
          uint32_t qh0, qh1, ql0 = 0u, ql1 = 0u, gh0, gh1, gl0 = 0u, gl1 = 0u;
          ldmatrix_x2_trans(qh0, qh1, &sm.Q[0][16 * ks + (lane & 15)][32 * dh + 8 * jd]);
          ldmatrix_x2_trans(gh0, gh1, &sm.dO[0][16 * ks + (lane & 15)][32 * dh + 8 * jd]);
          if (SPLIT) {
            ldmatrix_x2_trans(ql0, ql1, &sm.Q[LO][16 * ks + (lane & 15)][32 * dh + 8 * jd]);
            ldmatrix_x2_trans(gl0, gl1, &sm.dO[LO][16 * ks + (lane & 15)][32 * dh + 8 * jd]);
          }
          mma3<SPLIT>(dkacc[jd], sh, sl, qh0, qh1, ql0, ql1);
          mma3<SPLIT>(dvacc[jd], ph, pl, gh0, gh1, gl0, gl1);
        }
      }
      const int key0 = k0 + 16 * mt + g, key1 = key0 + 8;
#pragma unroll
      for (int jd = 0; jd < 4; ++jd) {
        const int col = head * AD + 32 * dh + 8 * jd + 2 * t;
        if (key0 < cf.nkv) {
          const int64_t off = ((int64_t)b * cf.nkv + key0) * cf.inner + col;
          atomicAdd(dk + off, dkacc[jd][0]); atomicAdd(dk + off + 1, dkacc[jd][1]);
          atomicAdd(dv + off, dvacc[jd][0]); atomicAdd(dv + off + 1, dvacc[jd][1]);
        }
        if (key1 < cf.nkv) {
          const int64_t off = ((int64_t)b * cf.nkv + key1) * cf.inner + col;
          atomicAdd(dk + off, dkacc[jd][2]); atomicAdd(dk + off + 1, dkacc[jd][3]);
          atomicAdd(dv + off, dvacc[jd][2]); atomicAdd(dv + off + 1, dvacc[jd][3]);
        }
      }
    }
    if (tid < BK * CD) {
      const int key = tid / CD, c = tid % CD;
      if (k0 + key < cf.nkv) atomicAdd(dvs + ((int64_t)bg * cf.nkv + k0 + key) * CD + c, sm.dkpos[key][c]);
    }
  }
  // ---- dQ
  if (ok0) {
    float* row = dq + qoff + (int64_t)(q0 + r0) * cf.inner + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(row + 8 * j) = make_float2(sm.dQ[j * 4][tid], sm.dQ[j * 4 + 1][tid]);
  }
  if (ok1) {
    float* row = dq + qoff + (int64_t)(q0 + r1) * cf.inner + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(row + 8 * j) = make_float2(sm.dQ[j * 4 + 2][tid], sm.dQ[j * 4 + 3][tid]);
  }
  // ---- CPB parameter partials of this CTA: [W1 32*CD][b1 32][W2 1024][b2 32][W3 OUT*32][b3 OUT]
  __syncthreads();
  float* const red = reinterpret_cast<float*>(&sm.P[0][0][0]);
  static_assert(sizeof(BwdSmem<SPLIT>::P) >= NRED * sizeof(float), "the reduction buffer aliases the P tile");
  for (int i = tid; i < NRED; i += MTHREADS) red[i] = 0.f;
  __syncthreads();
  {
    float* rW1 = red; float* rb1 = rW1 + CPB_H * 2; float* rW2 = rb1 + CPB_H; float* rb2 = rW2 + CPB_H * CPB_H;
    float* rw3 = rb2 + CPB_H; float* rb3 = rw3 + CPB_H;
    // layer 1: rows of the gL1 fragments are hidden units (distinct per g), columns (dW1x, dW1y | db1, -)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int kk = 16 * mt + g;
      if (t == 0) {
        atomicAdd(&rW1[kk * 2], gL1[mt][0]); atomicAdd(&rW1[kk * 2 + 1], gL1[mt][1]);
        atomicAdd(&rW1[(kk + 8) * 2], gL1[mt][2]); atomicAdd(&rW1[(kk + 8) * 2 + 1], gL1[mt][3]);
      } else if (t == 1) {
        atomicAdd(&rb1[kk], gL1[mt][0]); atomicAdd(&rb1[kk + 8], gL1[mt][2]);
      }
    }
    // values indexed by t only: add the 8 lanes (g) that share t first
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float a = gb2[j][e], x = gw3[j][e];
#pragma unroll
        for (int m = 4; m < 32; m <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, m); x += __shfl_xor_sync(0xffffffffu, x, m); }
        if (g == 0) { atomicAdd(&rb2[8 * j + 2 * t + e], a); atomicAdd(&rw3[8 * j + 2 * t + e], x); }
      }
    gb3 = warp_sum(gb3);
    if (lane == 0) atomicAdd(rb3, gb3);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {
        const int c = 16 * mt + g, kk = 8 * jn + 2 * t;
        atomicAdd(&rW2[c * CPB_H + kk], gW2[mt][jn][0]); atomicAdd(&rW2[c * CPB_H + kk + 1], gW2[mt][jn][1]);
        atomicAdd(&rW2[(c + 8) * CPB_H + kk], gW2[mt][jn][2]); atomicAdd(&rW2[(c + 8) * CPB_H + kk + 1], gW2[mt][jn][3]);
      }
  }
  __syncthreads();
  {
    const int np = CPB_H * CD + CPB_H + CPB_H * CPB_H + CPB_H + cf.OUT * CPB_H + cf.OUT;
    float* prow = part + ((int64_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * np;
    const float* rW1 = red; const float* rb1 = rW1 + CPB_H * 2; const float* rW2 = rb1 + CPB_H;
    const float* rb2 = rW2 + CPB_H * CPB_H; const float* rw3 = rb2 + CPB_H; const float* rb3 = rw3 + CPB_H;
    for (int i = tid; i < np; i += MTHREADS) {
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


// ------------------------------------------------------------------------------------------ lean bf16 kernels (round 2)
// Same fused attention as above for the plain-bf16 mode, with the per-pair CUDA-core work cut from ~140 to ~70 warp
// instructions per 16-pair tile (forward):
//   * signed-log distances come from per-key-tile tables in shared memory (the query lattice is regular: x distance per
//     (key, CTA row), y distance per (key, lattice row)): 4 LDS instead of 4 subtract + log + copysign chains per thread
//   * ReLU + bf16 packing is one cvt.rn.relu.bf16x2.f32 per two values
//   * layer 3 (32 -> 1) runs as two more MMAs whose B operand holds w3 in the column of the current key, so the bias lands
//     directly in the C fragment of the 8-key logit group (no thread-local dot product, no quad shuffle)
// tools/ubench_tmem.cu shows why this stays on warp-level MMAs: a tcgen05.mma costs >= 45 cycles per instruction whatever its
// N (measured), so the [128 pairs x 32] x [32 x 32] layer as 128 x 32 x 16 tcgen05 MMAs needs >= 90 cycles per 128 pairs for
// layer 2 alone (+ TMEM round trips), against 64 cycles for the same work as 16 m16n8k16 warp MMAs per SM sub-partition.
__device__ __forceinline__ uint32_t relu_pack(float lo, float hi) {      // max(x, 0) -> bf16x2, low half <- lo
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

template <int CD, int NK>
struct LeanTab {
  float SX[NK][MQ];                  // slog(qx(CTA row) - kx(key))            [key in tile][row]
  float SY[CD == 2 ? MQ : 1][NK];    // slog(qy(lattice-row slot) - ky(key))   [slot][key in tile]
  float qx[MQ], qy[MQ];
  float kpos[NK][2];
};
// fills the tables for the keys [0, nk) of the current tile (kpos already in shared memory)
template <int CD, int NK>
__device__ __forceinline__ void lean_tables(LeanTab<CD, NK>& tb, int ny) {
  for (int e = threadIdx.x; e < NK * MQ; e += MTHREADS) {
    const int key = e / MQ, r = e % MQ;
    tb.SX[key][r] = slog_fast(tb.qx[r] - tb.kpos[key][0]);
  }
  if (CD == 2)
    for (int e = threadIdx.x; e < NK * ny; e += MTHREADS) {
      const int slot = e / NK, key = e % NK;
      tb.SY[slot][key] = slog_fast(tb.qy[slot] - tb.kpos[key][1]);
    }
}
// query positions of the CTA rows (x per row; y per lattice-row slot).  Returns the number of y slots.
template <int CD, int NK>
__device__ __forceinline__ int lean_query_tables(LeanTab<CD, NK>& tb, const AttnCfg& cf, int q0) {
  const int last = min(q0 + MQ - 1, cf.n - 1);
  const int y0 = CD == 2 ? q0 / cf.W : 0;
  const int ny = CD == 2 ? last / cf.W - y0 + 1 : 1;
  if (threadIdx.x < MQ) {
    float px, py;
    query_pos(cf, min(q0 + (int)threadIdx.x, cf.n - 1), px, py);
    tb.qx[threadIdx.x] = px;
    if (CD == 2 && (int)threadIdx.x < ny) {
      query_pos(cf, min((y0 + (int)threadIdx.x) * cf.W, cf.n - 1), px, py);
      tb.qy[threadIdx.x] = py;
    }
  }
  return ny;
}

template <int CD>
struct LeanFwdSmem {
  __nv_bfloat16 Q[MQ][MLD], K[64][MLD], V[2][64][MLD];
  LeanTab<CD, 64> tb;
};

// grid (ceil(n / 64), heads, B), 128 threads; same contract as deform_attn_fwd_mma_kernel<CD, false>
template <int CD, int MINB>
__global__ void __launch_bounds__(MTHREADS, MINB) deform_attn_fwd_lean_kernel(const AttnCfg cf, const float* __restrict__ q,
                                                                            const float* __restrict__ k, const float* __restrict__ v,
                                                                            const float* __restrict__ vs, const float* __restrict__ cpb,
                                                                            float* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LeanFwdSmem<CD>& sm = *reinterpret_cast<LeanFwdSmem<CD>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * MQ;
  const int grp = head / cf.OUT, o = head % cf.OUT;
  const int bg = b * cf.G + grp;
  CpbRegs<CD, false> cr;
  cr.load(cpb, cf.OUT, o, lane);
  uint32_t w3f[2][2];                                   // B fragments of w3 (k = hidden unit), nonzero column chosen per key
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const float* W3 = cpb + CPB_H * CD + CPB_H + CPB_H * CPB_H + CPB_H + o * CPB_H + 16 * s + 2 * t;
    w3f[s][0] = pack_bf16(W3[0], W3[1]);
    w3f[s][1] = pack_bf16(W3[8], W3[9]);
  }
  load_tile<false, MQ>(sm.Q, nullptr, q + (int64_t)b * cf.n * cf.inner + head * AD, cf.inner, q0, cf.n);
  const int ny = lean_query_tables<CD, 64>(sm.tb, cf, q0);
  const int r0 = w * 16 + g, r1 = r0 + 8;
  const int ys0 = CD == 2 ? min(q0 + r0, cf.n - 1) / cf.W - q0 / cf.W : 0;
  const int ys1 = CD == 2 ? min(q0 + r1, cf.n - 1) / cf.W - q0 / cf.W : 0;
  float oacc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[j][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float keep_scale = cf.p_drop > 0.f ? 1.f / (1.f - cf.p_drop) : 1.f;
  const uint32_t thresh = (uint32_t)(cf.p_drop * 4294967296.0);
  const unsigned long long hrow0 = ((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + r0)) * cf.nkv;
  const unsigned long long hrow1 = ((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + r1)) * cf.nkv;
  const float* kb = k + (int64_t)b * cf.nkv * cf.inner + head * AD;
  const float* vb = v + (int64_t)b * cf.nkv * cf.inner + head * AD;

  for (int k0 = 0; k0 < cf.nkv; k0 += 64) {
    __syncthreads();
    load_tile<false, 64>(sm.K, nullptr, kb, cf.inner, k0, cf.nkv);
    load_tile<true, 64>(sm.V[0], sm.V[1], vb, cf.inner, k0, cf.nkv);
    if (tid < 64) {
      const bool ok = k0 + tid < cf.nkv;
      sm.tb.kpos[tid][0] = ok ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD] : 0.f;
      sm.tb.kpos[tid][1] = (ok && CD == 2) ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD + 1] : 0.f;
    }
    __syncthreads();
    lean_tables<CD, 64>(sm.tb, ny);
    // ---- S = b3 + Q K^T
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = cr.b3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4];
      ah[0] = *reinterpret_cast<const uint32_t*>(&sm.Q[r0][16 * ks + 2 * t]);
      ah[1] = *reinterpret_cast<const uint32_t*>(&sm.Q[r1][16 * ks + 2 * t]);
      ah[2] = *reinterpret_cast<const uint32_t*>(&sm.Q[r0][16 * ks + 2 * t + 8]);
      ah[3] = *reinterpret_cast<const uint32_t*>(&sm.Q[r1][16 * ks + 2 * t + 8]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&sm.K[8 * j + g][16 * ks + 2 * t]);
        const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&sm.K[8 * j + g][16 * ks + 2 * t + 8]);
        mma16816(s[j], ah, bh0, bh1);
      }
    }
    __syncthreads();                                    // tables complete
    // ---- + continuous position bias
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      float bias[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int key = 8 * j + u;
        const float ax0 = sm.tb.SX[key][r0], ax1 = sm.tb.SX[key][r1];
        const float ay0 = CD == 2 ? sm.tb.SY[ys0][key] : 0.f, ay1 = CD == 2 ? sm.tb.SY[ys1][key] : 0.f;
        uint32_t ah[2][4];
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          float h0[4], h1[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float u0 = fmaf(cr.w1x[s2][e], ax0, cr.b1[s2][e]), u1 = fmaf(cr.w1x[s2][e], ax1, cr.b1[s2][e]);
            if (CD == 2) { u0 = fmaf(cr.w1y[s2][e], ay0, u0); u1 = fmaf(cr.w1y[s2][e], ay1, u1); }
            h0[e] = u0; h1[e] = u1;
          }
          ah[s2][0] = relu_pack(h0[0], h0[1]); ah[s2][1] = relu_pack(h1[0], h1[1]);
          ah[s2][2] = relu_pack(h0[2], h0[3]); ah[s2][3] = relu_pack(h1[2], h1[3]);
        }
        float acc[4][4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          mma16816_c(acc[jj], ah[0], cr.wh[jj][0][0], cr.wh[jj][0][1], cr.b2[jj][0], cr.b2[jj][1], cr.b2[jj][0], cr.b2[jj][1]);
          mma16816(acc[jj], ah[1], cr.wh[jj][1][0], cr.wh[jj][1][1]);
        }
        const bool mine = g == u;                        // column of this key in the 8-key C fragment
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          uint32_t hh[4];
          hh[0] = relu_pack(acc[2 * s2][0], acc[2 * s2][1]);         hh[1] = relu_pack(acc[2 * s2][2], acc[2 * s2][3]);
          hh[2] = relu_pack(acc[2 * s2 + 1][0], acc[2 * s2 + 1][1]); hh[3] = relu_pack(acc[2 * s2 + 1][2], acc[2 * s2 + 1][3]);
          mma16816(bias, hh, mine ? w3f[s2][0] : 0u, mine ? w3f[s2][1] : 0u);
        }
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        if (jj == j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) s[jj][e] += bias[e];
        }
    }
    // ---- online softmax on the fragments
    float mt0 = -INFINITY, mt1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (k0 + 8 * j + 2 * t + e >= cf.nkv) { s[j][e] = -INFINITY; s[j][2 + e] = -INFINITY; }
        mt0 = fmaxf(mt0, s[j][e]); mt1 = fmaxf(mt1, s[j][2 + e]);
      }
    mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 1)); mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 1));
    mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 2)); mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 2));
    const float mn0 = fmaxf(m0, mt0), mn1 = fmaxf(m1, mt1);
    const float al0 = m0 == -INFINITY ? 0.f : __expf(m0 - mn0), al1 = m1 == -INFINITY ? 0.f : __expf(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= al0; l1 *= al1;
#pragma unroll
    for (int j = 0; j < 8; ++j) { oacc[j][0] *= al0; oacc[j][1] *= al0; oacc[j][2] *= al1; oacc[j][3] *= al1; }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float p0 = __expf(s[j][e] - mn0), p1 = __expf(s[j][2 + e] - mn1);
        l0 += p0; l1 += p1;
        if (cf.p_drop > 0.f) {
          const unsigned long long key = (unsigned long long)(k0 + 8 * j + 2 * t + e);
          p0 = pair_hash(salted(cf.seed, cf.salt), hrow0 + key) >= thresh ? p0 * keep_scale : 0.f;
          p1 = pair_hash(salted(cf.seed, cf.salt), hrow1 + key) >= thresh ? p1 * keep_scale : 0.f;
        }
        s[j][e] = p0; s[j][2 + e] = p1;
      }
    // ---- O += P V (P as bf16 hi + lo: the backward's delta = dO . O relies on an fp32-grade O)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4], al[4];
      split2<true>(s[2 * ks][0], s[2 * ks][1], ah[0], al[0]);
      split2<true>(s[2 * ks][2], s[2 * ks][3], ah[1], al[1]);
      split2<true>(s[2 * ks + 1][0], s[2 * ks + 1][1], ah[2], al[2]);
      split2<true>(s[2 * ks + 1][2], s[2 * ks + 1][3], ah[3], al[3]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t bh0, bh1, bl0, bl1;
        ldmatrix_x2_trans(bh0, bh1, &sm.V[0][16 * ks + (lane & 15)][8 * j]);
        ldmatrix_x2_trans(bl0, bl1, &sm.V[1][16 * ks + (lane & 15)][8 * j]);
        mma3<true>(oacc[j], ah, al, bh0, bh1, bl0, bl1);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  if (q0 + r0 < cf.n) {
    float* orow = out + ((int64_t)b * cf.n + q0 + r0) * cf.inner + head * AD + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(orow + 8 * j) = make_float2(oacc[j][0] * i0, oacc[j][1] * i0);
    if (t == 0) lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r0] = m0 + logf(l0);
  }
  if (q0 + r1 < cf.n) {
    float* orow = out + ((int64_t)b * cf.n + q0 + r1) * cf.inner + head * AD + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(orow + 8 * j) = make_float2(oacc[j][2] * i1, oacc[j][3] * i1);
    if (t == 0) lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r1] = m1 + logf(l1);
  }
}


// ---- 1-D bias MLP as an exact piecewise-linear table (csrc/cpb_table.cu; specification oracle/pwl.py): for one coordinate the
// MLP Linear(1,32) ReLU Linear(32,32) ReLU Linear(32,o) is a continuous piecewise-linear function of the signed-log distance
// with <= 1089 segments (65-70 for typical parameters), so the per-pair bias is slope[s] p + icpt[s]: no MMA per pair.
// Along one query row the distance moves by ~2 / n_kv per key, far less than a segment, so a (row, key tile) range contains
// zero or one breakpoint almost always: its two lines (slope, intercept) x 2 and the breakpoint live in REGISTERS for the
// tile and the per-pair bias is compare + select + FMA, with no table access.  Tiles whose range holds two or more
// breakpoints take a (warp-uniform) slow path that searches the shared-memory table per pair.
// Everything is kept in the base-2 domain: p2 = sign log2(|d| + 1) = p log2(e), logits scaled by log2(e) (the Q tile is
// scaled when it is converted to bf16), so that slope p + icpt becomes slope p2 + icpt log2(e) with an UNCHANGED slope, the
// distance costs one MUFU.LG2 and the softmax one MUFU.EX2 without multiplies.
constexpr int PWL_MAXSEG = 1089;
constexpr int PWL_MAXTILES = 512;        // 32-key tiles whose ranges are staged in shared memory (n_kv <= 16 384; beyond: global reads)
constexpr float PWL_LOG2E = 1.4426950408889634f, PWL_LN2 = 0.6931471805599453f;
struct PwlTable {
  float B[PWL_MAXSEG + 1];               // segment s covers [B[s], B[s + 1]) in p2 units; B[0] = -inf, B[nseg] = +inf
  const float2* co;                      // (slope, intercept) of this head's output: global memory, row stride `cs` float2
  int cs;
  __device__ __forceinline__ float2 line(int sg) const {          // (slope, intercept log2(e))
    const float2 c = __ldg(co + (int64_t)sg * cs);
    return make_float2(c.x, c.y * PWL_LOG2E);
  }
};
// segment of p, searched from the guess s (any p ends in the right segment)
__device__ __forceinline__ int pwl_seek(const PwlTable& tb, float p, int s) {
  while (p < tb.B[s]) --s;
  while (p >= tb.B[s + 1]) ++s;
  return s;
}
__device__ __forceinline__ void pwl_load(PwlTable& tb, const float* __restrict__ brk, const float* __restrict__ coef, int nb, int OUT,
                                         int o, int tid, int nthreads) {
  for (int i = tid; i <= nb + 1; i += nthreads) tb.B[i] = i == 0 ? -INFINITY : (i <= nb ? brk[i - 1] * PWL_LOG2E : INFINITY);
  if (tid == 0) { tb.co = reinterpret_cast<const float2*>(coef) + o; tb.cs = OUT; }
}
__device__ __forceinline__ float slog2_fast(float x) { return copysignf(__log2f(fabsf(x) + 1.f), x); }
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
// The two lines of one query row over one key tile: bias2(p) = p < b ? sl p + il : sh p + ih
struct PwlRow {
  float b, sl, il, sh, ih;
  int seg;                               // segment of the lower line (search guess of the next tile)
};
// kmin / kmax = range of the tile's key positions; returns true when the row's range holds more than one breakpoint
__device__ __forceinline__ bool pwl_row_setup(const PwlTable& tb, int ns, float qx, float kmin, float kmax, PwlRow& rw) {
  const float plo = slog2_fast(qx - kmax), phi = slog2_fast(qx - kmin);
  // common case: the lower end is still in the segment of the previous tile - three independent loads, no search
  const float b0 = tb.B[rw.seg], b1 = tb.B[rw.seg + 1], b2 = tb.B[min(rw.seg + 2, ns)];
  if (plo >= b0 && plo < b1 && rw.b == b1) return phi >= b2;
  const int sL = pwl_seek(tb, plo, rw.seg);
  const float2 cl = tb.line(sL), ch = tb.line(min(sL + 1, ns - 1));
  rw.seg = sL; rw.b = tb.B[sL + 1];
  rw.sl = cl.x; rw.il = cl.y; rw.sh = ch.x; rw.ih = ch.y;
  return phi >= tb.B[min(sL + 2, ns)];
}
__device__ __forceinline__ float pwl_bias(const PwlRow& rw, float p) {
  const bool lo = p < rw.b;
  return fmaf(p, lo ? rw.sl : rw.sh, lo ? rw.il : rw.ih);
}
// 16-byte global -> shared copies without register staging (rows past the end are zero-filled: src-size 0)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(d), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = ok ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" :: "r"(d), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }
// one stage of the key ring: NK keys x 64 dims of the bf16 K and V planes + the keys' positions
// Rows are 128 bytes, unpadded; the 16-byte chunk c of row r sits at chunk c ^ (r & 7), which keeps the fragment loads
// (8 rows x 4 words) and ldmatrix (8 rows x 16 bytes) free of bank conflicts.
template <int NK>
struct PwlStage {
  __nv_bfloat16 K[NK][AD], V[NK][AD];
  float kx[NK];
};
__device__ __forceinline__ const __nv_bfloat16* pwl_sw(const __nv_bfloat16 (*T)[AD], int row, int col) {
  return &T[row][((((col >> 3) ^ row) & 7) << 3) | (col & 7)];
}
template <int NK>
__device__ __forceinline__ void pwl_stage_load(PwlStage<NK>& st, const __nv_bfloat16* __restrict__ kb, const __nv_bfloat16* __restrict__ vb,
                                               const float* __restrict__ vsb, int64_t row_stride, int k0, int nkv, int tid) {
  for (int e = tid; e < NK * 8; e += MTHREADS) {
    const int r = e >> 3, c8 = (e & 7) * 8;
    const bool ok = k0 + r < nkv;
    const int64_t off = (int64_t)(ok ? k0 + r : 0) * row_stride + c8;
    const int cs = (((c8 >> 3) ^ r) & 7) << 3;
    cp_async16(&st.K[r][cs], kb + off, ok);
    cp_async16(&st.V[r][cs], vb + off, ok);
  }
  if (tid < NK) { const bool ok = k0 + tid < nkv; cp_async4(&st.kx[tid], vsb + (ok ? k0 + tid : 0), ok); }
}
// range (min, max) of the key positions of every 32-key tile, written once per call by a small kernel: the attention
// kernels read their tile's range one tile ahead instead of reducing it inside the key loop (a 10-step dependent shuffle chain)
__global__ void __launch_bounds__(256) pwl_key_range_kernel(const float* __restrict__ vs, int nkv, int nt, int64_t total, float2* __restrict__ kr) {
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= total) return;
  const int64_t bg = wid / nt;
  const int k = (int)(wid % nt) * 32 + lane;
  const float x = k < nkv ? vs[bg * nkv + k] : 0.f;
  float mn = k < nkv ? x : INFINITY, mx = k < nkv ? x : -INFINITY;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, m));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
  }
  if (lane == 0) kr[wid] = make_float2(mn, mx);
}
constexpr int PWL_FK = 64;                // keys per forward tile
struct PwlFwdSmem {
  __nv_bfloat16 Q[MQ][MLD];
  PwlStage<PWL_FK> st[3];
  float2 krng[PWL_MAXTILES];              // key-position range of every 32-key tile of this (b, g) row
  PwlTable pw;
};
// K and V arrive as bf16 planes (the mode's operand rounding, done once by the caller instead of once per query tile);
// tiles move through a three-stage cp.async ring with one barrier per tile.  P is rounded to bf16 for P V (as in every
// bf16 flash attention; the row sum l is accumulated from the unrounded values).
template <bool DROP>
__global__ void __launch_bounds__(MTHREADS, 3) deform_attn_fwd_pwl_kernel(const AttnCfg cf, const float* __restrict__ q,
                                                                            const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                                                                            const float* __restrict__ vs, const float2* __restrict__ kr,
                                                                            const float* __restrict__ brk,
                                                                            const float* __restrict__ coef, const int* __restrict__ nseg,
                                                                            float* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PwlFwdSmem& sm = *reinterpret_cast<PwlFwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * MQ;
  const int grp = head / cf.OUT, o = head % cf.OUT;
  const int bg = b * cf.G + grp;
  const int ns = nseg[0];
  const __nv_bfloat16* kb = k + (int64_t)b * cf.nkv * cf.inner + head * AD;
  const __nv_bfloat16* vb = v + (int64_t)b * cf.nkv * cf.inner + head * AD;
  const float* vsb = vs + (int64_t)bg * cf.nkv;
  const int ntiles = (cf.nkv + PWL_FK - 1) / PWL_FK;
  pwl_stage_load(sm.st[0], kb, vb, vsb, cf.inner, 0, cf.nkv, tid);
  cp_async_commit();
  if (ntiles > 1) pwl_stage_load(sm.st[1], kb, vb, vsb, cf.inner, PWL_FK, cf.nkv, tid);
  cp_async_commit();
  pwl_load(sm.pw, brk, coef, ns - 1, cf.OUT, o, tid, MTHREADS);
  {                                                     // Q tile, scaled by log2(e)
    const float* src = q + (int64_t)b * cf.n * cf.inner + head * AD;
    for (int e = tid; e < MQ * 16; e += MTHREADS) {
      const int r = e >> 4, c4 = (e & 15) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q0 + r < cf.n) x = *reinterpret_cast<const float4*>(src + (int64_t)(q0 + r) * cf.inner + c4);
      *reinterpret_cast<uint2*>(&sm.Q[r][c4]) = make_uint2(pack_bf16(x.x * PWL_LOG2E, x.y * PWL_LOG2E), pack_bf16(x.z * PWL_LOG2E, x.w * PWL_LOG2E));
    }
  }
  const int r0 = w * 16 + g, r1 = r0 + 8;
  float qx0, qx1, qdum;
  query_pos(cf, min(q0 + r0, cf.n - 1), qx0, qdum);
  query_pos(cf, min(q0 + r1, cf.n - 1), qx1, qdum);
  PwlRow rw0, rw1;
  rw0.seg = rw1.seg = ns >> 1;
  rw0.b = rw1.b = __int_as_float(0x7fc00000);          // NaN: forces the first line load
  float oacc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[j][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;          // base-2 running max, running sum
  const float keep_scale = DROP ? 1.f / (1.f - cf.p_drop) : 1.f;
  const uint32_t thresh = (uint32_t)(cf.p_drop * 4294967296.0);
  const unsigned long long hrow0 = ((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + r0)) * cf.nkv;
  const unsigned long long hrow1 = ((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + r1)) * cf.nkv;

  const int nt32 = (cf.nkv + 31) >> 5;
  const float2* krb = kr + (int64_t)bg * nt32;
  const bool rng_smem = nt32 <= PWL_MAXTILES;
  if (rng_smem)
    for (int i = tid; i < nt32; i += MTHREADS) sm.krng[i] = krb[i];

  for (int it = 0; it < ntiles; ++it) {
    const int k0 = it * PWL_FK;
    cp_async_wait<1>();                                 // my copies of tile `it` have landed
    __syncthreads();                                    // ... everyone's have, and everyone is done with tile it - 1
    if (it + 2 < ntiles) pwl_stage_load(sm.st[(it + 2) % 3], kb, vb, vsb, cf.inner, k0 + 2 * PWL_FK, cf.nkv, tid);
    cp_async_commit();
    const PwlStage<PWL_FK>& st = sm.st[it % 3];
    const int ia = 2 * it, ib = min(2 * it + 1, nt32 - 1);      // the two 32-key halves of the tile
    const float2 ra = rng_smem ? sm.krng[ia] : krb[ia], rb = rng_smem ? sm.krng[ib] : krb[ib];
    const float kmin = fminf(ra.x, rb.x), kmax = fmaxf(ra.y, rb.y);
    const bool many0 = pwl_row_setup(sm.pw, ns, qx0, kmin, kmax, rw0);
    const bool many1 = pwl_row_setup(sm.pw, ns, qx1, kmin, kmax, rw1);
    const bool slow = __any_sync(0xffffffffu, many0 || many1);
    // ---- S2 = log2(e) Q K^T
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
    // (fragments by ldmatrix: A = 16 rows x 16 dims of Q, B = two 8-key blocks x 16 dims of K per instruction)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4];
      ldmatrix_x4(ah[0], ah[1], ah[2], ah[3], &sm.Q[w * 16 + (lane & 15)][16 * ks + 8 * (lane >> 4)]);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        uint32_t b00, b01, b10, b11;
        ldmatrix_x4(b00, b01, b10, b11, pwl_sw(st.K, 8 * j + (lane & 7) + 8 * (lane >> 4), 16 * ks + 8 * ((lane >> 3) & 1)));
        mma16816(s[j], ah, b00, b01);
        mma16816(s[j + 1], ah, b10, b11);
      }
    }
    // ---- + continuous position bias
    if (!slow) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 kx = *reinterpret_cast<const float2*>(&st.kx[8 * j + 2 * t]);
        s[j][0] += pwl_bias(rw0, slog2_fast(qx0 - kx.x)); s[j][1] += pwl_bias(rw0, slog2_fast(qx0 - kx.y));
        s[j][2] += pwl_bias(rw1, slog2_fast(qx1 - kx.x)); s[j][3] += pwl_bias(rw1, slog2_fast(qx1 - kx.y));
      }
    } else {
      for (int j = 0; j < 8; ++j) {
        const float2 kx = *reinterpret_cast<const float2*>(&st.kx[8 * j + 2 * t]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p = slog2_fast(((e & 2) ? qx1 : qx0) - ((e & 1) ? kx.y : kx.x));
          const float2 c = sm.pw.line(pwl_seek(sm.pw, p, (e & 2) ? rw1.seg : rw0.seg));
          s[j][e] += fmaf(c.x, p, c.y);
        }
      }
    }
    if (k0 + PWL_FK > cf.nkv) {                         // ragged last tile
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (k0 + 8 * j + 2 * t + e >= cf.nkv) { s[j][e] = -INFINITY; s[j][2 + e] = -INFINITY; }
    }
    // ---- online softmax on the fragments (base 2)
    float mt0 = -INFINITY, mt1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mt0 = fmaxf(mt0, fmaxf(s[j][0], s[j][1])); mt1 = fmaxf(mt1, fmaxf(s[j][2], s[j][3]));
    }
    mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 1)); mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 1));
    mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 2)); mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 2));
    const float mn0 = fmaxf(m0, mt0), mn1 = fmaxf(m1, mt1);
    const float al0 = m0 == -INFINITY ? 0.f : ex2_fast(m0 - mn0), al1 = m1 == -INFINITY ? 0.f : ex2_fast(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= al0; l1 *= al1;
#pragma unroll
    for (int j = 0; j < 8; ++j) { oacc[j][0] *= al0; oacc[j][1] *= al0; oacc[j][2] *= al1; oacc[j][3] *= al1; }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float p0 = ex2_fast(s[j][e] - mn0), p1 = ex2_fast(s[j][2 + e] - mn1);
        l0 += p0; l1 += p1;
        if (DROP) {
          const unsigned long long key = (unsigned long long)(k0 + 8 * j + 2 * t + e);
          p0 = pair_hash(salted(cf.seed, cf.salt), hrow0 + key) >= thresh ? p0 * keep_scale : 0.f;
          p1 = pair_hash(salted(cf.seed, cf.salt), hrow1 + key) >= thresh ? p1 * keep_scale : 0.f;
        }
        s[j][e] = p0; s[j][2 + e] = p1;
      }
    // ---- O += P V
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
        ldmatrix_x4_trans(b00, b01, b10, b11, pwl_sw(st.V, 16 * ks + (lane & 15), 8 * j + 8 * (lane >> 4)));
        mma16816(oacc[j], ah, b00, b01);
        mma16816(oacc[j + 1], ah, b10, b11);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  if (q0 + r0 < cf.n) {
    float* orow = out + ((int64_t)b * cf.n + q0 + r0) * cf.inner + head * AD + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(orow + 8 * j) = make_float2(oacc[j][0] * i0, oacc[j][1] * i0);
    if (t == 0) lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r0] = (m0 + log2f(l0)) * PWL_LN2;
  }
  if (q0 + r1 < cf.n) {
    float* orow = out + ((int64_t)b * cf.n + q0 + r1) * cf.inner + head * AD + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(orow + 8 * j) = make_float2(oacc[j][2] * i1, oacc[j][3] * i1);
    if (t == 0) lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r1] = (m1 + log2f(l1)) * PWL_LN2;
  }
}


// ---- lean backward (plain bf16 operands).  Same phases and outputs as deform_attn_bwd_mma_kernel<CD, false>; per key the
// CUDA-core work drops from ~470 to ~250 warp instructions:
//   * distances from the shared-memory tables (lean_tables); d slog / dx = exp(-|slog|)
//   * layer 3 on the MMA pipe (w3 in column 0 of the B operand; lanes t == 0 hold the two biases, one shuffle each)
//   * second ReLU: h2 is kept only as the packed bf16 A fragments; dh2 = mask(h2 > 0) & (dS w3) in packed bf16x2 arithmetic
//   * db2 = dh2^T 1 and dw3 = h2^T dS ride on the MMA pipe as extra columns of the [16 pairs x 8] right operand
//     X = [ax | ay | 1 | dS | 0...] that the layer-1 gradients already use (dW1 | db1 = dpre^T X)
//   * first ReLU backward in packed form as well (mask from the bf16 h1 fragments)
__device__ __forceinline__ uint32_t nz_mask_bf16x2(uint32_t a) {       // 0xffff per half that is > 0 (a holds relu outputs)
  uint32_t d;
  asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(0u));
  return d;
}
__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

template <int CD>
struct LeanBwdSmem {
  __nv_bfloat16 Q[MQ][MLD], dO[2][MQ][MLD];
  __nv_bfloat16 K[BK][MLD], V[2][BK][MLD];
  __nv_bfloat16 P[MQ][PLD], dS[MQ][PLD];     // `red` aliases P + dS at the end of the kernel
  float dQ[32][MTHREADS];
  float stg[4][2][16][9];
  float kgs[4][8][16];
  float c1[4][24];
  float c2[4][8];                    // b2[4][2] by t
  float dkpos[BK][2];
  LeanTab<CD, BK> tb;
};

template <int CD, int UNR>
__global__ void __launch_bounds__(MTHREADS, 2) deform_attn_bwd_lean_kernel(
    const AttnCfg cf, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const float* __restrict__ vs, const float* __restrict__ cpb, const float* __restrict__ out,
    const float* __restrict__ lse, const float* __restrict__ dout, float* __restrict__ dq, float* __restrict__ dk,
    float* __restrict__ dv, float* __restrict__ dvs, float* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LeanBwdSmem<CD>& sm = *reinterpret_cast<LeanBwdSmem<CD>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * MQ;
  const int grp = head / cf.OUT, o = head % cf.OUT;
  const int bg = b * cf.G + grp;
  uint32_t wh[4][2][2];         // layer 2:  B[k][n = c] = W2[c][k]                  [n-tile j][k-step s][reg]
  uint32_t wth[4][2][2];        // dh1 = dh2 W2:  B[k' = c][n = k] = W2[c][k]        [n-tile jn][k-step s'][reg]
  uint32_t w1b[2][2];           // d pos = dpre W1:  B[k][n < CD] = W1[k][n]         [k-step s][reg]
  uint32_t w3f[2][2];           // w3 pairs of hidden units (16 s + 2 t + {0, 1}) and (+ 8): A-column / B-row order
  float b3;
  {
    const float* W1 = cpb; const float* B1 = W1 + CPB_H * CD; const float* W2 = B1 + CPB_H;
    const float* B2 = W2 + CPB_H * CPB_H; const float* W3 = B2 + CPB_H; const float* B3 = W3 + cf.OUT * CPB_H;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float* row = W2 + (8 * j + g) * CPB_H + 16 * s + 2 * t;
        wh[j][s][0] = pack_bf16(row[0], row[1]);
        wh[j][s][1] = pack_bf16(row[8], row[9]);
        const float* col = W2 + (16 * s + 2 * t) * CPB_H + 8 * j + g;
        wth[j][s][0] = pack_bf16(col[0], col[CPB_H]);
        wth[j][s][1] = pack_bf16(col[8 * CPB_H], col[9 * CPB_H]);
      }
    b3 = B3[o];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int k0_ = 16 * s + 2 * t;
      const bool on = g < CD;
      w1b[s][0] = pack_bf16(on ? W1[k0_ * CD + g] : 0.f, on ? W1[(k0_ + 1) * CD + g] : 0.f);
      w1b[s][1] = pack_bf16(on ? W1[(k0_ + 8) * CD + g] : 0.f, on ? W1[(k0_ + 9) * CD + g] : 0.f);
      w3f[s][0] = pack_bf16(W3[o * CPB_H + k0_], W3[o * CPB_H + k0_ + 1]);
      w3f[s][1] = pack_bf16(W3[o * CPB_H + k0_ + 8], W3[o * CPB_H + k0_ + 9]);
    }
    if (tid < 4 * 24) {
      const int tt = tid / 24, i = tid % 24, which = i / 8, s = (i % 8) / 4, e = i % 4;
      const int kk = 16 * s + 2 * tt + (e & 1) + 8 * (e >> 1);
      sm.c1[tt][i] = which == 0 ? W1[kk * CD] : (which == 1 ? (CD == 2 ? W1[kk * CD + 1] : 0.f) : B1[kk]);
    }
    if (tid < 4 * 8) {
      const int tt = tid / 8, i = tid % 8, j = i / 2, e = i % 2;
      sm.c2[tt][i] = B2[8 * j + 2 * tt + e];
    }
  }
  const int64_t qoff = (int64_t)b * cf.n * cf.inner + head * AD;
  load_tile<false, MQ>(sm.Q, nullptr, q + qoff, cf.inner, q0, cf.n);
  load_tile<true, MQ>(sm.dO[0], sm.dO[1], dout + qoff, cf.inner, q0, cf.n);
#pragma unroll
  for (int i = 0; i < 32; ++i) sm.dQ[i][tid] = 0.f;
  const int ny = lean_query_tables<CD, BK>(sm.tb, cf, q0);
  const int r0 = w * 16 + g, r1 = r0 + 8;
  const bool ok0 = q0 + r0 < cf.n, ok1 = q0 + r1 < cf.n;
  const int ys0 = CD == 2 ? min(q0 + r0, cf.n - 1) / cf.W - q0 / cf.W : 0;
  const int ys1 = CD == 2 ? min(q0 + r1, cf.n - 1) / cf.W - q0 / cf.W : 0;
  float delta0 = 0.f, delta1 = 0.f, lse0 = INFINITY, lse1 = INFINITY;
  {
    if (ok0) {
      const float* orow = out + qoff + (int64_t)(q0 + r0) * cf.inner + 16 * t;
      const float* grow = dout + qoff + (int64_t)(q0 + r0) * cf.inner + 16 * t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(orow + 4 * i), c = *reinterpret_cast<const float4*>(grow + 4 * i);
        delta0 = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, fmaf(a.w, c.w, delta0))));
      }
      lse0 = lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r0];
    }
    if (ok1) {
      const float* orow = out + qoff + (int64_t)(q0 + r1) * cf.inner + 16 * t;
      const float* grow = dout + qoff + (int64_t)(q0 + r1) * cf.inner + 16 * t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(orow + 4 * i), c = *reinterpret_cast<const float4*>(grow + 4 * i);
        delta1 = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, fmaf(a.w, c.w, delta1))));
      }
      lse1 = lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r1];
    }
    delta0 += __shfl_xor_sync(0xffffffffu, delta0, 1); delta1 += __shfl_xor_sync(0xffffffffu, delta1, 1);
    delta0 += __shfl_xor_sync(0xffffffffu, delta0, 2); delta1 += __shfl_xor_sync(0xffffffffu, delta1, 2);
  }
  // CPB parameter gradient accumulators (C fragments)
  float gW2[2][4][4];          // dW2[c][k]:                         [m-tile of c][n-tile of k]
  float gL1[2][4];             // dpre^T X = [dW1x | dW1y | db1 | .]: [m-tile of k]
  float gX2[2][4];             // dh2^T  X = [. | . | db2 | .]:       [m-tile of c]
  float gX3[2][4];             // h2^T   X = [. | . | . | dw3]:       [m-tile of c]
  float gb3 = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      gL1[a][c] = 0.f; gX2[a][c] = 0.f; gX3[a][c] = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) gW2[a][c][e] = 0.f;
    }
  const float keep_scale = cf.p_drop > 0.f ? 1.f / (1.f - cf.p_drop) : 1.f;
  const uint32_t thresh = (uint32_t)(cf.p_drop * 4294967296.0);
  const unsigned long long hrow0 = ((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + r0)) * cf.nkv;
  const unsigned long long hrow1 = ((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + r1)) * cf.nkv;
  const float* kb = k + (int64_t)b * cf.nkv * cf.inner + head * AD;
  const float* vb = v + (int64_t)b * cf.nkv * cf.inner + head * AD;
  const float* c1 = sm.c1[t];
  const float* c2 = sm.c2[t];
  float (*stS)[9] = sm.stg[w][0];
  float (*stD)[9] = sm.stg[w][1];
  const uint32_t w3c0 = g == 0 ? w3f[0][0] : 0u, w3c1 = g == 0 ? w3f[0][1] : 0u;     // layer 3 as an MMA: w3 in column 0
  const uint32_t w3c2 = g == 0 ? w3f[1][0] : 0u, w3c3 = g == 0 ? w3f[1][1] : 0u;
  const int qlane = lane & ~3;

  for (int k0 = 0; k0 < cf.nkv; k0 += BK) {
    __syncthreads();
    load_tile<false, BK>(sm.K, nullptr, kb, cf.inner, k0, cf.nkv);
    load_tile<true, BK>(sm.V[0], sm.V[1], vb, cf.inner, k0, cf.nkv);
    if (tid < BK) {
      const bool ok = k0 + tid < cf.nkv;
      sm.tb.kpos[tid][0] = ok ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD] : 0.f;
      sm.tb.kpos[tid][1] = (ok && CD == 2) ? vs[((int64_t)bg * cf.nkv + k0 + tid) * CD + 1] : 0.f;
      sm.dkpos[tid][0] = 0.f; sm.dkpos[tid][1] = 0.f;
    }
    __syncthreads();
    lean_tables<CD, BK>(sm.tb, ny);
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < BK / 8; ++j) {
      // ---- phase 1: S = Q K^T and dP = dO V^T (dO, V as hi + lo) of my 16 rows x this 8-key group -> warp-private staging
      {
        float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t ah[4], gh[4], gl[4];
          const int c0 = 16 * ks + 2 * t;
          ah[0] = *reinterpret_cast<const uint32_t*>(&sm.Q[r0][c0]);     ah[1] = *reinterpret_cast<const uint32_t*>(&sm.Q[r1][c0]);
          ah[2] = *reinterpret_cast<const uint32_t*>(&sm.Q[r0][c0 + 8]); ah[3] = *reinterpret_cast<const uint32_t*>(&sm.Q[r1][c0 + 8]);
          gh[0] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r0][c0]);     gh[1] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r1][c0]);
          gh[2] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r0][c0 + 8]); gh[3] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r1][c0 + 8]);
          gl[0] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r0][c0]);     gl[1] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r1][c0]);
          gl[2] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r0][c0 + 8]); gl[3] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r1][c0 + 8]);
          const uint32_t kh0 = *reinterpret_cast<const uint32_t*>(&sm.K[8 * j + g][c0]);
          const uint32_t kh1 = *reinterpret_cast<const uint32_t*>(&sm.K[8 * j + g][c0 + 8]);
          const uint32_t vh0 = *reinterpret_cast<const uint32_t*>(&sm.V[0][8 * j + g][c0]);
          const uint32_t vh1 = *reinterpret_cast<const uint32_t*>(&sm.V[0][8 * j + g][c0 + 8]);
          const uint32_t vl0 = *reinterpret_cast<const uint32_t*>(&sm.V[1][8 * j + g][c0]);
          const uint32_t vl1 = *reinterpret_cast<const uint32_t*>(&sm.V[1][8 * j + g][c0 + 8]);
          mma16816(s, ah, kh0, kh1);
          mma3<true>(dp, gh, gl, vh0, vh1, vl0, vl1);
        }
        __syncwarp();
        stS[g][2 * t] = s[0]; stS[g][2 * t + 1] = s[1]; stS[g + 8][2 * t] = s[2]; stS[g + 8][2 * t + 1] = s[3];
        stD[g][2 * t] = dp[0]; stD[g][2 * t + 1] = dp[1]; stD[g + 8][2 * t] = dp[2]; stD[g + 8][2 * t + 1] = dp[3];
        __syncwarp();
      }
      // ---- phase 2: CPB forward + backward, one key (16 pairs) per MMA tile
#pragma unroll UNR
      for (int u = 0; u < 8; ++u) {
        const int key = 8 * j + u;
        const float ax0 = sm.tb.SX[key][r0], ax1 = sm.tb.SX[key][r1];
        const float ay0 = CD == 2 ? sm.tb.SY[ys0][key] : 0.f, ay1 = CD == 2 ? sm.tb.SY[ys1][key] : 0.f;
        // layer 1 -> A fragments of h1 = relu(W1 a + b1)
        uint32_t ah[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const float4 wx = *reinterpret_cast<const float4*>(c1 + 4 * s), wy = *reinterpret_cast<const float4*>(c1 + 8 + 4 * s), bb = *reinterpret_cast<const float4*>(c1 + 16 + 4 * s);
          const float wxa[4] = {wx.x, wx.y, wx.z, wx.w}, wya[4] = {wy.x, wy.y, wy.z, wy.w}, bba[4] = {bb.x, bb.y, bb.z, bb.w};
          float h0[4], h1[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float u0 = fmaf(wxa[e], ax0, bba[e]), u1 = fmaf(wxa[e], ax1, bba[e]);
            if (CD == 2) { u0 = fmaf(wya[e], ay0, u0); u1 = fmaf(wya[e], ay1, u1); }
            h0[e] = u0; h1[e] = u1;
          }
          ah[s][0] = relu_pack(h0[0], h0[1]); ah[s][1] = relu_pack(h1[0], h1[1]);
          ah[s][2] = relu_pack(h0[2], h0[3]); ah[s][3] = relu_pack(h1[2], h1[3]);
        }
        // layer 2 -> A fragments of h2 = relu(W2 h1 + b2) (the fp32 accumulators die here)
        uint32_t hh[2][4];
        {
          const float4 ba = *reinterpret_cast<const float4*>(c2), bbv = *reinterpret_cast<const float4*>(c2 + 4);
          const float b2v[8] = {ba.x, ba.y, ba.z, ba.w, bbv.x, bbv.y, bbv.z, bbv.w};
          float acc[4][4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            mma16816_c(acc[jj], ah[0], wh[jj][0][0], wh[jj][0][1], b2v[2 * jj], b2v[2 * jj + 1], b2v[2 * jj], b2v[2 * jj + 1]);
            mma16816(acc[jj], ah[1], wh[jj][1][0], wh[jj][1][1]);
          }
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            hh[s][0] = relu_pack(acc[2 * s][0], acc[2 * s][1]);         hh[s][1] = relu_pack(acc[2 * s][2], acc[2 * s][3]);
            hh[s][2] = relu_pack(acc[2 * s + 1][0], acc[2 * s + 1][1]); hh[s][3] = relu_pack(acc[2 * s + 1][2], acc[2 * s + 1][3]);
          }
        }
        // layer 3: column 0 of the C fragment = bias of rows r0 / r1 (lanes t == 0) -> whole quad
        float p0, p1;
        {
          float d3[4];
          mma16816_c(d3, hh[0], w3c0, w3c1, 0.f, 0.f, 0.f, 0.f);
          mma16816(d3, hh[1], w3c2, w3c3);
          p0 = __shfl_sync(0xffffffffu, d3[0], qlane);
          p1 = __shfl_sync(0xffffffffu, d3[2], qlane);
        }
        const bool kok = k0 + key < cf.nkv;
        const float P0 = kok ? __expf(stS[g][u] + p0 + b3 - lse0) : 0.f;
        const float P1 = kok ? __expf(stS[g + 8][u] + p1 + b3 - lse1) : 0.f;
        float keep0 = 1.f, keep1 = 1.f;
        if (cf.p_drop > 0.f) {
          keep0 = pair_hash(salted(cf.seed, cf.salt), hrow0 + (unsigned long long)(k0 + key)) >= thresh ? keep_scale : 0.f;
          keep1 = pair_hash(salted(cf.seed, cf.salt), hrow1 + (unsigned long long)(k0 + key)) >= thresh ? keep_scale : 0.f;
        }
        const float dS0 = P0 * (stD[g][u] * keep0 - delta0);
        const float dS1 = P1 * (stD[g + 8][u] * keep1 - delta1);
        {                                                  // lanes t = 0 / 1 store rows r0 / r1 (branch-free: selects + predicated stores)
          const __nv_bfloat16 pv = __float2bfloat16_rn(t ? P1 * keep1 : P0 * keep0), sv = __float2bfloat16_rn(t ? dS1 : dS0);
          __nv_bfloat16* pp = &sm.P[t ? r1 : r0][key];
          if (t < 2) { *pp = pv; *(pp + (&sm.dS[0][0] - &sm.P[0][0])) = sv; }
        }
        gb3 += t == 0 ? dS0 + dS1 : 0.f;
        // dh2 = 1[h2 > 0] (dS w3): packed bf16x2, A-fragment order
        const uint32_t dS0p = pack_bf16(dS0, dS0), dS1p = pack_bf16(dS1, dS1);
        uint32_t dah[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          dah[s][0] = nz_mask_bf16x2(hh[s][0]) & mul_bf16x2(dS0p, w3f[s][0]);
          dah[s][1] = nz_mask_bf16x2(hh[s][1]) & mul_bf16x2(dS1p, w3f[s][0]);
          dah[s][2] = nz_mask_bf16x2(hh[s][2]) & mul_bf16x2(dS0p, w3f[s][1]);
          dah[s][3] = nz_mask_bf16x2(hh[s][3]) & mul_bf16x2(dS1p, w3f[s][1]);
        }
        // right operand X = [ax | ay | 1 | dS | 0 ...] ([16 pairs x 8], C-fragment layout: lanes t == 0 hold (ax, ay), t == 1 (1, dS))
        uint32_t bx0, bx1;
        {
          const uint32_t xa0 = pack_bf16(ax0, ay0), xa1 = pack_bf16(ax1, ay1), xb0 = pack_bf16(1.f, dS0), xb1 = pack_bf16(1.f, dS1);
          const uint32_t x0 = t == 0 ? xa0 : (t == 1 ? xb0 : 0u), x1 = t == 0 ? xa1 : (t == 1 ? xb1 : 0u);
          bx0 = movm_trans(x0); bx1 = movm_trans(x1);
        }
        // dW2[c][k] += sum_pairs dh2[p][c] h1[p][k];  [. . db2 .] += dh2^T X;  [. . . dw3] += h2^T X
        {
          uint32_t bh[4][2];
#pragma unroll
          for (int jn = 0; jn < 4; ++jn) {
            const int s = jn >> 1, rb = (jn & 1) * 2;
            bh[jn][0] = movm_trans(ah[s][rb]); bh[jn][1] = movm_trans(ah[s][rb + 1]);
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            uint32_t ta[4];
            ta[0] = movm_trans(dah[mt][0]); ta[1] = movm_trans(dah[mt][2]); ta[2] = movm_trans(dah[mt][1]); ta[3] = movm_trans(dah[mt][3]);
#pragma unroll
            for (int jn = 0; jn < 4; ++jn) mma16816(gW2[mt][jn], ta, bh[jn][0], bh[jn][1]);
            mma16816(gX2[mt], ta, bx0, bx1);
            uint32_t th[4];
            th[0] = movm_trans(hh[mt][0]); th[1] = movm_trans(hh[mt][2]); th[2] = movm_trans(hh[mt][1]); th[3] = movm_trans(hh[mt][3]);
            mma16816(gX3[mt], th, bx0, bx1);
          }
        }
        // dh1 = dh2 W2, first ReLU backward (mask from the bf16 h1 fragments), packed as A fragments
        uint32_t peh[4][2];                                 // [n-tile jn of k][row r0 / r1]
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          float d1[4];
          mma16816_c(d1, dah[0], wth[jn][0][0], wth[jn][0][1], 0.f, 0.f, 0.f, 0.f);
          mma16816(d1, dah[1], wth[jn][1][0], wth[jn][1][1]);
          peh[jn][0] = nz_mask_bf16x2(ah[jn >> 1][(jn & 1) * 2]) & pack_bf16(d1[0], d1[1]);
          peh[jn][1] = nz_mask_bf16x2(ah[jn >> 1][(jn & 1) * 2 + 1]) & pack_bf16(d1[2], d1[3]);
        }
        // d bias / d (signed-log distance) = dpre W1; lanes t == 0 hold (x, y) of rows r0 / r1
        float dpo[4];
        {
          const uint32_t pa0[4] = {peh[0][0], peh[0][1], peh[1][0], peh[1][1]};
          const uint32_t pa1[4] = {peh[2][0], peh[2][1], peh[3][0], peh[3][1]};
          mma16816_c(dpo, pa0, w1b[0][0], w1b[0][1], 0.f, 0.f, 0.f, 0.f);
          mma16816(dpo, pa1, w1b[1][0], w1b[1][1]);
        }
        // [dW1x | dW1y | db1 | .] += dpre^T X
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t ta[4];
          ta[0] = movm_trans(peh[2 * mt][0]); ta[1] = movm_trans(peh[2 * mt + 1][0]);
          ta[2] = movm_trans(peh[2 * mt][1]); ta[3] = movm_trans(peh[2 * mt + 1][1]);
          mma16816(gL1[mt], ta, bx0, bx1);
        }
        // key-coordinate gradient: d slog(x)/dx = 1 / (|x| + 1) = exp(-|slog(x)|), minus sign for the key
        {
          const float kgx = -(dpo[0] * __expf(-fabsf(ax0)) + dpo[2] * __expf(-fabsf(ax1)));
          const float kgy = CD == 2 ? -(dpo[1] * __expf(-fabsf(ay0)) + dpo[3] * __expf(-fabsf(ay1))) : 0.f;
          if (t == 0) *reinterpret_cast<float2*>(&sm.kgs[w][g][2 * u]) = make_float2(kgx, kgy);
        }
      }
      __syncwarp();
      if (lane < 16) {
        float r = 0.f;
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) r += sm.kgs[w][gg][lane];
        if (CD == 2 || (lane & 1) == 0) atomicAdd(&sm.dkpos[8 * j + (lane >> 1)][lane & 1], r);
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- phase 3: dQ += dS K
    {
      float dqacc[8][4];
#pragma unroll
      for (int jd = 0; jd < 8; ++jd)
#pragma unroll
        for (int e = 0; e < 4; ++e) dqacc[jd][e] = sm.dQ[jd * 4 + e][tid];
#pragma unroll
      for (int ks = 0; ks < BK / 16; ++ks) {
        uint32_t ah[4];
        const int c0 = 16 * ks + 2 * t;
        ah[0] = *reinterpret_cast<const uint32_t*>(&sm.dS[r0][c0]);     ah[1] = *reinterpret_cast<const uint32_t*>(&sm.dS[r1][c0]);
        ah[2] = *reinterpret_cast<const uint32_t*>(&sm.dS[r0][c0 + 8]); ah[3] = *reinterpret_cast<const uint32_t*>(&sm.dS[r1][c0 + 8]);
#pragma unroll
        for (int jd = 0; jd < 8; ++jd) {
          uint32_t bh0, bh1;
          ldmatrix_x2_trans(bh0, bh1, &sm.K[16 * ks + (lane & 15)][8 * jd]);
          mma16816(dqacc[jd], ah, bh0, bh1);
        }
      }
#pragma unroll
      for (int jd = 0; jd < 8; ++jd)
#pragma unroll
        for (int e = 0; e < 4; ++e) sm.dQ[jd * 4 + e][tid] = dqacc[jd][e];
    }
    // ---- dK += dS^T Q and dV += (P keep)^T dO: warp slice = 16 keys x 32 dims, contraction over the CTA's 64 queries
    {
      const int mt = w & 1, dh = w >> 1;
      float dkacc[4][4], dvacc[4][4];
#pragma unroll
      for (int jd = 0; jd < 4; ++jd)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dkacc[jd][e] = 0.f; dvacc[jd][e] = 0.f; }
      const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int qrow = 16 * ks + rr + 8 * (mi >> 1), kcol = 16 * mt + 8 * (mi & 1);
        uint32_t sh[4], ph[4];
        ldmatrix_x4_trans(sh[0], sh[1], sh[2], sh[3], &sm.dS[qrow][kcol]);
        ldmatrix_x4_trans(ph[0], ph[1], ph[2], ph[3], &sm.P[qrow][kcol]);
#pragma unroll
        for (int jd = 0; jd < 4; ++jd) {
          uint32_t qh0, qh1, gh0, gh1;
          ldmatrix_x2_trans(qh0, qh1, &sm.Q[16 * ks + (lane & 15)][32 * dh + 8 * jd]);
          ldmatrix_x2_trans(gh0, gh1, &sm.dO[0][16 * ks + (lane & 15)][32 * dh + 8 * jd]);
          mma16816(dkacc[jd], sh, qh0, qh1);
          mma16816(dvacc[jd], ph, gh0, gh1);
        }
      }
      const int key0 = k0 + 16 * mt + g, key1 = key0 + 8;
#pragma unroll
      for (int jd = 0; jd < 4; ++jd) {
        const int col = head * AD + 32 * dh + 8 * jd + 2 * t;
        if (key0 < cf.nkv) {
          const int64_t off = ((int64_t)b * cf.nkv + key0) * cf.inner + col;
          atomicAdd(reinterpret_cast<float2*>(dk + off), make_float2(dkacc[jd][0], dkacc[jd][1]));     // red.global.add.v2.f32
          atomicAdd(reinterpret_cast<float2*>(dv + off), make_float2(dvacc[jd][0], dvacc[jd][1]));
        }
        if (key1 < cf.nkv) {
          const int64_t off = ((int64_t)b * cf.nkv + key1) * cf.inner + col;
          atomicAdd(reinterpret_cast<float2*>(dk + off), make_float2(dkacc[jd][2], dkacc[jd][3]));
          atomicAdd(reinterpret_cast<float2*>(dv + off), make_float2(dvacc[jd][2], dvacc[jd][3]));
        }
      }
    }
    if (tid < BK * CD) {
      const int key = tid / CD, c = tid % CD;
      if (k0 + key < cf.nkv) atomicAdd(dvs + ((int64_t)bg * cf.nkv + k0 + key) * CD + c, sm.dkpos[key][c]);
    }
  }
  // ---- dQ
  if (ok0) {
    float* row = dq + qoff + (int64_t)(q0 + r0) * cf.inner + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(row + 8 * j) = make_float2(sm.dQ[j * 4][tid], sm.dQ[j * 4 + 1][tid]);
  }
  if (ok1) {
    float* row = dq + qoff + (int64_t)(q0 + r1) * cf.inner + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(row + 8 * j) = make_float2(sm.dQ[j * 4 + 2][tid], sm.dQ[j * 4 + 3][tid]);
  }
  // ---- CPB parameter partials of this CTA: [W1 32*CD][b1 32][W2 1024][b2 32][W3 OUT*32][b3 OUT]
  __syncthreads();
  float* const red = reinterpret_cast<float*>(&sm.P[0][0]);
  static_assert(sizeof(LeanBwdSmem<CD>::P) + sizeof(LeanBwdSmem<CD>::dS) >= NRED * sizeof(float), "the reduction buffer aliases P + dS");
  for (int i = tid; i < NRED; i += MTHREADS) red[i] = 0.f;
  __syncthreads();
  {
    float* rW1 = red; float* rb1 = rW1 + CPB_H * 2; float* rW2 = rb1 + CPB_H; float* rb2 = rW2 + CPB_H * CPB_H;
    float* rw3 = rb2 + CPB_H; float* rb3 = rw3 + CPB_H;
    // C fragments with rows = hidden unit (16 mt + g, + 8) and columns 2 t + {0, 1} of X = [ax | ay | 1 | dS]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int kk = 16 * mt + g;
      if (t == 0) {
        atomicAdd(&rW1[kk * 2], gL1[mt][0]); atomicAdd(&rW1[kk * 2 + 1], gL1[mt][1]);
        atomicAdd(&rW1[(kk + 8) * 2], gL1[mt][2]); atomicAdd(&rW1[(kk + 8) * 2 + 1], gL1[mt][3]);
      } else if (t == 1) {
        atomicAdd(&rb1[kk], gL1[mt][0]); atomicAdd(&rb1[kk + 8], gL1[mt][2]);
        atomicAdd(&rb2[kk], gX2[mt][0]); atomicAdd(&rb2[kk + 8], gX2[mt][2]);
        atomicAdd(&rw3[kk], gX3[mt][1]); atomicAdd(&rw3[kk + 8], gX3[mt][3]);
      }
    }
    gb3 = warp_sum(gb3);
    if (lane == 0) atomicAdd(rb3, gb3);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {
        const int c = 16 * mt + g, kk = 8 * jn + 2 * t;
        atomicAdd(&rW2[c * CPB_H + kk], gW2[mt][jn][0]); atomicAdd(&rW2[c * CPB_H + kk + 1], gW2[mt][jn][1]);
        atomicAdd(&rW2[(c + 8) * CPB_H + kk], gW2[mt][jn][2]); atomicAdd(&rW2[(c + 8) * CPB_H + kk + 1], gW2[mt][jn][3]);
      }
  }
  __syncthreads();
  {
    const int np = CPB_H * CD + CPB_H + CPB_H * CPB_H + CPB_H + cf.OUT * CPB_H + cf.OUT;
    float* prow = part + ((int64_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * np;
    const float* rW1 = red; const float* rb1 = rW1 + CPB_H * 2; const float* rW2 = rb1 + CPB_H;
    const float* rb2 = rW2 + CPB_H * CPB_H; const float* rw3 = rb2 + CPB_H; const float* rb3 = rw3 + CPB_H;
    for (int i = tid; i < np; i += MTHREADS) {
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


// ---- 1-D backward with the tabulated bias (see PwlTable above).  No per-key MLP tiles: the warp's 16 rows x 32 keys are
// processed as logit fragments, exactly like the forward.  Per element: p2 = sign log2(|d| + 1), the row's two lines from
// registers (PwlRow), P = 2^(S2 + bias2 - lse2), dS = P (dP keep - delta).  Bias-MLP gradients: per segment the sums
// G0[s] = sum dS and G1[s] = sum dS p.  A (row, key tile) range touches the segments {seg, seg + 1} only, so the tile's sums
// are four register accumulators (total and upper-line part); they are merged into per-row running sums that are flushed to
// shared-memory bins only when the row's lower segment changes (~ once per segment per row), and the bins go to the global
// [OUT, rows, 2] buffer at the end; sml_cpb1d_table_bwd maps them to the MLP parameters.  Tiles with >= 2 breakpoints in range
// search the table and bin per pair (warp-uniform slow path).  Key-coordinate gradient: d bias / d p = slope,
// d p / d kx = -1 / (|d| + 1) = -2^(-|p2|).  dQ / dK / dV as in the lean backward (dK is rescaled by ln 2: the Q tile carries log2 e).
struct PwlBwdSmem {
  __nv_bfloat16 Q[MQ][MLD], dO[MQ][MLD];
  PwlStage<BK> st[3];
  __nv_bfloat16 P[2][MQ][PLD], dS[2][MQ][PLD];
  float dkpos[2][MTHREADS / 32][BK];      // per-warp partial key-coordinate gradients (no shared-memory atomics)
  float2 krng[PWL_MAXTILES];              // key-position range of every tile of this (b, g) row
  PwlTable pw;
};
// running per-row sums of the segments {seg, seg + 1}; the four lanes of a quad hold the same row: they are reduced at flush time
struct PwlAcc { int seg; float a0, b0, a1, b1; };
// Adds the running sums of the lanes with `need` to the bins and clears them.  The 16 rows of a warp sit in nearly the same
// place of the table, so they almost always leave the same segment in the same tile: then the warp reduces the four sums
// with shuffles and one lane issues the four shared-memory atomics (fp32 shared atomics are compare-and-swap loops: 8 rows
// hitting one bin cost ~15 spins each).  Rows that disagree fall back to one atomic set per row.  Warp-uniform call;
// `need` is the same in the four lanes of a quad.
__device__ __forceinline__ void pwl_flush(float2* bins, PwlAcc& ac, bool need, int lane) {
  const unsigned who = __ballot_sync(0xffffffffu, need);
  if (who == 0u) return;
  const int segref = __shfl_sync(0xffffffffu, ac.seg, __ffs(who) - 1);
  const bool same = __all_sync(0xffffffffu, !need || ac.seg == segref);
  // the four lanes of a quad hold partial sums (different keys) of the same row: always reduced; all rows only when they agree
  float v0 = need ? ac.a0 : 0.f, v1 = need ? ac.b0 : 0.f, v2 = need ? ac.a1 : 0.f, v3 = need ? ac.b1 : 0.f;
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) {
    if (m < 4 || same) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, m); v1 += __shfl_xor_sync(0xffffffffu, v1, m);
      v2 += __shfl_xor_sync(0xffffffffu, v2, m); v3 += __shfl_xor_sync(0xffffffffu, v3, m);
    }
  }
  if (same ? lane == 0 : (need && (lane & 3) == 0)) {
    const int sg = same ? segref : ac.seg;
    if (v0 != 0.f || v1 != 0.f) { atomicAdd(&bins[sg].x, v0); atomicAdd(&bins[sg].y, v1 * PWL_LN2); }
    if (v2 != 0.f || v3 != 0.f) { atomicAdd(&bins[sg + 1].x, v2); atomicAdd(&bins[sg + 1].y, v3 * PWL_LN2); }
  }
  if (need) ac.a0 = ac.b0 = ac.a1 = ac.b1 = 0.f;
}
// one lane's sums (slow path: no assumption on the other lanes)
__device__ __forceinline__ void pwl_flush_lane(float2* bins, PwlAcc& ac) {
  if (ac.a0 != 0.f || ac.b0 != 0.f) { atomicAdd(&bins[ac.seg].x, ac.a0); atomicAdd(&bins[ac.seg].y, ac.b0 * PWL_LN2); }
  if (ac.a1 != 0.f || ac.b1 != 0.f) { atomicAdd(&bins[ac.seg + 1].x, ac.a1); atomicAdd(&bins[ac.seg + 1].y, ac.b1 * PWL_LN2); }
  ac.a0 = ac.b0 = ac.a1 = ac.b1 = 0.f;
}

// Operands: K, V as bf16 planes from the caller; dO is rounded to bf16 once and delta = bf16(dO) . O uses the same rounded
// values, so that sum_j P_ij dP_ij = delta holds to fp32 rounding with single-pass products (dP = bf16(dO) bf16(V)^T,
// O = sum_j P_ij bf16(V_j)).  Key tiles move through a three-stage cp.async ring, P / dS tiles are double-buffered: one
// barrier per tile; 69 KB of shared memory and <= 168 registers: three CTAs per SM.
template <bool DROP>
__global__ void __launch_bounds__(MTHREADS, 3) deform_attn_bwd_pwl_kernel(
    const AttnCfg cf, const float* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    const float* __restrict__ vs, const float2* __restrict__ kr, const float* __restrict__ brk, const float* __restrict__ coef,
    const int* __restrict__ nseg, const float* __restrict__ out, const float* __restrict__ lse, const float* __restrict__ dout,
    float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, float* __restrict__ dvs, float* __restrict__ gseg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PwlBwdSmem& sm = *reinterpret_cast<PwlBwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * MQ;
  const int grp = head / cf.OUT, o = head % cf.OUT;
  const int bg = b * cf.G + grp;
  const int ns = nseg[0];
  const __nv_bfloat16* kb = k + (int64_t)b * cf.nkv * cf.inner + head * AD;
  const __nv_bfloat16* vb = v + (int64_t)b * cf.nkv * cf.inner + head * AD;
  const float* vsb = vs + (int64_t)bg * cf.nkv;
  const int ntiles = (cf.nkv + BK - 1) / BK;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (i < ntiles) pwl_stage_load(sm.st[i], kb, vb, vsb, cf.inner, i * BK, cf.nkv, tid);
    cp_async_commit();
  }
  pwl_load(sm.pw, brk, coef, ns - 1, cf.OUT, o, tid, MTHREADS);
  if (ntiles <= PWL_MAXTILES)
    for (int i = tid; i < ntiles; i += MTHREADS) sm.krng[i] = kr[(int64_t)bg * ntiles + i];
  float2* bins = reinterpret_cast<float2*>(gseg) + (int64_t)o * (PWL_MAXSEG + 1);      // global: ~4 reductions per warp and segment
  const int64_t qoff = (int64_t)b * cf.n * cf.inner + head * AD;
  for (int e = tid; e < MQ * 16; e += MTHREADS) {          // Q tile scaled by log2(e); dO tile
    const int r = e >> 4, c4 = (e & 15) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
    if (q0 + r < cf.n) {
      x = *reinterpret_cast<const float4*>(q + qoff + (int64_t)(q0 + r) * cf.inner + c4);
      y = *reinterpret_cast<const float4*>(dout + qoff + (int64_t)(q0 + r) * cf.inner + c4);
    }
    *reinterpret_cast<uint2*>(&sm.Q[r][c4]) = make_uint2(pack_bf16(x.x * PWL_LOG2E, x.y * PWL_LOG2E), pack_bf16(x.z * PWL_LOG2E, x.w * PWL_LOG2E));
    *reinterpret_cast<uint2*>(&sm.dO[r][c4]) = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
  }
  const int r0 = w * 16 + g, r1 = r0 + 8;
  const bool ok0 = q0 + r0 < cf.n, ok1 = q0 + r1 < cf.n;
  float qx0, qx1, qdum;
  query_pos(cf, min(q0 + r0, cf.n - 1), qx0, qdum);
  query_pos(cf, min(q0 + r1, cf.n - 1), qx1, qdum);
  float delta0 = 0.f, delta1 = 0.f, lse0 = INFINITY, lse1 = INFINITY;      // lse in base-2 units
  {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      if (rr ? ok1 : ok0) {
        const int64_t ro = qoff + (int64_t)(q0 + (rr ? r1 : r0)) * cf.inner + 16 * t;
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(out + ro + 4 * i), c = *reinterpret_cast<const float4*>(dout + ro + 4 * i);
          const uint32_t c01 = pack_bf16(c.x, c.y), c23 = pack_bf16(c.z, c.w);
          d = fmaf(a.x, bf16_lo_f(c01), fmaf(a.y, bf16_hi_f(c01), fmaf(a.z, bf16_lo_f(c23), fmaf(a.w, bf16_hi_f(c23), d))));
        }
        const float l2 = lse[((int64_t)b * cf.heads + head) * cf.n + q0 + (rr ? r1 : r0)] * PWL_LOG2E;
        if (rr) { delta1 = d; lse1 = l2; } else { delta0 = d; lse0 = l2; }
      }
    }
    delta0 += __shfl_xor_sync(0xffffffffu, delta0, 1); delta1 += __shfl_xor_sync(0xffffffffu, delta1, 1);
    delta0 += __shfl_xor_sync(0xffffffffu, delta0, 2); delta1 += __shfl_xor_sync(0xffffffffu, delta1, 2);
  }
  const float keep_scale = DROP ? 1.f / (1.f - cf.p_drop) : 1.f;
  const uint32_t thresh = (uint32_t)(cf.p_drop * 4294967296.0);
  const unsigned long long hrow0 = ((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + r0)) * cf.nkv;
  const unsigned long long hrow1 = ((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + r1)) * cf.nkv;
  PwlRow rw0, rw1;
  rw0.seg = rw1.seg = ns >> 1;
  rw0.b = rw1.b = __int_as_float(0x7fc00000);          // NaN: forces the first line load
  PwlAcc ac0, ac1;
  ac0.seg = ac1.seg = ns >> 1;
  ac0.a0 = ac0.b0 = ac0.a1 = ac0.b1 = 0.f;
  ac1.a0 = ac1.b0 = ac1.a1 = ac1.b1 = 0.f;
  float dqacc[8][4];
#pragma unroll
  for (int jd = 0; jd < 8; ++jd)
#pragma unroll
    for (int e = 0; e < 4; ++e) dqacc[jd][e] = 0.f;
  // element offset of this thread's dK / dV slice (key 16 mt + g of tile 0, dims 32 dh + 2 t ...)
  const int64_t goff = ((int64_t)b * cf.nkv + 16 * (w & 1) + g) * cf.inner + head * AD + 32 * (w >> 1) + 2 * t;
  cp_async_wait<1>();                                   // tile 0
  __syncthreads();

  const float2* krb = kr + (int64_t)bg * ntiles;
  const bool rng_smem = ntiles <= PWL_MAXTILES;      // (staged before the first barrier below)

  for (int it = 0; it < ntiles; ++it) {
    const int k0 = it * BK;
    const float2 rng = rng_smem ? sm.krng[it] : krb[it];
    const float kmin = rng.x, kmax = rng.y;
    const PwlStage<BK>& st = sm.st[it % 3];
    __nv_bfloat16 (*Pt)[PLD] = sm.P[it & 1];
    __nv_bfloat16 (*dSt)[PLD] = sm.dS[it & 1];
    float* dkp = sm.dkpos[it & 1][w];
    const bool many0 = pwl_row_setup(sm.pw, ns, qx0, kmin, kmax, rw0);
    const bool many1 = pwl_row_setup(sm.pw, ns, qx1, kmin, kmax, rw1);
    const bool slow = __any_sync(0xffffffffu, many0 || many1);
    // ---- S2 = log2(e) Q K^T and dP = dO V^T of my 16 rows x the 32 keys of the tile
    float s[4][4], dp[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[j][e] = 0.f; dp[j][e] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ah[4], gh[4];
      ldmatrix_x4(ah[0], ah[1], ah[2], ah[3], &sm.Q[w * 16 + (lane & 15)][16 * ks + 8 * (lane >> 4)]);
      ldmatrix_x4(gh[0], gh[1], gh[2], gh[3], &sm.dO[w * 16 + (lane & 15)][16 * ks + 8 * (lane >> 4)]);
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const int krow = 8 * j + (lane & 7) + 8 * (lane >> 4), kcol = 16 * ks + 8 * ((lane >> 3) & 1);
        uint32_t k00, k01, k10, k11, v00, v01, v10, v11;
        ldmatrix_x4(k00, k01, k10, k11, pwl_sw(st.K, krow, kcol));
        ldmatrix_x4(v00, v01, v10, v11, pwl_sw(st.V, krow, kcol));
        mma16816(s[j], ah, k00, k01);
        mma16816(dp[j], gh, v00, v01);
        mma16816(s[j + 1], ah, k10, k11);
        mma16816(dp[j + 1], gh, v10, v11);
      }
    }
    // ---- per element: bias, P, dS, segment sums, key-coordinate gradient
    const bool ragged = k0 + BK > cf.nkv;
    if (!slow) {
      float aT0 = 0.f, bT0 = 0.f, aH0 = 0.f, bH0 = 0.f, aT1 = 0.f, bT1 = 0.f, aH1 = 0.f, bH1 = 0.f;
      float kgs[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 kx = *reinterpret_cast<const float2*>(&st.kx[8 * j + 2 * t]);
        float kg[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float kxe = e ? kx.y : kx.x;
          const float p0 = slog2_fast(qx0 - kxe), p1 = slog2_fast(qx1 - kxe);
          const bool lo0 = p0 < rw0.b, lo1 = p1 < rw1.b;
          const float sl0 = lo0 ? rw0.sl : rw0.sh, sl1 = lo1 ? rw1.sl : rw1.sh;
          float P0 = ex2_fast(s[j][e] + fmaf(p0, sl0, lo0 ? rw0.il : rw0.ih) - lse0);
          float P1 = ex2_fast(s[j][2 + e] + fmaf(p1, sl1, lo1 ? rw1.il : rw1.ih) - lse1);
          if (ragged && k0 + 8 * j + 2 * t + e >= cf.nkv) { P0 = 0.f; P1 = 0.f; }
          float keep0 = 1.f, keep1 = 1.f;
          if (DROP) {
            const unsigned long long key = (unsigned long long)(k0 + 8 * j + 2 * t + e);
            keep0 = pair_hash(salted(cf.seed, cf.salt), hrow0 + key) >= thresh ? keep_scale : 0.f;
            keep1 = pair_hash(salted(cf.seed, cf.salt), hrow1 + key) >= thresh ? keep_scale : 0.f;
          }
          const float dS0 = P0 * (DROP ? dp[j][e] * keep0 - delta0 : dp[j][e] - delta0);
          const float dS1 = P1 * (DROP ? dp[j][2 + e] * keep1 - delta1 : dp[j][2 + e] - delta1);
          s[j][e] = DROP ? P0 * keep0 : P0; s[j][2 + e] = DROP ? P1 * keep1 : P1;      // fragments now hold P keep and dS
          dp[j][e] = dS0; dp[j][2 + e] = dS1;
          const float h0 = lo0 ? 0.f : dS0, h1 = lo1 ? 0.f : dS1;
          aT0 += dS0; bT0 = fmaf(dS0, p0, bT0); aH0 += h0; bH0 = fmaf(h0, p0, bH0);
          aT1 += dS1; bT1 = fmaf(dS1, p1, bT1); aH1 += h1; bH1 = fmaf(h1, p1, bH1);
          kg[e] = -fmaf(dS0 * sl0, ex2_fast(-fabsf(p0)), dS1 * sl1 * ex2_fast(-fabsf(p1)));
        }
        kgs[j][0] = kg[0]; kgs[j][1] = kg[1];
        *reinterpret_cast<uint32_t*>(&Pt[r0][8 * j + 2 * t]) = pack_bf16(s[j][0], s[j][1]);
        *reinterpret_cast<uint32_t*>(&Pt[r1][8 * j + 2 * t]) = pack_bf16(s[j][2], s[j][3]);
        *reinterpret_cast<uint32_t*>(&dSt[r0][8 * j + 2 * t]) = pack_bf16(dp[j][0], dp[j][1]);
        *reinterpret_cast<uint32_t*>(&dSt[r1][8 * j + 2 * t]) = pack_bf16(dp[j][2], dp[j][3]);
      }
      // key-coordinate gradients of keys 8 j + 2 t + {0, 1}: sum over the warp's 16 rows (8 lanes share t); the eight chains
      // are reduced together (three shuffle rounds of eight independent values instead of four separate 3-step chains)
#pragma unroll
      for (int m = 4; m < 32; m <<= 1)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          kgs[j][0] += __shfl_xor_sync(0xffffffffu, kgs[j][0], m); kgs[j][1] += __shfl_xor_sync(0xffffffffu, kgs[j][1], m);
        }
      if (g == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float2*>(&dkp[8 * j + 2 * t]) = make_float2(kgs[j][0], kgs[j][1]);
      }
      // merge the tile's sums into the rows' running sums (flush when the lower segment moved)
      pwl_flush(bins, ac0, rw0.seg != ac0.seg, lane); ac0.seg = rw0.seg;
      pwl_flush(bins, ac1, rw1.seg != ac1.seg, lane); ac1.seg = rw1.seg;
      ac0.a0 += aT0 - aH0; ac0.b0 += bT0 - bH0; ac0.a1 += aH0; ac0.b1 += bH0;
      ac1.a0 += aT1 - aH1; ac1.b0 += bT1 - bH1; ac1.a1 += aH1; ac1.b1 += bH1;
    } else {
      for (int j = 0; j < 4; ++j) {
        const float2 kx = *reinterpret_cast<const float2*>(&st.kx[8 * j + 2 * t]);
        float kg[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float kxe = e ? kx.y : kx.x;
          const bool kok = k0 + 8 * j + 2 * t + e < cf.nkv;
          const float p0 = slog2_fast(qx0 - kxe), p1 = slog2_fast(qx1 - kxe);
          const int s0 = pwl_seek(sm.pw, p0, rw0.seg), s1 = pwl_seek(sm.pw, p1, rw1.seg);
          const float2 c0 = sm.pw.line(s0), c1 = sm.pw.line(s1);
          const float P0 = kok ? ex2_fast(s[j][e] + fmaf(c0.x, p0, c0.y) - lse0) : 0.f;
          const float P1 = kok ? ex2_fast(s[j][2 + e] + fmaf(c1.x, p1, c1.y) - lse1) : 0.f;
          float keep0 = 1.f, keep1 = 1.f;
          if (DROP) {
            const unsigned long long key = (unsigned long long)(k0 + 8 * j + 2 * t + e);
            keep0 = pair_hash(salted(cf.seed, cf.salt), hrow0 + key) >= thresh ? keep_scale : 0.f;
            keep1 = pair_hash(salted(cf.seed, cf.salt), hrow1 + key) >= thresh ? keep_scale : 0.f;
          }
          const float dS0 = P0 * (dp[j][e] * keep0 - delta0);
          const float dS1 = P1 * (dp[j][2 + e] * keep1 - delta1);
          s[j][e] = P0 * keep0; s[j][2 + e] = P1 * keep1;
          dp[j][e] = dS0; dp[j][2 + e] = dS1;
          if (s0 != ac0.seg) { pwl_flush_lane(bins, ac0); ac0.seg = s0; }          // run-length accumulation per lane
          if (s1 != ac1.seg) { pwl_flush_lane(bins, ac1); ac1.seg = s1; }
          ac0.a0 += dS0; ac0.b0 = fmaf(dS0, p0, ac0.b0);
          ac1.a0 += dS1; ac1.b0 = fmaf(dS1, p1, ac1.b0);
          kg[e] = -fmaf(dS0 * c0.x, ex2_fast(-fabsf(p0)), dS1 * c1.x * ex2_fast(-fabsf(p1)));
        }
#pragma unroll
        for (int m = 4; m < 32; m <<= 1) { kg[0] += __shfl_xor_sync(0xffffffffu, kg[0], m); kg[1] += __shfl_xor_sync(0xffffffffu, kg[1], m); }
        if (g == 0) *reinterpret_cast<float2*>(&dkp[8 * j + 2 * t]) = make_float2(kg[0], kg[1]);
        *reinterpret_cast<uint32_t*>(&Pt[r0][8 * j + 2 * t]) = pack_bf16(s[j][0], s[j][1]);
        *reinterpret_cast<uint32_t*>(&Pt[r1][8 * j + 2 * t]) = pack_bf16(s[j][2], s[j][3]);
        *reinterpret_cast<uint32_t*>(&dSt[r0][8 * j + 2 * t]) = pack_bf16(dp[j][0], dp[j][1]);
        *reinterpret_cast<uint32_t*>(&dSt[r1][8 * j + 2 * t]) = pack_bf16(dp[j][2], dp[j][3]);
      }
    }
    cp_async_wait<0>();                                 // my copies of tile it + 1 have landed
    __syncthreads();                                    // P / dS of this tile and tile it + 1 visible; everyone is past tile it - 1
    if (it + 2 < ntiles) pwl_stage_load(sm.st[(it + 2) % 3], kb, vb, vsb, cf.inner, k0 + 2 * BK, cf.nkv, tid);
    cp_async_commit();
    // ---- dQ += dS K
#pragma unroll
    for (int ks = 0; ks < BK / 16; ++ks) {
      uint32_t ah[4];
      ldmatrix_x4(ah[0], ah[1], ah[2], ah[3], &dSt[w * 16 + (lane & 15)][16 * ks + 8 * (lane >> 4)]);
#pragma unroll
      for (int jd = 0; jd < 8; jd += 2) {
        uint32_t b00, b01, b10, b11;
        ldmatrix_x4_trans(b00, b01, b10, b11, pwl_sw(st.K, 16 * ks + (lane & 15), 8 * jd + 8 * (lane >> 4)));
        mma16816(dqacc[jd], ah, b00, b01);
        mma16816(dqacc[jd + 1], ah, b10, b11);
      }
    }
    // ---- dK += dS^T Q and dV += (P keep)^T dO: warp slice = 16 keys x 32 dims, contraction over the CTA's 64 queries
    {
      const int mt = w & 1, dh = w >> 1;
      float dkacc[4][4], dvacc[4][4];
#pragma unroll
      for (int jd = 0; jd < 4; ++jd)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dkacc[jd][e] = 0.f; dvacc[jd][e] = 0.f; }
      const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int qrow = 16 * ks + rr + 8 * (mi >> 1), kcol = 16 * mt + 8 * (mi & 1);
        uint32_t sh[4], ph[4];
        ldmatrix_x4_trans(sh[0], sh[1], sh[2], sh[3], &dSt[qrow][kcol]);
        ldmatrix_x4_trans(ph[0], ph[1], ph[2], ph[3], &Pt[qrow][kcol]);
#pragma unroll
        for (int jd = 0; jd < 4; ++jd) {
          uint32_t qh0, qh1, gh0, gh1;
          ldmatrix_x2_trans(qh0, qh1, &sm.Q[16 * ks + (lane & 15)][32 * dh + 8 * jd]);
          ldmatrix_x2_trans(gh0, gh1, &sm.dO[16 * ks + (lane & 15)][32 * dh + 8 * jd]);
          mma16816(dkacc[jd], sh, qh0, qh1);
          mma16816(dvacc[jd], ph, gh0, gh1);
        }
      }
      const int key0 = k0 + 16 * mt + g, key1 = key0 + 8;
      const int64_t off0 = goff + (int64_t)k0 * cf.inner, off1 = off0 + 8 * (int64_t)cf.inner;
      if (key0 < cf.nkv) {
#pragma unroll
        for (int jd = 0; jd < 4; ++jd) {
          atomicAdd(reinterpret_cast<float2*>(dk + off0 + 8 * jd), make_float2(dkacc[jd][0] * PWL_LN2, dkacc[jd][1] * PWL_LN2));
          atomicAdd(reinterpret_cast<float2*>(dv + off0 + 8 * jd), make_float2(dvacc[jd][0], dvacc[jd][1]));
        }
      }
      if (key1 < cf.nkv) {
#pragma unroll
        for (int jd = 0; jd < 4; ++jd) {
          atomicAdd(reinterpret_cast<float2*>(dk + off1 + 8 * jd), make_float2(dkacc[jd][2] * PWL_LN2, dkacc[jd][3] * PWL_LN2));
          atomicAdd(reinterpret_cast<float2*>(dv + off1 + 8 * jd), make_float2(dvacc[jd][2], dvacc[jd][3]));
        }
      }
    }
    if (tid < BK && k0 + tid < cf.nkv)
      atomicAdd(dvs + (int64_t)bg * cf.nkv + k0 + tid, (sm.dkpos[it & 1][0][tid] + sm.dkpos[it & 1][1][tid]) +
                                                         (sm.dkpos[it & 1][2][tid] + sm.dkpos[it & 1][3][tid]));
  }
  // ---- flush the running sums, dQ, and this CTA's segment sums
  pwl_flush(bins, ac0, true, lane);
  pwl_flush(bins, ac1, true, lane);
  if (ok0) {
    float* row = dq + qoff + (int64_t)(q0 + r0) * cf.inner + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(row + 8 * j) = make_float2(dqacc[j][0], dqacc[j][1]);
  }
  if (ok1) {
    float* row = dq + qoff + (int64_t)(q0 + r1) * cf.inner + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(row + 8 * j) = make_float2(dqacc[j][2], dqacc[j][3]);
  }
}

// ---- warp-specialised backward (plain bf16 operands, 2-D or 1-D lattices whose 64-query tiles span <= 8 lattice rows).
// The lean backward above runs 8 warps per SM (220 registers, 56 of them parameter-gradient accumulators) and is bound by
// dependent-issue latency: each warp walks a ~290-instruction chain per key with two warps per scheduler.  Here a CTA has two
// warpgroups that pipeline over the keys:
//   producers (warps 0-3): bias MLP forward (layers 1-3), P, dS, the bf16 P / dS tiles, gb3; then dQ += dS K
//   consumers (warps 4-7): dh2, dW2 / db2 / dw3, dh1 = dh2 W2, first ReLU backward, dW1 / db1, key-coordinate gradients; then
//                          dK += dS^T Q, dV += (P keep)^T dO
// Producer i and consumer i own the same 16 query rows; per key the producer hands over its h1 / h2 A fragments, dS and the
// signed-log distances (22 registers per lane) through a double-buffered shared-memory slot guarded by two mbarriers.  Both
// roles fit 128 registers, so two CTAs = 16 warps share an SM and each scheduler interleaves four half-length chains.
constexpr int WS_THREADS = 256;
constexpr int WS_NY = 8;             // lattice-row slots of the y table (64 consecutive tokens span <= 8 rows when W >= 10)

__device__ __forceinline__ uint32_t ws_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ws_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(ws_smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();                          // a protocol bug must not hang the GPU
}

template <int CD>
struct WsBwdSmem {
  __nv_bfloat16 Q[MQ][MLD], dO[2][MQ][MLD];
  __nv_bfloat16 K[BK][MLD], V[2][BK][MLD];
  __nv_bfloat16 P[MQ][PLD], dS[MQ][PLD];     // `red` aliases P + dS at the end of the kernel
  float dQ[32][MTHREADS];
  float stg[4][2][16][9];
  float kgs[4][8][16];
  float c1[4][24];
  float c2[4][8];
  float dkpos[BK][2];
  float SX[BK][MQ];
  float SY[CD == 2 ? WS_NY : 1][BK];
  float qx[MQ], qy[WS_NY];
  float kpos[BK][2];
  uint32_t wth[16][32];              // dh1 = dh2 W2 operand fragments by lane: [(jn * 2 + s) * 2 + reg][lane]
  uint4 hb[4][2][6][32];             // hand-over slots [pair][slot][4-register group][lane]
  uint64_t full[4][2], empty[4][2];
};

template <int CD>
__global__ void __launch_bounds__(WS_THREADS, 2) deform_attn_bwd_ws_kernel(
    const AttnCfg cf, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const float* __restrict__ vs, const float* __restrict__ cpb, const float* __restrict__ out,
    const float* __restrict__ lse, const float* __restrict__ dout, float* __restrict__ dq, float* __restrict__ dk,
    float* __restrict__ dv, float* __restrict__ dvs, float* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WsBwdSmem<CD>& sm = *reinterpret_cast<WsBwdSmem<CD>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const bool producer = warp < 4;
  const int w = warp & 3;                        // pair index: producer w and consumer w own query rows 16 w .. 16 w + 15
  const int ptid = tid & 127;                    // thread index inside the warpgroup
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * MQ;
  const int grp = head / cf.OUT, o = head % cf.OUT;
  const int bg = b * cf.G + grp;
  const float* W1 = cpb; const float* B1 = W1 + CPB_H * CD; const float* W2 = B1 + CPB_H;
  const float* B2 = W2 + CPB_H * CPB_H; const float* W3 = B2 + CPB_H; const float* B3 = W3 + cf.OUT * CPB_H;
  // ---- shared constants
  if (tid < 4 * 24) {
    const int tt = tid / 24, i = tid % 24, which = i / 8, s = (i % 8) / 4, e = i % 4;
    const int kk = 16 * s + 2 * tt + (e & 1) + 8 * (e >> 1);
    sm.c1[tt][i] = which == 0 ? W1[kk * CD] : (which == 1 ? (CD == 2 ? W1[kk * CD + 1] : 0.f) : B1[kk]);
  }
  if (tid >= 128 && tid < 128 + 4 * 8) {
    const int i0 = tid - 128, tt = i0 / 8, i = i0 % 8, j = i / 2, e = i % 2;
    sm.c2[tt][i] = B2[8 * j + 2 * tt + e];
  }
  if (tid >= 160 && tid < 192) {                 // W2 operand fragments of dh1 = dh2 W2 (B[k' = c][n = k] = W2[c][k]), by lane
    const int ln = tid - 160, gg = ln >> 2, tt = ln & 3;
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float* col = W2 + (16 * s + 2 * tt) * CPB_H + 8 * jn + gg;
        sm.wth[(jn * 2 + s) * 2 + 0][ln] = pack_bf16(col[0], col[CPB_H]);
        sm.wth[(jn * 2 + s) * 2 + 1][ln] = pack_bf16(col[8 * CPB_H], col[9 * CPB_H]);
      }
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i)
      for (int s = 0; s < 2; ++s) { ws_mbar_init(&sm.full[i][s], 1); ws_mbar_init(&sm.empty[i][s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int64_t qoff = (int64_t)b * cf.n * cf.inner + head * AD;
  if (producer) {
    load_tile<false, MQ>(sm.Q, nullptr, q + qoff, cf.inner, q0, cf.n);
#pragma unroll
    for (int i = 0; i < 32; ++i) sm.dQ[i][ptid] = 0.f;
  } else {
    // consumer threads load dO (hi / lo) with their warpgroup index
    for (int e = ptid; e < MQ * 16; e += 128) {
      const int r = e >> 4, c4 = (e & 15) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q0 + r < cf.n) x = *reinterpret_cast<const float4*>(dout + qoff + (int64_t)(q0 + r) * cf.inner + c4);
      uint32_t h0, l0, h1, l1;
      split2<true>(x.x, x.y, h0, l0);
      split2<true>(x.z, x.w, h1, l1);
      *reinterpret_cast<uint2*>(&sm.dO[0][r][c4]) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(&sm.dO[1][r][c4]) = make_uint2(l0, l1);
    }
  }
  {                                              // query positions: x per CTA row, y per lattice-row slot
    const int last = min(q0 + MQ - 1, cf.n - 1);
    const int y0 = CD == 2 ? q0 / cf.W : 0;
    const int nyq = CD == 2 ? last / cf.W - y0 + 1 : 1;
    if (tid >= 192 && tid < 192 + MQ) {
      const int r = tid - 192;
      float px, py;
      query_pos(cf, min(q0 + r, cf.n - 1), px, py);
      sm.qx[r] = px;
      if (CD == 2 && r < nyq) {
        query_pos(cf, min((y0 + r) * cf.W, cf.n - 1), px, py);
        sm.qy[r] = py;
      }
    }
  }
  const int ny = CD == 2 ? min(q0 + MQ - 1, cf.n - 1) / cf.W - q0 / cf.W + 1 : 1;
  const int r0 = w * 16 + g, r1 = r0 + 8;
  const bool ok0 = q0 + r0 < cf.n, ok1 = q0 + r1 < cf.n;
  const float* kb = k + (int64_t)b * cf.nkv * cf.inner + head * AD;
  const float* vb = v + (int64_t)b * cf.nkv * cf.inner + head * AD;
  uint32_t kc = 0;                               // keys handed over so far by this pair (slot = kc & 1, phase = (kc >> 1) & 1)

  // ------------------------------------------------------------------------------------------------ role state
  // producer
  uint32_t wh[4][2][2];
  uint32_t w3c[4];
  float b3 = 0.f, delta0 = 0.f, delta1 = 0.f, lse0 = INFINITY, lse1 = INFINITY, gb3 = 0.f;
  int ys0 = 0, ys1 = 0;
  unsigned long long hrow0 = 0, hrow1 = 0;
  // consumer
  float gW2[2][4][4], gL1[2][4], gX2[2][4], gX3[2][4];
  uint32_t w1b[2][2], w3f[2][2];
  if (producer) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float* row = W2 + (8 * j + g) * CPB_H + 16 * s + 2 * t;
        wh[j][s][0] = pack_bf16(row[0], row[1]);
        wh[j][s][1] = pack_bf16(row[8], row[9]);
      }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int k0_ = 16 * s + 2 * t;
      w3c[2 * s] = g == 0 ? pack_bf16(W3[o * CPB_H + k0_], W3[o * CPB_H + k0_ + 1]) : 0u;
      w3c[2 * s + 1] = g == 0 ? pack_bf16(W3[o * CPB_H + k0_ + 8], W3[o * CPB_H + k0_ + 9]) : 0u;
    }
    b3 = B3[o];
    ys0 = CD == 2 ? min(q0 + r0, cf.n - 1) / cf.W - q0 / cf.W : 0;
    ys1 = CD == 2 ? min(q0 + r1, cf.n - 1) / cf.W - q0 / cf.W : 0;
    hrow0 = ((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + r0)) * cf.nkv;
    hrow1 = ((unsigned long long)(b * cf.heads + head) * cf.n + (q0 + r1)) * cf.nkv;
    if (ok0) {
      const float* orow = out + qoff + (int64_t)(q0 + r0) * cf.inner + 16 * t;
      const float* grow = dout + qoff + (int64_t)(q0 + r0) * cf.inner + 16 * t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(orow + 4 * i), c = *reinterpret_cast<const float4*>(grow + 4 * i);
        delta0 = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, fmaf(a.w, c.w, delta0))));
      }
      lse0 = lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r0];
    }
    if (ok1) {
      const float* orow = out + qoff + (int64_t)(q0 + r1) * cf.inner + 16 * t;
      const float* grow = dout + qoff + (int64_t)(q0 + r1) * cf.inner + 16 * t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(orow + 4 * i), c = *reinterpret_cast<const float4*>(grow + 4 * i);
        delta1 = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, fmaf(a.w, c.w, delta1))));
      }
      lse1 = lse[((int64_t)b * cf.heads + head) * cf.n + q0 + r1];
    }
    delta0 += __shfl_xor_sync(0xffffffffu, delta0, 1); delta1 += __shfl_xor_sync(0xffffffffu, delta1, 1);
    delta0 += __shfl_xor_sync(0xffffffffu, delta0, 2); delta1 += __shfl_xor_sync(0xffffffffu, delta1, 2);
  } else {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        gL1[a][c] = 0.f; gX2[a][c] = 0.f; gX3[a][c] = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) gW2[a][c][e] = 0.f;
      }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int k0_ = 16 * s + 2 * t;
      const bool on = g < CD;
      w1b[s][0] = pack_bf16(on ? W1[k0_ * CD + g] : 0.f, on ? W1[(k0_ + 1) * CD + g] : 0.f);
      w1b[s][1] = pack_bf16(on ? W1[(k0_ + 8) * CD + g] : 0.f, on ? W1[(k0_ + 9) * CD + g] : 0.f);
      w3f[s][0] = pack_bf16(W3[o * CPB_H + k0_], W3[o * CPB_H + k0_ + 1]);
      w3f[s][1] = pack_bf16(W3[o * CPB_H + k0_ + 8], W3[o * CPB_H + k0_ + 9]);
    }
  }
  const float keep_scale = cf.p_drop > 0.f ? 1.f / (1.f - cf.p_drop) : 1.f;
  const uint32_t thresh = (uint32_t)(cf.p_drop * 4294967296.0);
  const float* c1 = sm.c1[t];
  const float* c2 = sm.c2[t];
  float (*stS)[9] = sm.stg[w][0];
  float (*stD)[9] = sm.stg[w][1];
  const int qlane = lane & ~3;

  for (int k0 = 0; k0 < cf.nkv; k0 += BK) {
    __syncthreads();
    if (producer) {
      load_tile<false, BK>(sm.K, nullptr, kb, cf.inner, k0, cf.nkv);
      load_tile<true, BK>(sm.V[0], sm.V[1], vb, cf.inner, k0, cf.nkv);
    } else if (ptid < BK) {
      const bool ok = k0 + ptid < cf.nkv;
      sm.kpos[ptid][0] = ok ? vs[((int64_t)bg * cf.nkv + k0 + ptid) * CD] : 0.f;
      sm.kpos[ptid][1] = (ok && CD == 2) ? vs[((int64_t)bg * cf.nkv + k0 + ptid) * CD + 1] : 0.f;
      sm.dkpos[ptid][0] = 0.f; sm.dkpos[ptid][1] = 0.f;
    }
    __syncthreads();
    for (int e = tid; e < BK * MQ; e += WS_THREADS) {
      const int key = e / MQ, r = e % MQ;
      sm.SX[key][r] = slog_fast(sm.qx[r] - sm.kpos[key][0]);
    }
    if (CD == 2)
      for (int e = tid; e < BK * ny; e += WS_THREADS) {
        const int slot = e / BK, key = e % BK;
        sm.SY[slot][key] = slog_fast(sm.qy[slot] - sm.kpos[key][1]);
      }
    __syncthreads();
    if (producer) {
      // ================================================================================================ producer
#pragma unroll 1
      for (int j = 0; j < BK / 8; ++j) {
        {                                            // S = Q K^T and dP = dO V^T of my 16 rows x this 8-key group
          float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t ah[4], gh[4], gl[4];
            const int c0 = 16 * ks + 2 * t;
            ah[0] = *reinterpret_cast<const uint32_t*>(&sm.Q[r0][c0]);     ah[1] = *reinterpret_cast<const uint32_t*>(&sm.Q[r1][c0]);
            ah[2] = *reinterpret_cast<const uint32_t*>(&sm.Q[r0][c0 + 8]); ah[3] = *reinterpret_cast<const uint32_t*>(&sm.Q[r1][c0 + 8]);
            gh[0] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r0][c0]);     gh[1] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r1][c0]);
            gh[2] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r0][c0 + 8]); gh[3] = *reinterpret_cast<const uint32_t*>(&sm.dO[0][r1][c0 + 8]);
            gl[0] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r0][c0]);     gl[1] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r1][c0]);
            gl[2] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r0][c0 + 8]); gl[3] = *reinterpret_cast<const uint32_t*>(&sm.dO[1][r1][c0 + 8]);
            const uint32_t kh0 = *reinterpret_cast<const uint32_t*>(&sm.K[8 * j + g][c0]);
            const uint32_t kh1 = *reinterpret_cast<const uint32_t*>(&sm.K[8 * j + g][c0 + 8]);
            const uint32_t vh0 = *reinterpret_cast<const uint32_t*>(&sm.V[0][8 * j + g][c0]);
            const uint32_t vh1 = *reinterpret_cast<const uint32_t*>(&sm.V[0][8 * j + g][c0 + 8]);
            const uint32_t vl0 = *reinterpret_cast<const uint32_t*>(&sm.V[1][8 * j + g][c0]);
            const uint32_t vl1 = *reinterpret_cast<const uint32_t*>(&sm.V[1][8 * j + g][c0 + 8]);
            mma16816(s, ah, kh0, kh1);
            mma3<true>(dp, gh, gl, vh0, vh1, vl0, vl1);
          }
          __syncwarp();
          stS[g][2 * t] = s[0]; stS[g][2 * t + 1] = s[1]; stS[g + 8][2 * t] = s[2]; stS[g + 8][2 * t + 1] = s[3];
          stD[g][2 * t] = dp[0]; stD[g][2 * t + 1] = dp[1]; stD[g + 8][2 * t] = dp[2]; stD[g + 8][2 * t + 1] = dp[3];
          __syncwarp();
        }
#pragma unroll 1
        for (int u = 0; u < 8; ++u, ++kc) {
          const int key = 8 * j + u;
          const float ax0 = sm.SX[key][r0], ax1 = sm.SX[key][r1];
          const float ay0 = CD == 2 ? sm.SY[ys0][key] : 0.f, ay1 = CD == 2 ? sm.SY[ys1][key] : 0.f;
          uint32_t ah[2][4];
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const float4 wx = *reinterpret_cast<const float4*>(c1 + 4 * s), wy = *reinterpret_cast<const float4*>(c1 + 8 + 4 * s),
                         bb = *reinterpret_cast<const float4*>(c1 + 16 + 4 * s);
            const float wxa[4] = {wx.x, wx.y, wx.z, wx.w}, wya[4] = {wy.x, wy.y, wy.z, wy.w}, bba[4] = {bb.x, bb.y, bb.z, bb.w};
            float h0[4], h1[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float u0 = fmaf(wxa[e], ax0, bba[e]), u1 = fmaf(wxa[e], ax1, bba[e]);
              if (CD == 2) { u0 = fmaf(wya[e], ay0, u0); u1 = fmaf(wya[e], ay1, u1); }
              h0[e] = u0; h1[e] = u1;
            }
            ah[s][0] = relu_pack(h0[0], h0[1]); ah[s][1] = relu_pack(h1[0], h1[1]);
            ah[s][2] = relu_pack(h0[2], h0[3]); ah[s][3] = relu_pack(h1[2], h1[3]);
          }
          uint32_t hh[2][4];
          {
            const float4 ba = *reinterpret_cast<const float4*>(c2), bbv = *reinterpret_cast<const float4*>(c2 + 4);
            const float b2v[8] = {ba.x, ba.y, ba.z, ba.w, bbv.x, bbv.y, bbv.z, bbv.w};
            float acc[4][4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              mma16816_c(acc[jj], ah[0], wh[jj][0][0], wh[jj][0][1], b2v[2 * jj], b2v[2 * jj + 1], b2v[2 * jj], b2v[2 * jj + 1]);
              mma16816(acc[jj], ah[1], wh[jj][1][0], wh[jj][1][1]);
            }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              hh[s][0] = relu_pack(acc[2 * s][0], acc[2 * s][1]);         hh[s][1] = relu_pack(acc[2 * s][2], acc[2 * s][3]);
              hh[s][2] = relu_pack(acc[2 * s + 1][0], acc[2 * s + 1][1]); hh[s][3] = relu_pack(acc[2 * s + 1][2], acc[2 * s + 1][3]);
            }
          }
          float p0, p1;
          {
            float d3[4];
            mma16816_c(d3, hh[0], w3c[0], w3c[1], 0.f, 0.f, 0.f, 0.f);
            mma16816(d3, hh[1], w3c[2], w3c[3]);
            p0 = __shfl_sync(0xffffffffu, d3[0], qlane);
            p1 = __shfl_sync(0xffffffffu, d3[2], qlane);
          }
          const bool kok = k0 + key < cf.nkv;
          const float P0 = kok ? __expf(stS[g][u] + p0 + b3 - lse0) : 0.f;
          const float P1 = kok ? __expf(stS[g + 8][u] + p1 + b3 - lse1) : 0.f;
          float keep0 = 1.f, keep1 = 1.f;
          if (cf.p_drop > 0.f) {
            keep0 = pair_hash(salted(cf.seed, cf.salt), hrow0 + (unsigned long long)(k0 + key)) >= thresh ? keep_scale : 0.f;
            keep1 = pair_hash(salted(cf.seed, cf.salt), hrow1 + (unsigned long long)(k0 + key)) >= thresh ? keep_scale : 0.f;
          }
          const float dS0 = P0 * (stD[g][u] * keep0 - delta0);
          const float dS1 = P1 * (stD[g + 8][u] * keep1 - delta1);
          {
            const __nv_bfloat16 pv = __float2bfloat16_rn(t ? P1 * keep1 : P0 * keep0), sv = __float2bfloat16_rn(t ? dS1 : dS0);
            __nv_bfloat16* pp = &sm.P[t ? r1 : r0][key];
            if (t < 2) { *pp = pv; *(pp + (&sm.dS[0][0] - &sm.P[0][0])) = sv; }
          }
          gb3 += t == 0 ? dS0 + dS1 : 0.f;
          // hand over to the consumer of this pair
          const uint32_t slot = kc & 1u;
          ws_mbar_wait(&sm.empty[w][slot], ((kc >> 1) & 1u) ^ 1u);
          uint4* hb = &sm.hb[w][slot][0][lane];
          hb[0 * 32] = make_uint4(ah[0][0], ah[0][1], ah[0][2], ah[0][3]);
          hb[1 * 32] = make_uint4(ah[1][0], ah[1][1], ah[1][2], ah[1][3]);
          hb[2 * 32] = make_uint4(hh[0][0], hh[0][1], hh[0][2], hh[0][3]);
          hb[3 * 32] = make_uint4(hh[1][0], hh[1][1], hh[1][2], hh[1][3]);
          hb[4 * 32] = make_uint4(__float_as_uint(dS0), __float_as_uint(dS1), __float_as_uint(ax0), __float_as_uint(ay0));
          hb[5 * 32] = make_uint4(__float_as_uint(ax1), __float_as_uint(ay1), 0u, 0u);
          __syncwarp();
          if (lane == 0) ws_mbar_arrive(&sm.full[w][slot]);
        }
      }
    } else {
      // ================================================================================================ consumer
#pragma unroll 1
      for (int j = 0; j < BK / 8; ++j) {
#pragma unroll 1
        for (int u = 0; u < 8; ++u, ++kc) {
          const uint32_t slot = kc & 1u;
          ws_mbar_wait(&sm.full[w][slot], (kc >> 1) & 1u);
          uint32_t ah[2][4], hh[2][4];
          float dS0, dS1, ax0, ay0, ax1, ay1;
          {
            const uint4* hb = &sm.hb[w][slot][0][lane];
            const uint4 a0 = hb[0 * 32], a1 = hb[1 * 32], h0 = hb[2 * 32], h1 = hb[3 * 32], m0 = hb[4 * 32], m1 = hb[5 * 32];
            ah[0][0] = a0.x; ah[0][1] = a0.y; ah[0][2] = a0.z; ah[0][3] = a0.w;
            ah[1][0] = a1.x; ah[1][1] = a1.y; ah[1][2] = a1.z; ah[1][3] = a1.w;
            hh[0][0] = h0.x; hh[0][1] = h0.y; hh[0][2] = h0.z; hh[0][3] = h0.w;
            hh[1][0] = h1.x; hh[1][1] = h1.y; hh[1][2] = h1.z; hh[1][3] = h1.w;
            dS0 = __uint_as_float(m0.x); dS1 = __uint_as_float(m0.y); ax0 = __uint_as_float(m0.z); ay0 = __uint_as_float(m0.w);
            ax1 = __uint_as_float(m1.x); ay1 = __uint_as_float(m1.y);
          }
          __syncwarp();
          if (lane == 0) ws_mbar_arrive(&sm.empty[w][slot]);
          // dh2 = 1[h2 > 0] (dS w3): packed bf16x2, A-fragment order
          const uint32_t dS0p = pack_bf16(dS0, dS0), dS1p = pack_bf16(dS1, dS1);
          uint32_t dah[2][4];
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            dah[s][0] = nz_mask_bf16x2(hh[s][0]) & mul_bf16x2(dS0p, w3f[s][0]);
            dah[s][1] = nz_mask_bf16x2(hh[s][1]) & mul_bf16x2(dS1p, w3f[s][0]);
            dah[s][2] = nz_mask_bf16x2(hh[s][2]) & mul_bf16x2(dS0p, w3f[s][1]);
            dah[s][3] = nz_mask_bf16x2(hh[s][3]) & mul_bf16x2(dS1p, w3f[s][1]);
          }
          uint32_t bx0, bx1;
          {
            const uint32_t xa0 = pack_bf16(ax0, ay0), xa1 = pack_bf16(ax1, ay1), xb0 = pack_bf16(1.f, dS0), xb1 = pack_bf16(1.f, dS1);
            const uint32_t x0 = t == 0 ? xa0 : (t == 1 ? xb0 : 0u), x1 = t == 0 ? xa1 : (t == 1 ? xb1 : 0u);
            bx0 = movm_trans(x0); bx1 = movm_trans(x1);
          }
          {
            uint32_t bh[4][2];
#pragma unroll
            for (int jn = 0; jn < 4; ++jn) {
              const int s = jn >> 1, rb = (jn & 1) * 2;
              bh[jn][0] = movm_trans(ah[s][rb]); bh[jn][1] = movm_trans(ah[s][rb + 1]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              uint32_t ta[4];
              ta[0] = movm_trans(dah[mt][0]); ta[1] = movm_trans(dah[mt][2]); ta[2] = movm_trans(dah[mt][1]); ta[3] = movm_trans(dah[mt][3]);
#pragma unroll
              for (int jn = 0; jn < 4; ++jn) mma16816(gW2[mt][jn], ta, bh[jn][0], bh[jn][1]);
              mma16816(gX2[mt], ta, bx0, bx1);
              uint32_t th[4];
              th[0] = movm_trans(hh[mt][0]); th[1] = movm_trans(hh[mt][2]); th[2] = movm_trans(hh[mt][1]); th[3] = movm_trans(hh[mt][3]);
              mma16816(gX3[mt], th, bx0, bx1);
            }
          }
          uint32_t peh[4][2];
#pragma unroll
          for (int jn = 0; jn < 4; ++jn) {
            float d1[4];
            mma16816_c(d1, dah[0], sm.wth[(jn * 2 + 0) * 2][lane], sm.wth[(jn * 2 + 0) * 2 + 1][lane], 0.f, 0.f, 0.f, 0.f);
            mma16816(d1, dah[1], sm.wth[(jn * 2 + 1) * 2][lane], sm.wth[(jn * 2 + 1) * 2 + 1][lane]);
            peh[jn][0] = nz_mask_bf16x2(ah[jn >> 1][(jn & 1) * 2]) & pack_bf16(d1[0], d1[1]);
            peh[jn][1] = nz_mask_bf16x2(ah[jn >> 1][(jn & 1) * 2 + 1]) & pack_bf16(d1[2], d1[3]);
          }
          float dpo[4];
          {
            const uint32_t pa0[4] = {peh[0][0], peh[0][1], peh[1][0], peh[1][1]};
            const uint32_t pa1[4] = {peh[2][0], peh[2][1], peh[3][0], peh[3][1]};
            mma16816_c(dpo, pa0, w1b[0][0], w1b[0][1], 0.f, 0.f, 0.f, 0.f);
            mma16816(dpo, pa1, w1b[1][0], w1b[1][1]);
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            uint32_t ta[4];
            ta[0] = movm_trans(peh[2 * mt][0]); ta[1] = movm_trans(peh[2 * mt + 1][0]);
            ta[2] = movm_trans(peh[2 * mt][1]); ta[3] = movm_trans(peh[2 * mt + 1][1]);
            mma16816(gL1[mt], ta, bx0, bx1);
          }
          {
            const float kgx = -(dpo[0] * __expf(-fabsf(ax0)) + dpo[2] * __expf(-fabsf(ax1)));
            const float kgy = CD == 2 ? -(dpo[1] * __expf(-fabsf(ay0)) + dpo[3] * __expf(-fabsf(ay1))) : 0.f;
            if (t == 0) *reinterpret_cast<float2*>(&sm.kgs[w][g][2 * u]) = make_float2(kgx, kgy);
          }
        }
        __syncwarp();
        if (lane < 16) {
          float r = 0.f;
#pragma unroll
          for (int gg = 0; gg < 8; ++gg) r += sm.kgs[w][gg][lane];
          if (CD == 2 || (lane & 1) == 0) atomicAdd(&sm.dkpos[8 * j + (lane >> 1)][lane & 1], r);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (producer) {
      // ---- dQ += dS K
      float dqacc[8][4];
#pragma unroll
      for (int jd = 0; jd < 8; ++jd)
#pragma unroll
        for (int e = 0; e < 4; ++e) dqacc[jd][e] = sm.dQ[jd * 4 + e][ptid];
#pragma unroll
      for (int ks = 0; ks < BK / 16; ++ks) {
        uint32_t ah[4];
        const int c0 = 16 * ks + 2 * t;
        ah[0] = *reinterpret_cast<const uint32_t*>(&sm.dS[r0][c0]);     ah[1] = *reinterpret_cast<const uint32_t*>(&sm.dS[r1][c0]);
        ah[2] = *reinterpret_cast<const uint32_t*>(&sm.dS[r0][c0 + 8]); ah[3] = *reinterpret_cast<const uint32_t*>(&sm.dS[r1][c0 + 8]);
#pragma unroll
        for (int jd = 0; jd < 8; ++jd) {
          uint32_t bh0, bh1;
          ldmatrix_x2_trans(bh0, bh1, &sm.K[16 * ks + (lane & 15)][8 * jd]);
          mma16816(dqacc[jd], ah, bh0, bh1);
        }
      }
#pragma unroll
      for (int jd = 0; jd < 8; ++jd)
#pragma unroll
        for (int e = 0; e < 4; ++e) sm.dQ[jd * 4 + e][ptid] = dqacc[jd][e];
      if (ptid < BK * CD) {
        const int key = ptid / CD, c = ptid % CD;
        if (k0 + key < cf.nkv) atomicAdd(dvs + ((int64_t)bg * cf.nkv + k0 + key) * CD + c, sm.dkpos[key][c]);
      }
    } else {
      // ---- dK += dS^T Q and dV += (P keep)^T dO: warp slice = 16 keys x 32 dims, contraction over the CTA's 64 queries
      const int mt = w & 1, dh = w >> 1;
      float dkacc[4][4], dvacc[4][4];
#pragma unroll
      for (int jd = 0; jd < 4; ++jd)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dkacc[jd][e] = 0.f; dvacc[jd][e] = 0.f; }
      const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int qrow = 16 * ks + rr + 8 * (mi >> 1), kcol = 16 * mt + 8 * (mi & 1);
        uint32_t sh[4], ph[4];
        ldmatrix_x4_trans(sh[0], sh[1], sh[2], sh[3], &sm.dS[qrow][kcol]);
        ldmatrix_x4_trans(ph[0], ph[1], ph[2], ph[3], &sm.P[qrow][kcol]);
#pragma unroll
        for (int jd = 0; jd < 4; ++jd) {
          uint32_t qh0, qh1, gh0, gh1;
          ldmatrix_x2_trans(qh0, qh1, &sm.Q[16 * ks + (lane & 15)][32 * dh + 8 * jd]);
          ldmatrix_x2_trans(gh0, gh1, &sm.dO[0][16 * ks + (lane & 15)][32 * dh + 8 * jd]);
          mma16816(dkacc[jd], sh, qh0, qh1);
          mma16816(dvacc[jd], ph, gh0, gh1);
        }
      }
      const int key0 = k0 + 16 * mt + g, key1 = key0 + 8;
#pragma unroll
      for (int jd = 0; jd < 4; ++jd) {
        const int col = head * AD + 32 * dh + 8 * jd + 2 * t;
        if (key0 < cf.nkv) {
          const int64_t off = ((int64_t)b * cf.nkv + key0) * cf.inner + col;
          atomicAdd(reinterpret_cast<float2*>(dk + off), make_float2(dkacc[jd][0], dkacc[jd][1]));     // red.global.add.v2.f32
          atomicAdd(reinterpret_cast<float2*>(dv + off), make_float2(dvacc[jd][0], dvacc[jd][1]));
        }
        if (key1 < cf.nkv) {
          const int64_t off = ((int64_t)b * cf.nkv + key1) * cf.inner + col;
          atomicAdd(reinterpret_cast<float2*>(dk + off), make_float2(dkacc[jd][2], dkacc[jd][3]));
          atomicAdd(reinterpret_cast<float2*>(dv + off), make_float2(dvacc[jd][2], dvacc[jd][3]));
        }
      }
    }
  }
  // ---- dQ
  if (producer) {
    if (ok0) {
      float* row = dq + qoff + (int64_t)(q0 + r0) * cf.inner + 2 * t;
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(row + 8 * j) = make_float2(sm.dQ[j * 4][ptid], sm.dQ[j * 4 + 1][ptid]);
    }
    if (ok1) {
      float* row = dq + qoff + (int64_t)(q0 + r1) * cf.inner + 2 * t;
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(row + 8 * j) = make_float2(sm.dQ[j * 4 + 2][ptid], sm.dQ[j * 4 + 3][ptid]);
    }
  }
  // ---- CPB parameter partials of this CTA: [W1 32*CD][b1 32][W2 1024][b2 32][W3 OUT*32][b3 OUT]
  __syncthreads();
  float* const red = reinterpret_cast<float*>(&sm.P[0][0]);
  static_assert(sizeof(WsBwdSmem<CD>::P) + sizeof(WsBwdSmem<CD>::dS) >= NRED * sizeof(float), "the reduction buffer aliases P + dS");
  for (int i = tid; i < NRED; i += WS_THREADS) red[i] = 0.f;
  __syncthreads();
  {
    float* rW1 = red; float* rb1 = rW1 + CPB_H * 2; float* rW2 = rb1 + CPB_H; float* rb2 = rW2 + CPB_H * CPB_H;
    float* rw3 = rb2 + CPB_H; float* rb3 = rw3 + CPB_H;
    if (producer) {
      gb3 = warp_sum(gb3);
      if (lane == 0) atomicAdd(rb3, gb3);
    } else {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int kk = 16 * mt + g;
        if (t == 0) {
          atomicAdd(&rW1[kk * 2], gL1[mt][0]); atomicAdd(&rW1[kk * 2 + 1], gL1[mt][1]);
          atomicAdd(&rW1[(kk + 8) * 2], gL1[mt][2]); atomicAdd(&rW1[(kk + 8) * 2 + 1], gL1[mt][3]);
        } else if (t == 1) {
          atomicAdd(&rb1[kk], gL1[mt][0]); atomicAdd(&rb1[kk + 8], gL1[mt][2]);
          atomicAdd(&rb2[kk], gX2[mt][0]); atomicAdd(&rb2[kk + 8], gX2[mt][2]);
          atomicAdd(&rw3[kk], gX3[mt][1]); atomicAdd(&rw3[kk + 8], gX3[mt][3]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
          const int c = 16 * mt + g, kk = 8 * jn + 2 * t;
          atomicAdd(&rW2[c * CPB_H + kk], gW2[mt][jn][0]); atomicAdd(&rW2[c * CPB_H + kk + 1], gW2[mt][jn][1]);
          atomicAdd(&rW2[(c + 8) * CPB_H + kk], gW2[mt][jn][2]); atomicAdd(&rW2[(c + 8) * CPB_H + kk + 1], gW2[mt][jn][3]);
        }
    }
  }
  __syncthreads();
  {
    const int np = CPB_H * CD + CPB_H + CPB_H * CPB_H + CPB_H + cf.OUT * CPB_H + cf.OUT;
    float* prow = part + ((int64_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * np;
    const float* rW1 = red; const float* rb1 = rW1 + CPB_H * 2; const float* rW2 = rb1 + CPB_H;
    const float* rb2 = rW2 + CPB_H * CPB_H; const float* rw3 = rb2 + CPB_H; const float* rb3 = rw3 + CPB_H;
    for (int i = tid; i < np; i += WS_THREADS) {
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

template <int CD, bool SPLIT>
static int launch_fwd(const AttnCfg& cf, const float* q, const float* k, const float* v, const float* vs, const float* cpb,
                      float* out, float* lse, void* stream) {
  dim3 grid((cf.n + MQ - 1) / MQ, cf.heads, cf.B);
  const size_t smem = sizeof(FwdSmem<SPLIT>);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(deform_attn_fwd_mma_kernel<CD, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  deform_attn_fwd_mma_kernel<CD, SPLIT><<<grid, MTHREADS, smem, as_stream(stream)>>>(cf, q, k, v, vs, cpb, out, lse);
  SML_LAUNCH_OK("sml_deform_attn_fwd_tc");
  return 0;
}

template <int CD, int MINB>
static int launch_fwd_lean(const AttnCfg& cf, const float* q, const float* k, const float* v, const float* vs, const float* cpb,
                           float* out, float* lse, void* stream) {
  dim3 grid((cf.n + MQ - 1) / MQ, cf.heads, cf.B);
  const size_t smem = sizeof(LeanFwdSmem<CD>);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(deform_attn_fwd_lean_kernel<CD, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  deform_attn_fwd_lean_kernel<CD, MINB><<<grid, MTHREADS, smem, as_stream(stream)>>>(cf, q, k, v, vs, cpb, out, lse);
  SML_LAUNCH_OK("sml_deform_attn_fwd_tc (lean)");
  return 0;
}

/* floats of the key-range scratch of the table kernels: (min, max) per 32-key tile and (batch, group) */
extern "C" int64_t sml_deform_pwl_krange_floats(const int32_t* cfg8) {
  return 2 * (int64_t)cfg8[0] * cfg8[4] * ((cfg8[3] + 31) / 32);
}
/* krange [sml_deform_pwl_krange_floats] <- per-tile ranges of the key positions vs [(B G), nkv]; once per forward, reused by the backward */
extern "C" int sml_deform_pwl_key_ranges(const int32_t* cfg8, const float* vs, float* krange, void* stream) {
  AttnCfg cf;
  if (make_acfg(cf, cfg8, 0.f, 0)) return 1;
  SML_CHECK(cf.CD == 1 && vs && krange, "sml_deform_pwl_key_ranges: one coordinate dimension, non-null pointers");
  static_assert(BK == 32, "the key-range scratch is indexed by 32-key tiles");
  const int nt = (cf.nkv + 31) / 32;
  const int64_t total = (int64_t)cf.B * cf.G * nt;
  pwl_key_range_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(vs, cf.nkv, nt, total, reinterpret_cast<float2*>(krange));
  SML_LAUNCH_OK("sml_deform_pwl_key_ranges");
  return 0;
}
/* 1-D only (CD = 1): forward with the bias MLP replaced by its piecewise-linear table (sml_cpb1d_table_build).  kh, vh = bf16
 * planes of K and V (same [B, nkv, heads * 64] layout), q fp32; krange from sml_deform_pwl_key_ranges */
extern "C" int sml_deform_attn_fwd_pwl(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const void* kh, const void* vh,
                                       const float* vs, const float* krange, const float* brk, const float* coef, const int32_t* nseg,
                                       float* out, float* lse, void* stream) {
  AttnCfg cf;
  if (make_acfg(cf, cfg8, p_drop, seed)) return 1;
  SML_CHECK(cf.CD == 1, "sml_deform_attn_fwd_pwl: one coordinate dimension only");
  dim3 grid((cf.n + MQ - 1) / MQ, cf.heads, cf.B);
  const size_t smem = sizeof(PwlFwdSmem);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(deform_attn_fwd_pwl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(deform_attn_fwd_pwl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const __nv_bfloat16 *kp = (const __nv_bfloat16*)kh, *vp = (const __nv_bfloat16*)vh;
  const float2* kr = reinterpret_cast<const float2*>(krange);
  if (p_drop > 0.f) deform_attn_fwd_pwl_kernel<true><<<grid, MTHREADS, smem, as_stream(stream)>>>(cf, q, kp, vp, vs, kr, brk, coef, nseg, out, lse);
  else deform_attn_fwd_pwl_kernel<false><<<grid, MTHREADS, smem, as_stream(stream)>>>(cf, q, kp, vp, vs, kr, brk, coef, nseg, out, lse);
  SML_LAUNCH_OK("sml_deform_attn_fwd_pwl");
  return 0;
}

/* backward of sml_deform_attn_fwd_pwl (kh, vh, krange = the same planes and scratch).  dk, dv, dvs and gseg [OUT,
 * sml_cpb1d_seg_rows(), 2] zero-initialised by the caller; gseg[o][s] = (sum dS, sum dS p) over the pairs of the heads with
 * output o whose distance p lies in segment s */
extern "C" int sml_deform_attn_bwd_pwl(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const void* kh, const void* vh,
                                       const float* vs, const float* krange, const float* brk, const float* coef, const int32_t* nseg,
                                       const float* out, const float* lse, const float* dout, float* dq, float* dk, float* dv,
                                       float* dvs, float* gseg, void* stream) {
  AttnCfg cf;
  if (make_acfg(cf, cfg8, p_drop, seed)) return 1;
  SML_CHECK(cf.CD == 1, "sml_deform_attn_bwd_pwl: one coordinate dimension only");
  dim3 grid((cf.n + MQ - 1) / MQ, cf.heads, cf.B);
  const size_t smem = sizeof(PwlBwdSmem);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(deform_attn_bwd_pwl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(deform_attn_bwd_pwl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const __nv_bfloat16 *kp = (const __nv_bfloat16*)kh, *vp = (const __nv_bfloat16*)vh;
  const float2* kr = reinterpret_cast<const float2*>(krange);
  if (p_drop > 0.f)
    deform_attn_bwd_pwl_kernel<true><<<grid, MTHREADS, smem, as_stream(stream)>>>(cf, q, kp, vp, vs, kr, brk, coef, nseg, out, lse, dout, dq,
                                                                                  dk, dv, dvs, gseg);
  else
    deform_attn_bwd_pwl_kernel<false><<<grid, MTHREADS, smem, as_stream(stream)>>>(cf, q, kp, vp, vs, kr, brk, coef, nseg, out, lse, dout, dq,
                                                                                   dk, dv, dvs, gseg);
  SML_LAUNCH_OK("sml_deform_attn_bwd_pwl");
  return 0;
}

/* split: 0 = plain bf16 operands (lean kernel), 1 = bf16 hi/lo 3-pass (fp32-grade), 2 = plain bf16, round-1 kernel (kept for A/B) */
extern "C" int sml_deform_attn_fwd_tc(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const float* k,
                                      const float* v, const float* vs, const float* cpb, float* out, float* lse, int32_t split,
                                      void* stream) {
  AttnCfg cf;
  if (make_acfg(cf, cfg8, p_drop, seed)) return 1;
  if (split == 0) return cf.CD == 2 ? launch_fwd_lean<2, 3>(cf, q, k, v, vs, cpb, out, lse, stream)
                                    : launch_fwd_lean<1, 3>(cf, q, k, v, vs, cpb, out, lse, stream);
  if (cf.CD == 2) return split == 1 ? launch_fwd<2, true>(cf, q, k, v, vs, cpb, out, lse, stream)
                                    : launch_fwd<2, false>(cf, q, k, v, vs, cpb, out, lse, stream);
  return split == 1 ? launch_fwd<1, true>(cf, q, k, v, vs, cpb, out, lse, stream)
                    : launch_fwd<1, false>(cf, q, k, v, vs, cpb, out, lse, stream);
}

template <int CD, bool SPLIT>
static int launch_bwd(const AttnCfg& cf, const float* q, const float* k, const float* v, const float* vs, const float* cpb,
                      const float* out, const float* lse, const float* dout, float* dq, float* dk, float* dv, float* dvs,
                      float* part, void* stream) {
  dim3 grid((cf.n + MQ - 1) / MQ, cf.heads, cf.B);
  const size_t smem = sizeof(BwdSmem<SPLIT>);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(deform_attn_bwd_mma_kernel<CD, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  deform_attn_bwd_mma_kernel<CD, SPLIT><<<grid, MTHREADS, smem, as_stream(stream)>>>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk,
                                                                                     dv, dvs, part);
  SML_LAUNCH_OK("sml_deform_attn_bwd_tc");
  return 0;
}

template <int CD, int UNR>
static int launch_bwd_lean(const AttnCfg& cf, const float* q, const float* k, const float* v, const float* vs, const float* cpb,
                           const float* out, const float* lse, const float* dout, float* dq, float* dk, float* dv, float* dvs,
                           float* part, void* stream) {
  dim3 grid((cf.n + MQ - 1) / MQ, cf.heads, cf.B);
  const size_t smem = sizeof(LeanBwdSmem<CD>);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(deform_attn_bwd_lean_kernel<CD, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  deform_attn_bwd_lean_kernel<CD, UNR><<<grid, MTHREADS, smem, as_stream(stream)>>>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs,
                                                                              part);
  SML_LAUNCH_OK("sml_deform_attn_bwd_tc (lean)");
  return 0;
}

template <int CD>
static int launch_bwd_ws(const AttnCfg& cf, const float* q, const float* k, const float* v, const float* vs, const float* cpb,
                         const float* out, const float* lse, const float* dout, float* dq, float* dk, float* dv, float* dvs,
                         float* part, void* stream) {
  dim3 grid((cf.n + MQ - 1) / MQ, cf.heads, cf.B);
  const size_t smem = sizeof(WsBwdSmem<CD>);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(deform_attn_bwd_ws_kernel<CD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  deform_attn_bwd_ws_kernel<CD><<<grid, WS_THREADS, smem, as_stream(stream)>>>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs, part);
  SML_LAUNCH_OK("sml_deform_attn_bwd_tc (warp-specialised)");
  return 0;
}

/* same contract as sml_deform_attn_bwd (dk, dv, dvs zero-initialised; part [sml_deform_attn_bwd_blocks, nparams]);
 * split as in sml_deform_attn_fwd_tc; split = 5: plain bf16, warp-specialised producer / consumer variant (experimental) */
extern "C" int sml_deform_attn_bwd_tc(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const float* k,
                                      const float* v, const float* vs, const float* cpb, const float* out, const float* lse,
                                      const float* dout, float* dq, float* dk, float* dv, float* dvs, float* part, int32_t split,
                                      void* stream) {
  AttnCfg cf;
  if (make_acfg(cf, cfg8, p_drop, seed)) return 1;
  // split = 5: the warp-specialised producer / consumer variant (correct, parity-tested through this entry; measured 77 ms
  // against 49.5 ms for the lean kernel at B = 8 / 16 k: at 128 registers per thread both roles spill and the hand-over
  // costs more than the doubled occupancy returns - see DESIGN.md 3.1)
  if (split == 5 && (cf.CD == 1 || cf.W >= 10))      // y table: a 64-token tile spans <= 8 lattice rows
    return cf.CD == 2 ? launch_bwd_ws<2>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs, part, stream)
                      : launch_bwd_ws<1>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs, part, stream);
  if (split == 0 || split == 4 || split == 5)
    return cf.CD == 2 ? launch_bwd_lean<2, 1>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs, part, stream)
                      : launch_bwd_lean<1, 1>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs, part, stream);
  if (cf.CD == 2) return split == 1 ? launch_bwd<2, true>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs, part, stream)
                                    : launch_bwd<2, false>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs, part, stream);
  return split == 1 ? launch_bwd<1, true>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs, part, stream)
                    : launch_bwd<1, false>(cf, q, k, v, vs, cpb, out, lse, dout, dq, dk, dv, dvs, part, stream);
}

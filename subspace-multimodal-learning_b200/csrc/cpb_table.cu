// cpb_table.cu — the 1-D continuous-position-bias MLP (models/DeformableAttention1D.py:60-102) as an exact piecewise-linear
// table of t = sign(dx) log(|dx| + 1):  Linear(1,32) ReLU Linear(32,32) ReLU Linear(32,OUT) of ONE scalar is continuous and
// piecewise linear with at most 32 + 33*32 breakpoints.  sml_cpb1d_table_build finds them on the device (one block) and
// stores (slope, intercept) per segment and output; evaluation is a bounded binary search + one FMA per output instead of
// ~1100 FMAs per (query, key) pair.  Specification + CPU check: oracle/pwl.py.  (Round-1 status: table build + stand-alone
// evaluation; the table-based 1-D attention kernels are next)
#include "common.cuh"
#include <math_constants.h>

namespace sml {

constexpr int PWL_H = 32;
constexpr int PWL_NB = 2048;                 // sorted breakpoint slots (>= 32 + 33*32 = 1088), padded with +inf
constexpr int PWL_THREADS = 1024;

struct CpbView { const float *w1, *b1, *W2, *b2, *W3, *b3; };
__device__ __forceinline__ CpbView cpb_view(const float* cpb, int OUT) {
  CpbView v;
  v.w1 = cpb; v.b1 = v.w1 + PWL_H; v.W2 = v.b1 + PWL_H; v.b2 = v.W2 + PWL_H * PWL_H; v.W3 = v.b2 + PWL_H; v.b3 = v.W3 + OUT * PWL_H;
  return v;
}
__device__ __forceinline__ float seg_mid(float lo, float hi) {
  const bool li = lo == -CUDART_INF_F, hi_i = hi == CUDART_INF_F;
  if (li && hi_i) return 0.f;
  if (li) return hi - 1.f;
  if (hi_i) return lo + 1.f;
  return 0.5f * (lo + hi);
}
// layer-2 pre-activation of unit c as an affine function A t + C on the interval that contains tm
__device__ __forceinline__ void affine2(const float* w1, const float* b1, const float* W2row, float b2c, float tm, float& A, float& C) {
  float a = 0.f, cc = b2c;
#pragma unroll 8
  for (int k = 0; k < PWL_H; ++k) {
    const bool on = fmaf(w1[k], tm, b1[k]) > 0.f;
    if (on) { a = fmaf(W2row[k], w1[k], a); cc = fmaf(W2row[k], b1[k], cc); }
  }
  A = a; C = cc;
}

// brk [PWL_NB] ascending (+inf padded), coef [(PWL_NB + 1), OUT, 2] = (slope, intercept), nseg[0] = number of segments
__global__ void __launch_bounds__(PWL_THREADS) cpb1d_table_build_kernel(const float* __restrict__ cpb, int OUT,
                                                                        float* __restrict__ brk, float* __restrict__ coef,
                                                                        int* __restrict__ nseg) {
  __shared__ float sw1[PWL_H], sb1[PWL_H], sW2[PWL_H * PWL_H], sb2[PWL_H], sW3[2 * PWL_H], sb3[2];
  __shared__ float e1[PWL_H + 1];              // sorted layer-1 kinks
  __shared__ float sb[PWL_NB];
  __shared__ int n1s, nbs;
  const int tid = threadIdx.x;
  const CpbView v = cpb_view(cpb, OUT);
  for (int i = tid; i < PWL_H * PWL_H; i += PWL_THREADS) sW2[i] = v.W2[i];
  if (tid < PWL_H) { sw1[tid] = v.w1[tid]; sb1[tid] = v.b1[tid]; sb2[tid] = v.b2[tid]; }
  if (tid < OUT * PWL_H) sW3[tid] = v.W3[tid];
  if (tid < OUT) sb3[tid] = v.b3[tid];
  for (int i = tid; i < PWL_NB; i += PWL_THREADS) sb[i] = CUDART_INF_F;
  __syncthreads();
  // 1. layer-1 kinks, rank-sorted
  if (tid < PWL_H) {
    const float mine = sw1[tid] != 0.f ? -sb1[tid] / sw1[tid] : CUDART_INF_F;
    sb[tid] = mine;
  }
  __syncthreads();
  if (tid < PWL_H) {
    const float mine = sb[tid];
    int rank = 0;
    for (int j = 0; j < PWL_H; ++j) { const float o = sb[j]; rank += (o < mine) || (o == mine && j < tid); }
    e1[rank] = mine;
  }
  if (tid == 0) e1[PWL_H] = CUDART_INF_F;
  __syncthreads();
  if (tid == 0) { int c = 0; while (c < PWL_H && e1[c] != CUDART_INF_F) ++c; n1s = c; }
  __syncthreads();
  const int n1 = n1s;
  // 2. zero crossings of the layer-2 pre-activations inside every layer-1 interval
  for (int i = tid; i < (n1 + 1) * PWL_H; i += PWL_THREADS) {
    const int p = i / PWL_H, c = i % PWL_H;
    const float lo = p == 0 ? -CUDART_INF_F : e1[p - 1], hi = p == n1 ? CUDART_INF_F : e1[p];
    float A, C;
    affine2(sw1, sb1, sW2 + c * PWL_H, sb2[c], seg_mid(lo, hi), A, C);
    float tz = CUDART_INF_F;
    if (A != 0.f) { const float z = -C / A; if (z > lo && z < hi) tz = z; }
    sb[PWL_H + i] = tz;
  }
  __syncthreads();
  // 3. bitonic sort of the PWL_NB slots
  for (int k = 2; k <= PWL_NB; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < PWL_NB; i += PWL_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = sb[i], c = sb[ixj];
          if ((a > c) == ((i & k) == 0)) { sb[i] = c; sb[ixj] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = tid; i < PWL_NB; i += PWL_THREADS) brk[i] = sb[i];
  if (tid == 0) {
    int lo = 0, hi = PWL_NB;                   // number of finite breakpoints
    while (lo < hi) { const int m = (lo + hi) >> 1; if (sb[m] != CUDART_INF_F) lo = m + 1; else hi = m; }
    nbs = lo; nseg[0] = lo + 1;
  }
  __syncthreads();
  const int nb = nbs;
  // 4. per-segment slope / intercept of every output
  for (int s = tid; s <= nb; s += PWL_THREADS) {
    const float lo = s == 0 ? -CUDART_INF_F : sb[s - 1], hi = s == nb ? CUDART_INF_F : sb[s];
    const float tm = seg_mid(lo, hi);
    float sl[2] = {0.f, 0.f}, ic[2] = {sb3[0], OUT > 1 ? sb3[1] : 0.f};
    for (int c = 0; c < PWL_H; ++c) {
      float A, C;
      affine2(sw1, sb1, sW2 + c * PWL_H, sb2[c], tm, A, C);
      if (fmaf(A, tm, C) > 0.f) {
        for (int o = 0; o < OUT; ++o) { sl[o] = fmaf(sW3[o * PWL_H + c], A, sl[o]); ic[o] = fmaf(sW3[o * PWL_H + c], C, ic[o]); }
      }
    }
    for (int o = 0; o < OUT; ++o) { coef[((int64_t)s * OUT + o) * 2] = sl[o]; coef[((int64_t)s * OUT + o) * 2 + 1] = ic[o]; }
  }
}

// ---- parameter gradients from per-segment sums (oracle/pwl.py:param_grads_from_segments).  On segment s the ReLU pattern
// (m1 of layer 1, m2 of layer 2) is fixed, so bias_o(t) = sum_c W3[o,c] m2_c (A_c t + C_c) + b3_o with
// A_c = sum_k W2[c,k] m1_k w1_k, C_c = sum_k W2[c,k] m1_k b1_k + b2_c, and every d bias / d parameter is affine in t.  With
// G0[s,o] = sum dS and G1[s,o] = sum dS t over the segment's pairs (accumulated by deform_attn_bwd_pwl_kernel):
//   u0[s,c] = m2_c sum_o W3[o,c] G0[s,o], u1 likewise with G1          (gradient reaching layer 2's pre-activation c)
//   d b3_o = sum_s G0          d W3[o,c] = sum_s m2_c (A_c G1 + C_c G0)          d b2_c = sum_s u0
//   acc0[c,k] = sum_s m1_k u0[s,c], acc1 likewise       d W2[c,k] = w1_k acc1 + b1_k acc0
//   d b1_k = sum_c W2[c,k] acc0[c,k]                    d w1_k = sum_c W2[c,k] acc1[c,k]
// One block of 1024 threads = the (c, k) pairs; segments are walked in chunks of 32.
constexpr int PWL_SEG_STRIDE = 1090;         // rows of gseg per output: 32 + 33*32 + 1 segments, + 1 spare
__global__ void __launch_bounds__(PWL_THREADS) cpb1d_table_bwd_kernel(const float* __restrict__ cpb, int OUT,
                                                                      const float* __restrict__ brk, const int* __restrict__ nseg,
                                                                      const float* __restrict__ gseg, float* __restrict__ dcpb) {
  __shared__ float sw1[PWL_H], sb1[PWL_H], sW2[PWL_H * PWL_H], sb2[PWL_H], sW3[2 * PWL_H];
  __shared__ float u0[32][PWL_H], u1[32][PWL_H], Au[32][PWL_H], Cu[32][PWL_H];
  __shared__ float g0[32][2], g1[32][2];
  __shared__ unsigned m1s[32];
  __shared__ float red0[PWL_H][PWL_H + 1], red1[PWL_H][PWL_H + 1];
  const int tid = threadIdx.x;
  const int c = tid >> 5, k = tid & 31;        // my (layer-2 unit, layer-1 unit) pair; in phase A: (segment of the chunk, unit)
  const CpbView v = cpb_view(cpb, OUT);
  sW2[tid] = v.W2[tid];
  if (tid < PWL_H) { sw1[tid] = v.w1[tid]; sb1[tid] = v.b1[tid]; sb2[tid] = v.b2[tid]; }
  if (tid < OUT * PWL_H) sW3[tid] = v.W3[tid];
  const int ns = nseg[0], nb = ns - 1;
  float acc0 = 0.f, acc1 = 0.f;                // threads (c, k): sums for W2
  float aux0 = 0.f, aux1 = 0.f;                // threads < 32: d b2; threads 32 .. 32 + OUT*32: d W3; threads 128 .. 128 + OUT: d b3
  __syncthreads();
  for (int s0 = 0; s0 < ns; s0 += 32) {
    {                                          // phase A: thread (sl, cc) = (tid >> 5, tid & 31)
      const int sl = tid >> 5, cc = tid & 31, s = s0 + sl;
      float U0 = 0.f, U1 = 0.f, AA = 0.f, CC = 0.f;
      unsigned m1 = 0u;
      if (s < ns) {
        const float lo = s == 0 ? -CUDART_INF_F : brk[s - 1], hi = s == nb ? CUDART_INF_F : brk[s];
        const float tm = seg_mid(lo, hi);
        float A, C;
        affine2(sw1, sb1, sW2 + cc * PWL_H, sb2[cc], tm, A, C);
        m1 = __ballot_sync(0xffffffffu, fmaf(sw1[cc], tm, sb1[cc]) > 0.f);
        if (fmaf(A, tm, C) > 0.f) {
          AA = A; CC = C;
          for (int o = 0; o < OUT; ++o) {
            const float G0 = gseg[((int64_t)o * PWL_SEG_STRIDE + s) * 2], G1 = gseg[((int64_t)o * PWL_SEG_STRIDE + s) * 2 + 1];
            U0 = fmaf(sW3[o * PWL_H + cc], G0, U0); U1 = fmaf(sW3[o * PWL_H + cc], G1, U1);
          }
        }
      }
      u0[sl][cc] = U0; u1[sl][cc] = U1; Au[sl][cc] = AA; Cu[sl][cc] = CC;
      if (cc == 0) m1s[sl] = m1;
      if (cc < 2) {
        const bool ok = s < ns && cc < OUT;
        g0[sl][cc] = ok ? gseg[((int64_t)cc * PWL_SEG_STRIDE + s) * 2] : 0.f;
        g1[sl][cc] = ok ? gseg[((int64_t)cc * PWL_SEG_STRIDE + s) * 2 + 1] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int sl = 0; sl < 32; ++sl) {          // phase B
      if ((m1s[sl] >> k) & 1u) { acc0 += u0[sl][c]; acc1 += u1[sl][c]; }
    }
    if (tid < PWL_H) {
      for (int sl = 0; sl < 32; ++sl) aux0 += u0[sl][tid];
    } else if (tid < PWL_H + OUT * PWL_H) {
      const int o = (tid - PWL_H) >> 5, cc = tid & 31;
      for (int sl = 0; sl < 32; ++sl) aux0 = fmaf(Au[sl][cc], g1[sl][o], fmaf(Cu[sl][cc], g0[sl][o], aux0));
    } else if (tid >= 128 && tid < 128 + OUT) {
      for (int sl = 0; sl < 32; ++sl) aux1 += g0[sl][tid - 128];
    }
    __syncthreads();
  }
  // parameter order: w1 [32], b1 [32], W2 [32 x 32], b2 [32], W3 [OUT x 32], b3 [OUT]
  float* dw1 = dcpb; float* db1 = dw1 + PWL_H; float* dW2 = db1 + PWL_H; float* db2 = dW2 + PWL_H * PWL_H;
  float* dW3 = db2 + PWL_H; float* db3 = dW3 + OUT * PWL_H;
  dW2[tid] = fmaf(sw1[k], acc1, sb1[k] * acc0);
  red0[c][k] = sW2[tid] * acc0; red1[c][k] = sW2[tid] * acc1;
  if (tid < PWL_H) db2[tid] = aux0;
  else if (tid < PWL_H + OUT * PWL_H) dW3[tid - PWL_H] = aux0;
  else if (tid >= 128 && tid < 128 + OUT) db3[tid - 128] = aux1;
  __syncthreads();
  if (tid < PWL_H) {
    float a = 0.f, b = 0.f;
    for (int cc = 0; cc < PWL_H; ++cc) { a += red0[cc][tid]; b += red1[cc][tid]; }
    db1[tid] = a; dw1[tid] = b;
  }
}

// segment index of t: number of breakpoints <= t
__device__ __forceinline__ int pwl_segment(const float* __restrict__ brk, int nb, float t) {
  int lo = 0, hi = nb;
  while (lo < hi) { const int m = (lo + hi) >> 1; if (brk[m] <= t) lo = m + 1; else hi = m; }
  return lo;
}
__global__ void __launch_bounds__(256) cpb1d_table_eval_kernel(const float* __restrict__ t, int64_t n, const float* __restrict__ brk,
                                                               const float* __restrict__ coef, const int* __restrict__ nseg,
                                                               int OUT, float* __restrict__ out) {
  const int nb = nseg[0] - 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = t[i];
    const int s = pwl_segment(brk, nb, x);
    for (int o = 0; o < OUT; ++o) out[i * OUT + o] = fmaf(coef[((int64_t)s * OUT + o) * 2], x, coef[((int64_t)s * OUT + o) * 2 + 1]);
  }
}

}  // namespace sml

using namespace sml;

extern "C" int sml_cpb1d_table_slots(void) { return PWL_NB; }
extern "C" int sml_cpb1d_table_build(const float* cpb, int32_t OUT, float* brk, float* coef, int32_t* nseg, void* stream) {
  SML_CHECK(cpb && brk && coef && nseg, "sml_cpb1d_table_build: null pointer");
  SML_CHECK(OUT >= 1 && OUT <= 2, "sml_cpb1d_table_build: heads per group must be 1 or 2 (got %d)", OUT);
  cpb1d_table_build_kernel<<<1, PWL_THREADS, 0, as_stream(stream)>>>(cpb, OUT, brk, coef, nseg);
  SML_LAUNCH_OK("sml_cpb1d_table_build");
  return 0;
}
/* rows per output of the segment-sum buffer gseg [OUT, rows, 2] of sml_deform_attn_bwd_pwl / sml_cpb1d_table_bwd */
extern "C" int sml_cpb1d_seg_rows(void) { return PWL_SEG_STRIDE; }
/* dcpb [sml_deform_cpb_nparams(1, OUT)] (overwritten) = gradient of the bias-MLP parameters from the per-segment sums */
extern "C" int sml_cpb1d_table_bwd(const float* cpb, int32_t OUT, const float* brk, const int32_t* nseg, const float* gseg, float* dcpb,
                                   void* stream) {
  SML_CHECK(cpb && brk && nseg && gseg && dcpb, "sml_cpb1d_table_bwd: null pointer");
  SML_CHECK(OUT >= 1 && OUT <= 2, "sml_cpb1d_table_bwd: heads per group must be 1 or 2 (got %d)", OUT);
  cpb1d_table_bwd_kernel<<<1, PWL_THREADS, 0, as_stream(stream)>>>(cpb, OUT, brk, nseg, gseg, dcpb);
  SML_LAUNCH_OK("sml_cpb1d_table_bwd");
  return 0;
}
extern "C" int sml_cpb1d_table_eval(const float* t, int64_t n, const float* brk, const float* coef, const int32_t* nseg, int32_t OUT,
                                    float* out, void* stream) {
  if (n == 0) return 0;
  int64_t g = (n + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  cpb1d_table_eval_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>(t, n, brk, coef, nseg, OUT, out);
  SML_LAUNCH_OK("sml_cpb1d_table_eval");
  return 0;
}

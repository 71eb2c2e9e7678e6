// gemm_simt.cu — fp32 FFMA strided-batched GEMM (exact-fp32 device path; any strides).
// 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile, register-prefetch double buffering.
#include "common.cuh"

namespace sml {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

__global__ void __launch_bounds__(NT, 2) gemm_f32_kernel(const sml_gemm_desc d, const bool a_kfast,
                                                          const bool b_nfast) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int bz = blockIdx.z;
  const int b1 = bz / d.nb2, b2 = bz % d.nb2;
  const float* __restrict__ A = (const float*)d.A + b1 * d.sa_b1 + b2 * d.sa_b2;
  const float* __restrict__ B = (const float*)d.B + b1 * d.sb_b1 + b2 * d.sb_b2;
  float* __restrict__ C = (float*)d.C + b1 * d.sc_b1 + b2 * d.sc_b2;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int M = d.M, N = d.N, K = d.K;

  // per-thread load coordinates: 8 elements of A and 8 of B per k-tile
  int am[8], ak[8], bk[8], bn[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int e = tid + i * NT;                 // 0..2047
    if (a_kfast) { ak[i] = e % BK; am[i] = e / BK; } else { am[i] = e % BM; ak[i] = e / BM; }
    if (b_nfast) { bn[i] = e % BN; bk[i] = e / BN; } else { bk[i] = e % BK; bn[i] = e / BK; }
  }
  float ra[8], rb[8];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gm = m0 + am[i], gk = k0 + ak[i];
      ra[i] = (gm < M && gk < K) ? __ldg(A + (int64_t)gm * d.sa_m + (int64_t)gk * d.sa_k) : 0.f;
      const int gn = n0 + bn[i], gk2 = k0 + bk[i];
      rb[i] = (gn < N && gk2 < K) ? __ldg(B + (int64_t)gk2 * d.sb_k + (int64_t)gn * d.sb_n) : 0.f;
    }
  };
  const int tx = tid % 16, ty = tid / 16;       // micro-tile: rows ty*8.., cols tx*8..
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  gload(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) { As[ak[i]][am[i]] = ra[i]; Bs[bk[i]][bn[i]] = rb[i]; }
    __syncthreads();
    if (k0 + BK < K) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
      const float4 b1v = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1v.x; b[5] = b1v.y; b[6] = b1v.z; b[7] = b1v.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  const float alpha = d.alpha, beta = d.beta, diag = d.diag;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + tx * 8 + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      float* p = C + (int64_t)gm * d.sc_m + (int64_t)gn * d.sc_n;
      if (beta != 0.f) v = fmaf(beta, *p, v);
      if (gm == gn) v += diag;
      if (d.bias) v += __ldg(d.bias + gn);
      if (d.rowtab) v = fmaf(d.rowtab_scale, __ldg(d.rowtab + b1 * d.rt_b1 + b2 * d.rt_b2 + (int64_t)(gm / d.rowdiv) * d.rt_m + gn), v);
      if (d.relu) v = fmaxf(v, 0.f);
      *p = v;
    }
  }
}


// Skinny outputs (per-slide heads: [B,512] x [512,4] ...): one warp per output element, lanes stride over K.
// The tiled kernel above would run a single CTA through K/16 dependent global round trips (measured 89 us).
__global__ void __launch_bounds__(256) gemm_f32_small_kernel(const sml_gemm_desc d) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t per = (int64_t)d.M * d.N;
  if (w >= per * d.nb1 * d.nb2) return;
  const int bz = (int)(w / per);
  const int gm = (int)((w % per) / d.N), gn = (int)(w % d.N);
  const int b1 = bz / d.nb2, b2 = bz % d.nb2;
  const float* A = (const float*)d.A + b1 * d.sa_b1 + b2 * d.sa_b2 + (int64_t)gm * d.sa_m;
  const float* B = (const float*)d.B + b1 * d.sb_b1 + b2 * d.sb_b2 + (int64_t)gn * d.sb_n;
  float acc = 0.f;
  for (int k = lane; k < d.K; k += 32) acc = fmaf(__ldg(A + (int64_t)k * d.sa_k), __ldg(B + (int64_t)k * d.sb_k), acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    float* p = (float*)d.C + b1 * d.sc_b1 + b2 * d.sc_b2 + (int64_t)gm * d.sc_m + (int64_t)gn * d.sc_n;
    float v = d.alpha * acc;
    if (d.beta != 0.f) v = fmaf(d.beta, *p, v);
    if (gm == gn) v += d.diag;
    if (d.bias) v += __ldg(d.bias + gn);
    if (d.rowtab) v = fmaf(d.rowtab_scale, __ldg(d.rowtab + b1 * d.rt_b1 + b2 * d.rt_b2 + (int64_t)(gm / d.rowdiv) * d.rt_m + gn), v);
    if (d.relu) v = fmaxf(v, 0.f);
    *p = v;
  }
}

}  // namespace sml

extern "C" int sml_gemm_f32(const sml_gemm_desc* d, void* stream) {
  SML_CHECK(d && d->A && d->B && d->C, "sml_gemm_f32: null pointer");
  SML_CHECK(d->M > 0 && d->N > 0 && d->K > 0 && d->nb1 > 0 && d->nb2 > 0, "sml_gemm_f32: bad sizes M=%d N=%d K=%d nb=%d,%d",
            d->M, d->N, d->K, d->nb1, d->nb2);
  const long long nbz = (long long)d->nb1 * d->nb2;
  const long long outs = (long long)d->M * d->N * nbz;
  if (outs <= 32768 && d->K >= 32) {          // skinny: warp per output element
    sml::gemm_f32_small_kernel<<<(unsigned)((outs + 7) / 8), 256, 0, sml::as_stream(stream)>>>(*d);
    SML_LAUNCH_OK("sml_gemm_f32(small)");
    return 0;
  }
  SML_CHECK(nbz <= 65535, "sml_gemm_f32: batch %lld > 65535", nbz);
  dim3 grid((d->N + sml::BN - 1) / sml::BN, (d->M + sml::BM - 1) / sml::BM, (unsigned)nbz);
  SML_CHECK(grid.y <= 65535, "sml_gemm_f32: M too large for grid.y");
  const bool a_kfast = (d->sa_k == 1) || (d->sa_m != 1);
  const bool b_nfast = (d->sb_n == 1) || (d->sb_k != 1);
  sml::gemm_f32_kernel<<<grid, sml::NT, 0, sml::as_stream(stream)>>>(*d, a_kfast, b_nfast);
  SML_LAUNCH_OK("sml_gemm_f32");
  return 0;
}

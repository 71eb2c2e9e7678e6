// heads.cu — small per-slide kernels: batch-norm statistics / normalise (the statistics are what the ranks exchange),
// BatchLoss on Gram matrices (utils/loss.py:26-38), OrthogonalLoss (models/cmta_utils.py:1217-1228), elementwise product.
// All latency-bound ([B,128]-sized operands): one launch each, no intermediate tensors.
#include "common.cuh"

namespace sml {

__global__ void __launch_bounds__(256) mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                                  int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = a[i] * b[i];
}

// sums[c] = sum_r x[r,c]; sums[C+c] = sum_r x[r,c]^2   (thread per channel; R is a per-rank batch, i.e. small)
__global__ void bn_stats_kernel(const float* __restrict__ x, float* __restrict__ sums, int R, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f, q = 0.f;
  for (int r = 0; r < R; ++r) { const float v = x[(int64_t)r * C + c]; s += v; q = fmaf(v, v, q); }
  sums[c] = s; sums[C + c] = q;
}
__global__ void bn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                              const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ y, int R, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * C) return;
  const int c = (int)(i % C);
  y[i] = (x[i] - mean[c]) * invstd[c] * (w ? w[c] : 1.f) + (b ? b[c] : 0.f);
}
// sums[c] = sum dy ; sums[C+c] = sum dy * xhat
__global__ void bn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, float* __restrict__ sums, int R, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f, q = 0.f;
  for (int r = 0; r < R; ++r) {
    const float g = dy[(int64_t)r * C + c];
    s += g; q = fmaf(g, (x[(int64_t)r * C + c] - mean[c]) * invstd[c], q);
  }
  sums[c] = s; sums[C + c] = q;
}
// dx = w * invstd * (dy - sum_dy/n - xhat * sum_dyx/n)   with GLOBAL sums and count n
__global__ void bn_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                              const float* __restrict__ invstd, const float* __restrict__ w, const float* __restrict__ sums,
                              float inv_n, float* __restrict__ dx, int R, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * C) return;
  const int c = (int)(i % C);
  const float xh = (x[i] - mean[c]) * invstd[c];
  dx[i] = (w ? w[c] : 1.f) * invstd[c] * (dy[i] - sums[c] * inv_n - xh * sums[C + c] * inv_n);
}

// ---- BatchLoss on Gram matrices.  S [N,N], T [NS,N,N].  loss = (rownorm(S) - mean_s rownorm(T_s))^2 / N
// one block; warp per row.  rn [(1+NS), N] row norms saved for backward.
__global__ void __launch_bounds__(256) batchloss_fwd_kernel(const float* __restrict__ S, const float* __restrict__ T,
                                                            float* __restrict__ loss, float* __restrict__ rn, int N, int NS) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = wid; i < N; i += nw) {
    for (int s = 0; s <= NS; ++s) {
      const float* row = (s == 0 ? S : T + (int64_t)(s - 1) * N * N) + (int64_t)i * N;
      float q = 0.f;
      for (int j = lane; j < N; j += 32) q = fmaf(row[j], row[j], q);
      q = sqrtf(warp_sum(q));
      if (lane == 0) rn[s * N + i] = q;
    }
    __syncwarp();
    for (int j = lane; j < N; j += 32) {
      float tm = 0.f;
      for (int s = 0; s < NS; ++s) tm += T[((int64_t)s * N + i) * N + j] / rn[(s + 1) * N + i];
      const float dlt = S[(int64_t)i * N + j] / rn[i] - tm / NS;
      loss[(int64_t)i * N + j] = dlt * dlt / N;
    }
  }
}
__global__ void __launch_bounds__(256) batchloss_bwd_kernel(const float* __restrict__ S, const float* __restrict__ T,
                                                            const float* __restrict__ rn, const float* __restrict__ dloss,
                                                            float* __restrict__ dS, float* __restrict__ dT, int N, int NS) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = wid; i < N; i += nw) {
    // D = 2 (Sn - Tm) / N * dloss ; rowdots of D with the normalised rows
    float dotS = 0.f, dotT[16];
    for (int s = 0; s < NS; ++s) dotT[s] = 0.f;
    for (int j = lane; j < N; j += 32) {
      float tm = 0.f;
      for (int s = 0; s < NS; ++s) tm += T[((int64_t)s * N + i) * N + j] / rn[(s + 1) * N + i];
      const float sn = S[(int64_t)i * N + j] / rn[i];
      const float D = 2.f * (sn - tm / NS) / N * dloss[(int64_t)i * N + j];
      dotS = fmaf(D, sn, dotS);
      for (int s = 0; s < NS; ++s) dotT[s] = fmaf(D, T[((int64_t)s * N + i) * N + j] / rn[(s + 1) * N + i], dotT[s]);
    }
    dotS = warp_sum(dotS);
    for (int s = 0; s < NS; ++s) dotT[s] = warp_sum(dotT[s]);
    for (int j = lane; j < N; j += 32) {
      float tm = 0.f;
      for (int s = 0; s < NS; ++s) tm += T[((int64_t)s * N + i) * N + j] / rn[(s + 1) * N + i];
      const float sn = S[(int64_t)i * N + j] / rn[i];
      const float D = 2.f * (sn - tm / NS) / N * dloss[(int64_t)i * N + j];
      dS[(int64_t)i * N + j] = (D - sn * dotS) / rn[i];
      for (int s = 0; s < NS; ++s) {
        const float tn = T[((int64_t)s * N + i) * N + j] / rn[(s + 1) * N + i];
        dT[((int64_t)s * N + i) * N + j] = -(D - tn * dotT[s]) / (NS * rn[(s + 1) * N + i]);
      }
    }
  }
}

// ---- OrthogonalLoss: warp per row.  stats [B, 9]: |P|,|Ph|,|G|,|Gh| (clamped at eps), P.Ph, G.Gh, P.G, P.Gh, G.Ph
__device__ __forceinline__ float sgnf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
__global__ void __launch_bounds__(256) orth_fwd_kernel(const float* __restrict__ P, const float* __restrict__ Ph,
                                                       const float* __restrict__ G, const float* __restrict__ Gh, float gamma,
                                                       float* __restrict__ loss, float* __restrict__ stats, int B, int D) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= B) return;
  float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = lane; j < D; j += 32) {
    const float p = P[(int64_t)r * D + j], ph = Ph[(int64_t)r * D + j], g = G[(int64_t)r * D + j], gh = Gh[(int64_t)r * D + j];
    a[0] = fmaf(p, p, a[0]); a[1] = fmaf(ph, ph, a[1]); a[2] = fmaf(g, g, a[2]); a[3] = fmaf(gh, gh, a[3]);
    a[4] = fmaf(p, ph, a[4]); a[5] = fmaf(g, gh, a[5]); a[6] = fmaf(p, g, a[6]); a[7] = fmaf(p, gh, a[7]); a[8] = fmaf(g, ph, a[8]);
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) a[i] = warp_sum(a[i]);
  if (lane == 0) {
    const float eps = 1e-8f;
    const float nP = fmaxf(sqrtf(a[0]), eps), nPh = fmaxf(sqrtf(a[1]), eps), nG = fmaxf(sqrtf(a[2]), eps), nGh = fmaxf(sqrtf(a[3]), eps);
    const float c1 = a[4] / (nP * nPh), c2 = a[5] / (nG * nGh), c3 = a[6] / (nP * nG), c4 = a[7] / (nP * nGh), c5 = a[8] / (nG * nPh);
    loss[r] = (1.f - fabsf(c1)) + (1.f - fabsf(c2)) + gamma * (fabsf(c3) + fabsf(c4) + fabsf(c5));
    float* st = stats + (int64_t)r * 9;
    st[0] = nP; st[1] = nPh; st[2] = nG; st[3] = nGh; st[4] = c1; st[5] = c2; st[6] = c3; st[7] = c4; st[8] = c5;
  }
}
__global__ void __launch_bounds__(256) orth_bwd_kernel(const float* __restrict__ P, const float* __restrict__ Ph,
                                                       const float* __restrict__ G, const float* __restrict__ Gh, float gamma,
                                                       const float* __restrict__ stats, const float* __restrict__ dloss,
                                                       float* __restrict__ dP, float* __restrict__ dPh, float* __restrict__ dG,
                                                       float* __restrict__ dGh, int B, int D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * D) return;
  const int r = (int)(i / D);
  const float* st = stats + (int64_t)r * 9;
  const float nP = st[0], nPh = st[1], nG = st[2], nGh = st[3], c1 = st[4], c2 = st[5], c3 = st[6], c4 = st[7], c5 = st[8];
  const float p = P[i], ph = Ph[i], g = G[i], gh = Gh[i], dl = dloss[r];
  // d cos(a,b)/da = b/(|a||b|) - cos * a/|a|^2 ; barred (detached) operands get no gradient
  dPh[i] = dl * (-sgnf(c1) * (p / (nP * nPh) - c1 * ph / (nPh * nPh)) + gamma * sgnf(c5) * (g / (nG * nPh) - c5 * ph / (nPh * nPh)));
  dGh[i] = dl * (-sgnf(c2) * (g / (nG * nGh) - c2 * gh / (nGh * nGh)) + gamma * sgnf(c4) * (p / (nP * nGh) - c4 * gh / (nGh * nGh)));
  dP[i] = dl * gamma * sgnf(c3) * (g / (nP * nG) - c3 * p / (nP * nP));
  dG[i] = dl * gamma * sgnf(c3) * (p / (nP * nG) - c3 * g / (nG * nG));
}

}  // namespace sml

using namespace sml;

extern "C" int sml_mul(const float* a, const float* b, float* y, int64_t n, void* stream) {
  if (n == 0) return 0;
  int64_t g = (n + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  mul_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>(a, b, y, n);
  SML_LAUNCH_OK("sml_mul");
  return 0;
}
extern "C" int sml_bn_stats(const float* x, float* sums, int32_t R, int32_t C, void* stream) {
  bn_stats_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(x, sums, R, C);
  SML_LAUNCH_OK("sml_bn_stats");
  return 0;
}
extern "C" int sml_bn_fwd(const float* x, const float* mean, const float* invstd, const float* w, const float* b, float* y,
                          int32_t R, int32_t C, void* stream) {
  const int64_t n = (int64_t)R * C;
  if (n == 0) return 0;
  bn_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(x, mean, invstd, w, b, y, R, C);
  SML_LAUNCH_OK("sml_bn_fwd");
  return 0;
}
extern "C" int sml_bn_bwd_stats(const float* dy, const float* x, const float* mean, const float* invstd, float* sums, int32_t R,
                                int32_t C, void* stream) {
  bn_bwd_stats_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(dy, x, mean, invstd, sums, R, C);
  SML_LAUNCH_OK("sml_bn_bwd_stats");
  return 0;
}
extern "C" int sml_bn_bwd(const float* dy, const float* x, const float* mean, const float* invstd, const float* w,
                          const float* sums, float inv_n, float* dx, int32_t R, int32_t C, void* stream) {
  const int64_t n = (int64_t)R * C;
  if (n == 0) return 0;
  bn_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(dy, x, mean, invstd, w, sums, inv_n, dx, R, C);
  SML_LAUNCH_OK("sml_bn_bwd");
  return 0;
}
extern "C" int sml_batchloss_fwd(const float* S, const float* T, float* loss, float* rn, int32_t N, int32_t NS, void* stream) {
  SML_CHECK(NS >= 1 && NS <= 16, "sml_batchloss_fwd: NS=%d out of range", NS);
  batchloss_fwd_kernel<<<1, 256, 0, as_stream(stream)>>>(S, T, loss, rn, N, NS);
  SML_LAUNCH_OK("sml_batchloss_fwd");
  return 0;
}
extern "C" int sml_batchloss_bwd(const float* S, const float* T, const float* rn, const float* dloss, float* dS, float* dT,
                                 int32_t N, int32_t NS, void* stream) {
  SML_CHECK(NS >= 1 && NS <= 16, "sml_batchloss_bwd: NS=%d out of range", NS);
  batchloss_bwd_kernel<<<1, 256, 0, as_stream(stream)>>>(S, T, rn, dloss, dS, dT, N, NS);
  SML_LAUNCH_OK("sml_batchloss_bwd");
  return 0;
}
extern "C" int sml_orthogonal_loss_fwd(const float* P, const float* Ph, const float* G, const float* Gh, float gamma,
                                       float* loss, float* stats, int32_t B, int32_t D, void* stream) {
  orth_fwd_kernel<<<(B + 7) / 8, 256, 0, as_stream(stream)>>>(P, Ph, G, Gh, gamma, loss, stats, B, D);
  SML_LAUNCH_OK("sml_orthogonal_loss_fwd");
  return 0;
}
extern "C" int sml_orthogonal_loss_bwd(const float* P, const float* Ph, const float* G, const float* Gh, float gamma,
                                       const float* stats, const float* dloss, float* dP, float* dPh, float* dG, float* dGh,
                                       int32_t B, int32_t D, void* stream) {
  const int64_t n = (int64_t)B * D;
  orth_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(P, Ph, G, Gh, gamma, stats, dloss, dP, dPh, dG, dGh, B, D);
  SML_LAUNCH_OK("sml_orthogonal_loss_bwd");
  return 0;
}

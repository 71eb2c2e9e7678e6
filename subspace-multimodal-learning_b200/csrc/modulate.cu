// modulate.cu — gradient modulation of the fusion classifier (train_test.py:87-184, task types diag2021 / grade / subtype):
// after backward, the two halves of every classifier row gradient (tumor-branch columns [:hs], immune-branch columns [hs:])
// are compared; when they conflict (negative cosine) the half that belongs to the currently better-scoring branch
// (ratio < 1 marks the WEAKER branch, whose gradient is re-projected) is replaced as the reference does:
//   g <- g - proj ; perpen <- g - proj (the projection is subtracted twice, reference quirk kept) ; g <- |g| perpen / |perpen|.
// The reference runs this as Python loops with a host sync per row; here it is one single-block kernel without any sync.
#include "common.cuh"

namespace sml {

constexpr int MOD_THREADS = 256;

__global__ void __launch_bounds__(MOD_THREADS) grad_modulate_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                                    float* __restrict__ Wg, const float* __restrict__ ft,
                                                                    const float* __restrict__ fi, const int64_t* __restrict__ labels,
                                                                    int B, int L, int hs, float* __restrict__ info) {
  extern __shared__ float sm[];                 // out_t [B*L] | out_i [B*L] | red [32] | scal [8]
  float* out_t = sm; float* out_i = sm + B * L; float* red = out_i + B * L; float* scal = red + 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = MOD_THREADS / 32;
  // ---- branch logits: out_t = f_t W[:, :hs]^T + b / 2, out_i = f_i W[:, hs:]^T + b / 2 (one warp per (sample, class))
  for (int idx = warp; idx < 2 * B * L; idx += nwarp) {
    const int which = idx / (B * L), r = idx % (B * L), s = r / L, c = r % L;
    const float* f = (which ? fi : ft) + (int64_t)s * hs;
    const float* w = W + (int64_t)c * 2 * hs + which * hs;
    float acc = 0.f;
    for (int k = lane; k < hs; k += 32) acc = fmaf(f[k], w[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) (which ? out_i : out_t)[r] = acc + 0.5f * bias[c];
  }
  __syncthreads();
  // ---- scores: sum over samples of softmax(out)[label]
  if (warp < 2) {
    const float* o = warp ? out_i : out_t;
    float sc = 0.f;
    for (int s = lane; s < B; s += 32) {
      float mx = -INFINITY;
      for (int c = 0; c < L; ++c) mx = fmaxf(mx, o[s * L + c]);
      float den = 0.f;
      for (int c = 0; c < L; ++c) den += expf(o[s * L + c] - mx);
      const int lab = (int)labels[s];
      sc += expf(o[s * L + lab] - mx) / den;
    }
    sc = warp_sum(sc);
    if (lane == 0) scal[warp] = sc;
  }
  __syncthreads();
  const float score_t = scal[0], score_i = scal[1];
  const float ratio_t = score_t / score_i, ratio_i = 1.f / ratio_t;
  int n_t = 0, n_i = 0;
  // ---- per classifier row: conflict test and re-projection (whole block per row; rows are few)
  for (int r = 0; r < L; ++r) {
    float* gt = Wg + (int64_t)r * 2 * hs;
    float* gi = gt + hs;
    float dot = 0.f, nt = 0.f, ni = 0.f;
    for (int k = tid; k < hs; k += MOD_THREADS) { const float a = gt[k], c = gi[k]; dot = fmaf(a, c, dot); nt = fmaf(a, a, nt); ni = fmaf(c, c, ni); }
    dot = block_sum(dot, red); nt = block_sum(nt, red); ni = block_sum(ni, red);
    const float sim = dot / (sqrtf(nt) * sqrtf(ni));
    if (!(sim < 0.f)) continue;                                   // uniform across the block
    const bool fix_t = ratio_t < 1.f, fix_i = !fix_t && ratio_i < 1.f;
    if (!fix_t && !fix_i) continue;
    float* g = fix_t ? gt : gi;                                   // gradient to re-project
    const float* o = fix_t ? gi : gt;                             // onto the complement of this one
    const float ps = dot / (fix_t ? ni : nt);
    // g1 = g - ps o ; perpen = g1 - ps o ; result = |g1| perpen / |perpen|
    float n1 = 0.f, np = 0.f;
    for (int k = tid; k < hs; k += MOD_THREADS) {
      const float g1 = g[k] - ps * o[k], pp = g1 - ps * o[k];
      n1 = fmaf(g1, g1, n1); np = fmaf(pp, pp, np);
    }
    n1 = block_sum(n1, red); np = block_sum(np, red);
    const float sc = sqrtf(n1) / sqrtf(np);
    __syncthreads();
    for (int k = tid; k < hs; k += MOD_THREADS) g[k] = sc * (g[k] - 2.f * ps * o[k]);
    __syncthreads();
    if (fix_t) ++n_t; else ++n_i;
  }
  if (tid == 0 && info) { info[0] = score_t; info[1] = score_i; info[2] = (float)n_t; info[3] = (float)n_i; }
}

}  // namespace sml

using namespace sml;

extern "C" int sml_grad_modulate(const float* W, const float* bias, float* Wgrad, const float* feat_t, const float* feat_i,
                                 const int64_t* labels, int32_t B, int32_t L, int32_t hs, float* info, void* stream) {
  SML_CHECK(W && bias && Wgrad && feat_t && feat_i && labels, "sml_grad_modulate: null pointer");
  SML_CHECK(B > 0 && L > 0 && hs > 0 && B * L <= 4096, "sml_grad_modulate: bad sizes (B=%d L=%d hs=%d)", B, L, hs);
  const size_t smem = (size_t)(2 * B * L + 32 + 8) * sizeof(float);
  grad_modulate_kernel<<<1, MOD_THREADS, smem, as_stream(stream)>>>(W, bias, Wgrad, feat_t, feat_i, labels, B, L, hs, info);
  SML_LAUNCH_OK("sml_grad_modulate");
  return 0;
}

// deform_attn.cuh — definitions shared by the SIMT (deform.cu) and tensor-core (deform_mma.cu) fused deformable
// attention kernels: configuration, query-lattice coordinates, the signed-log distance and the dropout hash.
#pragma once
#include "common.cuh"

namespace sml {

// CPB MLP: Linear(CD,32) ReLU Linear(32,32) ReLU Linear(32,OUT)[o]; packed weights in global memory:
//   [W1: 32*CD][b1: 32][W2: 32*32 (row c, col k)][b2: 32][W3: OUT*32][b3: OUT]
struct AttnCfg {
  int B, heads, n, nkv, inner;    // q/out rows have `inner` = heads*64 floats
  int G, CD, OUT;                 // offset groups, coordinate dims, heads per group
  int W;                          // 2-D query lattice width (CD == 2): token i -> (x = i % W, y = i / W)
  float qden0, qden1;             // query-lattice normalisation denominators max(size-1,1)
  float p_drop; unsigned long long seed;
  const unsigned long long* salt;   // see api.cu (NULL: seed as passed)
};
constexpr int AD = 64;            // head dim (the only configuration on the reference path)
constexpr int CPB_H = 32;

__device__ __forceinline__ float slogf(float x) {
  const float l = logf(fabsf(x) + 1.f);
  return x > 0.f ? l : (x < 0.f ? -l : 0.f);
}
// counter-based dropout hash of one (batch, head, query, key) pair: 64 -> 32 bit fold + an avalanching 32-bit finaliser
// (two multiplies; the previous 64-bit splitmix cost ~25 integer instructions per pair, ~10 % of the attention kernels).
// Every deformable-attention kernel (SIMT, warp-MMA, tcgen05; forward and backward) regenerates the mask from this function.
__device__ __forceinline__ uint32_t pair_hash(unsigned long long seed, unsigned long long idx) {
  uint32_t x = (uint32_t)idx ^ ((uint32_t)(idx >> 32) * 0x9E3779B1u) ^ (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x85EBCA77u);
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ void query_pos(const AttnCfg& cf, int i, float& px, float& py) {
  if (cf.CD == 2) {
    px = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, (float)(i % cf.W)), cf.qden0), -1.f);
    py = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, (float)(i / cf.W)), cf.qden1), -1.f);
  } else {
    px = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, (float)i), cf.qden0), -1.f);
    py = 0.f;
  }
}


static inline int make_acfg(AttnCfg& cf, const int32_t* c, float p_drop, uint64_t seed) {
  cf.B = c[0]; cf.heads = c[1]; cf.n = c[2]; cf.nkv = c[3]; cf.inner = c[1] * AD; cf.G = c[4]; cf.CD = c[5];
  cf.OUT = c[1] / c[4]; cf.W = c[6];
  SML_CHECK(c[7] == AD, "sml_deform_attn: head dim must be %d (got %d)", AD, c[7]);
  SML_CHECK(cf.CD == 1 || cf.CD == 2, "sml_deform_attn: coordinate dims must be 1 or 2");
  SML_CHECK(cf.heads % cf.G == 0, "sml_deform_attn: heads must be a multiple of groups");
  if (cf.CD == 2) {
    const int Hq = cf.n / cf.W;
    SML_CHECK(Hq * cf.W == cf.n, "sml_deform_attn: n must be H*W");
    cf.qden0 = (float)(Hq - 1 > 1 ? Hq - 1 : 1);      // reference normalises x by (h-1), y by (w-1) (2D.py:100-108)
    cf.qden1 = (float)(cf.W - 1 > 1 ? cf.W - 1 : 1);
  } else {
    cf.qden0 = (float)(cf.n - 1 > 1 ? cf.n - 1 : 1); cf.qden1 = 1.f;
  }
  cf.p_drop = p_drop; cf.seed = seed; cf.salt = seed_salt();
  return 0;
}

}  // namespace sml

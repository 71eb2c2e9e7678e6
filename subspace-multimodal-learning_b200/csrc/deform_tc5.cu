// deform_tc5.cu — Blackwell-native (tcgen05 / TMEM) forward of the fused 2-D deformable attention
// (models/DeformableAttention2D.py:292-312 with the continuous-position-bias MLP of :120-157 evaluated per (query, key)
// pair).  bf16 operands, fp32 accumulation in tensor memory; the legacy warp-level kernels of deform_mma.cu remain the
// fp32-grade ("bf16x3") path.
//
// One CTA = 128 queries (= the 128 TMEM lanes) of one (batch, head); keys stream in tiles of 64.  Every matrix product of
// the pair pipeline is a tcgen05.mma whose M dimension is the 128 queries of the CTA paired with ONE key:
//   layer 1   D1[128 x 32] = A0[128 x 16] . W1e[16 x 32]      A0 row = (ax_hi, ax_hi, ax_lo, ay_hi, ay_hi, ay_lo, 1, 1, 0..)
//                                                             W1e   = (W1x_hi, W1x_lo, W1x_hi, W1y_hi, W1y_lo, W1y_hi, b1_hi, b1_lo, 0..)
//             (hi/lo bf16 splits of the signed-log distances, of W1 and of b1: fp32-grade pre-activations from one K = 16 step)
//   layer 2   D2[128 x 32] = relu(D1)[128 x 32] . W2^T + 1 . b2   (two K = 16 steps + one "ones" step carrying b2_hi + b2_lo)
//   layer 3   S[:, 16-column block of the key] += relu(D2)[128 x 32] . W3sel_j[32 x 16], where W3sel_j has w3 in column
//             (j mod 16) and zeros elsewhere: the bias lands directly in the attention-logit accumulator S = Q K^T in TMEM.
// The A operands relu(D1), relu(D2) never leave the SM: a compute thread (one per query) reads its accumulator row with
// tcgen05.ld, applies ReLU + bf16 packing with one F2FP.RELU per two values and writes the row back to tensor memory with
// tcgen05.st, from where the next MMA takes it as its A operand (A-from-TMEM form of tcgen05.mma).  The pipeline is three
// asynchronous round trips per key between the 8 compute warps and the single MMA-issuing thread (mbarriers both ways,
// tcgen05.commit towards the compute warps), software-pipelined over 6 key slots (2 warpgroups x 3).
// After the 64 keys of a tile: online softmax on the S row (registers), dropout by the same counter-based hash as the
// other kernels (the backward regenerates it), P as bf16 hi + lo into tensor memory, O += P V with three MMA passes
// (P_hi V_hi + P_hi V_lo + P_lo V_hi: the backward's delta = dO . O relies on an fp32-grade O, see deform_mma.cu).
//
// Tensor-memory map (512 columns):  S 0..63 | O 64..127 | P_hi 128..159 | P_lo 160..191 | ones 192..199 | key slots 200 + 48 s
// (slot: D 32 columns + A 16 columns).
#include "tc_common.cuh"
#include "deform_attn.cuh"

namespace sml {

constexpr int T5_Q = 128;            // queries per CTA
constexpr int T5_KT = 64;            // keys per tile
constexpr int T5_THREADS = 320;      // warp 0: spare, warp 1: TMEM allocator + MMA issuer, warps 2..9: compute (2 warpgroups)
constexpr int T5_SPW = 3;            // key slots per warpgroup
constexpr int T5_UPT = T5_KT / 2;    // keys per warpgroup per tile

constexpr uint32_t T5_COL_S = 0, T5_COL_O = 64, T5_COL_PH = 128, T5_COL_PL = 160, T5_COL_ONES = 192, T5_COL_SLOT = 200;
constexpr uint32_t T5_SLOT_COLS = 48, T5_SLOT_A = 32;

struct T5Smem {
  static constexpr int Q = 0;                          // [128 x 128 B] K-major A of S = Q K^T
  static constexpr int K = Q + T5_Q * 128;             // [64 x 128 B]  K-major B of S = Q K^T (rows = keys)
  static constexpr int VH = K + T5_KT * 128;           // [64 x 128 B]  N-major B of O += P V (rows = keys = K axis), hi
  static constexpr int VL = VH + T5_KT * 128;          // same, lo
  static constexpr int W1 = VL + T5_KT * 128;          // [32 x 128 B] rows = hidden unit, 16 K-elements used
  static constexpr int W2 = W1 + 32 * 128;             // [32 x 128 B] rows = output unit c, 32 K-elements
  static constexpr int B2 = W2 + 32 * 128;             // [32 x 128 B] rows = output unit c, K-elements (b2_hi, b2_lo, 0..)
  static constexpr int W3 = B2 + 32 * 128;             // 16 variants x [16 x 128 B]
  static constexpr int KPOS = W3 + 16 * 16 * 128;      // [64][2] floats
  static constexpr int STAT = KPOS + T5_KT * 8;        // [2][128] floats (row maxima of the two column halves)
  static constexpr int BAR = STAT + 2 * T5_Q * 4;      // mbarriers
  static constexpr int NBAR = 2 * T5_SPW * 6 + 4;
  static constexpr int TSLOT = BAR + NBAR * 8;
  static constexpr int TOTAL = TSLOT + 16 + 1024;      // + alignment slack
};

__device__ __forceinline__ uint32_t t5_sw(int row, int chunk16) { return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4)); }

__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_c), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// relu + bf16 pack of two fp32 values: low half <- lo, high half <- hi (one F2FP.RELU.BF16.F32.PACK_AB)
__device__ __forceinline__ uint32_t relu_pack_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float t5_slog(float x) { return copysignf(__logf(fabsf(x) + 1.f), x); }     // = deform_mma.cu:slog_fast
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// fp32 rows of a token-major tensor -> bf16 SWIZZLE_128B tile (64 columns = one 128-byte row); rows >= limit are zero
__device__ __forceinline__ void t5_load_rows(uint8_t* dst_hi, uint8_t* dst_lo, const float* __restrict__ src, int64_t row_stride,
                                             int row0, int limit, int rows, int tid, int nthreads) {
  for (int e = tid; e < rows * 8; e += nthreads) {
    const int r = e >> 3, c = e & 7;                 // 16-byte chunk c = 8 bf16
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
    if (row0 + r < limit) {
      const float* p = src + (int64_t)(row0 + r) * row_stride + c * 8;
      x0 = __ldg(reinterpret_cast<const float4*>(p)); x1 = __ldg(reinterpret_cast<const float4*>(p + 4));
    }
    const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
      if (dst_lo) l[i] = pack_bf16x2(v[2 * i] - __uint_as_float(h[i] << 16), v[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u));
    }
    *reinterpret_cast<uint4*>(dst_hi + t5_sw(r, c)) = make_uint4(h[0], h[1], h[2], h[3]);
    if (dst_lo) *reinterpret_cast<uint4*>(dst_lo + t5_sw(r, c)) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__device__ __forceinline__ void t5_put(uint8_t* tile, int row, int k, float v) {      // one bf16 element of a K-major SW128 tile
  *reinterpret_cast<__nv_bfloat16*>(tile + t5_sw(row, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(v);
}

// grid (ceil(n / 128), heads, B); q pre-scaled by d^-1/2; out [b, n, inner]; lse [b, heads, n].  CD = 2, one head per group.
__global__ void __launch_bounds__(T5_THREADS, 1)
deform_attn_fwd_tc5_kernel(const AttnCfg cf, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                           const float* __restrict__ vs, const float* __restrict__ cpb, float* __restrict__ out,
                           float* __restrict__ lse) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T5Smem::BAR);
  // per warpgroup g, slot s: compute -> MMA: a0, a1, a2 (count 4 warps); MMA -> compute: d1, d2, fr (tcgen05.commit)
  auto BA = [&](int kind, int g, int s) { return bars + (kind * 2 + g) * T5_SPW + s; };       // kind 0..5
  uint64_t* bar_kv = bars + 6 * 2 * T5_SPW;         // K / V / kpos tile in shared memory (count 8 warps)
  uint64_t* bar_sfull = bar_kv + 1;                 // all logit MMAs of the tile complete
  uint64_t* bar_pready = bar_kv + 2;                // P written, O rescaled (count 8 warps)
  uint64_t* bar_pvdone = bar_kv + 3;                // O += P V complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + T5Smem::TSLOT);
  float* kpos = reinterpret_cast<float*>(smem + T5Smem::KPOS);
  float* stat = reinterpret_cast<float*>(smem + T5Smem::STAT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * T5_Q;
  const int bg = b * cf.G + head;                   // OUT == 1: group == head
  const int ntiles = (cf.nkv + T5_KT - 1) / T5_KT;

  if (tid == 0) {
    for (int kind = 0; kind < 6; ++kind)
      for (int g = 0; g < 2; ++g)
        for (int s = 0; s < T5_SPW; ++s) mbar_init(BA(kind, g, s), kind < 3 ? 4u : 1u);
    mbar_init(bar_kv, 8); mbar_init(bar_sfull, 1); mbar_init(bar_pready, 8); mbar_init(bar_pvdone, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- constant operand tiles: W1e, W2, b2 block, the 16 column-selecting W3 variants; the Q tile
  {
    const float* W1 = cpb; const float* B1 = W1 + CPB_H * 2; const float* W2 = B1 + CPB_H;
    const float* B2 = W2 + CPB_H * CPB_H; const float* W3 = B2 + CPB_H;
    uint32_t* z = reinterpret_cast<uint32_t*>(smem + T5Smem::W1);
    for (int i = tid; i < (T5Smem::KPOS - T5Smem::W1) / 4; i += T5_THREADS) z[i] = 0u;
    __syncthreads();
    for (int i = tid; i < CPB_H; i += T5_THREADS) {              // row i = hidden unit
      const float wx = W1[i * 2], wy = W1[i * 2 + 1], bb = B1[i];
      const float wxh = bf16_round(wx), wyh = bf16_round(wy), bh = bf16_round(bb);
      uint8_t* t1 = smem + T5Smem::W1;
      t5_put(t1, i, 0, wxh); t5_put(t1, i, 1, wx - wxh); t5_put(t1, i, 2, wxh);
      t5_put(t1, i, 3, wyh); t5_put(t1, i, 4, wy - wyh); t5_put(t1, i, 5, wyh);
      t5_put(t1, i, 6, bh);  t5_put(t1, i, 7, bb - bh);
      const float b2 = B2[i], b2h = bf16_round(b2);
      t5_put(smem + T5Smem::B2, i, 0, b2h); t5_put(smem + T5Smem::B2, i, 1, b2 - b2h);
    }
    for (int i = tid; i < CPB_H * CPB_H; i += T5_THREADS) t5_put(smem + T5Smem::W2, i / CPB_H, i % CPB_H, W2[i]);   // row c, K = k
    for (int i = tid; i < 16 * CPB_H; i += T5_THREADS) {          // variant vv: row vv holds w3
      const int vv = i / CPB_H, kk = i % CPB_H;
      t5_put(smem + T5Smem::W3 + vv * 2048, vv, kk, W3[kk]);
    }
    t5_load_rows(smem + T5Smem::Q, nullptr, q + (int64_t)b * cf.n * cf.inner + head * AD, cf.inner, q0, cf.n, T5_Q, tid, T5_THREADS);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t id32 = make_idesc(128, 32, false, false);
      constexpr uint32_t id16 = make_idesc(128, 16, false, false);
      constexpr uint32_t id64 = make_idesc(128, 64, false, false);
      constexpr uint32_t id64n = make_idesc(128, 64, false, true);      // B = V, N-major
      const uint32_t sQ = smem_u32(smem + T5Smem::Q), sK = smem_u32(smem + T5Smem::K), sVh = smem_u32(smem + T5Smem::VH),
                     sVl = smem_u32(smem + T5Smem::VL), sW1 = smem_u32(smem + T5Smem::W1), sW2 = smem_u32(smem + T5Smem::W2),
                     sB2 = smem_u32(smem + T5Smem::B2), sW3 = smem_u32(smem + T5Smem::W3);
      const uint64_t dW1 = make_smem_desc(sW1, 16, 1024), dW2a = make_smem_desc(sW2, 16, 1024), dW2b = make_smem_desc(sW2 + 32, 16, 1024),
                     dB2 = make_smem_desc(sB2, 16, 1024);
      uint32_t qb = 0;
      for (int tile = 0; tile < ntiles; ++tile, qb += T5_UPT) {
        mbar_wait(bar_kv, tile & 1u);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < AD / 16; ++ks)
          tc_mma_bf16(tmem_base + T5_COL_S, make_smem_desc(sQ + ks * 32, 16, 1024), make_smem_desc(sK + ks * 32, 16, 1024), id64,
                      ks != 0 ? 1u : 0u);
        for (int t = 0; t < T5_UPT + 2; ++t) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (t < T5_UPT) {
              const uint32_t qq = qb + t, s = qq % T5_SPW, ph = (qq / T5_SPW) & 1u;
              const uint32_t slot = tmem_base + T5_COL_SLOT + (g * T5_SPW + s) * T5_SLOT_COLS;
              mbar_wait(BA(0, g, s), ph);
              tc_fence_after();
              tc_mma_bf16_ts(slot, slot + T5_SLOT_A, dW1, id32, 0u);
              tc_commit(BA(3, g, s));
            }
            if (t >= 1 && t < T5_UPT + 1) {
              const uint32_t qq = qb + t - 1, s = qq % T5_SPW, ph = (qq / T5_SPW) & 1u;
              const uint32_t slot = tmem_base + T5_COL_SLOT + (g * T5_SPW + s) * T5_SLOT_COLS;
              mbar_wait(BA(1, g, s), ph);
              tc_fence_after();
              tc_mma_bf16_ts(slot, slot + T5_SLOT_A, dW2a, id32, 0u);
              tc_mma_bf16_ts(slot, slot + T5_SLOT_A + 8, dW2b, id32, 1u);
              tc_mma_bf16_ts(slot, tmem_base + T5_COL_ONES, dB2, id32, 1u);
              tc_commit(BA(4, g, s));
            }
            if (t >= 2) {
              const uint32_t u = t - 2, qq = qb + u, s = qq % T5_SPW, ph = (qq / T5_SPW) & 1u;
              const uint32_t slot = tmem_base + T5_COL_SLOT + (g * T5_SPW + s) * T5_SLOT_COLS;
              const uint32_t key = 2 * u + g;
              mbar_wait(BA(2, g, s), ph);
              tc_fence_after();
              const uint32_t w3 = sW3 + (key & 15u) * 2048u;
              const uint32_t scol = tmem_base + T5_COL_S + (key & ~15u);
              tc_mma_bf16_ts(scol, slot + T5_SLOT_A, make_smem_desc(w3, 16, 1024), id16, 1u);
              tc_mma_bf16_ts(scol, slot + T5_SLOT_A + 8, make_smem_desc(w3 + 32, 16, 1024), id16, 1u);
              tc_commit(BA(5, g, s));
            }
          }
        }
        tc_commit(bar_sfull);
        mbar_wait(bar_pready, tile & 1u);
        tc_fence_after();
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {                    // P_hi V_lo + P_lo V_hi + P_hi V_hi (small terms first)
          const uint32_t pa = tmem_base + (pass == 1 ? T5_COL_PL : T5_COL_PH);
          const uint32_t vb = pass == 0 ? sVl : sVh;
#pragma unroll
          for (int ks = 0; ks < T5_KT / 16; ++ks)
            tc_mma_bf16_ts(tmem_base + T5_COL_O, pa + ks * 8, make_smem_desc(vb + ks * 2048, 8192, 1024), id64n,
                           (tile | pass | ks) != 0 ? 1u : 0u);
        }
        tc_commit(bar_pvdone);
      }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ compute warps: one thread per (query, warpgroup)
    const int cw = warp - 2, g = cw >> 2, qd = warp & 3;          // warpgroup, TMEM lane quarter of this warp
    const int row = qd * 32 + lane, ctid = cw * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(qd * 32) << 16);
    const int qi = q0 + row;
    float qx, qy;
    query_pos(cf, qi < cf.n ? qi : cf.n - 1, qx, qy);
    if (g == 0) {                                                  // the constant "ones" A block (b2_hi + b2_lo rows of the B2 tile)
      const uint32_t ones[8] = {0x3f803f80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      tc_st8(trow + T5_COL_ONES, ones);
      tc_wait_st();
    }
    float m_run = -INFINITY, l_part = 0.f;
    const float keep_scale = cf.p_drop > 0.f ? 1.f / (1.f - cf.p_drop) : 1.f;
    const uint32_t thresh = (uint32_t)(cf.p_drop * 4294967296.0);
    const unsigned long long hbase = ((unsigned long long)(b * cf.heads + head) * cf.n + (unsigned long long)qi) * cf.nkv;
    const float* kb = k + (int64_t)b * cf.nkv * cf.inner + head * AD;
    const float* vb = v + (int64_t)b * cf.nkv * cf.inner + head * AD;
    uint32_t qb = 0;
    uint32_t r[32];
    for (int tile = 0; tile < ntiles; ++tile, qb += T5_UPT) {
      const int k0 = tile * T5_KT;
      if (tile > 0) { mbar_wait(bar_pvdone, (tile - 1) & 1u); tc_fence_after(); }      // V, P free; O stable
      t5_load_rows(smem + T5Smem::K, nullptr, kb, cf.inner, k0, cf.nkv, T5_KT, ctid, 256);
      t5_load_rows(smem + T5Smem::VH, smem + T5Smem::VL, vb, cf.inner, k0, cf.nkv, T5_KT, ctid, 256);
      if (ctid < T5_KT) {
        const bool ok = k0 + ctid < cf.nkv;
        const float2 kp = ok ? *reinterpret_cast<const float2*>(vs + ((int64_t)bg * cf.nkv + k0 + ctid) * 2) : make_float2(0.f, 0.f);
        *reinterpret_cast<float2*>(kpos + 2 * ctid) = kp;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_kv);
      asm volatile("bar.sync 1, 256;" ::: "memory");               // kpos visible to every compute thread
      // ---- pair pipeline: iteration t runs stage A of key t, stage B of key t - 1, stage C of key t - 2 (this warpgroup's keys)
      for (int t = 0; t < T5_UPT + 2; ++t) {
        if (t < T5_UPT) {
          const uint32_t qq = qb + t, s = qq % T5_SPW, use = qq / T5_SPW;
          const uint32_t slot = trow + T5_COL_SLOT + (g * T5_SPW + s) * T5_SLOT_COLS;
          if (use > 0) { mbar_wait(BA(5, g, s), (use - 1) & 1u); tc_fence_after(); }
          const float2 kp = *reinterpret_cast<const float2*>(kpos + 2 * (2 * t + g));
          const float ax = t5_slog(qx - kp.x), ay = t5_slog(qy - kp.y);
          const float axh = bf16_round(ax), ayh = bf16_round(ay);
          uint32_t a0[8];
          a0[0] = pack_bf16x2(axh, axh); a0[1] = pack_bf16x2(ax - axh, ayh); a0[2] = pack_bf16x2(ayh, ay - ayh);
          a0[3] = 0x3f803f80u; a0[4] = 0u; a0[5] = 0u; a0[6] = 0u; a0[7] = 0u;
          tc_st8(slot + T5_SLOT_A, a0);
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BA(0, g, s));
        }
        if (t >= 1 && t < T5_UPT + 1) {
          const uint32_t qq = qb + t - 1, s = qq % T5_SPW, ph = (qq / T5_SPW) & 1u;
          const uint32_t slot = trow + T5_COL_SLOT + (g * T5_SPW + s) * T5_SLOT_COLS;
          mbar_wait(BA(3, g, s), ph);
          tc_fence_after();
          tc_ld32(slot, r);
          uint32_t a1[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) a1[c] = relu_pack_bf16(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1]));
          tc_st16(slot + T5_SLOT_A, a1);
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BA(1, g, s));
        }
        if (t >= 2) {
          const uint32_t qq = qb + t - 2, s = qq % T5_SPW, ph = (qq / T5_SPW) & 1u;
          const uint32_t slot = trow + T5_COL_SLOT + (g * T5_SPW + s) * T5_SLOT_COLS;
          mbar_wait(BA(4, g, s), ph);
          tc_fence_after();
          tc_ld32(slot, r);
          uint32_t a2[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) a2[c] = relu_pack_bf16(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1]));
          tc_st16(slot + T5_SLOT_A, a2);
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BA(2, g, s));
        }
      }
      // ---- online softmax of the tile: warpgroup g owns logit columns [32 g, 32 g + 32)
      mbar_wait(bar_sfull, tile & 1u);
      tc_fence_after();
      tc_ld32(trow + T5_COL_S + 32 * g, r);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (k0 + 32 * g + j >= cf.nkv) r[j] = 0xff800000u;         // -inf
        mx = fmaxf(mx, __uint_as_float(r[j]));
      }
      stat[g * T5_Q + row] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, stat[(g ^ 1) * T5_Q + row]);
      const float m_new = fmaxf(m_run, mx);
      const float alpha = m_run == -INFINITY ? 0.f : __expf(m_run - m_new);
      m_run = m_new;
      float lsum = 0.f;
      uint32_t ph_[16], pl_[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float p0 = __expf(__uint_as_float(r[2 * c]) - m_new), p1 = __expf(__uint_as_float(r[2 * c + 1]) - m_new);
        lsum += p0 + p1;
        if (cf.p_drop > 0.f) {
          const unsigned long long key = (unsigned long long)(k0 + 32 * g + 2 * c);
          p0 = pair_hash(salted(cf.seed, cf.salt), hbase + key) >= thresh ? p0 * keep_scale : 0.f;
          p1 = pair_hash(salted(cf.seed, cf.salt), hbase + key + 1) >= thresh ? p1 * keep_scale : 0.f;
        }
        ph_[c] = pack_bf16x2(p0, p1);
        pl_[c] = pack_bf16x2(p0 - __uint_as_float(ph_[c] << 16), p1 - __uint_as_float(ph_[c] & 0xffff0000u));
      }
      l_part = l_part * alpha + lsum;
      tc_st16(trow + T5_COL_PH + 16 * g, ph_);
      tc_st16(trow + T5_COL_PL + 16 * g, pl_);
      if (tile > 0) {                                              // rescale this warpgroup's half of the O row
        tc_ld32(trow + T5_COL_O + 32 * g, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
        tc_st32(trow + T5_COL_O + 32 * g, r);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pready);
    }
    // ---- epilogue: O / l, log-sum-exp (+ b3, a per-row constant that the softmax itself does not see)
    mbar_wait(bar_pvdone, (ntiles - 1) & 1u);
    tc_fence_after();
    stat[g * T5_Q + row] = l_part;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float l = l_part + stat[(g ^ 1) * T5_Q + row];
    tc_ld32(trow + T5_COL_O + 32 * g, r);
    if (qi < cf.n) {
      const float inv = 1.f / l;
      float* orow = out + ((int64_t)b * cf.n + qi) * cf.inner + head * AD + 32 * g;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(orow + j) = make_float4(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv,
                                                           __uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
      if (g == 0) {
        const float b3 = cpb[CPB_H * 2 + CPB_H + CPB_H * CPB_H + CPB_H + CPB_H];
        lse[((int64_t)b * cf.heads + head) * cf.n + qi] = m_run + logf(l) + b3;
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace sml

using namespace sml;

/* Blackwell-native forward of the fused 2-D deformable attention (same contract as sml_deform_attn_fwd_tc with split = 0):
 * CD = 2, one head per offset group, head dim 64. */
extern "C" int sml_deform_attn_fwd_tc5(const int32_t* cfg8, float p_drop, uint64_t seed, const float* q, const float* k,
                                       const float* v, const float* vs, const float* cpb, float* out, float* lse, void* stream) {
  AttnCfg cf;
  if (make_acfg(cf, cfg8, p_drop, seed)) return 1;
  SML_CHECK(cf.CD == 2 && cf.OUT == 1, "sml_deform_attn_fwd_tc5: 2-D coordinates and one head per offset group only");
  SML_CHECK(cf.nkv >= 1 && cf.n >= 1, "sml_deform_attn_fwd_tc5: empty input");
  dim3 grid((cf.n + T5_Q - 1) / T5_Q, cf.heads, cf.B);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(deform_attn_fwd_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T5Smem::TOTAL);
    attr = true;
  }
  deform_attn_fwd_tc5_kernel<<<grid, T5_THREADS, T5Smem::TOTAL, as_stream(stream)>>>(cf, q, k, v, vs, cpb, out, lse);
  SML_LAUNCH_OK("sml_deform_attn_fwd_tc5");
  return 0;
}

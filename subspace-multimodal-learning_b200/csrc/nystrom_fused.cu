// nystrom_fused.cu — fused tcgen05 kernels for the token-by-landmark similarity paths of Nystrom attention
// (models/NystromAttention.py:122-140 and their backward), so that the [n', m] tensors are not re-materialised.
//
// sml_nystrom_a1_bwd: backward of  out = softmax(q k~^T) W  for one 128-token tile at a time, all on chip:
//     dA1 = dout W^T            (tcgen05, fp32 in TMEM, 256 columns)
//     dS  = A1 * (dA1 - rowsum(dA1 * A1))   (registers; A1 tile staged by TMA, overwritten in place by bf16 dS)
//     dq  = dS k~  (+ landmark-mean term)   (tcgen05, A = dS from shared memory, K-major)
//     dk~ += dS^T q                          (tcgen05, the SAME shared-memory dS tile read as an MN-major operand,
//                                             accumulated in TMEM over all tiles of a (b, h), flushed with atomics)
// It replaces: the dA1 GEMM (1.09 GB fp32 write at B=8), the softmax-backward pass (2.7 GB), and the two GEMMs that
// re-read dS (1.1 GB).  Specialised to the TransMIL configuration d = 64, m = 256 (mil.py:176-184).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue.
#include "tc_common.cuh"

namespace sml {

constexpr int FD = 64, FM = 256, FT = 128;          // head dim, landmarks, tokens per tile
constexpr int F_THREADS = 320;
struct FusedSmem {
  static constexpr int WM = 0;                        // [256 x 128 B] K-major B of MMA1 (rows = landmarks)
  static constexpr int KL = WM + FM * 128;            // [256 x 128 B] N-major B of MMA2 (rows = landmarks = K)
  static constexpr int DOUT = KL + FM * 128;          // [128 x 128 B] K-major A of MMA1
  static constexpr int Q = DOUT + FT * 128;           // [128 x 128 B] N-major B of MMA3 (rows = tokens = K)
  static constexpr int A1 = Q + FT * 128;             // 4 k-blocks of [128 x 128 B]: A1 tile, then dS in place
  static constexpr int BAR = A1 + 4 * FT * 128;       // mbarriers
  static constexpr int STATS = BAR + 64;              // [2 bufs][2 halves][128] floats
  static constexpr int TOTAL = STATS + 2 * 2 * FT * 4 + 1024;
};

struct A1BwdParams {
  int B, h, npad, l;                 // l = npad / m (tokens per landmark)
  int segs;                          // each (b,h) group's tiles are cut into `segs` contiguous segments = work units
  const float* dql;                  // [B,h,m,d] landmark gradient added to dq (may be null)
  float* dq; __nv_bfloat16* dq_hi;   // q chunk of dqkv: element (b, t, head, c) at ((b*npad + t)*3*h + head)*64 + c
  float* dkl;                        // [B,h,m,d] accumulated with atomics
};

__device__ __forceinline__ uint32_t sw128_off(int row, int chunk16) { return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4)); }

__global__ void __launch_bounds__(F_THREADS, 1)
nystrom_a1_bwd_kernel(const __grid_constant__ CUtensorMap tmDout, const __grid_constant__ CUtensorMap tmWm,
                      const __grid_constant__ CUtensorMap tmKl, const __grid_constant__ CUtensorMap tmA1,
                      const __grid_constant__ CUtensorMap tmQ, const A1BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FusedSmem::BAR);
  uint64_t* bar_w = bars + 0;      // weights (W, k~) of the current (b,h) landed
  uint64_t* bar_in = bars + 1;     // dout / q / A1 tile landed
  uint64_t* bar_m1 = bars + 2;     // dA1 accumulator complete
  uint64_t* bar_ds = bars + 3;     // dS written to shared memory by the 8 epilogue warps
  uint64_t* bar_m2 = bars + 4;     // dq / dk~ MMAs complete (also: this tile's shared-memory operands are free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* stats = reinterpret_cast<float*>(smem + FusedSmem::STATS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_bh = p.npad / FT;
  const int nunits = p.B * p.h * p.segs;
  // work unit = one contiguous segment of the tiles of one (b,h); units are dealt round-robin to CTAs.  Within a unit
  // W / k~ stay resident in shared memory and the dk~ accumulator stays in TMEM; it is flushed (atomics) once per unit.
#define UNIT_DECODE(u)                                                        \
  const int bh = (u) / p.segs, seg = (u) % p.segs;                             \
  const int t_lo = (int)((long long)tiles_per_bh * seg / p.segs);              \
  const int t_hi = (int)((long long)tiles_per_bh * (seg + 1) / p.segs);
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1); mbar_init(bar_in, 1); mbar_init(bar_m1, 1); mbar_init(bar_ds, 8); mbar_init(bar_m2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t COL_DA1 = 0, COL_DQ = 256, COL_DKL = 320;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0, ig = 0;
      for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ig) {
        UNIT_DECODE(u)
        const int b = bh / p.h, head = bh % p.h;
        for (int tl = t_lo; tl < t_hi; ++tl, ++it) {
          if (it > 0 && !mbar_wait(bar_m2, (it - 1) & 1u)) return;      // previous tile's MMAs are done with shared memory
          if (tl == t_lo) {
            mbar_expect_tx(bar_w, 2 * FM * 128);
            tma_load_4d(smem + FusedSmem::WM, &tmWm, bar_w, 0, 0, head, b);
            for (int c = 0; c < FM / 64; ++c) tma_load_4d(smem + FusedSmem::KL + c * 8192, &tmKl, bar_w, 0, c * 64, head, b);
          }
          const int t0 = tl * FT;
          mbar_expect_tx(bar_in, (2 + 4) * FT * 128);
          tma_load_4d(smem + FusedSmem::DOUT, &tmDout, bar_in, 0, t0, head, b);
          for (int c = 0; c < FT / 64; ++c) tma_load_4d(smem + FusedSmem::Q + c * 8192, &tmQ, bar_in, 0, t0 + c * 64, head, b);
          for (int kb = 0; kb < 4; ++kb) tma_load_4d(smem + FusedSmem::A1 + kb * FT * 128, &tmA1, bar_in, kb * 64, t0, head, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t id_m1 = make_idesc(128, 256, false, false);   // dA1 = dout(K-major) x W(K-major)
      constexpr uint32_t id_m2 = make_idesc(128, 64, false, true);     // dq  = dS(K-major)  x k~(N-major)
      constexpr uint32_t id_m3 = make_idesc(128, 64, true, true);      // dk~ = dS(MN-major) x q(N-major)
      const uint32_t sWm = smem_u32(smem + FusedSmem::WM), sKl = smem_u32(smem + FusedSmem::KL);
      const uint32_t sDo = smem_u32(smem + FusedSmem::DOUT), sQ = smem_u32(smem + FusedSmem::Q), sA1 = smem_u32(smem + FusedSmem::A1);
      uint32_t it = 0, ig = 0;
      for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ig) {
        UNIT_DECODE(u)
        (void)bh;
        for (int tl = t_lo; tl < t_hi; ++tl, ++it) {
          if (tl == t_lo && !mbar_wait(bar_w, ig & 1u)) return;
          if (!mbar_wait(bar_in, it & 1u)) return;
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < FD / 16; ++k)
            tc_mma_bf16(tmem_base + COL_DA1, make_smem_desc(sDo + k * 32, 16, 1024), make_smem_desc(sWm + k * 32, 16, 1024), id_m1, k != 0 ? 1u : 0u);
          tc_commit(bar_m1);
          if (!mbar_wait(bar_ds, it & 1u)) return;
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < FM / 16; ++ks)     // K = 256 landmarks; A: 4 k-blocks x 4 k-steps, B: 16-row groups of k~
            tc_mma_bf16(tmem_base + COL_DQ, make_smem_desc(sA1 + (ks >> 2) * (FT * 128) + (ks & 3) * 32, 16, 1024),
                        make_smem_desc(sKl + ks * 2048, 8192, 1024), id_m2, ks != 0 ? 1u : 0u);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)           // landmarks [0,128) and [128,256): two 64-wide k-blocks each (LBO = 16 KB)
#pragma unroll
            for (int ks = 0; ks < FT / 16; ++ks)   // K = 128 tokens
              tc_mma_bf16(tmem_base + COL_DKL + hf * 64, make_smem_desc(sA1 + hf * 2 * (FT * 128) + ks * 2048, FT * 128, 1024),
                          make_smem_desc(sQ + ks * 2048, 8192, 1024), id_m3, (tl != t_lo || ks != 0) ? 1u : 0u);
          tc_commit(bar_m2);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int qd = warp & 3, half = (warp - 2) >> 2;
    const int row = qd * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(qd * 32) << 16);
    uint8_t* a1s = smem + FusedSmem::A1;
    uint32_t r[32];
    uint32_t it = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      UNIT_DECODE(u)
      const int b = bh / p.h, head = bh % p.h;
      for (int tl = t_lo; tl < t_hi; ++tl, ++it) {
        const int t = tl * FT + row;
        mbar_wait(bar_m1, it & 1u);
        tc_fence_after();
        // ---- pass 1: partial row dot over this warp's 128 landmarks
        float dot = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col = half * 128 + c * 32;
          tc_ld32(trow + COL_DA1 + col, r);
          const int kb = col >> 6, j0 = (col & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 pk = *reinterpret_cast<const uint4*>(a1s + kb * (FT * 128) + sw128_off(row, j0 + j));
            const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[u]));
              dot = fmaf(a.x, __uint_as_float(r[j * 8 + 2 * u]), dot);
              dot = fmaf(a.y, __uint_as_float(r[j * 8 + 2 * u + 1]), dot);
            }
          }
        }
        float* st = stats + (it & 1u) * 2 * FT;
        st[half * FT + row] = dot;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        dot += st[(half ^ 1) * FT + row];
        // ---- pass 2: dS = A1 * (dA1 - dot), bf16, in place over the A1 tile
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col = half * 128 + c * 32;
          tc_ld32(trow + COL_DA1 + col, r);
          const int kb = col >> 6, j0 = (col & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4* ptr = reinterpret_cast<uint4*>(a1s + kb * (FT * 128) + sw128_off(row, j0 + j));
            const uint4 pk = *ptr;
            uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[u]));
              const __nv_bfloat162 ds = __floats2bfloat162_rn(a.x * (__uint_as_float(r[j * 8 + 2 * u]) - dot),
                                                              a.y * (__uint_as_float(r[j * 8 + 2 * u + 1]) - dot));
              w[u] = *reinterpret_cast<const uint32_t*>(&ds);
            }
            *ptr = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar_ds)) : "memory");
        // ---- dq tile: 64 columns, 32 per warp half; + landmark-mean term; fp32 + bf16 stores
        mbar_wait(bar_m2, it & 1u);
        tc_fence_after();
        tc_ld32(trow + COL_DQ + half * 32, r);
        if (t < p.npad) {
          const int64_t o = (((int64_t)b * p.npad + t) * 3 * p.h + head) * FD + half * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.dql) {
            const float* tr = p.dql + (((int64_t)bh * FM) + t / p.l) * FD + half * 32;
            const float sc = 1.f / (float)p.l;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 tv = __ldg(reinterpret_cast<const float4*>(tr + j));
              v[j] = fmaf(sc, tv.x, v[j]); v[j + 1] = fmaf(sc, tv.y, v[j + 1]); v[j + 2] = fmaf(sc, tv.z, v[j + 2]); v[j + 3] = fmaf(sc, tv.w, v[j + 3]);
            }
          }
          if (p.dq) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(p.dq + o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (p.dq_hi) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __nv_bfloat162 hb[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) hb[u] = __floats2bfloat162_rn(v[j + 2 * u], v[j + 2 * u + 1]);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&hb[0]); pk.y = *reinterpret_cast<uint32_t*>(&hb[1]);
              pk.z = *reinterpret_cast<uint32_t*>(&hb[2]); pk.w = *reinterpret_cast<uint32_t*>(&hb[3]);
              *reinterpret_cast<uint4*>(p.dq_hi + o + j) = pk;
            }
          }
        }
        // ---- last tile of this (b,h): flush the dk~ accumulator (warps of half hf own landmarks hf*128 + row)
        if (tl == t_hi - 1) {
          float* dst = p.dkl + ((int64_t)bh * FM + half * 128 + row) * FD;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            tc_ld32(trow + COL_DKL + half * 64 + c * 32, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + c * 32 + j, __uint_as_float(r[j]));
          }
        }
        tc_fence_before();
      }
    }
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

int sml_fused_timeout_flag_internal(void) {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, g_tc_timeout_flag, sizeof(v));
  return (int)v;
}

extern "C" int sml_nystrom_a1_bwd(const void* dout_bf16, const void* wm_bf16, const void* a1_bf16, const void* kl_bf16,
                                  const void* qkv_bf16, const float* dql, float* dqkv, void* dqkv_hi, float* dkl, int32_t B,
                                  int32_t npad, int32_t h, int32_t d, int32_t m, void* stream) {
  SML_CHECK(d == FD && m == FM, "sml_nystrom_a1_bwd: specialised to d = 64, m = 256 (got d=%d m=%d)", d, m);
  SML_CHECK(npad % m == 0, "sml_nystrom_a1_bwd: npad must be a multiple of m");
  const int64_t hd = (int64_t)h * d;
  CUtensorMap tDout, tWm, tKl, tA1, tQ;
  // dout [B, npad, h*d]: head slice = inner 64, rows = tokens
  if (make_map(&tDout, dout_bf16, FD, npad, hd, h, FD, B, (int64_t)npad * hd, 64, FT, "dout")) return 1;
  if (make_map(&tWm, wm_bf16, FD, FM, FD, h, (int64_t)FM * FD, B, (int64_t)h * FM * FD, 64, FM, "wm")) return 1;
  if (make_map(&tKl, kl_bf16, FD, FM, FD, h, (int64_t)FM * FD, B, (int64_t)h * FM * FD, 64, 64, "kl")) return 1;
  if (make_map(&tA1, a1_bf16, FM, npad, FM, h, (int64_t)npad * FM, B, (int64_t)h * npad * FM, 64, FT, "a1")) return 1;
  if (make_map(&tQ, qkv_bf16, FD, npad, 3 * hd, h, FD, B, (int64_t)npad * 3 * hd, 64, 64, "q")) return 1;
  A1BwdParams p;
  p.B = B; p.h = h; p.npad = npad; p.l = npad / m;
  p.dql = dql; p.dq = dqkv; p.dq_hi = (__nv_bfloat16*)dqkv_hi; p.dkl = dkl;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(nystrom_a1_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedSmem::TOTAL);
    SML_CHECK(e == cudaSuccess, "sml_nystrom_a1_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int nbh = B * h, tiles = npad / FT;
  int segs = kNumSMs / nbh;                       // cut each (b,h) so that about one unit lands on every SM
  if (segs < 1) segs = 1;
  if (segs > tiles) segs = tiles;
  p.segs = segs;
  const int units = nbh * segs;
  const int grid = units < kNumSMs ? units : kNumSMs;
  nystrom_a1_bwd_kernel<<<grid, F_THREADS, FusedSmem::TOTAL, as_stream(stream)>>>(tDout, tWm, tKl, tA1, tQ, p);
  SML_LAUNCH_OK("sml_nystrom_a1_bwd");
  return 0;
}

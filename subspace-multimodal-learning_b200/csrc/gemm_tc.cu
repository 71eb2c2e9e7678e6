// gemm_tc.cu — tcgen05 / TMEM / TMA batched GEMM for sm_100a.
//
//   C[b1,b2] = epilogue( alpha * A[b1,b2] (MxK, bf16) * B[b1,b2] (KxN, bf16) ),  fp32 accumulation in TMEM.
//
// * operands are staged global -> shared by TMA (cp.async.bulk.tensor.4d, SWIZZLE_128B, zero OOB fill so
//   M/N/K tails need no special casing) into a STAGES-deep mbarrier ring;
// * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BN x 16) on shared-memory
//   descriptors; both K-major and MN-major operands are supported (4 layout combinations), so every
//   forward and backward contraction of the path runs without a transposed copy;
// * optional bf16x3 split (A_hi*B_hi + A_hi*B_lo + A_lo*B_hi) gives fp32-grade products (~2^-16) for the
//   fp32 parity mode and the Moore-Penrose chain;
// * 4 epilogue warps read the accumulator with tcgen05.ld (one TMEM lane = one output row per thread) and
//   apply alpha / beta*C / diag*I / bias / ReLU or a fused row softmax (row fits one tile: the three
//   Nystrom similarity softmaxes over m landmarks), writing fp32 and/or bf16 (hi, lo) results.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue
// (two warps per TMEM lane quarter, each taking half of the tile's columns).
#include "tc_common.cuh"

namespace sml {

struct EpiParams {
  int M, N, K, nb2;
  int a_c1, a_c2, b_c1, b_c2;   // 0 when that batch dim is broadcast for the operand (TMA coordinate stays 0)
  float* C; int64_t sc_m, sc_n, sc_b1, sc_b2;
  __nv_bfloat16* Cb_hi; __nv_bfloat16* Cb_lo;
  float alpha, beta, diag;
  const float* bias;
  int relu, softmax;
  int ksplit, kb_per_split;     // split-K: extra tile-list dimension; partial sums are red.add'ed into a zeroed C
  const float* rowtab; int64_t rt_m, rt_b1, rt_b2; int rowdiv; float rowtab_scale;
  int vec_ok;                   // rows 16-byte aligned, unit column stride: 128-bit epilogue loads / stores
  int tiles_n, tiles_m, total_tiles;   // persistent tile list: n fastest, then m, then split-K index, then batch
};

constexpr int TC_BM = 128, TC_BK = 64, TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;   // 16 KB

constexpr int TC_STG_LD = 36;                                   // staging row stride in floats (32 + 4 pad: conflict-free 128-bit access)
constexpr int TC_STG_BYTES = 8 * 32 * TC_STG_LD * 4;            // one 32x32 fp32 tile per epilogue warp

template <int BN, bool SPLIT, int STAGES>
struct TcSmem {
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = (TC_A_BYTES + B_BYTES) * (SPLIT ? 2 : 1);
  static constexpr bool STAGED = STAGES * STAGE_BYTES + TC_STG_BYTES + 8192 <= 227 * 1024;   // room for the coalescing tiles?
  static constexpr int TOTAL = STAGES * STAGE_BYTES + (2 * STAGES + 4) * 8 + 16 + 2 * 2 * TC_BM * 8 + (STAGED ? TC_STG_BYTES : 0) + 1024;
};

__device__ __forceinline__ void store_row_chunk(const EpiParams& ep, float* crow, __nv_bfloat16* hrow,
                                                __nv_bfloat16* lrow, int gm, int gn0, const float (&v)[32]) {
  // one thread owns 32 consecutive columns of one output row
  const bool vec = (ep.sc_n == 1) && (gn0 + 32 <= ep.N);
  if (crow) {
    if (vec && ((reinterpret_cast<uintptr_t>(crow + gn0) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(crow + gn0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (gn0 + j < ep.N) crow[(int64_t)(gn0 + j) * ep.sc_n] = v[j];
    }
  }
  if (hrow) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (gn0 + j < ep.N) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(v[j]);
        hrow[(int64_t)(gn0 + j) * ep.sc_n] = hi;
        if (lrow) lrow[(int64_t)(gn0 + j) * ep.sc_n] = __float2bfloat16_rn(v[j] - __bfloat162float(hi));
      }
    }
  }
  (void)gm;
}



// Coalesced epilogue stores: the warp's 32 rows x 32 columns go through a padded shared-memory tile so that every
// global store instruction writes 4 complete 128-byte row segments (row-per-thread stores touch 32 lines each and
// saturate the LSU's one-line-per-cycle address path: measured +8 us on a 128x256 tile with three output streams).
__device__ __forceinline__ void store_chunk_staged(const EpiParams& ep, float* stg, int lane, int row0, int64_t batch_off,
                                                   int gn0, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + lane * TC_STG_LD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int rr = lane >> 3, col = (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + rr;
    const int grow = row0 + r;
    if (grow < ep.M) {
      const float4 val = *reinterpret_cast<const float4*>(stg + r * TC_STG_LD + col);
      const int64_t off = batch_off + (int64_t)grow * ep.sc_m + gn0 + col;
      if (ep.C) *reinterpret_cast<float4*>(ep.C + off) = val;
      if (ep.Cb_hi) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(val.x, val.y), h1 = __floats2bfloat162_rn(val.z, val.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&h0); pk.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(ep.Cb_hi + off) = pk;
        if (ep.Cb_lo) {
          const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
          const __nv_bfloat162 l0 = __floats2bfloat162_rn(val.x - f0.x, val.y - f0.y), l1 = __floats2bfloat162_rn(val.z - f1.x, val.w - f1.y);
          pk.x = *reinterpret_cast<const uint32_t*>(&l0); pk.y = *reinterpret_cast<const uint32_t*>(&l1);
          *reinterpret_cast<uint2*>(ep.Cb_lo + off) = pk;
        }
      }
    }
  }
  __syncwarp();
}

struct TileCoord { int n0, m0, b1, b2, ks, kb_lo, nkb; };
__device__ __forceinline__ TileCoord decode_tile(const EpiParams& ep, int tile, int BN) {
  TileCoord t;
  const int nt = tile % ep.tiles_n;
  int rest = tile / ep.tiles_n;
  const int mt = rest % ep.tiles_m;
  rest /= ep.tiles_m;
  t.ks = rest % ep.ksplit;
  const int bz = rest / ep.ksplit;
  t.n0 = nt * BN; t.m0 = mt * TC_BM;
  t.b1 = bz / ep.nb2; t.b2 = bz % ep.nb2;
  const int nkb_all = (ep.K + TC_BK - 1) / TC_BK;
  t.kb_lo = t.ks * ep.kb_per_split;
  t.nkb = min(nkb_all, t.kb_lo + ep.kb_per_split) - t.kb_lo;
  return t;
}

// Epilogue of one 128 x BN tile: TMEM accumulator -> registers -> (alpha, beta*C, diag, bias, row-table, ReLU | row softmax)
// -> fp32 and/or bf16 hi/lo global stores.  Called by the 8 epilogue warps (warp index 2..9).
template <int BN, bool STAGED, bool LEAN>
__device__ __forceinline__ void epilogue_tile(const EpiParams& ep, const TileCoord& t, uint32_t tmem_acc, int warp, int lane,
                                              float2* stats_buf, float* staging) {
  const int q = warp & 3;                       // TMEM lane quarter this warp may access
  const int half = (warp - 2) >> 2;             // which half of the tile's columns this warp handles
  const int row = q * 32 + lane;
  constexpr int CH = BN / 2;                    // columns per epilogue warp
  const int cbeg = half * CH;
  uint32_t r[32];
  // LEAN (the Moore-Penrose chain: alpha, beta C, diag I, fp32 + bf16 hi / lo out, full vector chunks): split-K, softmax, bias,
  // row table, ReLU and the ragged-tail path are compiled out.  The generic kernel is ~9.7 k SASS instructions, more than
  // the instruction cache holds, and its cold epilogue showed up as instruction-fetch stalls in the 25 us chain launches.
  const int e_ksplit = LEAN ? 1 : ep.ksplit, e_softmax = LEAN ? 0 : ep.softmax, e_relu = LEAN ? 0 : ep.relu;
  const float* e_bias = LEAN ? nullptr : ep.bias;
  const float* e_rowtab = LEAN ? nullptr : ep.rowtab;
  const int e_vec = LEAN ? 1 : ep.vec_ok;
  {
      const int n0 = t.n0, m0 = t.m0, b1 = t.b1, b2 = t.b2, ks = t.ks;
      const int gm = m0 + row;
      const uint32_t trow = tmem_acc + ((uint32_t)(q * 32) << 16);
      const bool row_ok = gm < ep.M;
      const int64_t coff = (int64_t)b1 * ep.sc_b1 + (int64_t)b2 * ep.sc_b2 + (int64_t)gm * ep.sc_m;
      float* crow = ep.C ? ep.C + coff : nullptr;
      __nv_bfloat16* hrow = ep.Cb_hi ? ep.Cb_hi + coff : nullptr;
      __nv_bfloat16* lrow = ep.Cb_lo ? ep.Cb_lo + coff : nullptr;
      if (e_ksplit > 1) {
        // split-K partial: alpha * acc (+ bias / diag on split 0) accumulated with red.global.add into the zeroed C
        for (int c = cbeg; c < cbeg + CH; c += 32) {
          tc_ld32(trow + c, r);
          if (!(row_ok && n0 + c < ep.N)) continue;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gn = n0 + c + j;
            if (gn < ep.N) {
              float x = ep.alpha * __uint_as_float(r[j]);
              if (ks == 0) { if (gn == gm) x += ep.diag; if (e_bias) x += __ldg(e_bias + gn); }
              atomicAdd(crow + (int64_t)gn * ep.sc_n, x);
            }
          }
        }
      } else {
        float mx = -INFINITY, inv = 1.f;
        if (e_softmax && !ep.C && !ep.Cb_lo && ep.Cb_hi && e_vec && n0 + BN <= ep.N) {
          // bf16-only probabilities (bf16 mode), ~7 instructions per element (was ~18: this epilogue, not the MMAs or HBM, bounds
          // the [n', m] similarity + softmax launches).  Pass 1: row maximum of the raw accumulators (alpha > 0: one FMNMX per
          // element).  Pass 2: e = 2^(s * alpha log2 e - max * alpha log2 e) as one FFMA + one MUFU.EX2 (ftz, no range guards),
          // summed in fp32 and kept as packed bf16 pairs in registers.  Pass 3: scale by 1 / sum in fp32, repack, and store the
          // row's 64-byte pieces straight from registers (no shared-memory staging).
          const float sc = ep.alpha * 1.4426950408889634f;
          float m_ = -INFINITY;
          for (int c = cbeg; c < cbeg + CH; c += 32) {
            tc_ld32(trow + c, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) m_ = fmaxf(m_, __uint_as_float(r[j]));
          }
          float2* st = stats_buf;
          st[half * TC_BM + row].x = m_;
          asm volatile("bar.sync 1, 256;" ::: "memory");            // the 8 epilogue warps only
          mx = fmaxf(m_, st[(half ^ 1) * TC_BM + row].x);
          const float off = -sc * mx;
          uint32_t eb[CH / 2];
          float l_ = 0.f;
#pragma unroll
          for (int cc = 0; cc < CH / 32; ++cc) {
            tc_ld32(trow + cbeg + cc * 32, r);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float e0, e1;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(r[j]), sc, off)));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(r[j + 1]), sc, off)));
              l_ += e0 + e1;
              const __nv_bfloat162 pe = __floats2bfloat162_rn(e0, e1);
              eb[cc * 16 + j / 2] = *reinterpret_cast<const uint32_t*>(&pe);
            }
          }
          st[half * TC_BM + row].y = l_;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          inv = 1.f / (l_ + st[(half ^ 1) * TC_BM + row].y);
          if (row_ok) {
#pragma unroll
            for (int cc = 0; cc < CH / 32; ++cc) {
              __nv_bfloat16* dst = hrow + n0 + cbeg + cc * 32;
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                uint32_t w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const uint32_t pe = eb[cc * 16 + j + u];
                  const __nv_bfloat162 o = __floats2bfloat162_rn(__uint_as_float(pe << 16) * inv, __uint_as_float(pe & 0xffff0000u) * inv);
                  w[u] = *reinterpret_cast<const uint32_t*>(&o);
                }
                *reinterpret_cast<uint4*>(dst + 2 * j) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
          }
          return;
        }
        if (e_softmax) {
          // Row softmax fused in the epilogue.  A row lives in one TMEM lane; the two warps sharing a lane quarter each
          // scan their own half of the columns ONCE with an online (max, sum) update and swap the pair through shared
          // memory, so the statistics cost one TMEM pass.
          float m_ = -INFINITY, l_ = 0.f;
          for (int c = cbeg; c < cbeg + CH; c += 32) {
            if (n0 + c >= ep.N) break;
            tc_ld32(trow + c, r);
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c + j < ep.N) cm = fmaxf(cm, ep.alpha * __uint_as_float(r[j]));
            const float mn = fmaxf(m_, cm);
            float accs = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c + j < ep.N) accs += __expf(ep.alpha * __uint_as_float(r[j]) - mn);
            l_ = l_ * __expf(m_ - mn) + accs;            // exp(-inf) = 0 on the first chunk
            m_ = mn;
          }
          float2* st = stats_buf;
          st[half * TC_BM + row] = make_float2(m_, l_);
          asm volatile("bar.sync 1, 256;" ::: "memory");            // the 8 epilogue warps only
          const float2 o = st[(half ^ 1) * TC_BM + row];
          mx = fmaxf(m_, o.x);
          const float lsum = (m_ == -INFINITY ? 0.f : l_ * __expf(m_ - mx)) + (o.x == -INFINITY ? 0.f : o.y * __expf(o.x - mx));
          inv = 1.f / lsum;
        }
        for (int c = cbeg; c < cbeg + CH; c += 32) {
          tc_ld32(trow + c, r);
          const int gn0 = n0 + c;
          if (gn0 >= ep.N) break;
          float v[32];
          if (e_vec && gn0 + 32 <= ep.N) {
            // ---- lean path: full 32-column chunk, 16-byte aligned rows, unit column stride
            if (e_softmax) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __expf(ep.alpha * __uint_as_float(r[j]) - mx) * inv;
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = ep.alpha * __uint_as_float(r[j]);
              if (ep.beta != 0.f && row_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 cv = *reinterpret_cast<const float4*>(crow + gn0 + j);
                  v[j] = fmaf(ep.beta, cv.x, v[j]); v[j + 1] = fmaf(ep.beta, cv.y, v[j + 1]);
                  v[j + 2] = fmaf(ep.beta, cv.z, v[j + 2]); v[j + 3] = fmaf(ep.beta, cv.w, v[j + 3]);
                }
              }
              if (ep.diag != 0.f) {
                const int dj = gm - gn0;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += (j == dj) ? ep.diag : 0.f;
              }
              if (e_bias) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 bv = __ldg(reinterpret_cast<const float4*>(e_bias + gn0 + j));
                  v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                }
              }
              if (e_rowtab && row_ok) {                     // rows past M must not index the table
                const float* tr = e_rowtab + (int64_t)b1 * ep.rt_b1 + (int64_t)b2 * ep.rt_b2 + (int64_t)(gm / ep.rowdiv) * ep.rt_m + gn0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 tv = __ldg(reinterpret_cast<const float4*>(tr + j));
                  v[j] = fmaf(ep.rowtab_scale, tv.x, v[j]); v[j + 1] = fmaf(ep.rowtab_scale, tv.y, v[j + 1]);
                  v[j + 2] = fmaf(ep.rowtab_scale, tv.z, v[j + 2]); v[j + 3] = fmaf(ep.rowtab_scale, tv.w, v[j + 3]);
                }
              }
              if (e_relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
              }
            }
            if (STAGED) {
              store_chunk_staged(ep, staging + (warp - 2) * 32 * TC_STG_LD, lane, m0 + q * 32, (int64_t)b1 * ep.sc_b1 + (int64_t)b2 * ep.sc_b2, gn0, v);
            } else if (row_ok) {
              if (crow) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(crow + gn0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              }
              if (hrow) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  __nv_bfloat162 h[4];
                  uint4 pk;
#pragma unroll
                  for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(v[j + 2 * u], v[j + 2 * u + 1]);
                  pk.x = *reinterpret_cast<uint32_t*>(&h[0]); pk.y = *reinterpret_cast<uint32_t*>(&h[1]);
                  pk.z = *reinterpret_cast<uint32_t*>(&h[2]); pk.w = *reinterpret_cast<uint32_t*>(&h[3]);
                  *reinterpret_cast<uint4*>(hrow + gn0 + j) = pk;
                  if (lrow) {
                    __nv_bfloat162 l[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                      const float2 f = __bfloat1622float2(h[u]);
                      l[u] = __floats2bfloat162_rn(v[j + 2 * u] - f.x, v[j + 2 * u + 1] - f.y);
                    }
                    pk.x = *reinterpret_cast<uint32_t*>(&l[0]); pk.y = *reinterpret_cast<uint32_t*>(&l[1]);
                    pk.z = *reinterpret_cast<uint32_t*>(&l[2]); pk.w = *reinterpret_cast<uint32_t*>(&l[3]);
                    *reinterpret_cast<uint4*>(lrow + gn0 + j) = pk;
                  }
                }
              }
            }
          } else if (!LEAN) {
            // ---- general path (N tail, odd strides)
            if (!row_ok) continue;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int gn = gn0 + j;
              float x = ep.alpha * __uint_as_float(r[j]);
              if (e_softmax) x = __expf(x - mx) * inv;
              else if (gn < ep.N) {
                if (ep.beta != 0.f) x = fmaf(ep.beta, crow[(int64_t)gn * ep.sc_n], x);
                if (gn == gm) x += ep.diag;
                if (e_bias) x += __ldg(e_bias + gn);
                if (e_rowtab) x = fmaf(ep.rowtab_scale, __ldg(e_rowtab + (int64_t)b1 * ep.rt_b1 + (int64_t)b2 * ep.rt_b2 + (int64_t)(gm / ep.rowdiv) * ep.rt_m + gn), x);
                if (e_relu) x = fmaxf(x, 0.f);
              }
              v[j] = x;
            }
            store_row_chunk(ep, crow, hrow, lrow, gm, gn0, v);
          }
        }
      }
  }
}

// Persistent kernel: gridDim.x CTAs (<= #SMs) walk the tile list with stride gridDim.x.  The fp32 accumulator is
// double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the TMA + MMA main loop of tile i+1.
template <int BN, bool A_MN, bool B_MN, bool SPLIT, int STAGES, bool LEAN = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBl, const EpiParams ep) {
  using S = TcSmem<BN, SPLIT, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float2* stats = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(tmem_slot) + 16);   // [2 bufs][2 halves][128 rows]
  float* staging = reinterpret_cast<float*>(stats + 2 * 2 * TC_BM);                        // [8 warps][32][TC_STG_LD] if S::STAGED

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }   // 8 epilogue warps arrive
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation: two accumulator buffers of BN fp32 columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Programmatic dependent launch: let the next kernel in the stream start its own prologue now, and do not touch
  // global memory ourselves until every prerequisite kernel has completed and flushed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < ep.total_tiles && ok; tile += gridDim.x) {
        const TileCoord t = decode_tile(ep, tile, BN);
        const int ab1 = t.b1 * ep.a_c1, ab2 = t.b2 * ep.a_c2, bb1 = t.b1 * ep.b_c1, bb2 = t.b2 * ep.b_c2;
        for (int kb = 0; kb < t.nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          if (!mbar_wait(&empty_bar[s], ph ^ 1u)) { ok = false; break; }
          mbar_expect_tx(&full_bar[s], (uint32_t)S::STAGE_BYTES);
          uint8_t* a_dst = smem + s * S::STAGE_BYTES;
          uint8_t* b_dst = a_dst + TC_A_BYTES * (SPLIT ? 2 : 1);
          const int k0 = (t.kb_lo + kb) * TC_BK;
#pragma unroll
          for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
            const CUtensorMap* ta = part ? &tmAl : &tmA;
            const CUtensorMap* tb = part ? &tmBl : &tmB;
            uint8_t* ad = a_dst + part * TC_A_BYTES;
            uint8_t* bd = b_dst + part * S::B_BYTES;
            if (!A_MN) tma_load_4d(ad, ta, &full_bar[s], k0, t.m0, ab2, ab1);
            else {
#pragma unroll
              for (int c = 0; c < TC_BM / 64; ++c) tma_load_4d(ad + c * 8192, ta, &full_bar[s], t.m0 + c * 64, k0, ab2, ab1);
            }
            if (!B_MN) tma_load_4d(bd, tb, &full_bar[s], k0, t.n0, bb2, bb1);
            else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c) tma_load_4d(bd + c * 8192, tb, &full_bar[s], t.n0 + c * 64, k0, bb2, bb1);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TC_BM, BN, A_MN, B_MN);
      // K-major: 8-row groups 1024 B apart (SBO); LBO unused.  MN-major: 64-element chunks 8 KB apart (LBO),
      // 8-k-row groups 1024 B apart (SBO).
      constexpr uint32_t a_lbo = A_MN ? 8192u : 16u, b_lbo = B_MN ? 8192u : 16u;
      constexpr uint32_t a_kstep = A_MN ? 2048u : 32u, b_kstep = B_MN ? 2048u : 32u;   // bytes per UMMA_K = 16
      uint32_t it = 0, lt = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < ep.total_tiles && ok; tile += gridDim.x, ++lt) {
        const TileCoord t = decode_tile(ep, tile, BN);
        const uint32_t a = lt & 1u;
        if (!mbar_wait(&tmem_empty[a], ((lt >> 1) & 1u) ^ 1u)) break;    // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + a * BN;
        for (int kb = 0; kb < t.nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          if (!mbar_wait(&full_bar[s], ph)) { ok = false; break; }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
          const uint32_t b_addr = a_addr + TC_A_BYTES * (SPLIT ? 2 : 1);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * a_kstep, a_lbo, 1024u);
            const uint64_t db = make_smem_desc(b_addr + k * b_kstep, b_lbo, 1024u);
            tc_mma_bf16(acc, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            if (SPLIT) {
              const uint64_t dal = make_smem_desc(a_addr + TC_A_BYTES + k * a_kstep, a_lbo, 1024u);
              const uint64_t dbl = make_smem_desc(b_addr + S::B_BYTES + k * b_kstep, b_lbo, 1024u);
              tc_mma_bf16(acc, da, dbl, idesc, 1u);
              tc_mma_bf16(acc, dal, db, idesc, 1u);
            }
          }
          tc_commit(&empty_bar[s]);   // frees the smem slot once these MMAs have read it
        }
        if (ok) tc_commit(&tmem_full[a]);   // accumulator of this tile complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < ep.total_tiles; tile += gridDim.x, ++lt) {
      const TileCoord t = decode_tile(ep, tile, BN);
      const uint32_t a = lt & 1u;
      mbar_wait(&tmem_full[a], (lt >> 1) & 1u);   // on a (bounded) timeout keep going: all 8 warps must reach the named barrier
      tc_fence_after();
      epilogue_tile<BN, S::STAGED, LEAN>(ep, t, tmem_base + a * BN, warp, lane, stats + (lt & 1u) * 2 * TC_BM, staging);
      // this accumulator buffer may be overwritten by the MMA of tile lt + 2
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[a])) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ------------------------------------------------------------------------------------------ GEMM programs
// A "program" is a list of dependent GEMM steps (the Moore-Penrose chain: 24 forward / 66 backward 256^3 products per
// layer) executed by ONE persistent cooperative launch: every CTA walks the tile list of step s, then all CTAs meet
// at a grid barrier before any of them loads operands of step s+1.  This removes the per-launch floor (launch,
// TMEM alloc, barrier init, pipeline fill, teardown: ~8 us, measured) that dominated the chain's 13-24 us launches.
// Operand majors are run-time per step here (the chain mixes A, A^T, z, z^T ...).
struct ProgStep { CUtensorMap tm[4]; EpiParams ep; int a_mn, b_mn; };
constexpr int PROG_MAX_STEPS = 36;
struct Program { int nsteps; int pad_[15]; ProgStep st[PROG_MAX_STEPS]; };

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __threadfence();                 // this CTA's global stores of the finished step
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1u);
    unsigned int seen = 0;
    for (uint32_t it = 0; it < (1u << 24); ++it) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (seen >= target) break;
      __nanosleep(64);
    }
    if (seen < target) { atomicExch(&g_tc_timeout_flag, 2u); __threadfence_system(); __trap(); }   // fail loudly, never continue
    asm volatile("fence.proxy.async;" ::: "memory");    // later TMA (async proxy) reads see the other CTAs' generic-proxy stores
  }
  __syncthreads();
}

template <int BN, bool SPLIT, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_program_kernel(const __grid_constant__ Program prog, unsigned int* sync_counter) {
  using S = TcSmem<BN, SPLIT, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float2* stats = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(tmem_slot) + 16);
  float* staging = reinterpret_cast<float*>(stats + 2 * 2 * TC_BM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  uint32_t it = 0, lt = 0;           // ring / accumulator counters run on across steps (each role keeps its own copy)
  for (int step = 0; step < prog.nsteps; ++step) {
    const ProgStep& ps = prog.st[step];
    const EpiParams& ep = ps.ep;
    const bool a_mn = ps.a_mn != 0, b_mn = ps.b_mn != 0;
    if (warp == 0) {
      if (lane == 0) {
        bool ok = true;
        for (int tile = blockIdx.x; tile < ep.total_tiles && ok; tile += gridDim.x) {
          const TileCoord t = decode_tile(ep, tile, BN);
          const int ab1 = t.b1 * ep.a_c1, ab2 = t.b2 * ep.a_c2, bb1 = t.b1 * ep.b_c1, bb2 = t.b2 * ep.b_c2;
          for (int kb = 0; kb < t.nkb; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            if (!mbar_wait(&empty_bar[s], ph ^ 1u)) { ok = false; break; }
            mbar_expect_tx(&full_bar[s], (uint32_t)S::STAGE_BYTES);
            uint8_t* a_dst = smem + s * S::STAGE_BYTES;
            uint8_t* b_dst = a_dst + TC_A_BYTES * (SPLIT ? 2 : 1);
            const int k0 = (t.kb_lo + kb) * TC_BK;
#pragma unroll
            for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
              const CUtensorMap* ta = &ps.tm[part];
              const CUtensorMap* tb = &ps.tm[2 + part];
              uint8_t* ad = a_dst + part * TC_A_BYTES;
              uint8_t* bd = b_dst + part * S::B_BYTES;
              if (!a_mn) tma_load_4d(ad, ta, &full_bar[s], k0, t.m0, ab2, ab1);
              else
                for (int c = 0; c < TC_BM / 64; ++c) tma_load_4d(ad + c * 8192, ta, &full_bar[s], t.m0 + c * 64, k0, ab2, ab1);
              if (!b_mn) tma_load_4d(bd, tb, &full_bar[s], k0, t.n0, bb2, bb1);
              else
                for (int c = 0; c < BN / 64; ++c) tma_load_4d(bd + c * 8192, tb, &full_bar[s], t.n0 + c * 64, k0, bb2, bb1);
            }
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc(TC_BM, BN, a_mn, b_mn);
        const uint32_t a_lbo = a_mn ? 8192u : 16u, b_lbo = b_mn ? 8192u : 16u;
        const uint32_t a_kstep = a_mn ? 2048u : 32u, b_kstep = b_mn ? 2048u : 32u;
        bool ok = true;
        for (int tile = blockIdx.x; tile < ep.total_tiles && ok; tile += gridDim.x, ++lt) {
          const TileCoord t = decode_tile(ep, tile, BN);
          const uint32_t a = lt & 1u;
          if (!mbar_wait(&tmem_empty[a], ((lt >> 1) & 1u) ^ 1u)) break;
          tc_fence_after();
          const uint32_t acc = tmem_base + a * BN;
          for (int kb = 0; kb < t.nkb; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            if (!mbar_wait(&full_bar[s], ph)) { ok = false; break; }
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
            const uint32_t b_addr = a_addr + TC_A_BYTES * (SPLIT ? 2 : 1);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint64_t da = make_smem_desc(a_addr + k * a_kstep, a_lbo, 1024u);
              const uint64_t db = make_smem_desc(b_addr + k * b_kstep, b_lbo, 1024u);
              tc_mma_bf16(acc, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
              if (SPLIT) {
                const uint64_t dal = make_smem_desc(a_addr + TC_A_BYTES + k * a_kstep, a_lbo, 1024u);
                const uint64_t dbl = make_smem_desc(b_addr + S::B_BYTES + k * b_kstep, b_lbo, 1024u);
                tc_mma_bf16(acc, da, dbl, idesc, 1u);
                tc_mma_bf16(acc, dal, db, idesc, 1u);
              }
            }
            tc_commit(&empty_bar[s]);
          }
          if (ok) tc_commit(&tmem_full[a]);
        }
      }
    } else {
      for (int tile = blockIdx.x; tile < ep.total_tiles; tile += gridDim.x, ++lt) {
        const TileCoord t = decode_tile(ep, tile, BN);
        const uint32_t a = lt & 1u;
        mbar_wait(&tmem_full[a], (lt >> 1) & 1u);
        tc_fence_after();
        epilogue_tile<BN, S::STAGED, false>(ep, t, tmem_base + a * BN, warp, lane, stats + (lt & 1u) * 2 * TC_BM, staging);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[a])) : "memory");
      }
    }
    if (step + 1 < prog.nsteps) grid_barrier(sync_counter, (unsigned)(step + 1) * gridDim.x);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host side
template <int BN, bool A_MN, bool B_MN, bool SPLIT, int STAGES, bool LEAN = false>
static int launch_tc(const CUtensorMap* tms, const EpiParams& ep, dim3 grid, cudaStream_t st) {
  using S = TcSmem<BN, SPLIT, STAGES>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, SPLIT, STAGES, LEAN>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    SML_CHECK(e == cudaSuccess, "sml_gemm_tc: cudaFuncSetAttribute(%d B) failed: %s", S::TOTAL, cudaGetErrorString(e));
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr1[1];
  attr1[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr1[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr1; cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tms[0], tms[1], tms[2], tms[3], ep);
  SML_CHECK(le == cudaSuccess, "sml_gemm_tc: launch failed: %s", cudaGetErrorString(le));
  SML_LAUNCH_OK("sml_gemm_tc");
  return 0;
}

template <int BN, bool SPLIT, int STAGES, bool LEAN = false>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap* tms, const EpiParams& ep, dim3 grid, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_tc<BN, false, false, SPLIT, STAGES, LEAN>(tms, ep, grid, st);
  if (!a_mn && b_mn) return launch_tc<BN, false, true, SPLIT, STAGES, LEAN>(tms, ep, grid, st);
  if (a_mn && !b_mn) return launch_tc<BN, true, false, SPLIT, STAGES, LEAN>(tms, ep, grid, st);
  return launch_tc<BN, true, true, SPLIT, STAGES, LEAN>(tms, ep, grid, st);
}

}  // namespace sml

using namespace sml;

int sml_fused_timeout_flag_internal(void);     // nystrom_fused.cu's copy of the flag
extern "C" int sml_tc_timeout_flag(void) {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, g_tc_timeout_flag, sizeof(v));
  return (int)v | sml_fused_timeout_flag_internal();
}

struct TcPlan { CUtensorMap tms[4]; sml::EpiParams ep; bool a_mn, b_mn, split; int BN; dim3 grid; };

// validate a descriptor, build tensor maps + epilogue parameters.  force_bn > 0 pins the tile width (programs).
static int prepare_tc(const sml_gemm_desc* d, int force_bn, bool allow_splitk, cudaStream_t stream, TcPlan* plan) {
  SML_CHECK(d && d->A && d->B && (d->C || d->Cb_hi), "sml_gemm_tc: null pointer");
  SML_CHECK(d->M > 0 && d->N > 0 && d->K > 0 && d->nb1 > 0 && d->nb2 > 0, "sml_gemm_tc: bad sizes");
  const bool split = d->A_lo != nullptr || d->B_lo != nullptr;
  SML_CHECK(!split || (d->A_lo && d->B_lo), "sml_gemm_tc: bf16x3 split needs both A_lo and B_lo");
  // operand majors: A is [M,K]; K-major if sa_k == 1, M-major if sa_m == 1.  B is [K,N]; K-major if sb_k == 1.
  bool a_mn, b_mn;
  if (d->sa_k == 1 && (d->sa_m != 1 || d->K == 1)) a_mn = false;
  else if (d->sa_m == 1) a_mn = true;
  else { set_error("sml_gemm_tc: A needs unit stride along M or K (sa_m=%lld sa_k=%lld)", (long long)d->sa_m, (long long)d->sa_k); return 1; }
  if (d->sb_k == 1 && (d->sb_n != 1 || d->K == 1)) b_mn = false;
  else if (d->sb_n == 1) b_mn = true;
  else { set_error("sml_gemm_tc: B needs unit stride along N or K (sb_k=%lld sb_n=%lld)", (long long)d->sb_k, (long long)d->sb_n); return 1; }
  if (d->sa_k == 1 && d->sa_m == 1) a_mn = (d->M > 1 && d->K == 1);
  if (d->sb_k == 1 && d->sb_n == 1) b_mn = (d->N > 1 && d->K == 1);

  const int softmax = (d->relu & 2) ? 1 : 0;
  int BN = d->N > 128 ? 256 : (d->N > 64 ? 128 : 64);
  // bf16x3: 128-wide tiles (measured on the 64 x 256^3 pinv steps: 256-wide tiles 34 us vs 30 us, and a 3-stage ring that
  // forfeits the staged epilogue 33 us)
  if (split && BN == 256 && !softmax) BN = 128;
  if (force_bn > 0) BN = force_bn;
  SML_CHECK(!softmax || d->N <= BN, "sml_gemm_tc: fused row softmax needs N <= %d (N=%d)", BN, d->N);

  CUtensorMap* tms = plan->tms;
  const int64_t lda = a_mn ? d->sa_k : d->sa_m, ldb = b_mn ? d->sb_k : d->sb_n;
  for (int part = 0; part < (split ? 2 : 1); ++part) {
    const void* A = part ? d->A_lo : d->A;
    const void* B = part ? d->B_lo : d->B;
    int rc;
    if (!a_mn) rc = make_map(&tms[part], A, d->K, d->M, lda, d->nb2, d->sa_b2, d->nb1, d->sa_b1, 64, TC_BM, "A");
    else rc = make_map(&tms[part], A, d->M, d->K, lda, d->nb2, d->sa_b2, d->nb1, d->sa_b1, 64, 64, "A");
    if (rc) return rc;
    if (!b_mn) rc = make_map(&tms[2 + part], B, d->K, d->N, ldb, d->nb2, d->sb_b2, d->nb1, d->sb_b1, 64, BN, "B");
    else rc = make_map(&tms[2 + part], B, d->N, d->K, ldb, d->nb2, d->sb_b2, d->nb1, d->sb_b1, 64, 64, "B");
    if (rc) return rc;
  }
  if (!split) { tms[1] = tms[0]; tms[3] = tms[2]; }

  EpiParams& ep = plan->ep;
  ep.M = d->M; ep.N = d->N; ep.K = d->K; ep.nb2 = d->nb2;
  ep.a_c1 = (d->nb1 > 1 && d->sa_b1 != 0) ? 1 : 0; ep.a_c2 = (d->nb2 > 1 && d->sa_b2 != 0) ? 1 : 0;
  ep.b_c1 = (d->nb1 > 1 && d->sb_b1 != 0) ? 1 : 0; ep.b_c2 = (d->nb2 > 1 && d->sb_b2 != 0) ? 1 : 0;
  ep.C = (float*)d->C; ep.sc_m = d->sc_m; ep.sc_n = d->sc_n; ep.sc_b1 = d->sc_b1; ep.sc_b2 = d->sc_b2;
  ep.Cb_hi = (__nv_bfloat16*)d->Cb_hi; ep.Cb_lo = (__nv_bfloat16*)d->Cb_lo;
  ep.alpha = d->alpha; ep.beta = d->beta; ep.diag = d->diag; ep.bias = d->bias;
  ep.relu = d->relu & 1; ep.softmax = softmax;
  ep.rowtab = d->rowtab; ep.rt_m = d->rt_m; ep.rt_b1 = d->rt_b1; ep.rt_b2 = d->rt_b2; ep.rowdiv = d->rowdiv > 0 ? d->rowdiv : 1;
  ep.rowtab_scale = d->rowtab_scale;
  SML_CHECK(!(ep.beta != 0.f) || ep.C, "sml_gemm_tc: beta needs an fp32 C");

  long long nbz = (long long)d->nb1 * d->nb2;
  const int tiles_mn = ((d->N + BN - 1) / BN) * ((d->M + TC_BM - 1) / TC_BM);
  const int nkb_all = (d->K + TC_BK - 1) / TC_BK;
  ep.ksplit = 1; ep.kb_per_split = nkb_all;
  // split-K when the output has too few tiles to fill the 148 SMs and K is long (weight gradients: K = B*n')
  if (allow_splitk && tiles_mn * nbz < 74 && nkb_all >= 32 && ep.C && !ep.Cb_hi && !ep.relu && !softmax && ep.beta == 0.f && !ep.rowtab) {
    int want = (int)((2 * kNumSMs + tiles_mn * nbz - 1) / (tiles_mn * nbz));
    if (want > nkb_all / 8) want = nkb_all / 8;
    if (want > 1) {
      ep.kb_per_split = (nkb_all + want - 1) / want;
      ep.ksplit = (nkb_all + ep.kb_per_split - 1) / ep.kb_per_split;
      for (long long bz = 0; bz < nbz; ++bz) {           // zero C (strided views: row by row)
        float* cb = ep.C + (bz / d->nb2) * ep.sc_b1 + (bz % d->nb2) * ep.sc_b2;
        SML_CHECK(ep.sc_n == 1, "sml_gemm_tc: split-K needs unit column stride");
        cudaError_t e = cudaMemset2DAsync(cb, (size_t)ep.sc_m * 4, 0, (size_t)d->N * 4, (size_t)d->M, stream);
        SML_CHECK(e == cudaSuccess, "sml_gemm_tc: split-K memset failed: %s", cudaGetErrorString(e));
      }
    }
  }
  ep.vec_ok = (ep.sc_n == 1 && (ep.sc_m % 4) == 0 && (ep.sc_b1 % 4) == 0 && (ep.sc_b2 % 4) == 0 &&
               (!ep.C || (reinterpret_cast<uintptr_t>(ep.C) & 15) == 0) &&
               (!ep.Cb_hi || ((reinterpret_cast<uintptr_t>(ep.Cb_hi) & 15) == 0 && ep.sc_m % 8 == 0 && ep.sc_b1 % 8 == 0 && ep.sc_b2 % 8 == 0)) &&
               (!ep.Cb_lo || (reinterpret_cast<uintptr_t>(ep.Cb_lo) & 15) == 0) &&
               (!ep.bias || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0) &&
               (!ep.rowtab || ((reinterpret_cast<uintptr_t>(ep.rowtab) & 15) == 0 && ep.rt_m % 4 == 0 && ep.rt_b1 % 4 == 0 && ep.rt_b2 % 4 == 0))) ? 1 : 0;
  ep.tiles_n = (d->N + BN - 1) / BN;
  ep.tiles_m = (d->M + TC_BM - 1) / TC_BM;
  const long long total = (long long)ep.tiles_n * ep.tiles_m * nbz * ep.ksplit;
  SML_CHECK(total < (1ll << 31), "sml_gemm_tc: too many tiles");
  ep.total_tiles = (int)total;
  plan->grid = dim3((unsigned)(total < kNumSMs ? total : kNumSMs));      // persistent: one CTA per SM walks the tile list
  plan->a_mn = a_mn; plan->b_mn = b_mn; plan->split = split; plan->BN = BN;
  return 0;
}

extern "C" int sml_gemm_tc(const sml_gemm_desc* d, void* stream) {
  cudaStream_t st = as_stream(stream);
  TcPlan plan;
  if (int rc = prepare_tc(d, 0, true, st, &plan)) return rc;
  const CUtensorMap* tms = plan.tms;
  const EpiParams& ep = plan.ep;
  const bool a_mn = plan.a_mn, b_mn = plan.b_mn;
  const dim3 grid = plan.grid;
  // one persistent CTA per SM: operand ring + (where it fits) the epilogue's coalescing tiles
  if (!plan.split) {
    if (plan.BN == 256) return dispatch_major<256, false, 3>(a_mn, b_mn, tms, ep, grid, st);
    if (plan.BN == 128) return dispatch_major<128, false, 5>(a_mn, b_mn, tms, ep, grid, st);
    return dispatch_major<64, false, 6>(a_mn, b_mn, tms, ep, grid, st);
  } else {
    if (plan.BN == 256) return dispatch_major<256, true, 2>(a_mn, b_mn, tms, ep, grid, st);   // 192 KB ring: direct stores
    if (plan.BN == 128) {
      const bool lean = ep.ksplit <= 1 && !ep.softmax && !ep.bias && !ep.rowtab && !ep.relu && ep.vec_ok && ep.N % 32 == 0;
      if (lean) return dispatch_major<128, true, 2, true>(a_mn, b_mn, tms, ep, grid, st);
      return dispatch_major<128, true, 2>(a_mn, b_mn, tms, ep, grid, st);
    }
    return dispatch_major<64, true, 3>(a_mn, b_mn, tms, ep, grid, st);
  }
}

// Run n dependent bf16x3 GEMMs (all N <= 128-wide tiles, e.g. the pinv chain) as one cooperative persistent launch per
// PROG_MAX_STEPS steps.  `counter` = 4-byte device scratch used by the in-kernel grid barrier.
extern "C" int sml_gemm_tc_program(const sml_gemm_desc* descs, int32_t n, void* counter, void* stream) {
  cudaStream_t st = as_stream(stream);
  SML_CHECK(descs && n > 0 && counter, "sml_gemm_tc_program: bad arguments");
  using S = TcSmem<128, true, 2>;
  auto kern = gemm_tc_program_kernel<128, true, 2>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    SML_CHECK(e == cudaSuccess, "sml_gemm_tc_program: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr = true;
  }
  static thread_local Program prog;
  for (int base = 0; base < n; base += PROG_MAX_STEPS) {
    const int cnt = (n - base < PROG_MAX_STEPS) ? n - base : PROG_MAX_STEPS;
    unsigned maxgrid = 1;
    for (int i = 0; i < cnt; ++i) {
      TcPlan plan;
      if (int rc = prepare_tc(&descs[base + i], 128, false, st, &plan)) return rc;
      SML_CHECK(plan.split && !plan.ep.softmax, "sml_gemm_tc_program: step %d is not a plain bf16x3 GEMM", base + i);
      ProgStep& ps = prog.st[i];
      for (int t = 0; t < 4; ++t) ps.tm[t] = plan.tms[t];
      ps.ep = plan.ep; ps.a_mn = plan.a_mn ? 1 : 0; ps.b_mn = plan.b_mn ? 1 : 0;
      if (plan.grid.x > maxgrid) maxgrid = plan.grid.x;
    }
    prog.nsteps = cnt;
    SML_CHECK(cudaMemsetAsync(counter, 0, 4, st) == cudaSuccess, "sml_gemm_tc_program: memset failed");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(maxgrid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;     // all CTAs co-resident: the grid barrier cannot deadlock
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, prog, (unsigned int*)counter);
    SML_CHECK(le == cudaSuccess, "sml_gemm_tc_program: launch failed: %s", cudaGetErrorString(le));
    SML_LAUNCH_OK("sml_gemm_tc_program");
  }
  return 0;
}

namespace sml {
__device__ __forceinline__ void split4(const float4 v, uint2& h, uint2& l) {
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
  const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
  h.x = *reinterpret_cast<const uint32_t*>(&h0); h.y = *reinterpret_cast<const uint32_t*>(&h1);
  l.x = *reinterpret_cast<const uint32_t*>(&l0); l.y = *reinterpret_cast<const uint32_t*>(&l1);
}
// 4 elements per thread (128-bit loads, 64-bit stores); n4 = number of float4 groups, tail handled by the scalar kernel
__global__ void __launch_bounds__(256) split_bf16_vec_kernel(const float4* __restrict__ x, uint2* __restrict__ hi,
                                                             uint2* __restrict__ lo, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    uint2 h, l;
    split4(x[i], h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}
// rows [B, n, D] fp32 -> bf16 (hi / lo) rows [B, npad, D] with `pad` zero rows in front of every batch (D % 4 == 0)
__global__ void __launch_bounds__(256) split_bf16_pad_kernel(const float4* __restrict__ x, uint2* __restrict__ hi,
                                                             uint2* __restrict__ lo, int64_t total4, int64_t n, int64_t npad,
                                                             int64_t pad, int D4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / D4, c = i % D4;
    const int64_t b = row / npad, r = row % npad;
    uint2 h = make_uint2(0u, 0u), l = make_uint2(0u, 0u);
    if (r >= pad) split4(x[(b * n + (r - pad)) * D4 + c], h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}
}  // namespace sml

extern "C" int sml_split_bf16(const float* x, void* hi, void* lo, int64_t n, void* stream) {
  if (n == 0) return 0;
  const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(hi) & 7) == 0) &&
                   (!lo || (reinterpret_cast<uintptr_t>(lo) & 7) == 0);
  int64_t g = ((vec ? n / 4 : n) + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  if (vec) split_bf16_vec_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>((const float4*)x, (uint2*)hi, (uint2*)lo, n / 4);
  else split_bf16_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
  SML_LAUNCH_OK("sml_split_bf16");
  return 0;
}
extern "C" int sml_split_bf16_pad(const float* x, void* hi, void* lo, int64_t B, int64_t n, int64_t npad, int64_t pad, int32_t D,
                                  void* stream) {
  SML_CHECK(D % 4 == 0 && npad == n + pad && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "sml_split_bf16_pad: bad layout");
  const int64_t total4 = B * npad * (D / 4);
  if (total4 == 0) return 0;
  int64_t g = (total4 + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  split_bf16_pad_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>((const float4*)x, (uint2*)hi, (uint2*)lo, total4, n, npad, pad, D / 4);
  SML_LAUNCH_OK("sml_split_bf16_pad");
  return 0;
}

// ubench_tmem.cu — micro-benchmarks behind the deformable-attention design choices:
//   1. tcgen05.ld 32x32b.x32 throughput per SM with 4 / 8 / 16 warps        (bytes per clock)
//   2. tcgen05.st 32x32b.x16 throughput
//   3. issue rate of small tcgen05.mma (M=128, N=32, K=16), A from TMEM and A from shared memory
//   4. latency of one compute-warp -> MMA thread -> compute-warp round trip through mbarriers / tcgen05.commit
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tmem.bin tools/ubench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 22); ++it) if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t c, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(c), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(c), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
#define LD32(taddr, r)                                                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                 \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                 \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                 \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),        \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),  \
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
               : "r"(taddr) : "memory")
#define ST16(taddr, r)                                                                                                   \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"  \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),    \
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")

template <int N, bool TS>
__device__ __forceinline__ long long mma_loop(uint32_t tb, uint8_t* smem, int iters, uint64_t* bar) {
  constexpr uint32_t idn = make_idesc(128, N);
  const uint64_t db = make_smem_desc(smem_u32(smem), 16, 1024), da = make_smem_desc(smem_u32(smem + 16384), 16, 1024);
  const uint32_t d0 = tb, d1 = tb + 256, a0 = tb + 192, a1 = tb + 200;
  long long t0 = clock64();
  for (int it = 0; it < iters; it += 2) {
    if (TS) { mma_ts(d0, a0, db, idn, 0u); mma_ts(d1, a1, db, idn, 0u); }
    else { mma_ss(d0, da, db, idn, 0u); mma_ss(d1, da, db, idn, 0u); }
  }
  tc_commit(bar);
  mbar_wait(bar, 0);
  return clock64() - t0;
}

// mode 0: LDTM x32 throughput; mode 1: STTM x16; mode 2: MMA issue (A from TMEM); mode 3: MMA issue (A from smem);
// mode 4: round trip
__global__ void __launch_bounds__(512, 1) ubench(int mode, int nwarps, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bars[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], (i == 0 && mode == 4) ? 4u : 1u); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = tid; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  const uint32_t trow = tb + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = tid * 3 + i;
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (mode == 0 && warp < nwarps) {
    t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
      LD32(trow + ((it & 7) * 32 + (warp >> 2) * 256) % 512, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= r[i];
    }
    t1 = clock64();
    if (acc == 0x12345) out[100] = acc;
  } else if (mode == 1 && warp < nwarps) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      ST16(trow + ((it & 15) * 16 + (warp >> 2) * 256) % 512, r);
      if ((it & 3) == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    t1 = clock64();
  } else if ((mode == 13 || mode == 14) && warp == 0) {
    // the GEMM main loop's access pattern: descriptors change every iteration (stage / k offset), runtime N.
    // mode 13: one thread inside `if (lane == 0)` (gemm_tc.cu today); mode 14: whole warp, warp-uniform operands
    // (TMEM base through shfl), elected lane issues
    const uint32_t idn = make_idesc(128, nwarps);
    const uint32_t sb = smem_u32(smem);
    if (mode == 13) {
      if (lane == 0) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
          const uint32_t st = (it % 3) * 8192u, kk = (it & 3) * 32u;
          mma_ss(tb + (it & 1) * 256, make_smem_desc(sb + 16384 + st + kk, 16, 1024), make_smem_desc(sb + st + kk, 16, 1024), idn, (it & 15) ? 1u : 0u);
        }
        tc_commit(&bars[1]);
        mbar_wait(&bars[1], 0);
        t1 = clock64();
      }
    } else {
      const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t st = (it % 3) * 8192u, kk = (it & 3) * 32u;
        const uint64_t da = make_smem_desc(sb + 16384 + st + kk, 16, 1024), db = make_smem_desc(sb + st + kk, 16, 1024);
        if (elect_one()) mma_ss(tbu + (it & 1) * 256, da, db, idn, (it & 15) ? 1u : 0u);
      }
      if (elect_one()) tc_commit(&bars[1]);
      __syncwarp();
      mbar_wait(&bars[1], 0);
      t1 = clock64();
    }
  } else if ((mode == 11 || mode == 12) && tid == 0) {
    long long dt = 0;
    if (mode == 11) {
      if (nwarps == 16) dt = mma_loop<16, false>(tb, smem, iters, &bars[1]);
      if (nwarps == 32) dt = mma_loop<32, false>(tb, smem, iters, &bars[1]);
      if (nwarps == 64) dt = mma_loop<64, false>(tb, smem, iters, &bars[1]);
      if (nwarps == 128) dt = mma_loop<128, false>(tb, smem, iters, &bars[1]);
      if (nwarps == 256) dt = mma_loop<256, false>(tb, smem, iters, &bars[1]);
    } else {
      if (nwarps == 16) dt = mma_loop<16, true>(tb, smem, iters, &bars[1]);
      if (nwarps == 32) dt = mma_loop<32, true>(tb, smem, iters, &bars[1]);
      if (nwarps == 64) dt = mma_loop<64, true>(tb, smem, iters, &bars[1]);
      if (nwarps == 128) dt = mma_loop<128, true>(tb, smem, iters, &bars[1]);
      if (nwarps == 256) dt = mma_loop<256, true>(tb, smem, iters, &bars[1]);
    }
    t0 = 0; t1 = dt;
  } else if ((mode == 9 || mode == 10) && warp == 0) {
    // whole warp runs the loop with warp-uniform operands; one elected lane issues (the CUTLASS pattern)
    const uint32_t idn = make_idesc(128, nwarps);
    const uint64_t db = make_smem_desc(smem_u32(smem), 16, 1024), da = make_smem_desc(smem_u32(smem + 16384), 16, 1024);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        if (mode == 9) mma_ss(tb + (it & 1) * 256, da, db, idn, 0u);
        else mma_ts(tb + (it & 1) * 256, tb + 192 + (it & 3) * 8, db, idn, 0u);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bars[1]);
    __syncwarp();
    mbar_wait(&bars[1], 0);
    t1 = clock64();
  } else if (mode == 8 && lane == 0 && warp < nwarps) {
    // nwarps issuing threads (one per warp), each N=32 K=16 M=128 into its own D columns, own commit barrier
    constexpr uint32_t id32 = make_idesc(128, 32);
    const uint64_t db = make_smem_desc(smem_u32(smem), 16, 1024), da = make_smem_desc(smem_u32(smem + 16384), 16, 1024);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) mma_ss(tb + warp * 64 + (it & 1) * 32, da, db, id32, 0u);
    tc_commit(&bars[warp]);
    mbar_wait(&bars[warp], 0);
    t1 = clock64();
  } else if (mode >= 5 && mode <= 7 && tid == 0) {
    // mode 5: M=128, N = nwarps (A from smem); mode 6: M=64; mode 7: M=128, N=nwarps, A from TMEM, accumulate into ONE D
    const uint32_t idn = make_idesc(mode == 6 ? 64 : 128, nwarps);
    const uint64_t db = make_smem_desc(smem_u32(smem), 16, 1024), da = make_smem_desc(smem_u32(smem + 16384), 16, 1024);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 7) mma_ts(tb, tb + 256 + (it & 7) * 8, db, idn, 1u);
      else mma_ss(tb + ((it & 1) * 256 % 512), da, db, idn, 0u);
    }
    tc_commit(&bars[1]);
    mbar_wait(&bars[1], 0);
    t1 = clock64();
  } else if ((mode == 2 || mode == 3) && tid == 0) {
    constexpr uint32_t id32 = make_idesc(128, 32);
    const uint64_t db = make_smem_desc(smem_u32(smem), 16, 1024), da = make_smem_desc(smem_u32(smem + 16384), 16, 1024);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tb + (it & 7) * 32;
      if (mode == 2) mma_ts(d, tb + 256 + (it & 7) * 8, db, id32, 0u); else mma_ss(d, da, db, id32, 0u);
    }
    tc_commit(&bars[1]);
    mbar_wait(&bars[1], 0);
    t1 = clock64();
  } else if (mode == 4) {
    // warps 1..4 = compute (one warpgroup), thread 0 = MMA issuer
    constexpr uint32_t id32 = make_idesc(128, 32);
    const uint64_t db = make_smem_desc(smem_u32(smem), 16, 1024);
    if (tid == 0) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&bars[0], it & 1);
        tc_fence_after();
        mma_ts(tb, tb + 256, db, id32, 0u);
        tc_commit(&bars[2]);
      }
      t1 = clock64();
    } else if (warp >= 1 && warp <= 4) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        ST16(trow + 256, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[0]);
        mbar_wait(&bars[2], it & 1);
        tc_fence_after();
        LD32(trow, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      t1 = clock64();
      if (r[0] == 0x12345) out[100] = r[1];
    }
  }
  if (lane == 0 && t1 > t0) out[warp] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
  }
}

int main() {
  long long* d; cudaMalloc(&d, 128 * sizeof(long long));
  long long h[128];
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  auto run = [&](int mode, int nwarps, int iters, const char* what, double bytes_per_iter_per_warp) {
    cudaMemset(d, 0, 128 * sizeof(long long));
    ubench<<<1, 512, 100000>>>(mode, nwarps, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, d, 128 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 16; ++i) if (h[i] > mx) mx = h[i];
    printf("%-44s warps=%2d iters=%d  cycles=%lld  %.1f cyc/iter", what, nwarps, iters, mx, (double)mx / iters);
    if (bytes_per_iter_per_warp > 0) printf("  %.1f B/clk/SM", bytes_per_iter_per_warp * nwarps * iters / (double)mx);
    printf("\n");
  };
  for (int nw : {1, 4, 8, 16}) run(0, nw, 2000, "tcgen05.ld 32x32b.x32 (+wait each)", 4096.0);
  for (int nw : {1, 4, 8, 16}) run(1, nw, 2000, "tcgen05.st 32x32b.x16 (wait every 4)", 2048.0);
  run(2, 1, 2000, "tcgen05.mma 128x32x16 A from TMEM, 1 thread", 0);
  run(3, 1, 2000, "tcgen05.mma 128x32x16 A from smem, 1 thread", 0);
  for (int n : {16, 32, 64, 128, 256}) run(5, n, 2000, "tcgen05.mma M=128 N=<warps> K=16 SS", 0);
  for (int n : {32, 128, 256}) run(6, n, 2000, "tcgen05.mma M=64 N=<warps> K=16 SS", 0);
  for (int n : {32, 128, 256}) run(7, n, 2000, "tcgen05.mma M=128 N=<warps> K=16 TS acc same D", 0);
  for (int n : {1, 2, 4}) run(8, n, 2000, "tcgen05.mma 128x32x16 SS, <warps> issuing threads", 0);
  for (int n : {16, 32, 64, 128, 256}) run(9, n, 2000, "tcgen05.mma M=128 N=<warps> K=16 SS elect", 0);
  for (int n : {32, 128, 256}) run(10, n, 2000, "tcgen05.mma M=128 N=<warps> K=16 TS elect", 0);
  for (int n : {16, 32, 64, 128, 256}) run(11, n, 2000, "hoisted operands: mma M=128 N=<warps> K=16 SS", 0);
  for (int n : {16, 32, 64, 128, 256}) run(12, n, 2000, "hoisted operands: mma M=128 N=<warps> K=16 TS", 0);
  for (int n : {64, 128, 256}) run(13, n, 2000, "GEMM-loop pattern, if (lane == 0): mma M=128 N=<warps>", 0);
  for (int n : {64, 128, 256}) run(14, n, 2000, "GEMM-loop pattern, uniform + elect: mma M=128 N=<warps>", 0);
  run(4, 4, 1000, "round trip st16->arrive->mma->commit->ld32", 0);
  return 0;
}

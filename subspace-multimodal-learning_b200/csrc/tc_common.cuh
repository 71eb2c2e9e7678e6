// tc_common.cuh — shared tcgen05 / TMEM / TMA / mbarrier PTX wrappers, shared-memory matrix descriptors and the host-side
// tensor-map builder used by gemm_tc.cu and nystrom_fused.cu (sm_100a only).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace sml {

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
static __device__ unsigned int g_tc_timeout_flag = 0;   // one per translation unit (no -rdc); see sml_tc_timeout_flag()
// bounded wait: a protocol bug must never hang the GPU box, and must never be silent either.  On timeout the flag is
// raised (so that the host can tell a pipeline timeout from any other fault) and the kernel TRAPS: the launch fails, the
// CUDA context reports an error at the next synchronisation and no garbage result is ever consumed.  The flag is only a
// diagnostic; waits never short-circuit on it.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    if (mbar_try_wait(bar, parity)) return true;
  }
  atomicExch(&g_tc_timeout_flag, 1u);
  __threadfence_system();
  __trap();
  return false;
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (sm100 "version 1"), SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;   // descriptor version (Blackwell)
  d |= 2ull << 61;   // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: D=f32, A=B=bf16
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// ------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// operand viewed as [rows, inner] bf16 with `inner` contiguous, row stride ld, plus two batch dims
static inline int make_map(CUtensorMap* tm, const void* base, int64_t inner, int64_t rows, int64_t ld, int nb2, int64_t sb2,
                    int nb1, int64_t sb1, int box_inner, int box_rows, const char* what) {
  EncodeTiledFn enc = get_encode();
  SML_CHECK(enc, "sml_gemm_tc: cuTensorMapEncodeTiled not available (no CUDA driver?)");
  SML_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "sml_gemm_tc: %s base not 16-byte aligned", what);
  SML_CHECK((ld * 2) % 16 == 0, "sml_gemm_tc: %s row stride %lld not a multiple of 8 elements", what, (long long)ld);
  int64_t s2 = sb2, s1 = sb1;
  cuuint64_t d2 = (cuuint64_t)nb2, d1 = (cuuint64_t)nb1;
  if (nb2 <= 1 || s2 == 0) { d2 = 1; s2 = rows * ld; }     // broadcast / absent batch dim: size 1, any legal stride
  if (nb1 <= 1 || s1 == 0) { d1 = 1; s1 = rows * ld; }
  if (s2 <= 0) s2 = 8;
  if (s1 <= 0) s1 = 8;
  SML_CHECK((s2 * 2) % 16 == 0 && (s1 * 2) % 16 == 0, "sml_gemm_tc: %s batch strides must be multiples of 8 elements", what);
  cuuint64_t gdim[4] = {(cuuint64_t)inner, (cuuint64_t)rows, d2, d1};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)s2 * 2, (cuuint64_t)s1 * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SML_CHECK(r == CUDA_SUCCESS, "sml_gemm_tc: cuTensorMapEncodeTiled(%s) failed with %d (inner=%lld rows=%lld ld=%lld)", what,
            (int)r, (long long)inner, (long long)rows, (long long)ld);
  return 0;
}


}  // namespace sml

// common.cuh — shared helpers for libsml_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/sml_b200.h"

namespace sml {

void set_error(const char* fmt, ...);
const unsigned long long* seed_salt();      // device address of the current dropout salt, or NULL (api.cu)
__device__ __forceinline__ unsigned long long salted(unsigned long long seed, const unsigned long long* salt) {
  return salt ? seed ^ __ldg(salt) : seed;
}
extern std::atomic<long long> g_launches;

#define SML_CHECK(cond, ...)                                   \
  do {                                                         \
    if (!(cond)) { sml::set_error(__VA_ARGS__); return 1; }    \
  } while (0)

#define SML_LAUNCH_OK(name)                                                        \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      sml::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
      return 2;                                                                    \
    }                                                                              \
    sml::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` must hold >= 32 floats; all threads get the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

constexpr int kNumSMs = 148;

// optional bf16 copies of a value written by a kernel: hi = bf16(v), lo = bf16(v - hi) (operands of the tcgen05 engine)
__device__ __forceinline__ void store_split(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t i, float v) {
  if (hi) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

}  // namespace sml

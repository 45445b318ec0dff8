// Shared device / host helpers for libotto_covisit.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/otto_covisit.h"

#define FULL_MASK 0xffffffffu
#define AID_MASK 0x3fffffffu      // aid word = aid | type << 30
#define KEY_EMPTY 0xffffffffu
#define KEY_TAKEN 0x80000000u     // bit 31 set = not a candidate any more (aids are < 2^30)

void otto_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      otto_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
      return OTTO_ECUDA;                                                                       \
    }                                                                                          \
  } while (0)

extern unsigned long long g_otto_launches;   // kernels launched by this library (reported by bench.py)
#define LAUNCH_CHECK()              \
  do {                              \
    ++g_otto_launches;              \
    CUDA_TRY(cudaGetLastError());   \
  } while (0)

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
// second, independent mix (multi-pass selector in the reduce kernel)
__device__ __forceinline__ uint32_t hash32b(uint32_t x) {
  x *= 0x9e3779b1U;
  x ^= x >> 15;
  x *= 0x85ebca6bU;
  x ^= x >> 13;
  return x;
}
// sub-bin of aid_y inside a split aid_x row: uniform in [0, nb)
__device__ __forceinline__ uint32_t sub_bin(uint32_t y, uint32_t nb) { return __umulhi(hash32(y) , nb); }

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// streaming 8-byte / 16-byte accesses that do not pollute L1
__device__ __forceinline__ uint2 ld_stream_u2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u2(uint2* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(FULL_MASK, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(FULL_MASK, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint32_t lo = __shfl_xor_sync(FULL_MASK, (uint32_t)v, o);
    uint32_t hi = __shfl_xor_sync(FULL_MASK, (uint32_t)(v >> 32), o);
    uint64_t w = ((uint64_t)hi << 32) | lo;
    v = w > v ? w : v;
  }
  return v;
}

// Exclusive prefix sum over device arrays: out[i] = sum(in[0..i)), out[n] = total.
// Three launches (tile sums, scan of tile sums, apply); the arrays here are 2M..13M elements, so the
// scan is a few microseconds next to the multi-GB pair traffic and not worth a decoupled look-back.
#pragma once
#include "common.cuh"

namespace scan_detail {
constexpr int THREADS = 256;
constexpr int ITEMS = 8;
constexpr int TILE = THREADS * ITEMS;

template <typename TOut>
__device__ __forceinline__ TOut block_exclusive(TOut thread_total, TOut* smem /*[8+1]*/, TOut* block_total) {
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  TOut inc = thread_total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    TOut v = __shfl_up_sync(FULL_MASK, inc, o);
    if (lane >= (uint32_t)o) inc += v;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    TOut w = lane < THREADS / 32 ? smem[lane] : TOut(0);
    TOut winc = w;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      TOut v = __shfl_up_sync(FULL_MASK, winc, o);
      if (lane >= (uint32_t)o) winc += v;
    }
    if (lane < THREADS / 32) smem[lane] = winc - w;
    if (lane == THREADS / 32 - 1) smem[THREADS / 32] = winc;
  }
  __syncthreads();
  *block_total = smem[THREADS / 32];
  return smem[warp] + inc - thread_total;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(THREADS) tile_sums(const TIn* __restrict__ in, int64_t n, TOut* __restrict__ sums) {
  __shared__ TOut smem[THREADS / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
  TOut t = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
    if (base + i < n) t += (TOut)in[base + i];
  TOut total;
  block_exclusive<TOut>(t, smem, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename TOut>
__global__ void __launch_bounds__(THREADS) scan_sums(TOut* __restrict__ sums, int64_t n_tiles) {
  __shared__ TOut smem[THREADS / 32 + 1];
  TOut carry = 0;
  for (int64_t start = 0; start < n_tiles; start += TILE) {
    const int64_t base = start + (int64_t)threadIdx.x * ITEMS;
    TOut v[ITEMS];
    TOut t = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      v[i] = base + i < n_tiles ? sums[base + i] : TOut(0);
      t += v[i];
    }
    TOut total;
    TOut ex = block_exclusive<TOut>(t, smem, &total) + carry;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (base + i < n_tiles) sums[base + i] = ex;
      ex += v[i];
    }
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[n_tiles] = carry;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(THREADS)
    apply(const TIn* in, int64_t n, const TOut* __restrict__ sums, int64_t n_tiles, TOut* out) {
  __shared__ TOut smem[THREADS / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
  TOut v[ITEMS];
  TOut t = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    v[i] = base + i < n ? (TOut)in[base + i] : TOut(0);
    t += v[i];
  }
  TOut total;
  TOut ex = block_exclusive<TOut>(t, smem, &total) + sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[n_tiles];
}
}  // namespace scan_detail

// scratch must hold (ceil(n / 2048) + 1) TOut values. `in` and `out` may alias when the types match.
static inline int64_t scan_scratch_elems(int64_t n) { return ceil_div(n > 0 ? n : 1, scan_detail::TILE) + 1; }

template <typename TIn, typename TOut>
static inline int exclusive_scan(const TIn* in, int64_t n, TOut* out, TOut* scratch, cudaStream_t st) {
  using namespace scan_detail;
  const int64_t n_tiles = ceil_div(n > 0 ? n : 1, TILE);
  tile_sums<TIn, TOut><<<(unsigned)n_tiles, THREADS, 0, st>>>(in, n, scratch);
  scan_sums<TOut><<<1, THREADS, 0, st>>>(scratch, n_tiles);
  apply<TIn, TOut><<<(unsigned)n_tiles, THREADS, 0, st>>>(in, n, scratch, n_tiles, out);
  g_otto_launches += 2;  // three kernels, one check
  LAUNCH_CHECK();
  return OTTO_OK;
}

// Microbenchmark: shared-memory atomics vs plain RMW vs match_any throughput on one SM-full grid.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu && ./smem_atomics
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, uint32_t slots_mask) {
  extern __shared__ uint32_t tab[];
  for (uint32_t i = threadIdx.x; i <= slots_mask; i += blockDim.x) tab[i] = MODE == 1 ? 0xffffffffu : 0u;
  __syncthreads();
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
  for (int it = 0; it < iters; ++it) {
    const uint32_t h = rng(s) & slots_mask;
    if (MODE == 0) acc += atomicAdd(&tab[h], 1u);                       // ATOMS.ADD with return
    if (MODE == 1) acc += atomicCAS(&tab[h], 0xffffffffu, h);           // ATOMS.CAS
    if (MODE == 2) { uint32_t v = tab[h]; tab[h] = v + 1; acc += v; }   // plain LDS + STS
    if (MODE == 3) acc += __match_any_sync(0xffffffffu, h & 15u);       // MATCH.ANY (16 groups)
    if (MODE == 4) acc += __match_any_sync(0xffffffffu, h);             // MATCH.ANY (mostly unique)
    if (MODE == 5) atomicAdd(&tab[h], 1u);                              // ATOMS.ADD no return (RED)
    if (MODE == 6) { uint4 v = ((uint4*)tab)[h >> 2]; v.y += 1; ((uint4*)tab)[h >> 2] = v; acc += v.x; }  // LDS.128 + STS.128
    if (MODE == 7) acc += __ballot_sync(0xffffffffu, h & 1) + __shfl_sync(0xffffffffu, h, h & 31);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int blocks_per_sm, uint32_t slots) {
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int iters = 4096, blocks = n_sm * blocks_per_sm;
  uint32_t* out;
  cudaMalloc(&out, (size_t)blocks * 256 * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<blocks, 256, slots * 4>>>(out, 16, slots - 1);
  cudaEventRecord(a);
  k<MODE><<<blocks, 256, slots * 4>>>(out, iters, slots - 1);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double lane_ops = (double)blocks * 256 * iters;
  // cycles per warp-op per SM at 1.965 GHz
  const double cyc = ms * 1e-3 * 1.965e9 / (lane_ops / 32 / n_sm);
  printf("%-28s blocks/SM %d slots %5u: %8.3f ms  %7.1f Glane-op/s  %6.2f cyc per warp-op per SM  err=%s\n", name, blocks_per_sm, slots, ms,
         lane_ops / ms / 1e6, cyc, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int bps : {2, 4, 8}) {
    run<0>("ATOMS.ADD ret", bps, 4096);
    run<5>("ATOMS.ADD noret", bps, 4096);
    run<1>("ATOMS.CAS", bps, 4096);
    run<2>("LDS+STS u32", bps, 4096);
    run<6>("LDS.128+STS.128", bps, 4096);
    run<3>("MATCH.ANY 16 groups", bps, 4096);
    run<4>("MATCH.ANY unique", bps, 4096);
    run<7>("BALLOT+SHFL", bps, 4096);
  }
  return 0;
}

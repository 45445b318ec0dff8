// NVLink peer-memory throughput vs access granularity (2 GPUs of one box, cudaDeviceEnablePeerAccess).
// Each warp handles runs of `run` bytes at pseudo-random positions (8-byte aligned when aligned == 0, run-aligned
// otherwise) of a 2 GiB buffer that lives on the local or on the peer GPU, and either stores or loads them with one
// 8-byte element per lane per step - the access shape of the pair-record scatter (stores) and of the reduce (loads).
// The total is 4 GiB per configuration so that short-run configurations run for milliseconds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_rw peer_rw.cu && ./peer_rw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <bool STORE>
__global__ void runs_kernel(uint2* buf, uint64_t buf_elems, uint32_t run_elems, uint64_t n_runs, int aligned, uint64_t* sink) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  uint64_t acc = 0;
  if (run_elems >= 32) {
    for (uint64_t r = warp; r < n_runs; r += n_warps) {
      // multiply-high instead of a 64-bit modulo: the first version of this benchmark spent ~200 instructions per warp
      // iteration on address arithmetic and was issue bound at 4.2 M iterations (0.7 ms) for every run length <= 256 B
      uint64_t at = __umulhi(mix((uint32_t)r), (uint32_t)(buf_elems - run_elems));
      if (aligned) at &= ~(uint64_t)(run_elems - 1);
      for (uint32_t i = lane; i < run_elems; i += 32) {
        if (STORE) buf[at + i] = make_uint2((uint32_t)r, i);
        else { const uint2 v = buf[at + i]; acc += v.x + v.y; }
      }
    }
  } else {
    // several short runs per warp instruction: lane group g = lane / run_elems handles run (r * groups + g)
    const uint32_t groups = 32 / run_elems, g = lane / run_elems, i = lane % run_elems;
    for (uint64_t r = warp; r * groups < n_runs; r += n_warps) {
      const uint64_t id = r * groups + g;
      uint64_t at = __umulhi(mix((uint32_t)id), (uint32_t)(buf_elems - run_elems));
      if (aligned) at &= ~(uint64_t)(run_elems - 1);
      if (STORE) buf[at + i] = make_uint2((uint32_t)id, i);
      else { const uint2 v = buf[at + i]; acc += v.x + v.y; }
    }
  }
  if (!STORE && acc == 0x123456789ull) *sink = acc;
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  const bool have_peer = n >= 2;      // one GPU: local rows only (the scatter ceiling question of DESIGN.md §8)
  int can = 0;
  if (have_peer) CK(cudaDeviceCanAccessPeer(&can, 0, 1));
  printf("peer access 0 -> 1: %d\n", can);
  const uint64_t buf_bytes = 2ull << 30, buf_elems = buf_bytes / 8;
  uint2 *local = nullptr, *peer = nullptr;
  uint64_t* sink = nullptr;
  if (have_peer) {
    CK(cudaSetDevice(1));
    CK(cudaMalloc(&peer, buf_bytes));
    CK(cudaMemset(peer, 1, buf_bytes));
    CK(cudaDeviceSynchronize());
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
  }
  CK(cudaMalloc(&local, buf_bytes));
  CK(cudaMemset(local, 1, buf_bytes));
  CK(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const uint64_t total_bytes = 4ull << 30;
  const int runs[] = {8, 32, 64, 128, 256, 1024, 8192};
  printf("%-6s %-6s %-9s %8s %10s\n", "where", "op", "align", "run B", "GB/s");
  for (int where = 0; where < (have_peer ? 2 : 1); ++where)
    for (int store = 0; store < 2; ++store)
      for (int aligned = 0; aligned < 2; ++aligned)
        for (int run : runs) {
          uint2* buf = where ? peer : local;
          const uint32_t run_elems = run / 8;
          const uint64_t n_runs = total_bytes / run;
          float best = 1e30f;
          for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            if (store) runs_kernel<true><<<148 * 8, 256>>>(buf, buf_elems, run_elems, n_runs, aligned, sink);
            else runs_kernel<false><<<148 * 8, 256>>>(buf, buf_elems, run_elems, n_runs, aligned, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
          }
          printf("%-6s %-6s %-9s %8d %10.1f\n", where ? "peer" : "local", store ? "store" : "load", aligned ? "run" : "8 B", run,
                 total_bytes / best / 1e6);
          fflush(stdout);
        }
  return 0;
}

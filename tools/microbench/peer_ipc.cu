// Same access shapes as peer_rw.cu, but the peer buffer is reached the way the product reaches it: allocated with
// cudaMalloc by ANOTHER process on GPU 1 and mapped here through cudaIpcOpenMemHandle.  "mixed": odd runs go to the
// peer buffer, even runs to a local one, inside the same warp instruction (the owner-direct scatter's pattern).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_ipc peer_ipc.cu && ./peer_ipc
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <sys/wait.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// mode 0: all runs to a; 1: runs alternate between a and b
template <bool STORE>
__global__ void runs_kernel(uint2* a, uint2* b, int mode, uint64_t buf_elems, uint32_t run_elems, uint64_t n_runs, uint64_t* sink) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  uint64_t acc = 0;
  const uint32_t groups = run_elems >= 32 ? 1 : 32 / run_elems, g = run_elems >= 32 ? 0 : lane / run_elems;
  for (uint64_t r = warp; r * groups < n_runs; r += n_warps) {
    const uint64_t id = r * groups + g;
    uint64_t at = __umulhi(mix((uint32_t)id), (uint32_t)(buf_elems - run_elems));
    uint2* buf = (mode == 1 && (id & 1)) ? b : a;
    for (uint32_t i = run_elems >= 32 ? lane : lane % run_elems; i < run_elems; i += 32) {
      if (STORE) buf[at + i] = make_uint2((uint32_t)id, i);
      else { const uint2 v = buf[at + i]; acc += v.x + v.y; }
    }
  }
  if (!STORE && acc == 0x123456789ull) *sink = acc;
}

int main() {
  int to_parent[2], to_child[2];
  if (pipe(to_parent) || pipe(to_child)) return 1;
  const uint64_t buf_bytes = 4ull << 30, buf_elems = buf_bytes / 8;
  pid_t pid = fork();
  if (pid == 0) {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { cudaIpcMemHandle_t h = {}; (void)!write(to_parent[1], &h, sizeof(h)); return 0; }
    CK(cudaSetDevice(1));
    void* p = nullptr;
    CK(cudaMalloc(&p, buf_bytes));
    CK(cudaMemset(p, 1, buf_bytes));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, p));
    (void)!write(to_parent[1], &h, sizeof(h));
    char c;
    (void)!read(to_child[0], &c, 1);
    return 0;
  }
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  cudaIpcMemHandle_t h;
  if (read(to_parent[0], &h, sizeof(h)) != (ssize_t)sizeof(h)) return 1;
  CK(cudaSetDevice(0));
  uint2 *peer = nullptr, *local = nullptr;
  uint64_t* sink = nullptr;
  CK(cudaIpcOpenMemHandle((void**)&peer, h, cudaIpcMemLazyEnablePeerAccess));
  CK(cudaMalloc(&local, buf_bytes));
  CK(cudaMemset(local, 1, buf_bytes));
  CK(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const uint64_t total_bytes = 1ull << 30;
  const int runs[] = {8, 64, 256, 8192};
  const char* names[] = {"ipc-peer", "local", "mixed"};
  printf("%-9s %-6s %8s %10s\n", "where", "op", "run B", "GB/s");
  for (int where = 0; where < 3; ++where)
    for (int store = 0; store < 2; ++store)
      for (int run : runs) {
        uint2* a = where == 1 ? local : peer;
        const uint32_t run_elems = run / 8;
        const uint64_t n_runs = total_bytes / run;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaEventRecord(e0));
          if (store) runs_kernel<true><<<148 * 8, 256>>>(a, local, where == 2, buf_elems, run_elems, n_runs, sink);
          else runs_kernel<false><<<148 * 8, 256>>>(a, local, where == 2, buf_elems, run_elems, n_runs, sink);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms = 0;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (rep && ms < best) best = ms;
        }
        printf("%-9s %-6s %8d %10.1f\n", names[where], store ? "store" : "load", run, total_bytes / best / 1e6);
        fflush(stdout);
      }
  CK(cudaIpcCloseMemHandle(peer));
  (void)!write(to_child[1], "x", 1);
  waitpid(pid, nullptr, 0);
  return 0;
}

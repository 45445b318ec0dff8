// Two processes, one per GPU, each storing short runs into the OTHER's buffer at the same time (IPC mappings), the
// traffic pattern of an owner-direct scatter at N = 2.  Prints per process GB/s for 64-byte and 8 KiB runs, peer only
// and mixed (odd runs remote, even runs local).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_bidir peer_bidir.cu && ./peer_bidir
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

__global__ void store_runs(uint2* a, uint2* b, int mixed, uint64_t buf_elems, uint32_t run_elems, uint64_t n_runs) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t groups = run_elems >= 32 ? 1 : 32 / run_elems, g = run_elems >= 32 ? 0 : lane / run_elems;
  for (uint64_t r = warp; r * groups < n_runs; r += n_warps) {
    const uint64_t id = r * groups + g;
    const uint64_t at = __umulhi(mix((uint32_t)id), (uint32_t)(buf_elems - run_elems));
    uint2* buf = (mixed && (id & 1)) ? b : a;
    for (uint32_t i = run_elems >= 32 ? lane : lane % run_elems; i < run_elems; i += 32) buf[at + i] = make_uint2((uint32_t)id, i);
  }
}

static void barrier(int wfd, int rfd) {
  char c = 'x';
  (void)!write(wfd, &c, 1);
  (void)!read(rfd, &c, 1);
}

static int run(int dev, int wfd, int rfd, bool solo_first) {
  const uint64_t buf_bytes = 4ull << 30, buf_elems = buf_bytes / 8;
  CK(cudaSetDevice(dev));
  uint2 *local = nullptr, *peer = nullptr;
  CK(cudaMalloc(&local, buf_bytes));
  CK(cudaMemset(local, 1, buf_bytes));
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t mine, theirs;
  CK(cudaIpcGetMemHandle(&mine, local));
  (void)!write(wfd, &mine, sizeof(mine));
  if (read(rfd, &theirs, sizeof(theirs)) != (ssize_t)sizeof(theirs)) return 1;
  CK(cudaIpcOpenMemHandle((void**)&peer, theirs, cudaIpcMemLazyEnablePeerAccess));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const uint64_t total_bytes = 4ull << 30;
  // phase 0: only GPU 0 stores (GPU 1 idles); phase 1: both store at the same time
  for (int phase = 0; phase < 2; ++phase)
    for (int mixed = 0; mixed < 2; ++mixed)
      for (int run_b : {64, 8192}) {
        barrier(wfd, rfd);
        const bool active = phase == 1 || dev == 0;
        float best = 1e30f;
        if (active)
          for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            store_runs<<<148 * 8, 256>>>(peer, local, mixed, buf_elems, run_b / 8, total_bytes / run_b);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
          }
        barrier(wfd, rfd);
        if (active) {
          printf("gpu %d  %-13s %-6s run %5d B  %8.1f GB/s\n", dev, phase ? "both GPUs" : "GPU 0 alone", mixed ? "mixed" : "peer", run_b,
                 total_bytes / best / 1e6);
          fflush(stdout);
        }
      }
  barrier(wfd, rfd);
  CK(cudaIpcCloseMemHandle(peer));
  barrier(wfd, rfd);
  return 0;
}

int main() {
  int a[2], b[2];
  if (pipe(a) || pipe(b)) return 1;
  pid_t pid = fork();
  if (pid == 0) return run(1, b[1], a[0], false);
  int n = 0;
  const int rc = run(0, a[1], b[0], true);
  waitpid(pid, &n, 0);
  return rc;
}

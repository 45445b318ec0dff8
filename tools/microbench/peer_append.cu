// Two processes, one per GPU.  Every lane group of 8 lanes appends a 64-byte run to one of 1 M "rows": a returning
// atomicAdd on a LOCAL cursor reserves the slot, the run is stored into the row's region of the PEER's buffer (odd rows)
// or of the local buffer (even rows) - the structure of the owner-direct pair scatter (cursor atomics local, records
// remote).  Both GPUs run at the same time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_append peer_append.cu && ./peer_append
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

static uint32_t ROWS = 1u << 20;                      // argv[1] = log2(rows): 20 -> 4 GiB per GPU
constexpr uint32_t ROW_ELEMS = 512;                   // 4 KiB per row

// variant 0: slot from the atomic; 1: slot computed (no atomic); 2: atomic issued but slot computed (no dependency)
__global__ void append_runs(uint2* peer, uint2* local, uint32_t* cursor, int variant, int remote_share_256, uint64_t n_runs,
                            const uint32_t* __restrict__ src, uint32_t ROWS) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t g = lane >> 3, i = lane & 7;
  for (uint64_t r = warp; r * 4 < n_runs; r += n_warps) {
    const uint64_t id = r * 4 + g;
    const uint32_t h = mix((uint32_t)id);
    const uint32_t row = h & (ROWS - 1);
    const uint32_t payload = src ? src[(id * 8 + i) & 0xffffff] : (uint32_t)id;
    uint32_t slot = (h >> 20) & 0xff;
    if (variant != 1) {
      uint32_t s = 0;
      if (i == 0) s = atomicAdd(&cursor[row], 8u);
      s = __shfl_sync(0xffffffffu, s, g * 8);
      if (variant == 0) slot = s;
    }
    uint2* buf = ((h >> 12) & 0xff) < (uint32_t)remote_share_256 ? peer : local;
    buf[(uint64_t)row * ROW_ELEMS + ((slot + i) & (ROW_ELEMS - 1))] = make_uint2(payload, i);
  }
}

static void barrier(int wfd, int rfd) {
  char c = 'x';
  (void)!write(wfd, &c, 1);
  (void)!read(rfd, &c, 1);
}

static int run(int dev, int wfd, int rfd) {
  const uint64_t buf_bytes = (uint64_t)ROWS * ROW_ELEMS * 8;
  CK(cudaSetDevice(dev));
  uint2 *local = nullptr, *peer = nullptr;
  uint32_t *cursor = nullptr, *src = nullptr;
  CK(cudaMalloc(&local, buf_bytes));
  CK(cudaMemset(local, 1, buf_bytes));
  CK(cudaMalloc(&cursor, ROWS * 4));
  CK(cudaMemset(cursor, 0, ROWS * 4));
  CK(cudaMalloc(&src, 64 << 20));
  CK(cudaMemset(src, 3, 64 << 20));
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t mine, theirs;
  CK(cudaIpcGetMemHandle(&mine, local));
  (void)!write(wfd, &mine, sizeof(mine));
  if (read(rfd, &theirs, sizeof(theirs)) != (ssize_t)sizeof(theirs)) return 1;
  CK(cudaIpcOpenMemHandle((void**)&peer, theirs, cudaIpcMemLazyEnablePeerAccess));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const uint64_t n_runs = 32ull << 20;   // 2 GiB of records
  const char* vn[] = {"atomic slot", "no atomic", "atomic, unused"};
  for (int variant = 0; variant < 2; ++variant)
    for (int share : {0, 128, 256})
      for (int loads = 0; loads < 1; ++loads) {
        barrier(wfd, rfd);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaEventRecord(e0));
          append_runs<<<148 * 8, 256>>>(peer, local, cursor, variant, share, n_runs, loads ? src : nullptr, ROWS);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms = 0;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (ms < best) best = ms;
        }
        barrier(wfd, rfd);
        printf("gpu %d  %-15s remote %3d/256  loads %d  %8.2f ms  %8.1f GB/s\n", dev, vn[variant], share, loads, best,
               n_runs * 64 / best / 1e6);
        fflush(stdout);
      }
  barrier(wfd, rfd);
  CK(cudaIpcCloseMemHandle(peer));
  barrier(wfd, rfd);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1) ROWS = 1u << atoi(argv[1]);
  printf("rows %u = %.1f GiB per GPU\n", ROWS, ROWS * 4096.0 / (1 << 30));
  fflush(stdout);
  int a[2], b[2];
  if (pipe(a) || pipe(b)) return 1;
  pid_t pid = fork();
  if (pid == 0) return run(1, b[1], a[0]);
  int n = 0;
  const int rc = run(0, a[1], b[0]);
  waitpid(pid, &n, 0);
  return rc;
}

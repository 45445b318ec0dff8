// C-ABI entry points of the covisitation build (include/otto_covisit.h) and the small kernels around
// the two hot ones (pairgen.cuh, reduce.cuh): ingest, tail CSR, per-aid pair upper bounds, bins.
#include <limits.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "pairgen.cuh"
#include "reduce.cuh"
#include "otable.cuh"
#include "scan.cuh"

static thread_local char g_error[512] = "";
void otto_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
extern "C" const char* otto_last_error(void) { return g_error; }
extern "C" int otto_version(void) { return 100; }
unsigned long long g_otto_launches = 0;
extern "C" uint64_t otto_launch_count(void) { return g_otto_launches; }

// ------------------------------------------------------------------ ingest

__global__ void frame_sorted_kernel(const int32_t* __restrict__ session, const int32_t* __restrict__ ts, int64_t n,
                                    int32_t* flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 >= n) return;
  const int32_t s0 = session[i], s1 = session[i + 1];
  if (s0 > s1 || (s0 == s1 && ts[i] > ts[i + 1])) *flag = 1;
}

extern "C" int otto_frame_is_sorted(const int32_t* session, const int32_t* ts, int64_t n_events, int32_t* flag_dev,
                                    int32_t* sorted_host, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(flag_dev, 0, sizeof(int32_t), st));
  if (n_events > 1) {
    frame_sorted_kernel<<<(unsigned)ceil_div(n_events, 256), 256, 0, st>>>(session, ts, n_events, flag_dev);
    LAUNCH_CHECK();
  }
  int32_t flag = 0;
  CUDA_TRY(cudaMemcpyAsync(&flag, flag_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *sorted_host = flag ? 0 : 1;
  return OTTO_OK;
}

// Event contents the kernels index with: aid in [0, n_aids), type in {0, 1, 2}.  Counts the offending rows.
__global__ void frame_check_kernel(const int32_t* __restrict__ aid, const uint8_t* __restrict__ type, int64_t n, uint32_t n_aids,
                                   unsigned long long* bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool b = i < n && ((uint32_t)aid[i] >= n_aids || type[i] > 2);
  const uint32_t m = __ballot_sync(FULL_MASK, b);
  if (m && lane_id() == 0) atomicAdd(bad, (unsigned long long)__popc(m));
}

extern "C" int otto_frame_check(const int32_t* aid, const uint8_t* type, int64_t n_events, int32_t n_aids, void* count_dev,
                                int64_t* n_bad_host, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_aids <= 0 || !count_dev || !n_bad_host) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  CUDA_TRY(cudaMemsetAsync(count_dev, 0, 8, st));
  if (n_events > 0) {
    frame_check_kernel<<<(unsigned)ceil_div(n_events, 256), 256, 0, st>>>(aid, type, n_events, (uint32_t)n_aids,
                                                                          (unsigned long long*)count_dev);
    LAUNCH_CHECK();
  }
  unsigned long long bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&bad, count_dev, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *n_bad_host = (int64_t)bad;
  if (bad) {
    otto_set_error("%llu events have an aid outside [0, %d) or a type above 2", bad, n_aids);
    return OTTO_EINVAL;
  }
  return OTTO_OK;
}

// ---- frame -> session CSR without materialising per-event flags: boundaries are counted per tile, the tile counts
// scanned, and the second pass recomputes the flags of its tile (replaces torch.unique_consecutive + cumsum of round 1).
// Pass 1 also checks the (session, ts) order and the event contents, so a frame is read once before it is used.
constexpr int ING_THREADS = 256, ING_ITEMS = 8, ING_TILE = ING_THREADS * ING_ITEMS;

__global__ void __launch_bounds__(ING_THREADS)
    ingest_scan_kernel(const int32_t* __restrict__ session, const int32_t* __restrict__ aid, const int32_t* __restrict__ ts,
                       const uint8_t* __restrict__ type, int64_t n, uint32_t n_aids, uint32_t* __restrict__ tile_count,
                       unsigned long long* __restrict__ info /* [0] unsorted [1] bad events */) {
  const int64_t base = (int64_t)blockIdx.x * ING_TILE;
  uint32_t cnt = 0, unsorted = 0, bad = 0;
#pragma unroll
  for (int u = 0; u < ING_ITEMS; ++u) {
    const int64_t i = base + u * ING_THREADS + threadIdx.x;
    if (i < n) {
      const int32_t s1 = session[i];
      if (i == 0) {
        cnt += 1;
      } else {
        const int32_t s0 = session[i - 1];
        cnt += s0 != s1;
        unsorted |= (s0 > s1) || (s0 == s1 && ts[i - 1] > ts[i]);
      }
      if (aid) bad += ((uint32_t)aid[i] >= n_aids) || (type[i] > 2);
    }
  }
  __shared__ uint32_t s_cnt[ING_THREADS / 32], s_bad[ING_THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
    bad += __shfl_xor_sync(FULL_MASK, bad, o);
  }
  if (__any_sync(FULL_MASK, unsorted) && lane_id() == 0) atomicOr(&info[0], 1ull);
  if (lane_id() == 0) { s_cnt[threadIdx.x >> 5] = cnt; s_bad[threadIdx.x >> 5] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t c = 0, b = 0;
    for (int w = 0; w < ING_THREADS / 32; ++w) { c += s_cnt[w]; b += s_bad[w]; }
    tile_count[blockIdx.x] = c;
    if (b) atomicAdd(&info[1], (unsigned long long)b);
  }
}

__global__ void __launch_bounds__(ING_THREADS)
    ingest_offsets_kernel(const int32_t* __restrict__ session, int64_t n, const uint32_t* __restrict__ tile_base,
                          int32_t* __restrict__ ids, int32_t* __restrict__ offsets, int64_t n_sessions) {
  // thread t owns ING_ITEMS CONSECUTIVE events of the tile, so its session starts are written in order
  __shared__ uint32_t s_warp[ING_THREADS / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * ING_TILE + (int64_t)threadIdx.x * ING_ITEMS;
  uint32_t flags = 0, cnt = 0;
  int32_t prev = base > 0 && base <= n ? session[base - 1] : 0;
#pragma unroll
  for (int u = 0; u < ING_ITEMS; ++u) {
    const int64_t i = base + u;
    if (i < n) {
      const int32_t s1 = session[i];
      if (i == 0 || s1 != prev) { flags |= 1u << u; ++cnt; }
      prev = s1;
    }
  }
  uint32_t inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(FULL_MASK, inc, o);
    if ((int)lane_id() >= o) inc += v;
  }
  if (lane_id() == 31) s_warp[threadIdx.x >> 5] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int w = 0; w < ING_THREADS / 32; ++w) { const uint32_t v = s_warp[w]; s_warp[w] = run; run += v; }
  }
  __syncthreads();
  uint32_t at = tile_base[blockIdx.x] + s_warp[threadIdx.x >> 5] + inc - cnt;
#pragma unroll
  for (int u = 0; u < ING_ITEMS; ++u) {
    if ((flags >> u) & 1u) {
      ids[at] = session[base + u];
      offsets[at] = (int32_t)(base + u);
      ++at;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) offsets[n_sessions] = (int32_t)n;
}

__global__ void offsets_max_len_kernel(const int32_t* __restrict__ offsets, int64_t S, int32_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int32_t l = i < S ? offsets[i + 1] - offsets[i] : 0;
  for (int o = 16; o > 0; o >>= 1) l = max(l, __shfl_xor_sync(FULL_MASK, l, o));
  if (lane_id() == 0 && l > 0) atomicMax(out, l);
}

extern "C" int64_t otto_ingest_scratch_bytes(int64_t n_events) {
  const int64_t tiles = ceil_div(n_events > 0 ? n_events : 1, ING_TILE);
  return align_up((tiles + 2) * 4, 256) + scan_scratch_elems(tiles + 1) * 4 + 512;
}

// pass 1.  info_host: [0] n_sessions, [1] sorted by (session, ts) (1 / 0), [2] events with an aid outside [0, n_aids)
// or a type above 2 (aid == NULL skips the content check).  Synchronises.
extern "C" int otto_ingest_scan(const int32_t* session, const int32_t* aid, const int32_t* ts, const uint8_t* type,
                                int64_t n_events, int32_t n_aids, void* scratch, int64_t scratch_bytes, int64_t* info_host,
                                void* stream) {
  if (!session || !ts || !info_host || n_events < 0 || n_events >= (1ll << 31) || (aid && (!type || n_aids <= 0))) {
    otto_set_error("bad argument");
    return OTTO_EINVAL;
  }
  if (!scratch || scratch_bytes < otto_ingest_scratch_bytes(n_events)) { otto_set_error("ingest scratch too small"); return OTTO_ENOSPC; }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t tiles = ceil_div(n_events > 0 ? n_events : 1, ING_TILE);
  char* sc = (char*)scratch;
  uint32_t* tile_count = (uint32_t*)sc;
  uint32_t* scan_sc = (uint32_t*)(sc + align_up((tiles + 2) * 4, 256));
  unsigned long long* info = (unsigned long long*)(scan_sc + scan_scratch_elems(tiles + 1));
  info = (unsigned long long*)(((uintptr_t)info + 7) & ~(uintptr_t)7);
  CUDA_TRY(cudaMemsetAsync(info, 0, 16, st));
  CUDA_TRY(cudaMemsetAsync(tile_count, 0, (tiles + 2) * 4, st));
  if (n_events > 0) {
    ingest_scan_kernel<<<(unsigned)tiles, ING_THREADS, 0, st>>>(session, aid, ts, type, n_events, (uint32_t)n_aids, tile_count, info);
    LAUNCH_CHECK();
  }
  int rc = exclusive_scan<uint32_t, uint32_t>(tile_count, tiles, tile_count, scan_sc, st);
  if (rc) return rc;
  unsigned long long h[2];
  uint32_t total = 0;
  CUDA_TRY(cudaMemcpyAsync(h, info, 16, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&total, tile_count + tiles, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  info_host[0] = n_events > 0 ? (int64_t)total : 0;
  info_host[1] = h[0] ? 0 : 1;
  info_host[2] = (int64_t)h[1];
  return OTTO_OK;
}

// pass 2 (same scratch, untouched since otto_ingest_scan of the same session column): session ids [n_sessions],
// offsets int32 [n_sessions + 1], *max_len_dev = longest session
extern "C" int otto_ingest_offsets(const int32_t* session, int64_t n_events, int64_t n_sessions, void* scratch,
                                   int64_t scratch_bytes, int32_t* session_ids, int32_t* offsets, int32_t* max_len_dev,
                                   void* stream) {
  if (!session || !session_ids || !offsets || n_events < 0 || n_sessions < 0) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  if (!scratch || scratch_bytes < otto_ingest_scratch_bytes(n_events)) { otto_set_error("ingest scratch too small"); return OTTO_ENOSPC; }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t tiles = ceil_div(n_events > 0 ? n_events : 1, ING_TILE);
  if (max_len_dev) CUDA_TRY(cudaMemsetAsync(max_len_dev, 0, 4, st));
  if (n_events == 0) {
    CUDA_TRY(cudaMemsetAsync(offsets, 0, 4, st));
    return OTTO_OK;
  }
  ingest_offsets_kernel<<<(unsigned)tiles, ING_THREADS, 0, st>>>(session, n_events, (const uint32_t*)scratch, session_ids, offsets, n_sessions);
  LAUNCH_CHECK();
  if (max_len_dev && n_sessions > 0) {
    offsets_max_len_kernel<<<(unsigned)ceil_div(n_sessions, 256), 256, 0, st>>>(offsets, n_sessions, max_len_dev);
    LAUNCH_CHECK();
  }
  return OTTO_OK;
}

// One warp per session: reverse the ascending session so that ts is descending, keeping runs of equal
// ts in their original order (what the stable ts-descending sort of builder step 2 produces).
__global__ void __launch_bounds__(256)
    ingest_desc_kernel(const int32_t* __restrict__ off, int64_t n_sessions, const int32_t* __restrict__ aid,
                       const int32_t* __restrict__ ts, const uint8_t* __restrict__ type, int32_t* __restrict__ aid_out,
                       int32_t* __restrict__ ts_out, uint8_t* __restrict__ type_out) {
  const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (s >= n_sessions) return;
  const uint32_t lane = lane_id();
  const int32_t beg = off[s], end = off[s + 1];
  for (int32_t p = beg + (int32_t)lane; p < end; p += 32) {
    const int32_t t = ts[p];
    int32_t a = p, r = p + 1;  // run [a, r) of equal ts around p
    while (a > beg && ts[a - 1] == t) --a;
    while (r < end && ts[r] == t) ++r;
    const int32_t dst = beg + (end - r) + (p - a);
    aid_out[dst] = aid[p];
    ts_out[dst] = t;
    type_out[dst] = type[p];
  }
}

extern "C" int otto_ingest_desc(const int32_t* session_offsets, int64_t n_sessions, const int32_t* aid,
                                const int32_t* ts, const uint8_t* type, int64_t n_events, int32_t* aid_out,
                                int32_t* ts_out, uint8_t* type_out, void* stream) {
  (void)n_events;
  if (n_sessions <= 0) return OTTO_OK;
  ingest_desc_kernel<<<(unsigned)ceil_div(n_sessions, 8), 256, 0, (cudaStream_t)stream>>>(
      session_offsets, n_sessions, aid, ts, type, aid_out, ts_out, type_out);
  LAUNCH_CHECK();
  return OTTO_OK;
}

// ------------------------------------------------------------------ workspace layout

// What the tail kernels check about every event they copy: the pair kernels index arrays with the aid (row histogram,
// cursors, owner cuts - in owner-direct mode a PEER's memory), pack the type into two bits and store ts - ts_min as
// u32.  An offending event is replaced by (aid 0, type 0, ts_lo) so that nothing is written out of bounds, and
// flagged in stats[STAT_BAD_EVENTS]; count_finish then returns OTTO_EINVAL.
struct EventLimits {
  uint32_t n_aids;
  int32_t ts_lo, ts_hi;   // inclusive range (INT32_MIN .. INT32_MAX outside time mode)
};
constexpr int STAT_BAD_EVENTS = 6, STAT_MAX_BIN = 7;

__device__ __forceinline__ void check_event(const EventLimits& lim, int32_t& a, uint32_t& ty, int32_t& t, unsigned long long* stats) {
  if ((uint32_t)a >= lim.n_aids || ty > 2u || t < lim.ts_lo || t > lim.ts_hi) {
    a = 0;
    ty = 0;
    t = lim.ts_lo;
    atomicAdd(&stats[STAT_BAD_EVENTS], 1ull);
  }
}

struct Layout {
  int64_t S, E, Ecap, A, Bmax, Hmax;
  int64_t tail_off, tail_aw, tail_ts, winmask, row_count, row_total, bin_base, bin_x, bin_cnt, bin_off, row_off, hot_off, cursor,
      hot_rows, sub_cur, tile_row, tile_idx, Tmax, scan, stats, total;
};

// A row whose pair count (over all ranks) exceeds split_ub is "hot": it is split into aid_y-hash sub-bins of
// about split_ub / 3 records (2048 by default), which the 256-thread owner-table kernel takes (bins <= 3072).  The
// default split_ub is the largest bin the owner-table tiers take (6144 records), so that the hash-table kernel of
// reduce.cuh only ever sees hand-overs (round 1 / v5: 8192 and split_ub / 4).
static int32_t effective_split_ub(const OttoCovisitSpec* spec) { return spec->split_ub > 0 ? spec->split_ub : 6144; }
static int32_t sub_bin_target(const OttoCovisitSpec* spec) { return effective_split_ub(spec) >= 3 ? effective_split_ub(spec) / 3 : 1; }
constexpr uint32_t MAX_SUB_BINS = 4096;   // shared-memory cursors of the partition kernels

static int check_spec(const OttoCovisitSpec* spec) {
  if (!spec) { otto_set_error("spec is NULL"); return OTTO_EINVAL; }
  if (spec->n_aids <= 0 || spec->n_aids > (1 << 30)) { otto_set_error("n_aids must be in (0, 2^30]"); return OTTO_EINVAL; }
  if (spec->tail_n < 1 || spec->tail_n > OTTO_MAX_TAIL) { otto_set_error("tail_n must be in [1, 32]"); return OTTO_EINVAL; }
  if (spec->k < 1 || spec->k > OTTO_MAX_K) { otto_set_error("k must be in [1, 32]"); return OTTO_EINVAL; }
  if (spec->window_s <= 0) { otto_set_error("window_s must be positive"); return OTTO_EINVAL; }
  if (spec->weight_mode < 0 || spec->weight_mode > 2) { otto_set_error("bad weight_mode"); return OTTO_EINVAL; }
  if (spec->weight_mode == OTTO_WEIGHT_TIME) {
    if (spec->ts_max <= spec->ts_min || (int64_t)spec->ts_max - spec->ts_min >= (1 << 24)) {
      otto_set_error("time weights need 0 < ts_max - ts_min < 2^24");
      return OTTO_EINVAL;
    }
  }
  if (spec->weight_mode == OTTO_WEIGHT_TYPE)
    for (int i = 0; i < 3; ++i)
      if (spec->type_weight[i] < 0 || spec->type_weight[i] > 4096) { otto_set_error("type_weight must be in [0, 4096]"); return OTTO_EINVAL; }
  if ((spec->event_type_mask & 7u) == 0) { otto_set_error("event_type_mask selects nothing"); return OTTO_EINVAL; }
  return OTTO_OK;
}

static Layout make_layout(int64_t S, int64_t E, const OttoCovisitSpec* spec) {
  Layout L;
  L.S = S;
  L.E = E;
  L.A = spec->n_aids;
  L.Ecap = E < S * spec->tail_n ? E : S * spec->tail_n;
  if (L.Ecap < 1) L.Ecap = 1;
  // extra bins = sum over hot rows of (sub-bins - 1) <= (pairs of all ranks) / target <= events * (tail_n - 1) / target,
  // with the events of every rank when the row counts are all-reduced (multi-GPU); hot rows <= extra bins
  const int64_t Eg = spec->global_events > E ? spec->global_events : L.Ecap;
  L.Hmax = (Eg * (spec->tail_n - 1)) / sub_bin_target(spec) + 1;
  if (L.Hmax > L.A) L.Hmax = L.A + 1;
  L.Bmax = L.A + (Eg * (spec->tail_n - 1)) / sub_bin_target(spec) + 1;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t at = o; o = align_up(o + bytes, 256); return at; };
  L.tail_off = take((S + 2) * 4);
  L.tail_aw = take(L.Ecap * 4);
  L.tail_ts = take(L.Ecap * 4);
  L.winmask = take(L.Ecap * 4);
  L.row_count = take((L.A + 1) * 4);
  L.row_total = take((L.A + 1) * 4);
  L.bin_base = take((L.A + 2) * 4);
  L.bin_x = take(L.Bmax * 4);
  L.bin_cnt = take((L.Bmax + 1) * 4);
  L.bin_off = take((L.Bmax + 2) * 8);
  L.row_off = take((L.A + 2) * 8);
  L.hot_off = take((L.A + 2) * 8);
  L.cursor = take((L.A + 1) * 4);
  L.hot_rows = take((L.Hmax + 1) * 4);
  L.sub_cur = take((L.Bmax + 1) * 4);
  // tiles of the staging area: every hot row rounds up once
  L.Tmax = (Eg * (spec->tail_n - 1)) / 2048 + L.Hmax + 1;
  L.tile_row = take(L.Tmax * 4);
  L.tile_idx = take(L.Tmax * 4);
  int64_t scan_elems = scan_scratch_elems(S + 1);
  if (scan_scratch_elems(L.Bmax + 1) > scan_elems) scan_elems = scan_scratch_elems(L.Bmax + 1);
  if (scan_scratch_elems(L.A + 1) > scan_elems) scan_elems = scan_scratch_elems(L.A + 1);
  L.scan = take(scan_elems * 8);
  L.stats = take(512);
  L.total = o;
  return L;
}

extern "C" int otto_covisit_sizes(int64_t n_sessions, int64_t n_events, const OttoCovisitSpec* spec,
                                  OttoBuildSizes* out_host) {
  int rc = check_spec(spec);
  if (rc) return rc;
  if (n_sessions < 0 || n_events < 0 || n_events >= (1ll << 31)) { otto_set_error("n_events must be < 2^31"); return OTTO_EINVAL; }
  const Layout L = make_layout(n_sessions, n_events, spec);
  out_host->tail_capacity = L.Ecap;
  out_host->max_bins = L.Bmax;
  out_host->workspace_bytes = L.total;
  return OTTO_OK;
}

static int g_profile = 0;
static cudaEvent_t g_prof_sc[4];
static bool g_prof_sc_valid = false;
static int side_streams_init();
extern cudaStream_t g_side[3];
extern cudaEvent_t g_fork, g_join[3];

// ------------------------------------------------------------------ tail CSR + upper bounds

// events of each session that enter the self-join: the first tail_n of the (type-filtered) desc session
__global__ void tail_count_kernel(const int32_t* __restrict__ off, const uint8_t* __restrict__ type, int64_t S,
                                  uint32_t mask, int32_t tail_n, uint32_t* __restrict__ cnt) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int32_t beg = off[s], end = off[s + 1];
  int32_t n = 0;
  if ((mask & 7u) == 7u) {
    n = min(end - beg, tail_n);
  } else {
    for (int32_t p = beg; p < end && n < tail_n; ++p) n += type[p] < 32 ? (mask >> type[p]) & 1u : 0u;
  }
  cnt[s] = (uint32_t)n;
}

// Type-filtered variants (buy2buy keeps carts / orders only: about one event in ten): one THREAD per session
// walks its events until tail_n of the allowed types are taken.  Neighbouring sessions' tails are neighbours in
// the tail CSR, so the short runs written by neighbouring lanes share cache lines; the warp-per-session
// tail_copy_kernel below spent 3.9 ms of a 9.3 ms buy2buy build here (12.9 M warps for 22 M kept events).
__global__ void __launch_bounds__(256)
    tail_copy_filtered_kernel(const int32_t* __restrict__ off, const int32_t* __restrict__ aid, const int32_t* __restrict__ ts,
                              const uint8_t* __restrict__ type, int64_t S, uint32_t mask, const uint32_t* __restrict__ tail_off,
                              uint32_t* __restrict__ tail_aw, int32_t* __restrict__ tail_ts, const EventLimits lim,
                              unsigned long long* stats) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const uint32_t tb = tail_off[s];
  const uint32_t n = tail_off[s + 1] - tb;
  if (n == 0) return;
  const int32_t end = off[s + 1];
  uint32_t taken = 0;
  for (int32_t p = off[s]; p < end && taken < n; ++p) {
    uint32_t ty = type[p];
    if (ty < 32u && ((mask >> ty) & 1u)) {
      int32_t a = aid[p], t = ts[p];
      check_event(lim, a, ty, t, stats);
      tail_aw[tb + taken] = (uint32_t)a | (ty << 30);
      tail_ts[tb + taken] = t;
      ++taken;
    }
  }
}

// Unfiltered variants (every event type enters the tail): the tail of a session is simply its first
// min(len, tail_n) events, so a warp takes 32 consecutive sessions and one lane per OUTPUT event; the owning
// session of an output position is found by a 5-step binary descent over the 32 tail offsets held in
// registers.  Loads and stores are coalesced across session boundaries (tail_copy_kernel spends a whole
// warp per session, which leaves two thirds of the lanes idle at the median session length of 6).
__global__ void __launch_bounds__(256)
    tail_copy_all_kernel(const int32_t* __restrict__ off, const int32_t* __restrict__ aid, const int32_t* __restrict__ ts,
                         const uint8_t* __restrict__ type, int64_t S, const uint32_t* __restrict__ tail_off,
                         uint32_t* __restrict__ tail_aw, int32_t* __restrict__ tail_ts, const EventLimits lim,
                         unsigned long long* stats) {
  const int64_t s0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
  if (s0 >= S) return;
  const int lane = (int)lane_id();
  const int ns = (int)min((int64_t)32, S - s0);
  const int32_t src0 = off[s0 + min(lane, ns - 1)];
  const uint32_t to = tail_off[s0 + min(lane, ns)];
  const uint32_t to_next = tail_off[s0 + min(lane + 1, ns)];
  const uint32_t T0 = __shfl_sync(FULL_MASK, to, 0);
  const uint32_t total = __shfl_sync(FULL_MASK, to_next, ns - 1) - T0;
  for (uint32_t q0 = 0; q0 < total; q0 += 32) {
    const uint32_t q = q0 + lane;
    int u = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      const int c = u + step;
      const uint32_t tc = __shfl_sync(FULL_MASK, to, min(c, 31));
      if (c < ns && tc - T0 <= q) u = c;
    }
    const uint32_t tu = __shfl_sync(FULL_MASK, to, u);
    const int32_t src = __shfl_sync(FULL_MASK, src0, u) + (int32_t)(q - (tu - T0));
    if (q < total) {
      int32_t a = aid[src], t = ts[src];
      uint32_t ty = type[src];
      check_event(lim, a, ty, t, stats);
      tail_aw[T0 + q] = (uint32_t)a | (ty << 30);
      tail_ts[T0 + q] = t;
    }
  }
}

// ---- the same two kernels for a frame in FILE order (ts ascending inside a session), which is how the reference's
// frames are written (utilities/split_dataset_writer_parquet.py:17) and how a host frame arrives: the builder's
// stable sort_values(['session','ts'], ascending=[True, False]) reverses the runs of equal ts as runs and keeps the
// original order inside a run, so output position j of a session of L events (0 = most recent) mirrors source
// p = L - 1 - j inside the run [a, r) of equal ts around p: source = a + r - 1 - p.  Only the tails are read, so aid
// and type may stay in pinned host memory (the pointers are valid on the device under UVA): 5 bytes per TAIL event
// cross PCIe instead of 5 bytes per event, and the whole-frame reversal pass (otto_ingest_desc) is not needed.  ts is
// read with its neighbours and belongs in device memory.
__device__ __forceinline__ int32_t desc_source(const int32_t* __restrict__ ts, int32_t beg, int32_t end, int32_t p, int32_t& t) {
  t = ts[p];
  int32_t a = p, r = p + 1;
  while (a > beg && ts[a - 1] == t) --a;
  while (r < end && ts[r] == t) ++r;
  return a + r - 1 - p;
}

__global__ void tail_count_asc_kernel(const int32_t* __restrict__ off, const uint8_t* __restrict__ type, int64_t S,
                                      uint32_t mask, int32_t tail_n, uint32_t* __restrict__ cnt) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int32_t beg = off[s], end = off[s + 1];
  int32_t n = 0;
  if ((mask & 7u) == 7u) {
    n = min(end - beg, tail_n);
  } else {
    for (int32_t p = end - 1; p >= beg && n < tail_n; --p) n += type[p] < 32 ? (mask >> type[p]) & 1u : 0u;
  }
  cnt[s] = (uint32_t)n;
}

// one thread per session walks the runs of equal ts from the last one backwards, each run in file order
__global__ void __launch_bounds__(256)
    tail_copy_filtered_asc_kernel(const int32_t* __restrict__ off, const int32_t* __restrict__ aid, const int32_t* __restrict__ ts,
                                  const uint8_t* __restrict__ type, int64_t S, uint32_t mask, const uint32_t* __restrict__ tail_off,
                                  uint32_t* __restrict__ tail_aw, int32_t* __restrict__ tail_ts, const EventLimits lim,
                                  unsigned long long* stats) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const uint32_t tb = tail_off[s];
  const uint32_t n = tail_off[s + 1] - tb;
  if (n == 0) return;
  const int32_t beg = off[s];
  uint32_t taken = 0;
  for (int32_t r = off[s + 1]; r > beg && taken < n;) {
    const int32_t t0 = ts[r - 1];
    int32_t a = r - 1;
    while (a > beg && ts[a - 1] == t0) --a;
    for (int32_t p = a; p < r && taken < n; ++p) {
      uint32_t ty = type[p];
      if (ty < 32u && ((mask >> ty) & 1u)) {
        int32_t av = aid[p], t = t0;
        check_event(lim, av, ty, t, stats);
        tail_aw[tb + taken] = (uint32_t)av | (ty << 30);
        tail_ts[tb + taken] = t;
        ++taken;
      }
    }
    r = a;
  }
}

__global__ void __launch_bounds__(256)
    tail_copy_all_asc_kernel(const int32_t* __restrict__ off, const int32_t* __restrict__ aid, const int32_t* __restrict__ ts,
                             const uint8_t* __restrict__ type, int64_t S, const uint32_t* __restrict__ tail_off,
                             uint32_t* __restrict__ tail_aw, int32_t* __restrict__ tail_ts, const EventLimits lim,
                             unsigned long long* stats) {
  const int64_t s0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
  if (s0 >= S) return;
  const int lane = (int)lane_id();
  const int ns = (int)min((int64_t)32, S - s0);
  const int32_t beg0 = off[s0 + min(lane, ns - 1)];
  const int32_t end0 = off[s0 + min(lane, ns - 1) + 1];
  const uint32_t to = tail_off[s0 + min(lane, ns)];
  const uint32_t to_next = tail_off[s0 + min(lane + 1, ns)];
  const uint32_t T0 = __shfl_sync(FULL_MASK, to, 0);
  const uint32_t total = __shfl_sync(FULL_MASK, to_next, ns - 1) - T0;
  for (uint32_t q0 = 0; q0 < total; q0 += 32) {
    const uint32_t q = q0 + lane;
    int u = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      const int c = u + step;
      const uint32_t tc = __shfl_sync(FULL_MASK, to, min(c, 31));
      if (c < ns && tc - T0 <= q) u = c;
    }
    const uint32_t tu = __shfl_sync(FULL_MASK, to, u);
    const int32_t beg = __shfl_sync(FULL_MASK, beg0, u), end = __shfl_sync(FULL_MASK, end0, u);
    if (q < total) {
      int32_t t;
      const int32_t src = desc_source(ts, beg, end, end - 1 - (int32_t)(q - (tu - T0)), t);
      int32_t a = aid[src];
      uint32_t ty = type[src];
      check_event(lim, a, ty, t, stats);
      tail_aw[T0 + q] = (uint32_t)a | (ty << 30);
      tail_ts[T0 + q] = t;
    }
  }
}

// Bins of a row: 1, or ceil(total / target) aid_y-hash sub-bins when the row is hot (total over ALL ranks above
// split_ub).  Hot rows are also listed (any order) for the partition kernels; hot_cnt = the rank's own pairs of a
// hot row (its share of the staging area), 0 for ordinary rows.
__global__ void bins_count_kernel(const uint32_t* __restrict__ row_total, const uint32_t* __restrict__ row_count, int64_t A,
                                  uint32_t split_ub, uint32_t target, uint32_t* __restrict__ nb,
                                  unsigned long long* __restrict__ hot_cnt, uint32_t* __restrict__ hot_rows, int64_t hot_cap,
                                  unsigned long long* stats) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= A) return;
  const uint32_t tot = row_total[x];
  uint32_t n = 1;
  unsigned long long hc = 0;
  if (tot > split_ub) {
    n = (tot + target - 1) / target;
    if (n > MAX_SUB_BINS) n = MAX_SUB_BINS;
    if (n > 1) {
      hc = row_count[x];
      const unsigned long long at = atomicAdd(&stats[3], 1ull);
      if ((int64_t)at < hot_cap) hot_rows[at] = (uint32_t)x;
    }
  }
  nb[x] = n;
  hot_cnt[x] = hc;
  // largest bin this layout can produce (sub-bins: four times the mean covers the aid_y-hash imbalance); the
  // hash-table kernel of reduce.cuh holds a 24-bit count and a 40-bit time sum per entry (STAT_MAX_BIN guard)
  const unsigned long long est = n > 1 ? 4ull * ((tot + n - 1) / n) : tot;
  if (est > TIER3_MAX) atomicMax(&stats[STAT_MAX_BIN], est);   // bins the owner-table tiers take cannot overflow
}

__global__ void bins_fill_kernel(const uint32_t* __restrict__ bin_base, int64_t A, uint32_t* __restrict__ bin_x) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= A) return;
  for (uint32_t b = bin_base[x]; b < bin_base[x + 1]; ++b) bin_x[b] = (uint32_t)x;
}

// scatter cursor of a row: its final position for an ordinary row, its slice of the staging area (behind the
// P final records) for a hot row
__global__ void init_cursor_kernel(const unsigned long long* __restrict__ row_off, const unsigned long long* __restrict__ hot_off,
                                   const uint32_t* __restrict__ bin_base, int64_t A, uint32_t* __restrict__ cursor) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= A) return;
  const bool hot = bin_base[x + 1] - bin_base[x] > 1;
  cursor[x] = hot ? (uint32_t)(row_off[A] + hot_off[x]) : (uint32_t)row_off[x];
}

// stats[0] tail events, [1] pairs, [2] bins, [3] hot rows (counted by bins_count_kernel), [4] staged (hot) pairs
__global__ void count_stats_kernel(const uint32_t* tail_off, int64_t S, const uint32_t* bin_base, int64_t A,
                                   const unsigned long long* row_off, const unsigned long long* hot_off,
                                   unsigned long long* stats) {
  stats[0] = tail_off[S];
  stats[1] = row_off[A];
  stats[2] = bin_base[A];
  stats[4] = hot_off[A];
}

// ---- owner-direct scatter (multi-GPU): every rank lays out ALL owners' buffers the same way, from the row totals.
// E = exclusive scan of the row totals, Hs = exclusive scan of the hot rows' totals.  Owner o holds rows
// [cut[o], cut[o + 1]): P_o = E[cut[o + 1]] - E[cut[o]] final records followed by its staging area.
struct OwnerCuts {
  int32_t n, rank;
  uint32_t cut[OTTO_MAX_OWNERS + 1];
};
constexpr int STAT_CUT_E = 8, STAT_CUT_H = 8 + OTTO_MAX_OWNERS + 1;   // stats[] slots of E / Hs at the cuts
constexpr int STAT_CUT_B = STAT_CUT_H + OTTO_MAX_OWNERS + 1;          // first bin of every owner's range
constexpr int STAT_WORDS = STAT_CUT_B + OTTO_MAX_OWNERS + 1;
static_assert(STAT_WORDS * 8 <= 512, "stats region");

__global__ void owner_cut_values_kernel(const OwnerCuts c, const unsigned long long* __restrict__ E,
                                        const unsigned long long* __restrict__ Hs, const uint32_t* __restrict__ bin_base,
                                        unsigned long long* stats) {
  const int o = threadIdx.x;
  if (o <= c.n) {
    stats[STAT_CUT_E + o] = E[c.cut[o]];
    stats[STAT_CUT_H + o] = Hs[c.cut[o]];
    stats[STAT_CUT_B + o] = bin_base[c.cut[o]];
  }
}

// cursor of row x = its position inside its OWNER's buffer + the pairs that lower ranks hold of it.  The inputs of
// the local layout are then rewritten to "my rows with their totals over all ranks, nothing else": with them the
// single-GPU phases (offsets, partition, bin counts) describe exactly what the peers are about to write here.
__global__ void init_cursor_owned_kernel(const OwnerCuts c, const uint32_t* __restrict__ row_total,
                                         const uint32_t* __restrict__ row_before, const uint32_t* __restrict__ bin_base,
                                         int64_t A, const unsigned long long* __restrict__ E, unsigned long long* hot_off,
                                         const unsigned long long* __restrict__ stats, uint32_t* __restrict__ cursor,
                                         uint32_t* __restrict__ row_count) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= A) return;
  int o = 0;
#pragma unroll
  for (int g = 1; g < OTTO_MAX_OWNERS; ++g)
    if (g < c.n && (uint32_t)x >= c.cut[g]) o = g;
  const bool hot = bin_base[x + 1] - bin_base[x] > 1;
  const unsigned long long P_o = stats[STAT_CUT_E + o + 1] - stats[STAT_CUT_E + o];
  const unsigned long long pos = hot ? P_o + (hot_off[x] - stats[STAT_CUT_H + o]) : E[x] - stats[STAT_CUT_E + o];
  cursor[x] = (uint32_t)(pos + row_before[x]);
  const bool mine = o == c.rank;
  const uint32_t tot = row_total[x];
  row_count[x] = mine ? tot : 0u;
  hot_off[x] = (mine && hot) ? tot : 0ull;
}

// records per bin of the ordinary rows (the sub-bins of hot rows are counted by partition_count_kernel)
__global__ void bin_cnt_rows_kernel(const uint32_t* __restrict__ row_count, const uint32_t* __restrict__ bin_base, int64_t A,
                                    uint32_t* __restrict__ bin_cnt) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= A) return;
  const uint32_t b0 = bin_base[x];
  if (bin_base[x + 1] - b0 == 1) bin_cnt[b0] = row_count[x];
}

// ---- hot rows: staging area -> aid_y-hash sub-bins.  Work unit = a tile of PART_TILE staged records of one hot
// row, so the hottest rows (millions of records) spread over the whole GPU.  Pass 1 histograms a tile in shared
// memory and adds the non-zero counts to the bins; an exclusive scan over all bins gives the final offsets; pass 2
// histograms again (rank of every record inside its tile and sub-bin), reserves each sub-bin's chunk with one
// global atomic per tile and writes.  The sub-bin chunks of a tile are written within microseconds by one CTA, so
// their lines complete in L2.
constexpr int PART_THREADS = 256;
constexpr int PART_PER_THREAD = 8;
constexpr int PART_TILE = PART_THREADS * PART_PER_THREAD;

struct PartParams {
  const uint32_t* hot_rows;
  const unsigned long long* n_hot;     // stats[3]
  int64_t hot_cap;
  const uint32_t* bin_base;
  const uint32_t* row_count;
  const unsigned long long* row_off;   // row_off[A] = P: start of the staging area
  const unsigned long long* hot_off;
  int64_t A;
  uint32_t* bin_cnt;
  uint32_t* sub_cur;                   // [B] next free record slot of every bin (pass 2)
  uint32_t* tile_row;                  // [tiles] index into hot_rows
  uint32_t* tile_idx;                  // [tiles] tile number inside the row
  unsigned long long* n_tiles;         // stats[5]
  int64_t tile_cap;
  uint2* records;
};

// one thread per hot row: its tiles, appended to the tile list (any order)
__global__ void partition_tiles_kernel(const PartParams p) {
  unsigned long long n_hot = *p.n_hot;
  if ((int64_t)n_hot > p.hot_cap) n_hot = (unsigned long long)p.hot_cap;
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_hot) return;
  const uint32_t n = p.row_count[p.hot_rows[i]];
  const uint32_t nt = (n + PART_TILE - 1) / PART_TILE;
  if (nt == 0) return;
  const unsigned long long at = atomicAdd(p.n_tiles, (unsigned long long)nt);
  for (uint32_t t = 0; t < nt && (int64_t)(at + t) < p.tile_cap; ++t) {
    p.tile_row[at + t] = (uint32_t)i;
    p.tile_idx[at + t] = t;
  }
}

__global__ void init_sub_cur_kernel(const unsigned long long* __restrict__ bin_off, const uint32_t* __restrict__ bin_base,
                                    int64_t A, uint32_t* __restrict__ sub_cur) {
  const int64_t B = bin_base[A];
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x)
    sub_cur[b] = (uint32_t)bin_off[b];
}

// A tile is 16 KB of CONTIGUOUS staged records: it is brought into shared memory by one bulk asynchronous copy
// (cp.async.bulk.shared::cluster.global, the 1-D TMA path: UBLKCP in SASS) that thread 0 issues one tile ahead into
// the other of two buffers and that completes on an mbarrier - the loads of a tile cost the SM no issue slots and no
// registers, and the next tile streams in while this one is histogrammed and moved (round 1 / early round 2: eight
// LDG.64 per thread up front, which held 16 registers per thread and left the copy engine idle between tiles).
// Bulk copies need 16-byte aligned addresses and sizes and a record is 8 bytes: the copy covers the 16-byte aligned
// span inside the tile's bytes; a first or last record outside it is read with a plain load.
constexpr int PART_STAGES = 2;
constexpr uint32_t PART_BUF_BYTES = PART_TILE * 8 + 16;

struct PartTile {
  uint32_t b0, nb, m, head;      // first bin of the row, its sub-bins, records of the tile, bytes in front of record 0 in the buffer
  uint32_t bulk_bytes, pad;
  const uint2* src;              // record 0 of the tile in the staging area
};

__device__ __forceinline__ void mbar_init(uint32_t bar_s, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar_s, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar_s) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar_s, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar_s), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_s, const void* src, uint32_t bytes, uint32_t bar_s) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s), "l"(src),
               "r"(bytes), "r"(bar_s) : "memory");
}

template <bool MOVE>
constexpr size_t partition_smem() { return (size_t)PART_STAGES * PART_BUF_BYTES + (size_t)MAX_SUB_BINS * 4 * (MOVE ? 2 : 1); }

template <bool MOVE>
__global__ void __launch_bounds__(PART_THREADS) partition_kernel(const PartParams p) {
  extern __shared__ __align__(128) unsigned char part_smem[];
  __shared__ __align__(8) unsigned long long s_bar[PART_STAGES];
  __shared__ PartTile s_tile[PART_STAGES];
  uint32_t* s_hist = (uint32_t*)(part_smem + PART_STAGES * PART_BUF_BYTES);
  uint32_t* s_base = s_hist + MAX_SUB_BINS;     // MOVE only
  const uint32_t buf_s = (uint32_t)__cvta_generic_to_shared(part_smem);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(s_bar);
  unsigned long long n_tiles = *p.n_tiles;
  if ((int64_t)n_tiles > p.tile_cap) n_tiles = (unsigned long long)p.tile_cap;
  const uint2* stage_area = p.records + p.row_off[p.A];
  // thread 0: metadata of tile t into s_tile[st], bulk copy of its aligned span into buffer st
  auto issue = [&](unsigned long long t, int st) {
    const uint32_t x = p.hot_rows[p.tile_row[t]];
    PartTile ti;
    ti.b0 = p.bin_base[x];
    ti.nb = p.bin_base[x + 1] - ti.b0;
    const uint32_t n = p.row_count[x];
    const uint32_t t0 = p.tile_idx[t] * PART_TILE;
    ti.m = min((uint32_t)PART_TILE, n - t0);
    ti.src = stage_area + p.hot_off[x] + t0;
    const uintptr_t a = (uintptr_t)ti.src, e = a + (uintptr_t)ti.m * 8;
    const uintptr_t a0 = (a + 15) & ~(uintptr_t)15, a1 = e & ~(uintptr_t)15;     // aligned span inside [a, e)
    ti.head = (uint32_t)(a0 - a);                                                 // 0 or 8: record 0 lies in front of the span
    ti.bulk_bytes = a1 > a0 ? (uint32_t)(a1 - a0) : 0u;
    ti.pad = 0;
    s_tile[st] = ti;
    if (ti.bulk_bytes) {
      mbar_expect_tx(bar_s + st * 8, ti.bulk_bytes);
      bulk_g2s(buf_s + st * PART_BUF_BYTES, (const void*)a0, ti.bulk_bytes, bar_s + st * 8);
    } else {
      mbar_arrive(bar_s + st * 8);
    }
  };
  if (threadIdx.x == 0) {
    for (int i = 0; i < PART_STAGES; ++i) mbar_init(bar_s + i * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (blockIdx.x < n_tiles) issue(blockIdx.x, 0);
  }
  __syncthreads();
  uint32_t it = 0;
  for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
    const int st = it & 1;
    // the other buffer was consumed in the previous iteration (trailing barrier): refill it while this tile is processed
    if (threadIdx.x == 0 && t + gridDim.x < n_tiles) issue(t + gridDim.x, st ^ 1);
    const PartTile ti = s_tile[st];
    for (uint32_t j = threadIdx.x; j < ti.nb; j += PART_THREADS) s_hist[j] = 0u;
    mbar_wait(bar_s + st * 8, (it >> 1) & 1u);
    const uint32_t tile_s = buf_s + st * PART_BUF_BYTES;
    // record j: inside the copied span at byte 8 j - head, else (first / last record of a misaligned tile) from global memory
    auto record = [&](uint32_t j) {
      const uint32_t off = 8u * j - ti.head;          // wraps to 0xfffffff8 for j = 0 with head = 8: outside the span
      uint2 r;
      if (ti.bulk_bytes >= 8u && off <= ti.bulk_bytes - 8u) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(tile_s + off) : "memory");
      else r = ld_stream_u2(ti.src + j);
      return r;
    };
    __syncthreads();
    uint32_t sub[PART_PER_THREAD], rank[PART_PER_THREAD];
#pragma unroll
    for (int u = 0; u < PART_PER_THREAD; ++u) {
      const uint32_t j = u * PART_THREADS + threadIdx.x;
      sub[u] = 0;
      rank[u] = 0;
      if (j < ti.m) {
        sub[u] = sub_bin(record(j).x, ti.nb);
        rank[u] = atomicAdd(&s_hist[sub[u]], 1u);
      }
    }
    __syncthreads();
    if (!MOVE) {
      for (uint32_t j = threadIdx.x; j < ti.nb; j += PART_THREADS)
        if (s_hist[j]) atomicAdd(&p.bin_cnt[ti.b0 + j], s_hist[j]);
    } else {
      for (uint32_t j = threadIdx.x; j < ti.nb; j += PART_THREADS)
        if (s_hist[j]) s_base[j] = atomicAdd(&p.sub_cur[ti.b0 + j], s_hist[j]);
      __syncthreads();
#pragma unroll
      for (int u = 0; u < PART_PER_THREAD; ++u) {
        const uint32_t j = u * PART_THREADS + threadIdx.x;
        if (j < ti.m) p.records[s_base[sub[u]] + rank[u]] = record(j);
      }
    }
    __syncthreads();
  }
}

#define WS(type, field) ((type*)((char*)workspace + L.field))

static int check_ws(const Layout& L, void* workspace, int64_t workspace_bytes) {
  if (!workspace || workspace_bytes < L.total) {
    otto_set_error("workspace too small: need %lld bytes, got %lld", (long long)L.total, (long long)workspace_bytes);
    return OTTO_ENOSPC;
  }
  if (((uintptr_t)workspace & 255) != 0) { otto_set_error("workspace must be 256-byte aligned"); return OTTO_EINVAL; }
  return OTTO_OK;
}

static PairGenParams make_pairgen(const Layout& L, const OttoCovisitSpec* spec, void* workspace) {
  PairGenParams p;
  p.tail_off = WS(uint32_t, tail_off);
  p.tail_aw = WS(uint32_t, tail_aw);
  p.tail_ts = WS(int32_t, tail_ts);
  p.winmask = WS(uint32_t, winmask);
  p.row_hist = WS(uint32_t, row_count);
  p.cursor = WS(uint32_t, cursor);
  p.records = nullptr;
  p.n_sessions = L.S;
  p.window = (uint32_t)spec->window_s;
  p.x_type_mask = spec->x_type_mask;
  p.y_type_mask = spec->y_type_mask;
  p.weight_mode = spec->weight_mode;
  p.ts_min = spec->ts_min;
  for (int i = 0; i < 3; ++i) p.type_weight[i] = (uint32_t)spec->type_weight[i];
  p.n_owners = 1;
  for (int o = 0; o <= OTTO_MAX_OWNERS; ++o) p.cuts[o] = 0;
  for (int o = 0; o < OTTO_MAX_OWNERS; ++o) p.owner_rec[o] = nullptr;
  return p;
}

// tail CSR (steps 1-3) + in-session dedupe (steps 4-5: row masks) + pairs per aid_x row
static int count_begin_impl(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                            void* stream, bool ascending) {
  int rc = check_spec(spec);
  if (rc) return rc;
  if (!ev) { otto_set_error("events is NULL"); return OTTO_EINVAL; }
  const Layout L = make_layout(ev->n_sessions, ev->n_events, spec);
  if ((rc = check_ws(L, workspace, workspace_bytes))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t S = L.S;
  CUDA_TRY(cudaMemsetAsync(WS(uint32_t, row_count), 0, (L.A + 1) * 4, st));
  CUDA_TRY(cudaMemsetAsync(WS(char, stats), 0, 512, st));
  CUDA_TRY(cudaMemsetAsync(WS(uint32_t, tail_off), 0, (S + 2) * 4, st));
  if (S > 0) {
    if (ascending)
      tail_count_asc_kernel<<<(unsigned)ceil_div(S, 256), 256, 0, st>>>(ev->session_offsets, ev->type, S, spec->event_type_mask,
                                                                         spec->tail_n, WS(uint32_t, tail_off));
    else
      tail_count_kernel<<<(unsigned)ceil_div(S, 256), 256, 0, st>>>(ev->session_offsets, ev->type, S, spec->event_type_mask,
                                                                     spec->tail_n, WS(uint32_t, tail_off));
    LAUNCH_CHECK();
  }
  if ((rc = exclusive_scan<uint32_t, uint32_t>(WS(uint32_t, tail_off), S, WS(uint32_t, tail_off), WS(uint32_t, scan), st)))
    return rc;
  EventLimits lim;
  lim.n_aids = (uint32_t)spec->n_aids;
  const bool time_mode = spec->weight_mode == OTTO_WEIGHT_TIME;
  lim.ts_lo = time_mode ? spec->ts_min : INT32_MIN;
  lim.ts_hi = time_mode ? spec->ts_max : INT32_MAX;
  if (S > 0 && (spec->event_type_mask & 7u) == 7u) {
    auto kern = ascending ? tail_copy_all_asc_kernel : tail_copy_all_kernel;
    kern<<<(unsigned)ceil_div(S, 256), 256, 0, st>>>(ev->session_offsets, ev->aid, ev->ts, ev->type, S, WS(uint32_t, tail_off),
                                                     WS(uint32_t, tail_aw), WS(int32_t, tail_ts), lim, WS(unsigned long long, stats));
    LAUNCH_CHECK();
  } else if (S > 0) {
    auto kern = ascending ? tail_copy_filtered_asc_kernel : tail_copy_filtered_kernel;
    kern<<<(unsigned)ceil_div(S, 256), 256, 0, st>>>(ev->session_offsets, ev->aid, ev->ts, ev->type, S, spec->event_type_mask,
                                                     WS(uint32_t, tail_off), WS(uint32_t, tail_aw), WS(int32_t, tail_ts), lim,
                                                     WS(unsigned long long, stats));
    LAUNCH_CHECK();
  }
  if (S > 0) {
    PairGenParams p = make_pairgen(L, spec, workspace);
    const int64_t warps = ceil_div(S, 32);
    pairgen_kernel<0><<<(unsigned)ceil_div(warps, PAIRGEN_WARPS), PAIRGEN_WARPS * 32, 0, st>>>(p);
    LAUNCH_CHECK();
  }
  // the copy that a multi-GPU host all-reduces; the rank's own counts stay in row_count
  CUDA_TRY(cudaMemcpyAsync(WS(uint32_t, row_total), WS(uint32_t, row_count), (L.A + 1) * 4, cudaMemcpyDeviceToDevice, st));
  return OTTO_OK;
}

// The owner-table tiers take bins of up to 6144 records, where nothing can overflow.  Larger bins (split_ub raised,
// or a row beyond 4096 sub-bins) go to the hash-table kernel, whose entries hold a 24-bit count and, in time mode, a
// 40-bit sum of ts_x - ts_min: refuse layouts in which an entry could exceed either.
static int check_max_bin(const OttoCovisitSpec* spec, unsigned long long max_bin, int64_t n_sessions) {
  // a session contributes a pair (aid_x, aid_y) at most once (in-session dedupe), so an entry counts at most the sessions of
  // all ranks (bounded by their events when only those are known)
  const unsigned long long sessions = spec->global_events > 0 ? (unsigned long long)spec->global_events : (unsigned long long)n_sessions;
  if (max_bin > sessions) max_bin = sessions;
  if (max_bin >= (1ull << 24)) {
    otto_set_error("a bin may hold %llu records; the accumulators count to 2^24 per (aid_x, aid_y): lower split_ub", max_bin);
    return OTTO_EINVAL;
  }
  if (spec->weight_mode == OTTO_WEIGHT_TIME && max_bin * (unsigned long long)((int64_t)spec->ts_max - spec->ts_min) >= (1ull << 40)) {
    otto_set_error("a bin may hold %llu records over a time range of %lld s; the time sums hold 40 bits: lower split_ub", max_bin,
                   (long long)((int64_t)spec->ts_max - spec->ts_min));
    return OTTO_EINVAL;
  }
  return OTTO_OK;
}

extern "C" int otto_covisit_count_begin(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
  return count_begin_impl(ev, spec, workspace, workspace_bytes, stream, false);
}

extern "C" int otto_covisit_count_begin_asc(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                                            int64_t workspace_bytes, void* stream) {
  return count_begin_impl(ev, spec, workspace, workspace_bytes, stream, true);
}

static int bad_events_error(const OttoCovisitSpec* spec, unsigned long long n) {
  if (spec->weight_mode == OTTO_WEIGHT_TIME)
    otto_set_error("%llu tail events have an aid outside [0, %d), a type above 2 or a ts outside [ts_min, ts_max] = [%d, %d] "
                   "(seconds expected; the pickles carry milliseconds)", n, spec->n_aids, spec->ts_min, spec->ts_max);
  else
    otto_set_error("%llu tail events have an aid outside [0, %d) or a type above 2", n, spec->n_aids);
  return OTTO_EINVAL;
}

// bins from the (all-reduced) row totals, record offsets of this rank's rows, scatter cursors
extern "C" int otto_covisit_count_finish(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                                         int64_t workspace_bytes, OttoBuildStats* stats_host, void* stream) {
  int rc = check_spec(spec);
  if (rc) return rc;
  const Layout L = make_layout(ev->n_sessions, ev->n_events, spec);
  if ((rc = check_ws(L, workspace, workspace_bytes))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t A = L.A;
  CUDA_TRY(cudaMemsetAsync(WS(uint32_t, bin_base), 0, (A + 2) * 4, st));
  CUDA_TRY(cudaMemsetAsync(WS(unsigned long long, stats) + 3, 0, 8, st));
  CUDA_TRY(cudaMemsetAsync(WS(unsigned long long, stats) + STAT_MAX_BIN, 0, 8, st));
  bins_count_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, st>>>(
      WS(uint32_t, row_total), WS(uint32_t, row_count), A, (uint32_t)effective_split_ub(spec), (uint32_t)sub_bin_target(spec),
      WS(uint32_t, bin_base), WS(unsigned long long, hot_off), WS(uint32_t, hot_rows), L.Hmax, WS(unsigned long long, stats));
  LAUNCH_CHECK();
  if ((rc = exclusive_scan<uint32_t, uint32_t>(WS(uint32_t, bin_base), A, WS(uint32_t, bin_base), WS(uint32_t, scan), st)))
    return rc;
  if ((rc = exclusive_scan<uint32_t, unsigned long long>(WS(uint32_t, row_count), A, WS(unsigned long long, row_off),
                                                         WS(unsigned long long, scan), st)))
    return rc;
  if ((rc = exclusive_scan<unsigned long long, unsigned long long>(WS(unsigned long long, hot_off), A, WS(unsigned long long, hot_off),
                                                                   WS(unsigned long long, scan), st)))
    return rc;
  count_stats_kernel<<<1, 1, 0, st>>>(WS(uint32_t, tail_off), L.S, WS(uint32_t, bin_base), A, WS(unsigned long long, row_off),
                                      WS(unsigned long long, hot_off), WS(unsigned long long, stats));
  LAUNCH_CHECK();
  unsigned long long h[8];
  CUDA_TRY(cudaMemcpyAsync(h, WS(char, stats), sizeof(h), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (h[STAT_BAD_EVENTS]) return bad_events_error(spec, h[STAT_BAD_EVENTS]);
  if ((rc = check_max_bin(spec, h[STAT_MAX_BIN], L.S))) return rc;
  // the bin arrays were sized from the event count: refuse before anything is written past them
  if ((int64_t)h[2] > L.Bmax || (int64_t)h[3] > L.Hmax) {
    otto_set_error("%llu bins / %llu hot rows exceed the %lld / %lld the workspace was sized for: set spec.global_events to the "
                   "event count of all ranks", h[2], h[3], (long long)L.Bmax, (long long)L.Hmax);
    return OTTO_ENOSPC;
  }
  if (h[1] + h[4] >= (1ull << 32)) {
    otto_set_error("a rank is limited to 2^32 - 1 pair records including its staged hot rows (32 GiB); shard the sessions");
    return OTTO_EINVAL;
  }
  bins_fill_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, st>>>(WS(uint32_t, bin_base), A, WS(uint32_t, bin_x));
  LAUNCH_CHECK();
  init_cursor_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, st>>>(WS(unsigned long long, row_off), WS(unsigned long long, hot_off),
                                                                 WS(uint32_t, bin_base), A, WS(uint32_t, cursor));
  LAUNCH_CHECK();
  if (stats_host) {
    stats_host->tail_events = (int64_t)h[0];
    stats_host->pairs = (int64_t)h[1];
    stats_host->bins = (int64_t)h[2];
    stats_host->split_rows = (int64_t)h[3];
    stats_host->hot_pairs = (int64_t)h[4];
  }
  return OTTO_OK;
}

extern "C" int otto_covisit_count(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                                  OttoBuildStats* stats_host, void* stream) {
  int rc = otto_covisit_count_begin(ev, spec, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return otto_covisit_count_finish(ev, spec, workspace, workspace_bytes, stats_host, stream);
}

static int check_plan(const OttoOwnerPlan* plan, const OttoCovisitSpec* spec, bool need_ptrs) {
  if (!plan) { otto_set_error("owner plan is NULL"); return OTTO_EINVAL; }
  if (plan->n_owners < 1 || plan->n_owners > OTTO_MAX_OWNERS || plan->rank < 0 || plan->rank >= plan->n_owners) {
    otto_set_error("owner plan: n_owners must be in [1, %d] and rank below it", OTTO_MAX_OWNERS);
    return OTTO_EINVAL;
  }
  if (plan->aid_cuts[0] != 0 || plan->aid_cuts[plan->n_owners] != spec->n_aids) {
    otto_set_error("owner plan: aid_cuts must run from 0 to n_aids");
    return OTTO_EINVAL;
  }
  for (int o = 0; o < plan->n_owners; ++o) {
    if (plan->aid_cuts[o + 1] < plan->aid_cuts[o]) { otto_set_error("owner plan: aid_cuts must be non-decreasing"); return OTTO_EINVAL; }
    if (need_ptrs && !plan->owner_records[o]) { otto_set_error("owner plan: owner_records[%d] is NULL", o); return OTTO_EINVAL; }
  }
  return OTTO_OK;
}

static OwnerCuts make_cuts(const OttoOwnerPlan* plan) {
  OwnerCuts c;
  c.n = plan->n_owners;
  c.rank = plan->rank;
  for (int o = 0; o <= OTTO_MAX_OWNERS; ++o) c.cut[o] = (uint32_t)plan->aid_cuts[o <= plan->n_owners ? o : plan->n_owners];
  return c;
}

// bins from the row totals (workspace row_total = sum over ranks), the layout of MY rows, and scatter cursors that
// point into the owners' buffers
extern "C" int otto_covisit_count_finish_owned(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                                               int64_t workspace_bytes, const OttoOwnerPlan* plan,
                                               const uint32_t* row_before, OttoBuildStats* stats_host, void* stream) {
  int rc = check_spec(spec);
  if (rc) return rc;
  if ((rc = check_plan(plan, spec, false))) return rc;
  if (!ev) { otto_set_error("events is NULL"); return OTTO_EINVAL; }
  if (!row_before) { otto_set_error("row_before is NULL"); return OTTO_EINVAL; }
  const Layout L = make_layout(ev->n_sessions, ev->n_events, spec);
  if ((rc = check_ws(L, workspace, workspace_bytes))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t A = L.A;
  const OwnerCuts cuts = make_cuts(plan);
  unsigned long long* stats = WS(unsigned long long, stats);
  CUDA_TRY(cudaMemsetAsync(WS(uint32_t, bin_base), 0, (A + 2) * 4, st));
  CUDA_TRY(cudaMemsetAsync(stats + 3, 0, 8, st));
  CUDA_TRY(cudaMemsetAsync(stats + STAT_MAX_BIN, 0, 8, st));
  // hot_off := totals of the hot rows (the "own count" input is the total here)
  bins_count_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, st>>>(
      WS(uint32_t, row_total), WS(uint32_t, row_total), A, (uint32_t)effective_split_ub(spec), (uint32_t)sub_bin_target(spec),
      WS(uint32_t, bin_base), WS(unsigned long long, hot_off), WS(uint32_t, hot_rows), L.Hmax, stats);
  LAUNCH_CHECK();
  if ((rc = exclusive_scan<uint32_t, uint32_t>(WS(uint32_t, bin_base), A, WS(uint32_t, bin_base), WS(uint32_t, scan), st)))
    return rc;
  if ((rc = exclusive_scan<uint32_t, unsigned long long>(WS(uint32_t, row_total), A, WS(unsigned long long, row_off),
                                                         WS(unsigned long long, scan), st)))
    return rc;
  if ((rc = exclusive_scan<unsigned long long, unsigned long long>(WS(unsigned long long, hot_off), A, WS(unsigned long long, hot_off),
                                                                   WS(unsigned long long, scan), st)))
    return rc;
  owner_cut_values_kernel<<<1, 32, 0, st>>>(cuts, WS(unsigned long long, row_off), WS(unsigned long long, hot_off),
                                            WS(uint32_t, bin_base), stats);
  LAUNCH_CHECK();
  init_cursor_owned_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, st>>>(
      cuts, WS(uint32_t, row_total), row_before, WS(uint32_t, bin_base), A, WS(unsigned long long, row_off),
      WS(unsigned long long, hot_off), stats, WS(uint32_t, cursor), WS(uint32_t, row_count));
  LAUNCH_CHECK();
  // the local layout (my rows only) - from here on identical to a single-GPU count_finish
  if ((rc = exclusive_scan<uint32_t, unsigned long long>(WS(uint32_t, row_count), A, WS(unsigned long long, row_off),
                                                         WS(unsigned long long, scan), st)))
    return rc;
  if ((rc = exclusive_scan<unsigned long long, unsigned long long>(WS(unsigned long long, hot_off), A, WS(unsigned long long, hot_off),
                                                                   WS(unsigned long long, scan), st)))
    return rc;
  count_stats_kernel<<<1, 1, 0, st>>>(WS(uint32_t, tail_off), L.S, WS(uint32_t, bin_base), A, WS(unsigned long long, row_off),
                                      WS(unsigned long long, hot_off), stats);
  LAUNCH_CHECK();
  unsigned long long h[STAT_WORDS];
  CUDA_TRY(cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (h[STAT_BAD_EVENTS]) return bad_events_error(spec, h[STAT_BAD_EVENTS]);
  if ((rc = check_max_bin(spec, h[STAT_MAX_BIN], L.S))) return rc;
  if ((int64_t)h[2] > L.Bmax || (int64_t)h[3] > L.Hmax) {
    otto_set_error("%llu bins / %llu hot rows exceed the %lld / %lld the workspace was sized for: set spec.global_events to the "
                   "event count of all ranks", h[2], h[3], (long long)L.Bmax, (long long)L.Hmax);
    return OTTO_ENOSPC;
  }
  unsigned long long need_max = 0;
  for (int o = 0; o < plan->n_owners; ++o) {      // the same verdict on every rank
    const unsigned long long need = (h[STAT_CUT_E + o + 1] - h[STAT_CUT_E + o]) + (h[STAT_CUT_H + o + 1] - h[STAT_CUT_H + o]);
    if (need > need_max) need_max = need;
    if (need >= (1ull << 32)) {
      otto_set_error("owner %d would hold %llu pair records including its staged hot rows; the limit is 2^32 - 1 (32 GiB): use more owners", o, need);
      return OTTO_EINVAL;
    }
  }
  bins_fill_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, st>>>(WS(uint32_t, bin_base), A, WS(uint32_t, bin_x));
  LAUNCH_CHECK();
  if (stats_host) {
    stats_host->tail_events = (int64_t)h[0];
    stats_host->pairs = (int64_t)h[1];
    stats_host->bins = (int64_t)h[2];
    stats_host->split_rows = (int64_t)h[3];
    stats_host->hot_pairs = (int64_t)h[4];
    stats_host->owner_records_max = (int64_t)need_max;
    for (int o = 0; o <= OTTO_MAX_OWNERS; ++o)
      stats_host->owner_bin_cuts[o] = (int64_t)h[STAT_CUT_B + (o <= plan->n_owners ? o : plan->n_owners)];
  }
  return OTTO_OK;
}

// ---- owner plan on the device: totals, counts of lower ranks and balanced aid cuts from the all-gathered row counts ----
__global__ void plan_rows_kernel(const uint32_t* __restrict__ counts, int G, int rank, int64_t A, uint32_t* __restrict__ row_total,
                                 uint32_t* __restrict__ row_before, unsigned long long* __restrict__ total64, unsigned long long* flags) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= A) return;
  unsigned long long tot = 0, before = 0;
  for (int g = 0; g < G; ++g) {
    const unsigned long long c = counts[(int64_t)g * A + x];
    tot += c;
    if (g < rank) before += c;
  }
  if (tot >= (1ull << 32)) atomicOr(flags, 1ull);
  row_total[x] = (uint32_t)tot;
  row_before[x] = (uint32_t)before;
  total64[x] = tot;
}

// cut g = first row x with (pairs in rows < x) >= total * g / G: contiguous ranges with (nearly) equal pair counts
__global__ void plan_cuts_kernel(const unsigned long long* __restrict__ prefix, int64_t A, int G, int32_t* __restrict__ cuts) {
  const int g = threadIdx.x;
  if (g > G) return;
  int64_t cut = g == 0 ? 0 : A;
  if (g > 0 && g < G) {
    const unsigned long long total = prefix[A];
    const unsigned long long target = (unsigned long long)(((unsigned __int128)total * (unsigned)g) / (unsigned)G);
    int64_t lo = 0, hi = A;     // first x in [0, A] with prefix[x] >= target
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (prefix[mid] >= target) hi = mid;
      else lo = mid + 1;
    }
    cut = lo;
  }
  cuts[g] = (int32_t)cut;
}

extern "C" int64_t otto_covisit_plan_scratch_bytes(int32_t n_aids) {
  return align_up(((int64_t)n_aids + 2) * 8, 256) + scan_scratch_elems((int64_t)n_aids + 1) * 8 + 512;
}

extern "C" int otto_covisit_plan_owners(const uint32_t* counts_all, int32_t n_owners, int32_t rank, int32_t n_aids,
                                        uint32_t* row_total, uint32_t* row_before, void* scratch, int64_t scratch_bytes,
                                        int32_t* aid_cuts_host, void* stream) {
  if (!counts_all || !row_total || !row_before || !aid_cuts_host || n_aids <= 0 || n_owners < 1 || n_owners > OTTO_MAX_OWNERS ||
      rank < 0 || rank >= n_owners) {
    otto_set_error("bad argument");
    return OTTO_EINVAL;
  }
  if (!scratch || scratch_bytes < otto_covisit_plan_scratch_bytes(n_aids)) { otto_set_error("plan scratch too small"); return OTTO_ENOSPC; }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t A = n_aids;
  char* sc = (char*)scratch;
  unsigned long long* prefix = (unsigned long long*)sc;
  unsigned long long* scan_sc = (unsigned long long*)(sc + align_up((A + 2) * 8, 256));
  unsigned long long* flags = scan_sc + scan_scratch_elems(A + 1);
  int32_t* cuts_dev = (int32_t*)(flags + 1);
  CUDA_TRY(cudaMemsetAsync(flags, 0, 8, st));
  plan_rows_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, st>>>(counts_all, n_owners, rank, A, row_total, row_before, prefix, flags);
  LAUNCH_CHECK();
  int rc = exclusive_scan<unsigned long long, unsigned long long>(prefix, A, prefix, scan_sc, st);
  if (rc) return rc;
  plan_cuts_kernel<<<1, 32, 0, st>>>(prefix, A, n_owners, cuts_dev);
  LAUNCH_CHECK();
  struct { unsigned long long flag; int32_t cuts[OTTO_MAX_OWNERS + 2]; } h;
  static_assert(sizeof(h) == 8 + 4 * (OTTO_MAX_OWNERS + 2), "layout");
  CUDA_TRY(cudaMemcpyAsync(&h, flags, 8 + 4 * (n_owners + 1), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (h.flag) { otto_set_error("an aid_x row holds 2^32 or more pairs"); return OTTO_EINVAL; }
  for (int o = 0; o <= OTTO_MAX_OWNERS; ++o) aid_cuts_host[o] = h.cuts[o <= n_owners ? o : n_owners];
  for (int o = 1; o <= OTTO_MAX_OWNERS; ++o)
    if (aid_cuts_host[o] < aid_cuts_host[o - 1]) aid_cuts_host[o] = aid_cuts_host[o - 1];
  return OTTO_OK;
}

extern "C" int otto_covisit_views(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                                  int64_t workspace_bytes, uint64_t** bin_offsets, uint32_t** bin_base,
                                  uint32_t** bin_x, uint32_t** row_total) {
  int rc = check_spec(spec);
  if (rc) return rc;
  const Layout L = make_layout(ev->n_sessions, ev->n_events, spec);
  if ((rc = check_ws(L, workspace, workspace_bytes))) return rc;
  if (bin_offsets) *bin_offsets = WS(uint64_t, bin_off);
  if (bin_base) *bin_base = WS(uint32_t, bin_base);
  if (bin_x) *bin_x = WS(uint32_t, bin_x);
  if (row_total) *row_total = WS(uint32_t, row_total);
  return OTTO_OK;
}

// pair records: ordinary rows at their final positions, hot rows into the staging area.  plan == NULL: into `records`
// (single GPU, or multi-GPU with the owners reading the slabs afterwards); else into the owners' buffers.
static int scatter_records(const Layout& L, const OttoCovisitSpec* spec, void* workspace, void* records,
                           const OttoOwnerPlan* plan, cudaStream_t st) {
  if (g_profile) CUDA_TRY(cudaEventRecord(g_prof_sc[0], st));
  if (L.S > 0) {
    PairGenParams p = make_pairgen(L, spec, workspace);
    p.records = (uint2*)records;
    const int64_t warps = ceil_div(L.S, 32);
    if (plan) {
      p.n_owners = plan->n_owners;
      for (int o = 0; o <= OTTO_MAX_OWNERS; ++o) p.cuts[o] = (uint32_t)plan->aid_cuts[o <= plan->n_owners ? o : plan->n_owners];
      for (int o = 0; o < OTTO_MAX_OWNERS; ++o) p.owner_rec[o] = (uint2*)plan->owner_records[o < plan->n_owners ? o : 0];
      pairgen_kernel<2><<<(unsigned)ceil_div(warps, PAIRGEN_WARPS), PAIRGEN_WARPS * 32, 0, st>>>(p);
    } else {
      pairgen_kernel<1><<<(unsigned)ceil_div(warps, PAIRGEN_WARPS), PAIRGEN_WARPS * 32, 0, st>>>(p);
    }
    LAUNCH_CHECK();
  }
  if (g_profile) CUDA_TRY(cudaEventRecord(g_prof_sc[1], st));
  return OTTO_OK;
}

// records per bin; hot rows from the staging area into their aid_y-hash sub-bins; bin offsets
static int partition_records(const Layout& L, void* workspace, void* records, cudaStream_t st) {
  int rc;
  const int64_t A = L.A;
  CUDA_TRY(cudaMemsetAsync(WS(uint32_t, bin_cnt), 0, (L.Bmax + 1) * 4, st));
  bin_cnt_rows_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, st>>>(WS(uint32_t, row_count), WS(uint32_t, bin_base), A,
                                                                  WS(uint32_t, bin_cnt));
  LAUNCH_CHECK();
  PartParams pp;
  pp.hot_rows = WS(uint32_t, hot_rows);
  pp.n_hot = WS(unsigned long long, stats) + 3;
  pp.hot_cap = L.Hmax;
  pp.bin_base = WS(uint32_t, bin_base);
  pp.row_count = WS(uint32_t, row_count);
  pp.row_off = WS(unsigned long long, row_off);
  pp.hot_off = WS(unsigned long long, hot_off);
  pp.A = A;
  pp.bin_cnt = WS(uint32_t, bin_cnt);
  pp.sub_cur = WS(uint32_t, sub_cur);
  pp.tile_row = WS(uint32_t, tile_row);
  pp.tile_idx = WS(uint32_t, tile_idx);
  pp.n_tiles = WS(unsigned long long, stats) + 5;
  pp.tile_cap = L.Tmax;
  pp.records = (uint2*)records;
  int dev = 0, n_sm = 148;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  CUDA_TRY(cudaMemsetAsync(WS(unsigned long long, stats) + 5, 0, 8, st));
  partition_tiles_kernel<<<(unsigned)ceil_div(L.Hmax, 256), 256, 0, st>>>(pp);
  LAUNCH_CHECK();
  CUDA_TRY(cudaFuncSetAttribute(partition_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)partition_smem<false>()));
  CUDA_TRY(cudaFuncSetAttribute(partition_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)partition_smem<true>()));
  partition_kernel<false><<<n_sm * 4, PART_THREADS, partition_smem<false>(), st>>>(pp);
  LAUNCH_CHECK();
  if ((rc = exclusive_scan<uint32_t, unsigned long long>(WS(uint32_t, bin_cnt), L.Bmax, WS(unsigned long long, bin_off),
                                                         WS(unsigned long long, scan), st)))
    return rc;
  init_sub_cur_kernel<<<592, 256, 0, st>>>(WS(unsigned long long, bin_off), WS(uint32_t, bin_base), A, WS(uint32_t, sub_cur));
  LAUNCH_CHECK();
  if (g_profile) CUDA_TRY(cudaEventRecord(g_prof_sc[2], st));
  partition_kernel<true><<<n_sm * 3, PART_THREADS, partition_smem<true>(), st>>>(pp);
  LAUNCH_CHECK();
  if (g_profile) {
    CUDA_TRY(cudaEventRecord(g_prof_sc[3], st));
    g_prof_sc_valid = true;
  }
  return OTTO_OK;
}

static int scatter_args(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes, Layout* L) {
  int rc = check_spec(spec);
  if (rc) return rc;
  if (!ev) { otto_set_error("events is NULL"); return OTTO_EINVAL; }
  *L = make_layout(ev->n_sessions, ev->n_events, spec);
  return check_ws(*L, workspace, workspace_bytes);
}

static int check_records(void* records, int64_t records_capacity) {
  if (!records && records_capacity > 0) { otto_set_error("records is NULL"); return OTTO_EINVAL; }
  if (records_capacity >= (1ll << 32)) { otto_set_error("a rank is limited to 2^32 - 1 pair records (32 GiB); shard the sessions"); return OTTO_EINVAL; }
  return OTTO_OK;
}

extern "C" int otto_covisit_scatter(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                                    int64_t workspace_bytes, void* records, int64_t records_capacity, void* stream) {
  Layout L;
  int rc = scatter_args(ev, spec, workspace, workspace_bytes, &L);
  if (rc) return rc;
  if ((rc = check_records(records, records_capacity))) return rc;
  if ((rc = scatter_records(L, spec, workspace, records, nullptr, (cudaStream_t)stream))) return rc;
  return partition_records(L, workspace, records, (cudaStream_t)stream);
}

extern "C" int otto_covisit_scatter_owned(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                                          int64_t workspace_bytes, const OttoOwnerPlan* plan, void* stream) {
  Layout L;
  int rc = scatter_args(ev, spec, workspace, workspace_bytes, &L);
  if (rc) return rc;
  if ((rc = check_plan(plan, spec, true))) return rc;
  return scatter_records(L, spec, workspace, nullptr, plan, (cudaStream_t)stream);
}

extern "C" int otto_covisit_partition(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                                      int64_t workspace_bytes, void* records, int64_t records_capacity, void* stream) {
  Layout L;
  int rc = scatter_args(ev, spec, workspace, workspace_bytes, &L);
  if (rc) return rc;
  if ((rc = check_records(records, records_capacity))) return rc;
  return partition_records(L, workspace, records, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ staged scatter (sender-side combining)
//
// The direct scatter appends ~63-byte runs at 1.86 M row frontiers: on one GPU the partial sectors cost DRAM fills
// (profiles/r02_experiments.md), across NVLink every run is a small store packet (330 GB/s per direction against
// 550-590 GB/s for large transfers).  Staged form: pass A appends the same runs to COARSE buckets of this rank's own
// staging buffer (a bucket = 2^STAGE_LOGA consecutive aid_x rows, ~900 write frontiers for OTTO: L2 combines the runs
// into full lines), each record carrying its row relative to the bucket; pass B runs at the OWNER of the rows: it
// streams the segments of its buckets out of every sender's staging buffer (large contiguous reads - over NVLink for
// the peers' buffers) and places the records at their final position (ordinary rows) or in the staging area of the
// hot rows, with one cursor atomic per run.  Everything behind it (hot-row partition, reduce) is unchanged.
constexpr int PLACE_THREADS = 256, PLACE_PER_THREAD = 4, PLACE_TILE = PLACE_THREADS * PLACE_PER_THREAD;
// info words: [2] tiles of the place pass, [STAGE_INFO_TOTAL + g] records rank g stages
constexpr int STAGE_INFO_TOTAL = 8, STAGE_INFO_WORDS = STAGE_INFO_TOTAL + OTTO_MAX_OWNERS;

struct StageLayout {
  int64_t A, G, NB, Tmax;
  int64_t bcnt, bo, bcur, tiles, info, total;
};

static StageLayout stage_layout(const OttoCovisitSpec* spec, int64_t n_sessions, int64_t n_events, int n_ranks) {
  StageLayout S;
  S.A = spec->n_aids;
  S.G = n_ranks;
  S.NB = ceil_div(S.A, (int64_t)1 << STAGE_LOGA);
  int64_t Ecap = n_events < n_sessions * spec->tail_n ? n_events : n_sessions * spec->tail_n;
  const int64_t Eg = spec->global_events > n_events ? spec->global_events : (Ecap > 1 ? Ecap : 1);
  const int64_t Pmax = Eg * (spec->tail_n - 1);                  // every tail event pairs with at most tail_n - 1 others
  S.Tmax = Pmax / PLACE_TILE + (int64_t)n_ranks * S.NB + 1;      // tiles of a place pass <= sum over (rank, bucket) of ceil(n / tile)
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t at = o; o = align_up(o + bytes, 256); return at; };
  S.bcnt = take(S.G * (S.NB + 1) * 8);
  S.bo = take(S.G * (S.NB + 1) * 8);
  S.bcur = take((S.NB + 1) * 4 * STAGE_CUR_STRIDE);
  S.tiles = take(S.Tmax * 16);
  S.info = take(STAGE_INFO_WORDS * 8);
  S.total = o;
  return S;
}
#define SP(type, field) ((type*)((char*)plan + S.field))

// bcnt[g][b] := records of bucket b that rank g stages = sum of its counts over the bucket's rows; one block per bucket
__global__ void __launch_bounds__(256) stage_count_kernel(const uint32_t* __restrict__ counts, int G, int64_t A, int64_t nb_stride,
                                                            unsigned long long* __restrict__ bo) {
  const int64_t b = blockIdx.x;
  const int64_t lo = b << STAGE_LOGA, hi = min(A, (b + 1) << STAGE_LOGA);
  __shared__ unsigned long long s_part[8];
  for (int g = 0; g < G; ++g) {
    unsigned long long sum = 0;
    for (int64_t x = lo + threadIdx.x; x < hi; x += 256) sum += counts[(int64_t)g * A + x];
    for (int o = 16; o > 0; o >>= 1) sum += shfl_u64(sum, lane_id() ^ o);
    if (lane_id() == 0) s_part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
      for (int w = 0; w < 8; ++w) t += s_part[w];
      bo[(int64_t)g * nb_stride + b] = t;
    }
    __syncthreads();
  }
}

// bo[g][b] := where rank g's segment of bucket b starts in its staging buffer: exclusive scan of the counts rounded up
// to even (a segment starts on a 16-byte boundary, which the bulk copies of the place pass need); one block per rank;
// bo[g][NB] and info[STAGE_INFO_TOTAL + g] = the size of the rank's staging buffer; the staging cursors of `rank`
// start at its offsets
__global__ void __launch_bounds__(1024) stage_scan_kernel(const unsigned long long* __restrict__ bcnt, unsigned long long* __restrict__ bo,
                                                           int64_t nb, int rank, uint32_t* __restrict__ bcur, unsigned long long* info) {
  __shared__ unsigned long long s_carry, s_warp[32];
  const int g = blockIdx.x;
  const unsigned long long* cnt = bcnt + (int64_t)g * (nb + 1);
  unsigned long long* row = bo + (int64_t)g * (nb + 1);
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 <= nb; b0 += 1024) {
    const int64_t b = b0 + threadIdx.x;
    const unsigned long long v = b < nb ? (cnt[b] + 1ull) & ~1ull : 0ull;
    unsigned long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = shfl_u64(inc, (int)lane_id() - o >= 0 ? (int)lane_id() - o : (int)lane_id());
      if ((int)lane_id() >= o) inc += up;
    }
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    unsigned long long base = s_carry;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) base += s_warp[w];
    const unsigned long long excl = base + inc - v;
    if (b <= nb) row[b] = excl;
    if (b < nb && g == rank) bcur[b * STAGE_CUR_STRIDE] = (uint32_t)excl;
    if (b == nb) info[STAGE_INFO_TOTAL + g] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = base + inc;
    __syncthreads();
  }
}

extern "C" int64_t otto_covisit_stage_plan_bytes(const OttoCovisitSpec* spec, int64_t n_sessions, int64_t n_events, int32_t n_ranks) {
  if (check_spec(spec) || n_ranks < 1 || n_ranks > OTTO_MAX_OWNERS) return -1;
  return stage_layout(spec, n_sessions, n_events, n_ranks).total;
}

// record packing: aid_y in y_bits, v in v_bits, the row inside its bucket in STAGE_LOGA bits
static int stage_bits(const OttoCovisitSpec* spec, uint32_t* y_bits, uint32_t* v_bits) {
  uint32_t yb = 1;
  while (yb < 31 && (1ll << yb) < (int64_t)spec->n_aids) ++yb;
  unsigned long long vmax = 1;
  if (spec->weight_mode == OTTO_WEIGHT_TIME) vmax = (unsigned long long)((int64_t)spec->ts_max - spec->ts_min);
  if (spec->weight_mode == OTTO_WEIGHT_TYPE)
    for (int i = 0; i < 3; ++i)
      if ((unsigned long long)spec->type_weight[i] > vmax) vmax = (unsigned long long)spec->type_weight[i];
  uint32_t vb = 1;
  while (vb < 32 && (1ull << vb) <= vmax) ++vb;
  *y_bits = yb;
  *v_bits = vb;
  if (yb + vb + STAGE_LOGA > 64) {
    otto_set_error("a staged record cannot carry %u + %u + %d bits: use the direct scatter", yb, vb, STAGE_LOGA);
    return OTTO_EOVERFLOW;
  }
  return OTTO_OK;
}

// staged_records_host[g] := records rank g stages (the size of its staging buffer), from a plan that
// otto_covisit_stage_plan has been enqueued for on `stream`.  Synchronises (cheap behind a synchronising call such as
// otto_covisit_count_finish_owned).
extern "C" int otto_covisit_stage_totals(const OttoCovisitSpec* spec, int64_t n_sessions, int64_t n_events, void* plan, int32_t n_ranks,
                                         int64_t* staged_records_host, void* stream) {
  int rc = check_spec(spec);
  if (rc) return rc;
  if (!plan || !staged_records_host || n_ranks < 1 || n_ranks > OTTO_MAX_OWNERS) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  const StageLayout S = stage_layout(spec, n_sessions, n_events, n_ranks);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long h[STAGE_INFO_WORDS];
  CUDA_TRY(cudaMemcpyAsync(h, SP(char, info), sizeof(h), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int g = 0; g < n_ranks; ++g) {
    const unsigned long long total = h[STAGE_INFO_TOTAL + g];
    if (total >= (1ull << 32)) { otto_set_error("rank %d would stage %llu records; the limit is 2^32 - 1", g, total); return OTTO_EINVAL; }
    staged_records_host[g] = (int64_t)total;
  }
  return OTTO_OK;
}

// Plan of the staged scatter.  counts_all: [n_ranks][n_aids] pairs per row of every rank (one rank: NULL = the
// workspace's own row counts).  staged_records_host[g] = records rank g will stage (every rank computes all of them,
// so the staging buffers can be sized without a collective; NULL = do not synchronise, read them later with
// otto_covisit_stage_totals); OTTO_EOVERFLOW when the packed record cannot carry the row (fall back to the direct
// scatter).  Needs nothing of the workspace but the row counts, so it can be enqueued before otto_covisit_count_finish_owned.
extern "C" int otto_covisit_stage_plan(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                                       const uint32_t* counts_all, int32_t n_ranks, int32_t rank, void* plan, int64_t plan_bytes,
                                       int64_t* staged_records_host, void* stream) {
  int rc = check_spec(spec);
  if (rc) return rc;
  if (!ev || n_ranks < 1 || n_ranks > OTTO_MAX_OWNERS || rank < 0 || rank >= n_ranks) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  const Layout L = make_layout(ev->n_sessions, ev->n_events, spec);
  if ((rc = check_ws(L, workspace, workspace_bytes))) return rc;
  const StageLayout S = stage_layout(spec, ev->n_sessions, ev->n_events, n_ranks);
  if (!plan || plan_bytes < S.total) { otto_set_error("stage plan scratch too small: need %lld bytes", (long long)S.total); return OTTO_ENOSPC; }
  uint32_t yb, vb;
  if ((rc = stage_bits(spec, &yb, &vb))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!counts_all) {
    if (n_ranks != 1) { otto_set_error("counts_all is NULL"); return OTTO_EINVAL; }
    counts_all = WS(uint32_t, row_count);
  }
  stage_count_kernel<<<(unsigned)S.NB, 256, 0, st>>>(counts_all, n_ranks, L.A, S.NB + 1, SP(unsigned long long, bcnt));
  LAUNCH_CHECK();
  stage_scan_kernel<<<n_ranks, 1024, 0, st>>>(SP(unsigned long long, bcnt), SP(unsigned long long, bo), S.NB, rank, SP(uint32_t, bcur),
                                              SP(unsigned long long, info));
  LAUNCH_CHECK();
  if (!staged_records_host) return OTTO_OK;     // asynchronous form: otto_covisit_stage_totals reads the sizes later
  return otto_covisit_stage_totals(spec, ev->n_sessions, ev->n_events, plan, n_ranks, staged_records_host, stream);
}

// pass A: this rank's pairs into the coarse buckets of its staging buffer (`staged`: at least staged_records_host[rank]
// records of 8 bytes)
extern "C" int otto_covisit_scatter_staged(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                                           void* plan, int32_t n_ranks, void* staged, void* stream) {
  Layout L;
  int rc = scatter_args(ev, spec, workspace, workspace_bytes, &L);
  if (rc) return rc;
  if (!plan || !staged || n_ranks < 1 || n_ranks > OTTO_MAX_OWNERS) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  const StageLayout S = stage_layout(spec, ev->n_sessions, ev->n_events, n_ranks);
  cudaStream_t st = (cudaStream_t)stream;
  if (L.S > 0) {
    PairGenParams p = make_pairgen(L, spec, workspace);
    p.records = (uint2*)staged;
    p.bcur = SP(uint32_t, bcur);
    if ((rc = stage_bits(spec, &p.y_bits, &p.v_bits))) return rc;
    const int64_t warps = ceil_div(L.S, 32);
    pairgen_kernel<3><<<(unsigned)ceil_div(warps, PAIRGEN_WARPS), PAIRGEN_WARPS * 32, 0, st>>>(p);
    LAUNCH_CHECK();
  }
  return OTTO_OK;
}

static int occupancy_blocks(const void* kernel, int threads, size_t smem);
template <typename K>
static int set_smem(K kernel, size_t bytes);

struct StagedPtrs {
  const uint2* rank[OTTO_MAX_OWNERS];
};

struct PlaceParams {
  const uint4* tiles;                // {address of the tile's first record in its sender's buffer (lo, hi), records, bucket}
  const unsigned long long* info;    // [2] tiles
  int64_t tile_cap;
  uint32_t aid_lo, aid_hi;
  uint32_t y_bits, v_bits;
  uint32_t* cursor;
  uint2* records;
};

// one thread per bucket of my row range: the tiles of its G segments, appended to the tile list as ONE run, so that the
// blocks of the place pass, which walk the list in order, write into a few buckets' worth of rows at a time (L2 then
// combines the short runs into full lines)
__global__ void place_tiles_kernel(const StagedPtrs staged, int G, int64_t b_lo, int64_t b_hi, const unsigned long long* __restrict__ bcnt,
                                   const unsigned long long* __restrict__ bo, int64_t nb_stride, uint4* __restrict__ tiles,
                                   unsigned long long* info, int64_t tile_cap) {
  const int64_t b = b_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= b_hi) return;
  unsigned long long nt = 0;
  for (int g = 0; g < G; ++g) nt += (bcnt[(int64_t)g * nb_stride + b] + PLACE_TILE - 1) / PLACE_TILE;
  if (nt == 0) return;
  unsigned long long at = atomicAdd(&info[2], nt);
  for (int g = 0; g < G; ++g) {
    const unsigned long long beg = bo[(int64_t)g * nb_stride + b], n = bcnt[(int64_t)g * nb_stride + b];
    for (unsigned long long i = 0; i < n && (int64_t)at < tile_cap; i += PLACE_TILE, ++at) {
      const uint32_t m = (uint32_t)min((unsigned long long)PLACE_TILE, n - i);
      const unsigned long long src = (unsigned long long)(staged.rank[g] + beg + i);       // 16-byte aligned: beg is even
      tiles[at] = make_uint4((uint32_t)src, (uint32_t)(src >> 32), m, (uint32_t)b);
    }
  }
}

// Persistent blocks over the tile list.  A tile (1024 records, 8 KB) reaches shared memory by ONE bulk copy
// (cp.async.bulk, completion on an mbarrier) issued PLACE_STAGES - 1 tiles ahead by thread 0: the pass is bound by the
// latency of its loads - a few microseconds across NVLink - so the bytes in flight per SM decide its rate, and a ring
// of bulk copies keeps them in flight without holding registers or issue slots.  Per tile: records to registers, the
// stage goes back to the producer, then one cursor atomic per run of equal rows in the warp (the records of one
// (session, row) were staged as one run), all atomics of the tile issued before the first result is needed.
constexpr int PLACE_STAGES = 4;
constexpr uint32_t PLACE_BUF_BYTES = PLACE_TILE * 8;
constexpr size_t place_smem_bytes() { return (size_t)PLACE_STAGES * PLACE_BUF_BYTES; }

__global__ void __launch_bounds__(PLACE_THREADS) place_kernel(const PlaceParams p) {
  extern __shared__ __align__(128) unsigned char place_smem[];
  __shared__ __align__(8) unsigned long long s_bar[PLACE_STAGES];
  __shared__ uint2 s_meta[PLACE_STAGES];                       // {records of the tile, its bucket}
  const uint32_t buf_s = (uint32_t)__cvta_generic_to_shared(place_smem);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(s_bar);
  unsigned long long n_tiles = p.info[2];
  if ((int64_t)n_tiles > p.tile_cap) n_tiles = (unsigned long long)p.tile_cap;
  const uint32_t lane = lane_id();
  const unsigned long long ymask = (1ull << p.y_bits) - 1ull, vmask = (1ull << p.v_bits) - 1ull;
  const unsigned long long stride = gridDim.x;
  const uint4 none = make_uint4(0u, 0u, 0u, 0u);
  uint4 d_pref = none;                                          // thread 0: descriptor of the next tile to issue
  auto issue = [&](const uint4 d, int st) {
    s_meta[st] = make_uint2(d.z, d.w);
    const uint32_t bytes = ((d.z + 1u) & ~1u) * 8u;            // an odd tile ends its segment: the pad record is there
    mbar_expect_tx(bar_s + st * 8, bytes);
    bulk_g2s(buf_s + st * PLACE_BUF_BYTES, (const void*)(((unsigned long long)d.y << 32) | d.x), bytes, bar_s + st * 8);
  };
  if (threadIdx.x == 0) {
    for (int i = 0; i < PLACE_STAGES; ++i) mbar_init(bar_s + i * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < PLACE_STAGES; ++i) {
      const unsigned long long ti = blockIdx.x + i * stride;
      if (ti < n_tiles) issue(p.tiles[ti], i);
    }
    const unsigned long long tp = blockIdx.x + PLACE_STAGES * stride;
    if (tp < n_tiles) d_pref = p.tiles[tp];
  }
  __syncthreads();
  uint32_t it = 0;
  for (unsigned long long t = blockIdx.x; t < n_tiles; t += stride, ++it) {
    const int st = it % PLACE_STAGES;
    mbar_wait(bar_s + st * 8, (it / PLACE_STAGES) & 1u);
    const uint2 meta = s_meta[st];
    const uint32_t m = meta.x, x0 = meta.y << STAGE_LOGA;
    const uint32_t tile_s = buf_s + st * PLACE_BUF_BYTES;
    uint2 r[PLACE_PER_THREAD];
#pragma unroll
    for (int u = 0; u < PLACE_PER_THREAD; ++u) {
      const uint32_t j = u * PLACE_THREADS + threadIdx.x;
      r[u] = make_uint2(0xffffffffu, 0xffffffffu);
      if (j < m) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r[u].x), "=r"(r[u].y) : "r"(tile_s + 8u * j) : "memory");
    }
    __syncthreads();                                            // the stage is free: refill it PLACE_STAGES tiles ahead
    if (threadIdx.x == 0) {
      const unsigned long long tn = t + PLACE_STAGES * stride;
      if (tn < n_tiles) {
        issue(d_pref, st);
        d_pref = tn + stride < n_tiles ? p.tiles[tn + stride] : none;
      }
    }
    uint32_t slot[PLACE_PER_THREAD], rel[PLACE_PER_THREAD];
#pragma unroll
    for (int u = 0; u < PLACE_PER_THREAD; ++u) {
      const uint32_t j = u * PLACE_THREADS + threadIdx.x;
      const unsigned long long rec = ((unsigned long long)r[u].y << 32) | r[u].x;
      const uint32_t x = x0 + (uint32_t)(rec >> (p.y_bits + p.v_bits));
      const bool mine = j < m && x >= p.aid_lo && x < p.aid_hi;      // a bucket across an owner cut is read by both owners
      const uint32_t key = mine ? x : 0xffffffffu;
      const uint32_t prev = __shfl_up_sync(FULL_MASK, key, 1);
      const uint32_t heads = __ballot_sync(FULL_MASK, lane == 0 || key != prev);
      const int head = 31 - __clz(heads & (FULL_MASK >> (31 - lane)));
      const uint32_t above = lane == 31 ? 0u : heads & (FULL_MASK << (lane + 1));
      const int next = above ? __ffs(above) - 1 : 32;
      slot[u] = 0;
      if (mine && (int)lane == head) slot[u] = atomicAdd(&p.cursor[x], (uint32_t)(next - head));
      rel[u] = mine ? ((uint32_t)head << 8) | (lane - (uint32_t)head) : 0xffffffffu;
    }
#pragma unroll
    for (int u = 0; u < PLACE_PER_THREAD; ++u) {
      const uint32_t at = __shfl_sync(FULL_MASK, slot[u], (rel[u] >> 8) & 31u);
      if (rel[u] != 0xffffffffu) {
        const unsigned long long rec = ((unsigned long long)r[u].y << 32) | r[u].x;
        st_stream_u2(p.records + (at + (rel[u] & 0xffu)),
                     make_uint2((uint32_t)(rec & ymask), (uint32_t)((rec >> p.y_bits) & vmask)));
      }
    }
  }
}

// pass B at the owner of rows [aid_lo, aid_hi): the records every rank staged for them -> `records` (final positions of the
// ordinary rows, staging area of the hot rows; otto_covisit_partition and otto_covisit_reduce follow).  staged_host[g] =
// rank g's staging buffer as this process sees it (its own, or the peer mapping); every rank's otto_covisit_scatter_staged
// must have completed (any collective between the two calls orders that).  `plan` is the caller's own plan scratch: every
// rank computes the same bucket table.  No synchronisation.
extern "C" int otto_covisit_place_staged(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                                         void* plan, int32_t n_ranks, const void* const* staged_host, int32_t aid_lo, int32_t aid_hi,
                                         void* records, int64_t records_capacity, void* stream) {
  Layout L;
  int rc = scatter_args(ev, spec, workspace, workspace_bytes, &L);
  if (rc) return rc;
  if ((rc = check_records(records, records_capacity))) return rc;
  if (!plan || !staged_host || n_ranks < 1 || n_ranks > OTTO_MAX_OWNERS || aid_lo < 0 || aid_hi > spec->n_aids || aid_lo > aid_hi) {
    otto_set_error("bad argument");
    return OTTO_EINVAL;
  }
  for (int g = 0; g < n_ranks; ++g)
    if (!staged_host[g]) { otto_set_error("staging buffer of rank %d is NULL", g); return OTTO_EINVAL; }
  const StageLayout S = stage_layout(spec, ev->n_sessions, ev->n_events, n_ranks);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t A = L.A;
  // cursors of MY layout: final position of an ordinary row, slice of the staging area of a hot row
  init_cursor_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, st>>>(WS(unsigned long long, row_off), WS(unsigned long long, hot_off),
                                                                 WS(uint32_t, bin_base), A, WS(uint32_t, cursor));
  LAUNCH_CHECK();
  if (aid_hi == aid_lo) return OTTO_OK;
  const int64_t b_lo = (int64_t)aid_lo >> STAGE_LOGA, b_hi = (((int64_t)aid_hi - 1) >> STAGE_LOGA) + 1;
  CUDA_TRY(cudaMemsetAsync(SP(unsigned long long, info) + 2, 0, 8, st));
  StagedPtrs sp;
  for (int g = 0; g < OTTO_MAX_OWNERS; ++g) sp.rank[g] = (const uint2*)staged_host[g < n_ranks ? g : 0];
  place_tiles_kernel<<<(unsigned)ceil_div(b_hi - b_lo, 256), 256, 0, st>>>(
      sp, n_ranks, b_lo, b_hi, SP(unsigned long long, bcnt), SP(unsigned long long, bo), S.NB + 1, SP(uint4, tiles),
      SP(unsigned long long, info), S.Tmax);
  LAUNCH_CHECK();
  PlaceParams pp;
  pp.tiles = SP(uint4, tiles);
  pp.info = SP(unsigned long long, info);
  pp.tile_cap = S.Tmax;
  pp.aid_lo = (uint32_t)aid_lo;
  pp.aid_hi = (uint32_t)aid_hi;
  if ((rc = stage_bits(spec, &pp.y_bits, &pp.v_bits))) return rc;
  pp.cursor = WS(uint32_t, cursor);
  pp.records = (uint2*)records;
  int dev = 0, n_sm = 148;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  if ((rc = set_smem(place_kernel, place_smem_bytes()))) return rc;
  static const int occ = occupancy_blocks((const void*)place_kernel, PLACE_THREADS, place_smem_bytes());
  place_kernel<<<n_sm * occ, PLACE_THREADS, place_smem_bytes(), st>>>(pp);
  LAUNCH_CHECK();
  return OTTO_OK;
}

// ------------------------------------------------------------------ reduce

struct ScratchLayout {
  int64_t p_key, p_sum, p_cnt, p_len, list[N_TIERS], counters, stats, total;
};

static ScratchLayout make_scratch(int k, int64_t n_bins, int64_t n_aids_range) {
  ScratchLayout s;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t at = o; o = align_up(o + bytes, 256); return at; };
  int64_t extra = n_bins - n_aids_range;
  if (extra < 0) extra = 0;
  const int64_t slots = 2 * extra + 2;
  s.p_key = take(slots * k * 8);
  s.p_sum = take(slots * k * 8);
  s.p_cnt = take(slots * k * 4);
  s.p_len = take(slots * 4);
  for (int t = 0; t < N_TIERS; ++t) s.list[t] = take((n_bins + 1) * 4);
  s.counters = take(64);
  s.stats = take(64);
  s.total = o;
  return s;
}

extern "C" int64_t otto_covisit_reduce_scratch_bytes(const OttoCovisitSpec* spec, int64_t n_bins, int64_t n_aids_range) {
  if (check_spec(spec)) return -1;
  return make_scratch(spec->k, n_bins, n_aids_range).total;
}

static cudaEvent_t g_prof_ev[6];
static bool g_prof_ready = false, g_prof_valid = false;

extern "C" int otto_profile_enable(int on) {
  if (on && !g_prof_ready) {
    for (auto& e : g_prof_ev) CUDA_TRY(cudaEventCreate(&e));
    for (auto& e : g_prof_sc) CUDA_TRY(cudaEventCreate(&e));
    g_prof_ready = true;
  }
  g_profile = on;
  return OTTO_OK;
}
extern "C" int otto_profile_reduce_ms(float* ms_host) {
  if (!g_prof_valid) { otto_set_error("no profiled otto_covisit_reduce call yet"); return OTTO_EINVAL; }
  CUDA_TRY(cudaEventSynchronize(g_prof_ev[5]));
  for (int i = 0; i < 5; ++i) CUDA_TRY(cudaEventElapsedTime(&ms_host[i], g_prof_ev[i], g_prof_ev[i + 1]));
  return OTTO_OK;
}
extern "C" int otto_profile_scatter_ms(float* ms_host) {
  if (!g_prof_sc_valid) { otto_set_error("no profiled otto_covisit_scatter call yet"); return OTTO_EINVAL; }
  CUDA_TRY(cudaEventSynchronize(g_prof_sc[3]));
  for (int i = 0; i < 3; ++i) CUDA_TRY(cudaEventElapsedTime(&ms_host[i], g_prof_sc[i], g_prof_sc[i + 1]));
  return OTTO_OK;
}
#define PROF_MARK(i)                                                  \
  do {                                                                \
    if (g_profile) CUDA_TRY(cudaEventRecord(g_prof_ev[i], st));       \
  } while (0)

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return OTTO_OK;
}

// The tier kernels work on disjoint bin lists and are persistent (blocks take work until their list is empty), so they
// run concurrently on three side streams + the caller's stream, launched largest-footprint first: the 512-thread tier
// leaves half of an SM's shared memory and three quarters of its warp slots unused, which the blocks of the later
// launches fill; as a tier runs dry the next one's waiting blocks take its place.  OTTO_REDUCE_SERIAL=1 (and profiled
// calls) keep everything on the caller's stream.
cudaStream_t g_side[3];
cudaEvent_t g_fork, g_join[3];
static bool g_side_ready = false;

static int side_streams_init() {
  if (g_side_ready) return OTTO_OK;
  for (auto& s : g_side) CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&g_fork, cudaEventDisableTiming));
  for (auto& e : g_join) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  g_side_ready = true;
  return OTTO_OK;
}

static int occupancy_blocks(const void* kernel, int threads, size_t smem) {
  int n = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
  return n;
}

template <bool TIME, int TIER>
static int launch_tier(const ReduceParams& p, int n_sm, cudaStream_t st) {
  int rc;
  if (TIER == 0) {
    auto kern = otable_warp_kernel<TIME>;
    static const int occ = occupancy_blocks((const void*)kern, OTW_WARPS * 32, 0);
    kern<<<n_sm * occ, OTW_WARPS * 32, 0, st>>>(p);
  } else if (TIER == 1) {
    auto kern = otable_block_kernel<TIME, 128, 11, 1>;
    constexpr size_t smem = otable_block_smem<TIME, 11>();
    if ((rc = set_smem(kern, smem))) return rc;
    static const int occ = occupancy_blocks((const void*)kern, 128, smem);
    kern<<<n_sm * occ, 128, smem, st>>>(p);
  } else if (TIER == 2) {
    auto kern = otable_block_kernel<TIME, 256, 12, 2>;
    constexpr size_t smem = otable_block_smem<TIME, 12>();
    if ((rc = set_smem(kern, smem))) return rc;
    static const int occ = occupancy_blocks((const void*)kern, 256, smem);
    kern<<<n_sm * occ, 256, smem, st>>>(p);
  } else if (TIER == 3) {
    auto kern = otable_block_kernel<TIME, 512, 13, 3>;
    constexpr size_t smem = otable_block_smem<TIME, 13>();
    if ((rc = set_smem(kern, smem))) return rc;
    static const int occ = occupancy_blocks((const void*)kern, 512, smem);
    kern<<<n_sm * occ, 512, smem, st>>>(p);
  } else {
    // bins beyond 6144 records and bins handed over by the owner-table tiers: the v5 hash-table kernel (multi-pass)
    auto kern = reduce_block_kernel<TIME, 512, 13, 4, true>;
    constexpr size_t smem = reduce_block_smem<TIME, 512, 13, true>();
    if ((rc = set_smem(kern, smem))) return rc;
    static const int occ = occupancy_blocks((const void*)kern, 512, smem);
    kern<<<n_sm * occ, 512, smem, st>>>(p);
  }
  LAUNCH_CHECK();
  return OTTO_OK;
}

// profile slots: [0] classify + warp tier 0, [1] block tier 3, [2] block tier 2, [3] block tier 1, [4] hash-table tier
// (hand-overs) + split-row merge
template <bool TIME>
static int launch_reduce(const ReduceParams& p, int n_sm, cudaStream_t st) {
  int rc;
  static const bool serial = getenv("OTTO_REDUCE_SERIAL") != nullptr;
  const bool fork = !serial && !g_profile;
  if (fork && (rc = side_streams_init())) return rc;
  const int64_t n_bins = p.bin_hi - p.bin_lo;
  PROF_MARK(0);
  reduce_classify_kernel<<<(unsigned)ceil_div(n_bins, 256), 256, 0, st>>>(p);
  LAUNCH_CHECK();
  if (fork) {
    CUDA_TRY(cudaEventRecord(g_fork, st));
    for (auto& s : g_side) CUDA_TRY(cudaStreamWaitEvent(s, g_fork, 0));
    if ((rc = launch_tier<TIME, 3>(p, n_sm, g_side[2]))) return rc;
    if ((rc = launch_tier<TIME, 2>(p, n_sm, g_side[1]))) return rc;
    if ((rc = launch_tier<TIME, 1>(p, n_sm, g_side[0]))) return rc;
    if ((rc = launch_tier<TIME, 0>(p, n_sm, st))) return rc;
    for (int i = 0; i < 3; ++i) {
      CUDA_TRY(cudaEventRecord(g_join[i], g_side[i]));
      CUDA_TRY(cudaStreamWaitEvent(st, g_join[i], 0));
    }
  } else {
    if ((rc = launch_tier<TIME, 0>(p, n_sm, st))) return rc;
    PROF_MARK(1);
    if ((rc = launch_tier<TIME, 3>(p, n_sm, st))) return rc;
    PROF_MARK(2);
    if ((rc = launch_tier<TIME, 2>(p, n_sm, st))) return rc;
    PROF_MARK(3);
    if ((rc = launch_tier<TIME, 1>(p, n_sm, st))) return rc;
    PROF_MARK(4);
  }
  // after every owner-table tier: their hand-overs join the oversized bins on list 4
  if ((rc = launch_tier<TIME, 4>(p, n_sm, st))) return rc;
  merge_split_rows_kernel<<<n_sm * 8, MERGE_WARPS * 32, 0, st>>>(p);
  LAUNCH_CHECK();
  PROF_MARK(5);
  g_prof_valid = g_profile != 0;
  return OTTO_OK;
}

extern "C" int otto_covisit_reduce(const OttoCovisitSpec* spec, const uint32_t* bin_base, const uint32_t* bin_x,
                                   int64_t bin_lo, int64_t bin_hi, int32_t aid_lo, int32_t aid_hi,
                                   const OttoPairSegment* segments_host, int32_t n_segments, void* scratch,
                                   int64_t scratch_bytes, const OttoTopK* out, OttoBuildStats* stats_host,
                                   void* stream) {
  int rc = check_spec(spec);
  if (rc) return rc;
  if (n_segments != 1 || !segments_host) {
    otto_set_error("otto_covisit_reduce takes ONE bin-contiguous segment: merge received segments first (otto_covisit_merge_segments)");
    return OTTO_EINVAL;
  }
  if (!out || out->k != spec->k || out->n_aids != spec->n_aids) { otto_set_error("output table shape does not match the spec"); return OTTO_EINVAL; }
  if (bin_hi < bin_lo || aid_hi < aid_lo) { otto_set_error("empty or inverted range"); return OTTO_EINVAL; }
  const ScratchLayout SL = make_scratch(spec->k, bin_hi - bin_lo, aid_hi - aid_lo);
  if (!scratch || scratch_bytes < SL.total) {
    otto_set_error("reduce scratch too small: need %lld bytes", (long long)SL.total);
    return OTTO_ENOSPC;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ReduceParams p;
  memset(&p, 0, sizeof(p));
  p.records = (const uint2*)segments_host[0].records;
  p.offsets = segments_host[0].offsets;
  p.bin_x = bin_x;
  p.bin_base = bin_base;
  p.bin_lo = bin_lo;
  p.bin_hi = bin_hi;
  p.aid_lo = aid_lo;
  p.aid_hi = aid_hi;
  p.k = spec->k;
  p.time_mode = spec->weight_mode == OTTO_WEIGHT_TIME;
  p.w_scale = p.time_mode ? 3.0 / (double)(spec->ts_max - spec->ts_min) : 0.0;
  p.range = p.time_mode ? (uint32_t)(spec->ts_max - spec->ts_min) : 0u;
  p.y_bits = 0;
  while (p.y_bits < 31 && (1ll << p.y_bits) < (int64_t)spec->n_aids) ++p.y_bits;
  p.max_v = 1;
  if (spec->weight_mode == OTTO_WEIGHT_TYPE)
    for (int i = 0; i < 3; ++i)
      if ((uint32_t)spec->type_weight[i] > p.max_v) p.max_v = (uint32_t)spec->type_weight[i];
  p.out_y = out->aid_y;
  p.out_w = out->wgt;
  p.out_len = out->len;
  p.out_cnt = out->cnt;
  p.out_tsum = out->tsum;
  char* sc = (char*)scratch;
  p.p_key = (uint64_t*)(sc + SL.p_key);
  p.p_sum = (uint64_t*)(sc + SL.p_sum);
  p.p_cnt = (uint32_t*)(sc + SL.p_cnt);
  p.p_len = (int32_t*)(sc + SL.p_len);
  for (int t = 0; t < N_TIERS; ++t) p.list[t] = (uint32_t*)(sc + SL.list[t]);
  p.counters = (uint32_t*)(sc + SL.counters);
  p.stats = (unsigned long long*)(sc + SL.stats);
  CUDA_TRY(cudaMemsetAsync(p.counters, 0, 64, st));
  CUDA_TRY(cudaMemsetAsync(p.stats, 0, 64, st));
  int dev = 0, n_sm = 148;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  if (bin_hi > bin_lo) {
    rc = p.time_mode ? launch_reduce<true>(p, n_sm, st) : launch_reduce<false>(p, n_sm, st);
    if (rc) return rc;
  }
  if (stats_host) {
    unsigned long long h[8];
    CUDA_TRY(cudaMemcpyAsync(h, p.stats, sizeof(h), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    stats_host->distinct = (int64_t)h[0];
    stats_host->pair_checksum = (int64_t)h[1];
    stats_host->table_overflow = (int64_t)h[2];
    for (int i = 0; i < 4; ++i) stats_host->tier_records[i] = (int64_t)h[4 + i];
    if (h[2]) {
      otto_set_error("a shared-memory hash table overflowed; lower split_ub");
      return OTTO_EOVERFLOW;
    }
  }
  return OTTO_OK;
}

// ------------------------------------------------------------------ multi-GPU: merge the received segments

struct MergeParams {
  OttoPairSegment seg[OTTO_MAX_SEGMENTS];
  int32_t n_seg;
  int64_t n_bins;
};

__global__ void merge_count_kernel(const MergeParams p, unsigned long long* __restrict__ total) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.n_bins) return;
  unsigned long long n = 0;
  for (int s = 0; s < p.n_seg; ++s) n += p.seg[s].offsets[b + 1] - p.seg[s].offsets[b];
  total[b] = n;
}

// one warp per bin: the runs of the segments, in segment order, back to back
__global__ void __launch_bounds__(256)
    merge_copy_kernel(const MergeParams p, const unsigned long long* __restrict__ merged_off, uint2* __restrict__ dst) {
  const uint32_t lane = lane_id();
  const int64_t n_warps = (int64_t)gridDim.x * 8;
  for (int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); b < p.n_bins; b += n_warps) {
    unsigned long long at = merged_off[b];
    for (int s = 0; s < p.n_seg; ++s) {
      const uint64_t beg = p.seg[s].offsets[b] - p.seg[s].offsets[0], end = p.seg[s].offsets[b + 1] - p.seg[s].offsets[0];
      const uint2* src = (const uint2*)p.seg[s].records;
      for (uint64_t i = beg + lane; i < end; i += 32) st_stream_u2(dst + at + (i - beg), ld_stream_u2(src + i));
      at += end - beg;
    }
  }
}

extern "C" int64_t otto_covisit_merge_scratch_bytes(int64_t n_bins) { return scan_scratch_elems(n_bins + 1) * 8 + 256; }

extern "C" int otto_covisit_merge_segments(const OttoPairSegment* segments_host, int32_t n_segments, int64_t n_bins,
                                           void* merged_records, int64_t merged_capacity, uint64_t* merged_offsets,
                                           void* scratch, int64_t scratch_bytes, int64_t* n_records_host, void* stream) {
  if (n_segments < 1 || n_segments > OTTO_MAX_SEGMENTS) { otto_set_error("n_segments must be in [1, 8]"); return OTTO_EINVAL; }
  if (n_bins < 0 || !merged_offsets) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  if (scratch_bytes < otto_covisit_merge_scratch_bytes(n_bins)) { otto_set_error("merge scratch too small"); return OTTO_ENOSPC; }
  cudaStream_t st = (cudaStream_t)stream;
  MergeParams p;
  memset(&p, 0, sizeof(p));
  for (int s = 0; s < n_segments; ++s) p.seg[s] = segments_host[s];
  p.n_seg = n_segments;
  p.n_bins = n_bins;
  unsigned long long* off = (unsigned long long*)merged_offsets;
  if (n_bins > 0) {
    merge_count_kernel<<<(unsigned)ceil_div(n_bins, 256), 256, 0, st>>>(p, off);
    LAUNCH_CHECK();
  }
  int rc = exclusive_scan<unsigned long long, unsigned long long>(off, n_bins, off, (unsigned long long*)scratch, st);
  if (rc) return rc;
  int64_t total = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, off + n_bins, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (total > merged_capacity) { otto_set_error("merged_records too small: need %lld records", (long long)total); return OTTO_ENOSPC; }
  if (n_bins > 0 && total > 0) {
    int dev = 0, n_sm = 148;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    merge_copy_kernel<<<n_sm * 8, 256, 0, st>>>(p, off, (uint2*)merged_records);
    LAUNCH_CHECK();
  }
  if (n_records_host) *n_records_host = total;
  return OTTO_OK;
}

// ------------------------------------------------------------------ peer memory (CUDA IPC over NVLink)

extern "C" int otto_peer_alloc(int64_t bytes, void** ptr_host) {
  if (bytes <= 0 || !ptr_host) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  // The size is rounded up to a multiple of 32 MiB.  Measured on this pool's B200s (profiles/r01_microbench_peer.txt):
  // a peer that maps an IPC-exported cudaMalloc allocation whose size is NOT a multiple of 2 MiB reaches it through
  // small pages; scattered 64-byte stores over a 6.2 GB buffer then run at 6 GB/s instead of 313 GB/s (TLB misses
  // on every run), which made the owner-direct scatter 35x slower than the link allows.
  bytes = (bytes + (32ll << 20) - 1) & ~((32ll << 20) - 1);
  const cudaError_t e = cudaMalloc(ptr_host, (size_t)bytes);
  if (e != cudaSuccess) {
    size_t free_b = 0, total_b = 0;
    int dev = -1;
    cudaGetDevice(&dev);
    cudaMemGetInfo(&free_b, &total_b);
    otto_set_error("otto_peer_alloc: cudaMalloc of %lld bytes on device %d failed (%s); %zu of %zu bytes free", (long long)bytes,
                   dev, cudaGetErrorString(e), free_b, total_b);
    return OTTO_ECUDA;
  }
  return OTTO_OK;
}
extern "C" int otto_peer_free(void* ptr) {
  CUDA_TRY(cudaFree(ptr));
  return OTTO_OK;
}
extern "C" int otto_peer_get_handle(void* ptr, uint8_t* handle_host) {
  static_assert(sizeof(cudaIpcMemHandle_t) == OTTO_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle_host, &h, sizeof(h));
  return OTTO_OK;
}
extern "C" int otto_peer_open(const uint8_t* handle_host, void** ptr_host) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  CUDA_TRY(cudaIpcOpenMemHandle(ptr_host, h, cudaIpcMemLazyEnablePeerAccess));
  return OTTO_OK;
}
extern "C" int otto_peer_close(void* ptr) {
  CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return OTTO_OK;
}

// ------------------------------------------------------------------ one-shot build

extern "C" int otto_covisit_build(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                                  int64_t workspace_bytes, const OttoTopK* out, OttoBuildStats* stats_host,
                                  void* stream) {
  OttoBuildStats local;
  memset(&local, 0, sizeof(local));
  OttoBuildStats* stp = stats_host ? stats_host : &local;
  int rc = otto_covisit_count(ev, spec, workspace, workspace_bytes, stp, stream);
  if (rc) return rc;
  const Layout L = make_layout(ev->n_sessions, ev->n_events, spec);
  const ScratchLayout SL = make_scratch(spec->k, stp->bins, L.A);
  const int64_t rec_at = align_up(L.total, 256);
  const int64_t scratch_at = align_up(rec_at + (stp->pairs + stp->hot_pairs) * 8, 256);
  const int64_t need = scratch_at + SL.total;
  if (workspace_bytes < need) {
    otto_set_error("workspace too small for %lld pair records: need %lld bytes", (long long)stp->pairs, (long long)need);
    return OTTO_ENOSPC;
  }
  char* ws = (char*)workspace;
  if ((rc = otto_covisit_scatter(ev, spec, workspace, workspace_bytes, ws + rec_at, stp->pairs + stp->hot_pairs, stream))) return rc;
  OttoPairSegment seg;
  seg.records = ws + rec_at;
  seg.offsets = WS(uint64_t, bin_off);
  return otto_covisit_reduce(spec, WS(uint32_t, bin_base), WS(uint32_t, bin_x), 0, stp->bins, 0, spec->n_aids, &seg, 1,
                             ws + scratch_at, SL.total, out, stp, stream);
}

// bytes otto_covisit_build needs once the pair count is known
extern "C" int64_t otto_covisit_build_bytes(int64_t n_sessions, int64_t n_events, const OttoCovisitSpec* spec,
                                            int64_t pairs, int64_t bins) {
  if (check_spec(spec)) return -1;
  const Layout L = make_layout(n_sessions, n_events, spec);
  const ScratchLayout SL = make_scratch(spec->k, bins, L.A);
  return align_up(align_up(L.total, 256) + pairs * 8, 256) + SL.total;
}

// ------------------------------------------------------------------ table <-> file rows

__global__ void len_to_i64_kernel(const int32_t* __restrict__ len, int64_t n, unsigned long long* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (unsigned long long)len[i];
}

extern "C" int otto_topk_row_offsets(const OttoTopK* table, int64_t* row_offsets, int64_t* n_rows_host, void* scratch,
                                     int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t A = table->n_aids;
  if (scratch_bytes < scan_scratch_elems(A) * 8) { otto_set_error("scan scratch too small"); return OTTO_ENOSPC; }
  len_to_i64_kernel<<<(unsigned)ceil_div(A, 256), 256, 0, st>>>(table->len, A, (unsigned long long*)row_offsets);
  LAUNCH_CHECK();
  int rc = exclusive_scan<unsigned long long, unsigned long long>((unsigned long long*)row_offsets, A,
                                                                  (unsigned long long*)row_offsets,
                                                                  (unsigned long long*)scratch, st);
  if (rc) return rc;
  if (n_rows_host) {
    CUDA_TRY(cudaMemcpyAsync(n_rows_host, row_offsets + A, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return OTTO_OK;
}

__global__ void topk_to_rows_kernel(const int32_t* __restrict__ ty, const float* __restrict__ tw,
                                    const int32_t* __restrict__ len, const int64_t* __restrict__ row_off, int64_t A, int k,
                                    int32_t* __restrict__ ax, int32_t* __restrict__ ay, float* __restrict__ w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A * k) return;
  const int64_t x = i / k;
  const int r = (int)(i % k);
  if (r < len[x]) {
    const int64_t at = row_off[x] + r;
    ax[at] = (int32_t)x;
    ay[at] = ty[i];
    w[at] = tw[i];
  }
}

extern "C" int otto_topk_to_rows(const OttoTopK* table, const int64_t* row_offsets, int32_t* aid_x, int32_t* aid_y,
                                 float* wgt, void* stream) {
  const int64_t n = (int64_t)table->n_aids * table->k;
  topk_to_rows_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      table->aid_y, table->wgt, table->len, row_offsets, table->n_aids, table->k, aid_x, aid_y, wgt);
  LAUNCH_CHECK();
  return OTTO_OK;
}

// rows grouped by aid_x (any group order), ranked inside a group by file order
__global__ void rows_to_topk_kernel(const int32_t* __restrict__ ax, const int32_t* __restrict__ ay,
                                    const float* __restrict__ w, int64_t n, int32_t A, int k, int32_t* __restrict__ ty,
                                    float* __restrict__ tw, int32_t* __restrict__ len) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t x = ax[i];
  if (x < 0 || x >= A) return;
  if (i > 0 && ax[i - 1] == x) return;  // only the first row of a group walks it
  int r = 0;
  for (int64_t j = i; j < n && ax[j] == x && r < k; ++j, ++r) {
    ty[(int64_t)x * k + r] = ay[j];
    if (tw) tw[(int64_t)x * k + r] = w ? w[j] : 0.f;
  }
  len[x] = r;
}

extern "C" int otto_rows_to_topk(const int32_t* aid_x, const int32_t* aid_y, const float* wgt, int64_t n_rows,
                                 const OttoTopK* table, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)table->n_aids * table->k;
  CUDA_TRY(cudaMemsetAsync(table->aid_y, 0xff, n * 4, st));
  if (table->wgt) CUDA_TRY(cudaMemsetAsync(table->wgt, 0, n * 4, st));
  CUDA_TRY(cudaMemsetAsync(table->len, 0, (int64_t)table->n_aids * 4, st));
  if (n_rows > 0) {
    rows_to_topk_kernel<<<(unsigned)ceil_div(n_rows, 256), 256, 0, st>>>(aid_x, aid_y, wgt, n_rows, table->n_aids,
                                                                          table->k, table->aid_y, table->wgt, table->len);
    LAUNCH_CHECK();
  }
  return OTTO_OK;
}

// Covisitation candidate generation: per test session gather the neighbours of its history aids from the
// top-K tables, count votes and keep the most common N (count desc, first-seen asc), then drop history aids.
//
// Replaces the per-session Python loop of src/ranker/covisitation_candidate_generation.py:108-141 /
// :248-281 and src/covisitation/inference.py:204-236 / :396-428:
//   session_unique_aids = list(dict.fromkeys(aids[::-1]))                         -> H (recency order)
//   np.unique(aids[types <= 1]) etc.                                              -> ascending subsets
//   itertools.chain(*[table[aid] for aid in <set> if aid in table])               -> gather, in set order
//   Counter(concat).most_common(N) ... if aid not in session_unique_aids          -> vote + cut + filter
// Counter.most_common sorts by (count desc, first insertion asc).  v2 counts the votes in an open-addressing
// table (shared memory; key = aid_y, payload = count and the smallest position in the concatenation) instead of
// sorting all gathered items (v1: two bitonic sorts per target were 60 % of the instructions,
// profiles/r01_cand_v1), then ranks only the entries that can reach the top N:
//   entry  = count << 48 | (0xffff - first position) << 32 | aid_y          (descending = most_common order)
//   N <= 32: the N-th largest of the 32 lane-group maxima is a lower bound of the N-th largest entry; the
//            entries at or above it (about N .. 2N) are sorted in registers by one warp
//   N  > 32: all entries are sorted (bitonic, shared memory)
// One cooperative routine serves three tiers that differ only in group size and where the arrays live:
// a warp with shared-memory arrays (sessions whose bound is <= 256 items), a 256-thread block with
// shared-memory arrays (<= 2048 items) and a 256-thread block over a global scratch slab (anything else;
// OTTO's longest test session has 458 events).
#include "common.cuh"

struct CandParams {
  const int32_t* off;
  const int32_t* aid;
  const uint8_t* type;
  int64_t n_sessions;
  OttoCandidateSpec spec;
  int32_t* out_aid;
  int32_t* out_score;
  int32_t* out_len;
  // tiers
  uint32_t* list_block;   // sessions for the 128-thread shared-memory tier
  uint32_t* list_large;   // sessions for the 256-thread shared-memory tier
  uint32_t* list_global;  // sessions for the block-global tier
  uint32_t* counters;     // [0] n_block [1] n_global [2] next_block [3] next_global [4] n_large [5] next_large
  uint64_t* slab;         // global tier: per-block slab
  int64_t slab_words;     // u64 words per block
  int32_t max_k_sum;      // max over targets of the sum of table_k over its sources
  int32_t max_len;        // longest session (events)
  int32_t n_run;          // targets with distinct source lists (the others are copies: carts == orders in the reference)
  int32_t run_target[OTTO_MAX_TARGETS];
};

constexpr int W_LCAP = 32, W_MCAP = 256;        // warp tier: events, gathered items
constexpr int B_LCAP = 64, B_MCAP = 1024;       // 128-thread shared-memory tier (5 CTAs per SM)
constexpr int X_LCAP = 256, X_MCAP = 4096;      // 256-thread shared-memory tier (1 CTA per SM)
constexpr int CAND_WARPS = 4;

template <int T>
__device__ __forceinline__ void gsync() {
  if (T == 32) __syncwarp();
  else __syncthreads();
}

// ascending bitonic sort of n (power of two) keys by a group of T threads
template <int T>
__device__ __forceinline__ void bitonic_sort(uint64_t* a, int n, int tid) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n >> 1); t += T) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const uint64_t x = a[i], y = a[l];
        const bool up = (i & k) == 0;
        if ((x > y) == up) { a[i] = y; a[l] = x; }
      }
      gsync<T>();
    }
  }
}

// bitonic sort of one u64 per lane, descending (lane 0 ends with the largest)
__device__ __forceinline__ uint64_t cand_warp_sort_desc(uint64_t v) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t o = shfl_u64(v, lane ^ j);
      const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);
      v = keep_max ? (o > v ? o : v) : (o < v ? o : v);
    }
  }
  return v;
}

__device__ __forceinline__ int pow2_at_least(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

// arrays of one session's work area
struct Work {
  int32_t* ev_aid;   // [Lcap] events, most recent first
  int32_t* ev_ty;    // [Lcap]
  int32_t* uidx;     // [Lcap] index among unique aids, or -1
  int32_t* H;        // [Lcap] unique aids in recency order
  int32_t* tmask;    // [Lcap] per unique aid: OR of 1 << type
  int32_t* elen;     // [Lcap] per element of the current source: table row length
  int32_t* estart;   // [Lcap] exclusive prefix of elen
  uint64_t* ord;     // [pow2(Lcap)] subset sort buffer
  uint32_t* keys;    // [HS] vote table: aid_y (KEY_EMPTY = free)
  uint32_t* cnt;     // [HS] votes
  uint32_t* first;   // [HS] smallest position in the concatenation
  uint32_t* occ;     // [HS / 2] claimed slots
  uint64_t* ent;     // [Mcap] inverted entries of the survivors, best first
  int32_t* scal;     // [8] scalars: 0 U, 1 n_ord, 2 n_occ, 3 kept, 4 n_cand;  [6..7] threshold (u64)
  uint32_t hmask;    // HS - 1
  int hshift;        // 32 - log2(HS)
};

__device__ __forceinline__ uint32_t hist_mask(int sel) {
  return sel == OTTO_HIST_TYPE_LE1 ? 3u : sel == OTTO_HIST_TYPE_GE1 ? 6u : sel == OTTO_HIST_TYPE_EQ0 ? 1u : 7u;
}

// one vote for aid y at position pos of the concatenation (the table never fills: items <= HS / 2)
__device__ __forceinline__ void vote(const Work& w, uint32_t y, uint32_t pos) {
  uint32_t h = (y * 0x9E3779B1u) >> w.hshift;
  while (true) {
    const uint32_t prev = atomicCAS(&w.keys[h], KEY_EMPTY, y);
    if (prev == KEY_EMPTY) {
      w.occ[atomicAdd(&w.scal[2], 1)] = h;
      break;
    }
    if (prev == y) break;
    h = (h + 1) & w.hmask;
  }
  atomicAdd(&w.cnt[h], 1u);
  atomicMin(&w.first[h], pos);
}

__device__ __forceinline__ uint64_t entry_key(const Work& w, uint32_t h) {
  return ((uint64_t)w.cnt[h] << 48) | ((uint64_t)(0xffffu - w.first[h]) << 32) | w.keys[h];
}

template <int T>
__device__ void process_session(const CandParams& p, int64_t s, int tid, const Work& w, uint64_t* s_gmax) {
  const OttoCandidateSpec& sp = p.spec;
  const int32_t beg = p.off[s], end = p.off[s + 1];
  const int L = end - beg;
  const int N = sp.top_n;
  const uint32_t lane = tid & 31, lt = lanemask_lt();
  // 1. events, most recent first
  for (int i = tid; i < L; i += T) {
    w.ev_aid[i] = p.aid[end - 1 - i];
    w.ev_ty[i] = p.type[end - 1 - i];
  }
  gsync<T>();
  // 2. unique aids in recency order: dict.fromkeys(aids[::-1])
  for (int i = tid; i < L; i += T) {
    const int32_t a = w.ev_aid[i];
    bool first = true;
    for (int j = 0; j < i; ++j)
      if (w.ev_aid[j] == a) { first = false; break; }
    w.uidx[i] = first ? 0 : -1;
  }
  gsync<T>();
  for (int i = tid; i < L; i += T) {
    if (w.uidx[i] == 0) {
      int u = 0;
      for (int j = 0; j < i; ++j) u += w.uidx[j] >= 0;   // uidx[j] is 0 or -1 until rewritten below (only own slot)
      w.H[u] = w.ev_aid[i];
    }
  }
  gsync<T>();
  if (tid == 0) {
    int u = 0;
    for (int j = 0; j < L; ++j) u += w.uidx[j] >= 0;
    w.scal[0] = u;
  }
  gsync<T>();
  const int U = w.scal[0];
  // 3. event types seen per unique aid
  for (int u = tid; u < U; u += T) {
    const int32_t a = w.H[u];
    int m = 0;
    for (int i = 0; i < L; ++i)
      if (w.ev_aid[i] == a) m |= 1 << w.ev_ty[i];
    w.tmask[u] = m;
  }
  gsync<T>();

  for (int ri = 0; ri < p.n_run; ++ri) {
    const int tg = p.run_target[ri];
    // 4. gather + vote: walk the table rows of every source in concatenation order
    int base = 0;
    for (int si = 0; si < sp.target_n_sources[tg]; ++si) {
      const int src = sp.target_sources[tg][si];
      const int tb = sp.source_table[src], sel = sp.source_hist[src];
      const int K = sp.table_k[tb];
      const int32_t* tlen = sp.table_len[tb];
      const int32_t* ty = sp.table_aid_y[tb];
      const int32_t* E;   // ordered history set of this source
      int ne;
      if (sel == OTTO_HIST_RECENCY) {
        E = w.H;
        ne = U;
      } else {
        // np.unique(aids[type filter]): ascending subset of the unique aids
        const uint32_t mask = hist_mask(sel);
        const int np = pow2_at_least(U > 1 ? U : 1);
        for (int u = tid; u < np; u += T)
          w.ord[u] = (u < U && (w.tmask[u] & mask)) ? (uint64_t)(uint32_t)w.H[u] : ~0ull;
        gsync<T>();
        bitonic_sort<T>(w.ord, np, tid);
        if (tid == 0) {
          int c = 0;
          while (c < np && w.ord[c] != ~0ull) ++c;
          w.scal[1] = c;
        }
        gsync<T>();
        ne = w.scal[1];
        for (int u = tid; u < ne; u += T) w.uidx[u] = (int32_t)w.ord[u];
        gsync<T>();
        E = w.uidx;
      }
      for (int e = tid; e < ne; e += T) {
        const int32_t a = E[e];
        w.elen[e] = (a >= 0 && a < sp.n_aids) ? tlen[a] : 0;
      }
      gsync<T>();
      for (int e = tid; e < ne; e += T) {
        int o = 0;
        for (int j = 0; j < e; ++j) o += w.elen[j];
        w.estart[e] = o;
      }
      gsync<T>();
      const int total = ne > 0 ? w.estart[ne - 1] + w.elen[ne - 1] : 0;
      for (int idx = tid; idx < ne * K; idx += T) {
        const int e = idx / K, r = idx - e * K;
        if (r < w.elen[e]) vote(w, (uint32_t)ty[(int64_t)E[e] * K + r], (uint32_t)(base + w.estart[e] + r));
      }
      gsync<T>();
      base += total;
    }
    const int d = w.scal[2];
    const int n_top = d < N ? d : N;
    // 5. the n_top best entries, inverted (~key), best first, into w.ent
    if (d <= 32) {
      if (tid < 32) {
        const uint64_t v = (int)lane < d ? entry_key(w, w.occ[lane]) : 0ull;
        const uint64_t sorted = cand_warp_sort_desc(v);
        if ((int)lane < d) w.ent[lane] = ~sorted;
      }
    } else if (N <= 32) {
      uint64_t best = 0;
      for (int i = tid; i < d; i += T) {
        const uint64_t k = entry_key(w, w.occ[i]);
        best = k > best ? k : best;
      }
      if (T > 32) {
        s_gmax[tid] = best;
        __syncthreads();
      }
      if (tid < 32) {
        uint64_t g = best;
        if (T > 32)
          for (int wv = 1; wv < T / 32; ++wv) g = s_gmax[wv * 32 + lane] > g ? s_gmax[wv * 32 + lane] : g;
        const uint64_t kth = shfl_u64(cand_warp_sort_desc(g), N - 1);   // d > 32: every lane holds an entry
        if (lane == 0) {
          *(uint64_t*)(w.scal + 6) = kth;
          w.scal[4] = 0;
        }
      }
      gsync<T>();
      const uint64_t thr = *(const uint64_t*)(w.scal + 6);
      for (int i = tid; i < d; i += T) {
        const uint64_t k = entry_key(w, w.occ[i]);
        if (k >= thr) w.ent[atomicAdd(&w.scal[4], 1)] = ~k;
      }
      gsync<T>();
      const int n_c = w.scal[4];
      if (n_c <= 32) {
        if (tid < 32) {
          const uint64_t v = (int)lane < n_c ? ~w.ent[lane] : 0ull;
          const uint64_t sorted = cand_warp_sort_desc(v);
          if ((int)lane < n_c) w.ent[lane] = ~sorted;
        }
      } else {
        const int np = pow2_at_least(n_c);
        for (int i = n_c + tid; i < np; i += T) w.ent[i] = ~0ull;
        gsync<T>();
        bitonic_sort<T>(w.ent, np, tid);
      }
    } else {
      for (int i = tid; i < d; i += T) w.ent[i] = ~entry_key(w, w.occ[i]);
      const int np = pow2_at_least(d);
      for (int i = d + tid; i < np; i += T) w.ent[i] = ~0ull;
      gsync<T>();
      bitonic_sort<T>(w.ent, np, tid);
    }
    gsync<T>();
    // reset the claimed slots for the next target / session
    for (int i = tid; i < d; i += T) {
      const uint32_t h = w.occ[i];
      w.keys[h] = KEY_EMPTY;
      w.cnt[h] = 0;
      w.first[h] = 0xffffffffu;
    }
    // 6. most_common(N), then drop history aids; order is preserved.  Dropped entries become ~0.
    for (int r = tid; r < n_top; r += T) {
      const uint32_t a = (uint32_t)(~w.ent[r]);
      bool keep = true;
      if (sp.drop_history)
        for (int u = 0; u < U; ++u)
          if ((uint32_t)w.H[u] == a) { keep = false; break; }
      if (!keep) w.ent[r] = ~0ull;
    }
    gsync<T>();
    int32_t* oa = p.out_aid + ((int64_t)tg * p.n_sessions + s) * N;
    int32_t* os = p.out_score + ((int64_t)tg * p.n_sessions + s) * N;
    int kept;
    if (n_top <= 32) {
      // one warp: ballot compaction (the usual case: N = 20)
      if (tid < 32) {
        const bool has = (int)lane < n_top && w.ent[lane] != ~0ull;
        const uint32_t m = __ballot_sync(FULL_MASK, has);
        if (has) {
          const uint64_t key = ~w.ent[lane];
          const int at = __popc(m & lt);
          oa[at] = (int32_t)(uint32_t)key;
          os[at] = (int32_t)(key >> 48);
        }
        if (lane == 0) w.scal[3] = __popc(m);
      }
    } else {
      for (int r = tid; r < n_top; r += T) {
        if (w.ent[r] != ~0ull) {
          int at = 0;
          for (int j = 0; j < r; ++j) at += w.ent[j] != ~0ull;
          const uint64_t key = ~w.ent[r];
          oa[at] = (int32_t)(uint32_t)key;
          os[at] = (int32_t)(key >> 48);
        }
      }
      if (tid == 0) {
        int c = 0;
        for (int j = 0; j < n_top; ++j) c += w.ent[j] != ~0ull;
        w.scal[3] = c;
      }
    }
    gsync<T>();
    kept = w.scal[3];
    if (tid == 0) {
      p.out_len[(int64_t)tg * p.n_sessions + s] = kept;
      w.scal[2] = 0;
    }
    for (int r = kept + tid; r < N; r += T) {
      oa[r] = -1;
      os[r] = 0;
    }
    gsync<T>();
  }
}

// layout of a work area: u64 arrays first (ent[MCAP], ord[pow2 LCAP]), then the u32 / i32 arrays
template <int LCAP, int MCAP>
__host__ __device__ constexpr size_t work_bytes() {
  return (size_t)MCAP * 8 + (size_t)LCAP * 8 + (size_t)(2 * MCAP) * 12 + (size_t)MCAP * 4 + (size_t)LCAP * 4 * 7 + 32;
}

__device__ __forceinline__ Work carve_rt(unsigned char* base, int64_t lcap, int64_t mcap, int32_t* scal) {
  Work w;
  const int64_t hs = 2 * mcap;
  w.ent = (uint64_t*)base;
  w.ord = w.ent + mcap;
  w.keys = (uint32_t*)(w.ord + lcap);
  w.cnt = w.keys + hs;
  w.first = w.cnt + hs;
  w.occ = w.first + hs;
  w.ev_aid = (int32_t*)(w.occ + mcap);
  w.ev_ty = w.ev_aid + lcap;
  w.uidx = w.ev_ty + lcap;
  w.H = w.uidx + lcap;
  w.tmask = w.H + lcap;
  w.elen = w.tmask + lcap;
  w.estart = w.elen + lcap;
  w.scal = scal ? scal : w.estart + lcap;
  w.hmask = (uint32_t)(hs - 1);
  w.hshift = 32 - (63 - __clzll((long long)hs));
  return w;
}

template <int T>
__device__ __forceinline__ void clear_table(const Work& w, int tid) {
  for (uint32_t h = tid; h <= w.hmask; h += T) {
    w.keys[h] = KEY_EMPTY;
    w.cnt[h] = 0;
    w.first[h] = 0xffffffffu;
  }
  if (tid == 0) w.scal[2] = 0;
  gsync<T>();
}

__device__ __forceinline__ int64_t item_bound(const CandParams& p, int L) { return (int64_t)L * p.max_k_sum; }

// tier 1: one warp per session; larger sessions are appended to the block tiers' lists
__global__ void __launch_bounds__(CAND_WARPS * 32) candidates_warp_kernel(const CandParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Work w = carve_rt(smem_raw + warp * work_bytes<W_LCAP, W_MCAP>(), W_LCAP, W_MCAP, nullptr);
  clear_table<32>(w, lane);
  const int64_t n_warps = (int64_t)gridDim.x * CAND_WARPS;
  for (int64_t s = (int64_t)blockIdx.x * CAND_WARPS + warp; s < p.n_sessions; s += n_warps) {
    const int L = p.off[s + 1] - p.off[s];
    const int64_t bound = item_bound(p, L);
    if (L > W_LCAP || bound > W_MCAP) {
      if (lane == 0) {
        if (L <= B_LCAP && bound <= B_MCAP) p.list_block[atomicAdd(&p.counters[0], 1u)] = (uint32_t)s;
        else if (L <= X_LCAP && bound <= X_MCAP) p.list_large[atomicAdd(&p.counters[4], 1u)] = (uint32_t)s;
        else p.list_global[atomicAdd(&p.counters[1], 1u)] = (uint32_t)s;
      }
      continue;
    }
    process_session<32>(p, s, lane, w, nullptr);
  }
}

static int64_t global_mcap(int max_len, int max_k_sum) {
  const int64_t bound = (int64_t)max_len * max_k_sum;
  int64_t mcap = 1;
  while (mcap < bound) mcap <<= 1;
  return mcap;
}
static int64_t global_lcap(int max_len) {
  int64_t lcap = 1;
  while (lcap < max_len) lcap <<= 1;
  return lcap;
}
// u64 words of one block's slab in the global tier (same layout as work_bytes)
static int64_t global_slab_words(int max_len, int max_k_sum) {
  const int64_t mcap = global_mcap(max_len, max_k_sum), lcap = global_lcap(max_len);
  return (mcap * 8 + lcap * 8 + 2 * mcap * 12 + mcap * 4 + lcap * 4 * 7 + 32 + 7) / 8;
}

// tiers 2 and 3: one block per session (64 threads over shared memory: sessions of 6 .. 40 events keep few
// lanes busy and a wide block mostly waits at its barriers - profiles/r01_cand_v2; 256 threads over the slab)
// TIER 0: 128-thread shared memory, 1: 256-thread shared memory, 2: 256 threads over the global slab
template <int TIER, int T>
__global__ void __launch_bounds__(T) candidates_block_kernel(const CandParams p, int64_t g_lcap, int64_t g_mcap) {
  constexpr bool GLOBAL = TIER == 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_item;
  __shared__ uint64_t s_gmax[T];
  __shared__ __align__(8) int32_t s_scal[8];
  Work w;
  if (GLOBAL) w = carve_rt((unsigned char*)(p.slab + (int64_t)blockIdx.x * p.slab_words), g_lcap, g_mcap, s_scal);
  else if (TIER == 1) w = carve_rt(smem_raw, X_LCAP, X_MCAP, s_scal);
  else w = carve_rt(smem_raw, B_LCAP, B_MCAP, s_scal);
  const uint32_t* list = GLOBAL ? p.list_global : TIER == 1 ? p.list_large : p.list_block;
  const uint32_t n_items = p.counters[GLOBAL ? 1 : TIER == 1 ? 4 : 0];
  if (n_items == 0) return;
  clear_table<T>(w, threadIdx.x);
  while (true) {
    if (threadIdx.x == 0) s_item = atomicAdd(&p.counters[GLOBAL ? 3 : TIER == 1 ? 5 : 2], 1u);
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    process_session<T>(p, (int64_t)list[item], threadIdx.x, w, s_gmax);
    __syncthreads();
  }
}

static int spec_max_k_sum(const OttoCandidateSpec* sp) {
  int best = 0;
  for (int tg = 0; tg < sp->n_targets; ++tg) {
    int sum = 0;
    for (int i = 0; i < sp->target_n_sources[tg]; ++i) sum += sp->table_k[sp->source_table[sp->target_sources[tg][i]]];
    if (sum > best) best = sum;
  }
  return best;
}

static int check_cand_spec(const OttoCandidateSpec* sp) {
  if (!sp) { otto_set_error("candidate spec is NULL"); return OTTO_EINVAL; }
  if (sp->n_tables < 1 || sp->n_tables > OTTO_MAX_TABLES || sp->n_sources < 1 || sp->n_sources > OTTO_MAX_SOURCES ||
      sp->n_targets < 1 || sp->n_targets > OTTO_MAX_TARGETS) { otto_set_error("bad table / source / target count"); return OTTO_EINVAL; }
  if (sp->top_n < 1 || sp->top_n > 4096) { otto_set_error("top_n must be in [1, 4096]"); return OTTO_EINVAL; }
  for (int i = 0; i < sp->n_tables; ++i)
    if (sp->table_k[i] < 1 || sp->table_k[i] > OTTO_MAX_K) { otto_set_error("table_k must be in [1, 32]"); return OTTO_EINVAL; }
  for (int i = 0; i < sp->n_sources; ++i)
    if (sp->source_table[i] < 0 || sp->source_table[i] >= sp->n_tables || sp->source_hist[i] < 0 || sp->source_hist[i] > 3) {
      otto_set_error("bad source %d", i);
      return OTTO_EINVAL;
    }
  for (int tg = 0; tg < sp->n_targets; ++tg) {
    if (sp->target_n_sources[tg] < 1 || sp->target_n_sources[tg] > OTTO_MAX_SOURCES) { otto_set_error("bad target %d", tg); return OTTO_EINVAL; }
    for (int i = 0; i < sp->target_n_sources[tg]; ++i)
      if (sp->target_sources[tg][i] < 0 || sp->target_sources[tg][i] >= sp->n_sources) { otto_set_error("bad target %d", tg); return OTTO_EINVAL; }
  }
  return OTTO_OK;
}

constexpr int GLOBAL_BLOCKS = 296;

extern "C" int64_t otto_candidates_scratch_bytes(int64_t n_sessions, int32_t max_session_len, const OttoCandidateSpec* spec) {
  if (check_cand_spec(spec)) return -1;
  if ((int64_t)max_session_len * spec_max_k_sum(spec) >= 65535) {
    otto_set_error("session too long: positions are 16-bit (max_session_len * sum of k must be < 65535)");
    return -1;
  }
  const int64_t lists = align_up((n_sessions + 1) * 4, 256) * 3 + 256;
  return lists + GLOBAL_BLOCKS * global_slab_words(max_session_len, spec_max_k_sum(spec)) * 8 + 256;
}

extern "C" int otto_candidates(const OttoSessions* sessions, int32_t max_session_len, const OttoCandidateSpec* spec,
                               void* scratch, int64_t scratch_bytes, const OttoCandidates* out, void* stream) {
  int rc = check_cand_spec(spec);
  if (rc) return rc;
  if (!sessions || !out) { otto_set_error("NULL argument"); return OTTO_EINVAL; }
  const int64_t need = otto_candidates_scratch_bytes(sessions->n_sessions, max_session_len, spec);
  if (need < 0) return OTTO_EINVAL;
  if (!scratch || scratch_bytes < need) { otto_set_error("candidate scratch too small: need %lld bytes", (long long)need); return OTTO_ENOSPC; }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t S = sessions->n_sessions;
  if (S == 0) return OTTO_OK;
  CandParams p;
  p.off = sessions->session_offsets;
  p.aid = sessions->aid;
  p.type = sessions->type;
  p.n_sessions = S;
  p.spec = *spec;
  p.out_aid = out->aid;
  p.out_score = out->score;
  p.out_len = out->len;
  char* sc = (char*)scratch;
  const int64_t list_bytes = align_up((S + 1) * 4, 256);
  p.list_block = (uint32_t*)sc;
  p.list_global = (uint32_t*)(sc + list_bytes);
  p.list_large = (uint32_t*)(sc + 2 * list_bytes);
  p.counters = (uint32_t*)(sc + 3 * list_bytes);
  p.slab = (uint64_t*)(sc + 3 * list_bytes + 256);
  p.max_k_sum = spec_max_k_sum(spec);
  p.max_len = max_session_len;
  p.slab_words = global_slab_words(max_session_len, p.max_k_sum);
  // targets with the same ordered source list give the same lists (ranker/covisitation_candidate_generation.py
  // :133,:138: carts and orders are the same concatenation): compute once, copy the slab
  int dup_of[OTTO_MAX_TARGETS];
  p.n_run = 0;
  for (int tg = 0; tg < spec->n_targets; ++tg) {
    dup_of[tg] = -1;
    for (int u = 0; u < tg && dup_of[tg] < 0; ++u) {
      bool same = spec->target_n_sources[u] == spec->target_n_sources[tg];
      for (int i = 0; same && i < spec->target_n_sources[tg]; ++i) same = spec->target_sources[u][i] == spec->target_sources[tg][i];
      if (same) dup_of[tg] = dup_of[u] >= 0 ? dup_of[u] : u;
    }
    if (dup_of[tg] < 0) p.run_target[p.n_run++] = tg;
  }
  CUDA_TRY(cudaMemsetAsync(p.counters, 0, 64, st));
  int dev = 0, n_sm = 148;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  {
    constexpr size_t smem = CAND_WARPS * work_bytes<W_LCAP, W_MCAP>();
    CUDA_TRY(cudaFuncSetAttribute(candidates_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ceil_div(S, CAND_WARPS);
    if (blocks > (int64_t)n_sm * 40) blocks = (int64_t)n_sm * 40;
    candidates_warp_kernel<<<(unsigned)blocks, CAND_WARPS * 32, smem, st>>>(p);
    LAUNCH_CHECK();
  }
  {
    constexpr size_t smem = work_bytes<B_LCAP, B_MCAP>();
    CUDA_TRY(cudaFuncSetAttribute(candidates_block_kernel<0, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    candidates_block_kernel<0, 128><<<n_sm * 5, 128, smem, st>>>(p, 0, 0);
    LAUNCH_CHECK();
  }
  {
    constexpr size_t smem = work_bytes<X_LCAP, X_MCAP>();
    CUDA_TRY(cudaFuncSetAttribute(candidates_block_kernel<1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    candidates_block_kernel<1, 256><<<n_sm, 256, smem, st>>>(p, 0, 0);
    LAUNCH_CHECK();
  }
  candidates_block_kernel<2, 256><<<GLOBAL_BLOCKS, 256, 0, st>>>(p, global_lcap(max_session_len), global_mcap(max_session_len, p.max_k_sum));
  LAUNCH_CHECK();
  for (int tg = 0; tg < spec->n_targets; ++tg) {
    if (dup_of[tg] < 0) continue;
    const int64_t slab = S * spec->top_n;
    CUDA_TRY(cudaMemcpyAsync(out->aid + tg * slab, out->aid + dup_of[tg] * slab, slab * 4, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(out->score + tg * slab, out->score + dup_of[tg] * slab, slab * 4, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(out->len + tg * S, out->len + dup_of[tg] * S, S * 4, cudaMemcpyDeviceToDevice, st));
  }
  return OTTO_OK;
}

// covisitation/inference.py:238-243: history + votes[:n - |H|] + popular[:n - len], cut to n
__global__ void assemble_kernel(const int32_t* __restrict__ off, const int32_t* __restrict__ aid, int64_t S, int n_targets,
                                int top_n, const int32_t* __restrict__ cand_aid, const int32_t* __restrict__ cand_len,
                                const int32_t* __restrict__ popular, int n_pop, int n, int32_t* __restrict__ pred,
                                uint8_t* __restrict__ long_session) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int32_t beg = off[s], end = off[s + 1];
  for (int tg = 0; tg < n_targets; ++tg) {
    int32_t* out = pred + ((int64_t)tg * S + s) * n;
    int len = 0, uniq = 0;
    // unique aids, most recent first.  Only the first n matter (and whether there are at least n), so an aid is
    // checked against the <= n kept so far and the walk stops at n: a 458-event session costs a few hundred
    // compares instead of 458^2 / 2, which used to stall its whole warp.
    for (int32_t i = end - 1; i >= beg && len < n; --i) {
      const int32_t a = aid[i];
      bool seen = false;
      for (int j = 0; j < len; ++j)
        if (out[j] == a) { seen = true; break; }
      if (!seen) out[len++] = a;
    }
    uniq = len;     // == n means "n or more"
    if (tg == 0 && long_session) long_session[s] = uniq >= n;
    // sorted_aids[:n - len(unique)]: only when the history is shorter than n
    const int32_t* ca = cand_aid + ((int64_t)tg * S + s) * top_n;
    const int cl = cand_len[(int64_t)tg * S + s];
    const int take = uniq < n ? min(cl, n - uniq) : 0;
    for (int r = 0; r < take; ++r) out[len++] = ca[r];
    // most_frequent[:n - len(predictions)]
    const int fill = min(n_pop, n - len);
    for (int r = 0; r < fill; ++r) out[len + r] = popular[tg * n_pop + r];
    len += fill > 0 ? fill : 0;
    for (int r = len; r < n; ++r) out[r] = -1;
  }
}

extern "C" int otto_assemble_predictions(const OttoSessions* sessions, const OttoCandidates* cand, int32_t n_targets,
                                         int32_t top_n, const int32_t* popular, int32_t n_popular, int32_t n,
                                         int32_t* pred, uint8_t* long_session, void* stream) {
  if (!sessions || !cand || !pred || n < 1 || n_targets < 1) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  const int64_t S = sessions->n_sessions;
  if (S == 0) return OTTO_OK;
  assemble_kernel<<<(unsigned)ceil_div(S, 128), 128, 0, (cudaStream_t)stream>>>(
      sessions->session_offsets, sessions->aid, S, n_targets, top_n, cand->aid, cand->len, popular, n_popular, n, pred,
      long_session);
  LAUNCH_CHECK();
  return OTTO_OK;
}

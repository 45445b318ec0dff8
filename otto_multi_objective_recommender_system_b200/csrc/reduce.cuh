// Accumulate + per-aid top-K: bins of 8-byte pair records -> rows of the top-K table.
//
// Replaces builder steps 7-8 of SURVEY.md Appendix A: groupby(['aid_x','aid_y']).wgt.sum() followed by
// the stable sort (aid_x asc, wgt desc) and cumcount() < K, whose tie-break is aid_y ascending.
//
// A bin is either a whole aid_x row or one aid_y-hash slice of a hot row.  Its records are streamed once
// from HBM into an open-addressing hash table in shared memory keyed by aid_y (atomicCAS claim, atomicAdd
// of the integer payloads), so the accumulated weights are exact integers (count, sum of ts_x - ts_min,
// or sum of type weights) regardless of arrival order; the float weight is formed once per distinct pair:
//   time  wgt = float(count + 3 * tsum / (ts_max - ts_min))   (fp64 -> fp32)
//   type / unit  wgt = float(sum)                               (exact below 2^24)
// Selection key = (float bits of wgt) << 32 | ~aid_y, so a plain max implements (wgt desc, aid_y asc).
//
// Top-K selection (round-1 profile: K rounds of warp arg-max cost 46 warp-instructions per record at 8
// active lanes).  Now: every thread takes the max key of its slots; the K-th largest of a warp's 32 lane
// maxima is a lower bound T on the K-th largest key of the bin (>= K entries reach it); a second sweep
// pushes the entries >= T (about 1.5 K of them) into a small candidate list that one warp rank-sorts.
// If an adversarial layout overflows the list, the old exact K-round selection runs instead.
//
// Three size classes share the code: a warp with a 512-slot table per bin, a 128-thread block (2048
// slots) and a 256-thread block (4096 slots) that falls back to several aid_y-hash passes when a bin
// exceeds its table.  Slices of split rows write partial top-K lists; merge_split_rows() picks the final
// K (slices hold disjoint aid_y, so the merge is a pure selection).
#pragma once
#include "common.cuh"

struct ReduceParams {
  OttoPairSegment seg[OTTO_MAX_SEGMENTS];
  int32_t n_seg;
  const uint32_t* bin_x;     // [B] bin -> aid_x (global bin ids)
  const uint32_t* bin_base;  // [A + 1]
  int64_t bin_lo, bin_hi;    // this call's bins
  int32_t aid_lo, aid_hi;
  int32_t k;
  int32_t time_mode;
  double w_scale;            // 3 / (ts_max - ts_min)
  int32_t* out_y;
  float* out_w;
  int32_t* out_len;
  uint32_t* out_cnt;
  uint64_t* out_tsum;
  // partial top-K lists of split-row slices, slot = 2 * extra(x) + j (see partial_slot)
  uint64_t* p_key;
  uint64_t* p_sum;
  uint32_t* p_cnt;
  int32_t* p_len;
  // work lists for the block kernels
  uint32_t* list_m;
  uint32_t* list_l;
  uint32_t* counters;        // [0] n_m  [1] n_l  [2] next_m  [3] next_l
  unsigned long long* stats; // [0] distinct  [1] checksum  [2] overflow  [3] slow-path selections
};

constexpr uint32_t SMALL_MAX = 256;    // records: warp kernel, 512-slot table
constexpr uint32_t MEDIUM_MAX = 1024;  // records: 128-thread kernel, 2048-slot table
constexpr uint32_t LARGE_SLOTS = 4096; // 256-thread kernel; more than LARGE_CAP records -> multi-pass
constexpr uint32_t LARGE_CAP = 3072;

__device__ __forceinline__ int64_t partial_slot(const ReduceParams& p, uint32_t x, uint32_t j) {
  const int64_t extra = ((int64_t)p.bin_base[x] - x) - (p.bin_lo - p.aid_lo);
  return 2 * extra + j;
}

__device__ __forceinline__ uint32_t bin_records(const ReduceParams& p, int64_t b) {
  uint32_t n = 0;
  for (int s = 0; s < p.n_seg; ++s) n += (uint32_t)(p.seg[s].offsets[b - p.bin_lo + 1] - p.seg[s].offsets[b - p.bin_lo]);
  return n;
}

template <bool TIME, typename SumT>
struct Table {
  uint32_t* keys;
  SumT* sum;
  uint32_t* cnt;  // TIME only
  uint32_t mask;  // slots - 1

  __device__ __forceinline__ void clear(uint32_t tid, uint32_t nthreads) {
    for (uint32_t h = tid; h <= mask; h += nthreads) {
      keys[h] = KEY_EMPTY;
      sum[h] = 0;
      if (TIME) cnt[h] = 0;
    }
  }
  // returns false on overflow
  __device__ __forceinline__ bool insert(uint32_t y, uint32_t v) {
    uint32_t h = hash32(y) & mask;
    for (uint32_t probe = 0; probe <= mask; ++probe) {
      const uint32_t prev = atomicCAS(&keys[h], KEY_EMPTY, y);
      if (prev == KEY_EMPTY || prev == y) {
        if (TIME) atomicAdd(&cnt[h], 1u);
        atomicAdd(&sum[h], (SumT)v);
        return true;
      }
      h = (h + 1) & mask;
    }
    return false;
  }
  __device__ __forceinline__ float weight(uint32_t h, double w_scale) const {
    if (TIME) return (float)((double)cnt[h] + w_scale * (double)sum[h]);
    return (float)sum[h];
  }
  // selection key of slot h, 0 when empty or already taken
  __device__ __forceinline__ uint64_t key(uint32_t h, double w_scale) const {
    const uint32_t y = keys[h];
    if (y & KEY_TAKEN) return 0;
    return ((uint64_t)__float_as_uint(weight(h, w_scale)) << 32) | (uint32_t)(~y);
  }
};

// Candidate lists in shared memory: key (0 = none), cnt, sum.
struct Cands {
  uint64_t* key;
  uint64_t* sum;
  uint32_t* cnt;
};

// K-th largest of the 32 lane values (0 if fewer than k lanes are non-zero). Keys are distinct or 0.
__device__ __forceinline__ uint64_t warp_kth_largest(uint64_t v, int k) {
  const uint32_t lane = lane_id();
  int rank = 0;
#pragma unroll 8
  for (int l = 0; l < 32; ++l) {
    const uint64_t o = shfl_u64(v, l);
    rank += (o > v) || (o == v && l < (int)lane);
  }
  const uint32_t m = __ballot_sync(FULL_MASK, rank == k - 1);
  return shfl_u64(v, __ffs(m) - 1);
}

// One warp: ranks the n_c candidates (distinct non-zero keys, unordered) and hands every candidate with
// rank < k to `emit(rank, key, cnt, sum)`.  Returns min(n_c, k).
template <typename Emit>
__device__ __forceinline__ int warp_rank_emit(const Cands& c, int n_c, int k, Emit emit) {
  const uint32_t lane = lane_id();
  for (int i = lane; i < n_c; i += 32) {
    const uint64_t mine = c.key[i];
    int rank = 0;
    for (int j = 0; j < n_c; ++j) rank += c.key[j] > mine;
    if (rank < k) emit(rank, mine, c.cnt[i], c.sum[i]);
  }
  return n_c < k ? n_c : k;
}

// ---- exact K-round selection (slow path, kept for candidate-list overflow) ----
template <bool TIME, typename SumT>
__device__ __forceinline__ int warp_select_table_slow(Table<TIME, SumT>& t, uint32_t lo, uint32_t hi, uint32_t stride_lanes,
                                                      int k, double w_scale, Cands dst, int dst_base) {
  const uint32_t lane = lane_id();
  (void)stride_lanes;
  uint64_t best = 0;
  uint32_t best_h = 0;
  for (uint32_t h = lo + lane; h < hi; h += 32) {
    const uint64_t kk = t.key(h, w_scale);
    if (kk > best) { best = kk; best_h = h; }
  }
  int found = 0;
  for (; found < k; ++found) {
    const uint64_t m = warp_max_u64(best);
    if (m == 0) break;
    if (best == m) {
      dst.key[dst_base + found] = m;
      dst.sum[dst_base + found] = (uint64_t)t.sum[best_h];
      dst.cnt[dst_base + found] = TIME ? t.cnt[best_h] : 0u;
      t.keys[best_h] |= KEY_TAKEN;
      best = 0;
      for (uint32_t h = lo + lane; h < hi; h += 32) {
        const uint64_t kk = t.key(h, w_scale);
        if (kk > best) { best = kk; best_h = h; }
      }
    }
  }
  __syncwarp();
  return found;
}

// Writes one finished entry of bin b: a table row (ordinary bin) or a partial list (slice of a split row).
struct BinOut {
  bool whole;      // ordinary bin: write the table row
  int64_t row;     // x * k   or   slot * k
  uint32_t x;
};

__device__ __forceinline__ BinOut bin_out(const ReduceParams& p, int64_t b) {
  BinOut o;
  o.x = p.bin_x[b];
  const uint32_t bb0 = p.bin_base[o.x];
  o.whole = (p.bin_base[o.x + 1] - bb0) == 1;
  o.row = o.whole ? (int64_t)o.x * p.k : partial_slot(p, o.x, (uint32_t)(b - bb0)) * p.k;
  return o;
}

__device__ __forceinline__ void emit_entry(const ReduceParams& p, const BinOut& o, int r, uint64_t kk, uint32_t cnt,
                                           uint64_t sum) {
  if (o.whole) {
    p.out_y[o.row + r] = (int32_t)(~(uint32_t)kk);
    p.out_w[o.row + r] = __uint_as_float((uint32_t)(kk >> 32));
    if (p.out_cnt) p.out_cnt[o.row + r] = cnt;
    if (p.out_tsum) p.out_tsum[o.row + r] = sum;
  } else {
    p.p_key[o.row + r] = kk;
    p.p_sum[o.row + r] = sum;
    p.p_cnt[o.row + r] = cnt;
  }
}

// pads the row beyond `found` and stores the length (one warp)
__device__ __forceinline__ void emit_finish(const ReduceParams& p, const BinOut& o, int found) {
  const uint32_t lane = lane_id();
  if (o.whole) {
    for (int r = found + (int)lane; r < p.k; r += 32) {
      p.out_y[o.row + r] = -1;
      p.out_w[o.row + r] = 0.f;
      if (p.out_cnt) p.out_cnt[o.row + r] = 0u;
      if (p.out_tsum) p.out_tsum[o.row + r] = 0ull;
    }
    if (lane == 0) p.out_len[o.x] = found;
  } else if (lane == 0) {
    p.p_len[o.row / p.k] = found;
  }
}

// ---- small bins: one warp per bin ----
constexpr int SMALL_WARPS = 8;
constexpr uint32_t SMALL_SLOTS = 512;
constexpr int SMALL_CANDS = 96;
// per warp: cand key[96] sum[96] (u64) | table sum[512] keys[512] cnt[512] (u32) | cand cnt[96] | counter
constexpr uint32_t SMALL_PER_WARP = SMALL_CANDS * 16 + SMALL_SLOTS * 12 + SMALL_CANDS * 4 + 16;

template <bool TIME>
__global__ void __launch_bounds__(SMALL_WARPS * 32) reduce_small_kernel(const ReduceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned char* base = smem_raw + warp * SMALL_PER_WARP;
  Cands c;
  c.key = (uint64_t*)base;
  c.sum = c.key + SMALL_CANDS;
  Table<TIME, uint32_t> t;
  t.sum = (uint32_t*)(c.sum + SMALL_CANDS);
  t.keys = t.sum + SMALL_SLOTS;
  t.cnt = t.keys + SMALL_SLOTS;
  c.cnt = t.cnt + SMALL_SLOTS;
  uint32_t* n_cand = c.cnt + SMALL_CANDS;

  uint32_t st_occ = 0, st_slow = 0;
  uint64_t st_pay = 0;
  bool overflow = false;
  const int64_t n_warps = (int64_t)gridDim.x * SMALL_WARPS;
  for (int64_t b = p.bin_lo + (int64_t)blockIdx.x * SMALL_WARPS + warp; b < p.bin_hi; b += n_warps) {
    const uint32_t n = bin_records(p, b);
    if (n > SMALL_MAX) {  // hand over to a block kernel
      if (lane == 0) {
        if (n <= MEDIUM_MAX) p.list_m[atomicAdd(&p.counters[0], 1u)] = (uint32_t)(b - p.bin_lo);
        else p.list_l[atomicAdd(&p.counters[1], 1u)] = (uint32_t)(b - p.bin_lo);
      }
      continue;
    }
    const BinOut o = bin_out(p, b);
    if (n == 0) {
      emit_finish(p, o, 0);
      continue;
    }
    uint32_t slots = 32;
    while (slots < 2 * n) slots <<= 1;
    t.mask = slots - 1;
    t.clear(lane, 32);
    if (lane == 0) *n_cand = 0;
    __syncwarp();
    for (int s = 0; s < p.n_seg; ++s) {
      const uint64_t o0 = p.seg[s].offsets[0];
      const uint64_t beg = p.seg[s].offsets[b - p.bin_lo] - o0, end = p.seg[s].offsets[b - p.bin_lo + 1] - o0;
      const uint2* rec = (const uint2*)p.seg[s].records;
      for (uint64_t i = beg + lane; i < end; i += 32) {
        const uint2 r = ld_stream_u2(rec + i);
        if (!t.insert(r.x, r.y)) overflow = true;
      }
    }
    __syncwarp();
    // sweep 1: lane maxima -> threshold
    uint64_t best = 0;
    for (uint32_t h = lane; h < slots; h += 32) {
      if (t.keys[h] != KEY_EMPTY) {
        ++st_occ;
        st_pay += TIME ? (uint64_t)t.cnt[h] : (uint64_t)t.sum[h];
        const uint64_t kk = t.key(h, p.w_scale);
        best = kk > best ? kk : best;
      }
    }
    const uint64_t thr = warp_kth_largest(best, p.k);
    // sweep 2: candidates
    for (uint32_t h = lane; h < slots; h += 32) {
      if (t.keys[h] != KEY_EMPTY) {
        const uint64_t kk = t.key(h, p.w_scale);
        if (kk >= thr) {
          const uint32_t at = atomicAdd(n_cand, 1u);
          if (at < SMALL_CANDS) {
            c.key[at] = kk;
            c.sum[at] = (uint64_t)t.sum[h];
            c.cnt[at] = TIME ? t.cnt[h] : 0u;
          }
        }
      }
    }
    __syncwarp();
    int n_c = (int)*n_cand;
    if (n_c > SMALL_CANDS) {  // adversarial layout: exact K-round selection straight from the table
      ++st_slow;
      n_c = warp_select_table_slow<TIME, uint32_t>(t, 0, slots, 32, p.k, p.w_scale, c, 0);
    }
    const int found = warp_rank_emit(c, n_c, p.k, [&](int r, uint64_t kk, uint32_t cnt, uint64_t sum) {
      emit_entry(p, o, r, kk, cnt, sum);
    });
    emit_finish(p, o, found);
    __syncwarp();
  }
  // one stats update per warp
  for (int off = 16; off > 0; off >>= 1) {
    st_occ += __shfl_xor_sync(FULL_MASK, st_occ, off);
    st_slow += __shfl_xor_sync(FULL_MASK, st_slow, off);
    st_pay += shfl_u64(st_pay, lane ^ off);
  }
  if (lane == 0 && (st_occ || st_pay)) {
    atomicAdd(&p.stats[0], (unsigned long long)st_occ);
    atomicAdd(&p.stats[1], (unsigned long long)st_pay);
    if (st_slow) atomicAdd(&p.stats[3], (unsigned long long)(st_slow / 32));
  }
  if (overflow) atomicOr(&p.stats[2], 1ull);
}

// ---- medium / large bins: one block per bin, work taken from a list through an atomic cursor ----
constexpr int BLOCK_CANDS = 256;

template <bool TIME, int THREADS, uint32_t SLOTS, typename SumT, bool LARGE>
__global__ void __launch_bounds__(THREADS) reduce_block_kernel(const ReduceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int WARPS = THREADS / 32;
  constexpr int NC = BLOCK_CANDS + OTTO_MAX_K;     // candidates + room for the carried best list
  __shared__ uint32_t s_item, s_ncand;
  __shared__ uint64_t s_thr[WARPS];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  // carve: cand key[NC] sum[NC] | best key[K] sum[K] (u64) | table sum | keys | cnt | cand cnt | best cnt
  Cands c, best;
  c.key = (uint64_t*)smem_raw;
  c.sum = c.key + NC;
  best.key = c.sum + NC;
  best.sum = best.key + OTTO_MAX_K;
  Table<TIME, SumT> t;
  t.sum = (SumT*)(best.sum + OTTO_MAX_K);
  t.keys = (uint32_t*)(t.sum + SLOTS);
  t.cnt = t.keys + SLOTS;
  c.cnt = t.cnt + (TIME ? SLOTS : 0);
  best.cnt = c.cnt + NC;

  const uint32_t* list = LARGE ? p.list_l : p.list_m;
  const uint32_t n_items = p.counters[LARGE ? 1 : 0];
  uint32_t st_occ = 0, st_slow = 0;
  uint64_t st_pay = 0;
  bool overflow = false;
  while (true) {
    if (threadIdx.x == 0) s_item = atomicAdd(&p.counters[LARGE ? 3 : 2], 1u);
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    const int64_t b = p.bin_lo + list[item];
    const uint32_t n = bin_records(p, b);
    const uint32_t n_pass = (LARGE && n > LARGE_CAP) ? (n + LARGE_CAP - 1) / LARGE_CAP : 1;
    uint32_t slots = SLOTS;
    if (n_pass == 1) {
      slots = 64;
      while (slots < n + n / 3 + 1) slots <<= 1;
      if (slots > SLOTS) slots = SLOTS;
    }
    t.mask = slots - 1;
    const BinOut o = bin_out(p, b);
    int n_best = 0;  // meaningful in warp 0
    for (uint32_t pass = 0; pass < n_pass; ++pass) {
      const bool last = pass + 1 == n_pass;
      t.clear(threadIdx.x, THREADS);
      if (threadIdx.x == 0) s_ncand = 0;
      __syncthreads();
      for (int s = 0; s < p.n_seg; ++s) {
        const uint64_t o0 = p.seg[s].offsets[0];
        const uint64_t beg = p.seg[s].offsets[b - p.bin_lo] - o0, end = p.seg[s].offsets[b - p.bin_lo + 1] - o0;
        const uint2* rec = (const uint2*)p.seg[s].records;
        for (uint64_t i = beg + threadIdx.x; i < end; i += THREADS) {
          const uint2 r = ld_stream_u2(rec + i);
          if (n_pass > 1 && __umulhi(hash32b(r.x), n_pass) != pass) continue;
          if (!t.insert(r.x, r.y)) overflow = true;
        }
      }
      __syncthreads();
      // sweep 1: thread maxima -> per-warp thresholds -> block threshold
      uint64_t tbest = 0;
      for (uint32_t h = threadIdx.x; h < slots; h += THREADS) {
        if (t.keys[h] != KEY_EMPTY) {
          ++st_occ;
          st_pay += TIME ? (uint64_t)t.cnt[h] : (uint64_t)t.sum[h];
          const uint64_t kk = t.key(h, p.w_scale);
          tbest = kk > tbest ? kk : tbest;
        }
      }
      const uint64_t wthr = warp_kth_largest(tbest, p.k);
      if (lane == 0) s_thr[warp] = wthr;
      __syncthreads();
      uint64_t thr = 0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) thr = s_thr[w] > thr ? s_thr[w] : thr;
      // sweep 2: candidates
      for (uint32_t h = threadIdx.x; h < slots; h += THREADS) {
        if (t.keys[h] != KEY_EMPTY) {
          const uint64_t kk = t.key(h, p.w_scale);
          if (kk >= thr) {
            const uint32_t at = atomicAdd(&s_ncand, 1u);
            if (at < BLOCK_CANDS) {
              c.key[at] = kk;
              c.sum[at] = (uint64_t)t.sum[h];
              c.cnt[at] = TIME ? t.cnt[h] : 0u;
            }
          }
        }
      }
      __syncthreads();
      int n_c = (int)s_ncand;
      if (n_c > BLOCK_CANDS) {
        // adversarial layout: exact K-round selection per warp slice, lists land in c at warp * K
        if (lane == 0 && warp == 0) ++st_slow;
        const uint32_t per = slots / WARPS;
        const int f = warp_select_table_slow<TIME, SumT>(t, warp * per, (warp + 1) * per, 32, p.k, p.w_scale, c,
                                                         warp * OTTO_MAX_K);
        for (int r = f + (int)lane; r < OTTO_MAX_K; r += 32) c.key[warp * OTTO_MAX_K + r] = 0;
        __syncthreads();
        // compact the non-zero keys to the front (one warp; WARPS * 32 <= BLOCK_CANDS)
        if (warp == 0) {
          int w = 0;
          for (int i0 = 0; i0 < WARPS * OTTO_MAX_K; i0 += 32) {
            const int i = i0 + lane;
            const uint64_t kk = c.key[i];
            const uint64_t ss = c.sum[i];
            const uint32_t cc = c.cnt[i];
            const uint32_t m = __ballot_sync(FULL_MASK, kk != 0);
            __syncwarp();
            if (kk != 0) {
              const int at = w + __popc(m & lanemask_lt());
              c.key[at] = kk; c.sum[at] = ss; c.cnt[at] = cc;
            }
            w += __popc(m);
            __syncwarp();
          }
          n_c = w;
        }
      }
      if (warp == 0) {
        n_c = __shfl_sync(FULL_MASK, n_c, 0);
        // entries carried from earlier passes join the candidates
        for (int r = lane; r < n_best; r += 32) {
          c.key[n_c + r] = best.key[r];
          c.sum[n_c + r] = best.sum[r];
          c.cnt[n_c + r] = best.cnt[r];
        }
        __syncwarp();
        n_c += n_best;
        if (last) {
          const int found = warp_rank_emit(c, n_c, p.k, [&](int r, uint64_t kk, uint32_t cnt, uint64_t sum) {
            emit_entry(p, o, r, kk, cnt, sum);
          });
          emit_finish(p, o, found);
        } else {
          n_best = warp_rank_emit(c, n_c, p.k, [&](int r, uint64_t kk, uint32_t cnt, uint64_t sum) {
            best.key[r] = kk; best.sum[r] = sum; best.cnt[r] = cnt;
          });
          __syncwarp();
        }
      }
      __syncthreads();
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    st_occ += __shfl_xor_sync(FULL_MASK, st_occ, off);
    st_slow += __shfl_xor_sync(FULL_MASK, st_slow, off);
    st_pay += shfl_u64(st_pay, lane ^ off);
  }
  if (lane == 0 && (st_occ || st_pay)) {
    atomicAdd(&p.stats[0], (unsigned long long)st_occ);
    atomicAdd(&p.stats[1], (unsigned long long)st_pay);
    if (st_slow) atomicAdd(&p.stats[3], (unsigned long long)st_slow);
  }
  if (overflow) atomicOr(&p.stats[2], 1ull);
}

template <bool TIME, int THREADS, uint32_t SLOTS, typename SumT>
constexpr size_t reduce_block_smem() {
  constexpr int NC = BLOCK_CANDS + OTTO_MAX_K;
  return (size_t)NC * 16 + OTTO_MAX_K * 16 + SLOTS * sizeof(SumT) + SLOTS * 4 + (TIME ? SLOTS * 4 : 0) + NC * 4 +
         OTTO_MAX_K * 4;
}

// ---- split rows: merge the slices' partial lists (disjoint aid_y) into the final row ----
__global__ void __launch_bounds__(256) merge_split_rows_kernel(const ReduceParams p) {
  const uint32_t lane = lane_id();
  const int64_t n_warps = (int64_t)gridDim.x * 8;
  const int k = p.k;
  for (int64_t x0 = p.aid_lo + ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32; x0 < p.aid_hi; x0 += n_warps * 32) {
    const int64_t xl = x0 + lane;
    uint32_t nb = 1;
    if (xl < p.aid_hi) nb = p.bin_base[xl + 1] - p.bin_base[xl];
    uint32_t todo = __ballot_sync(FULL_MASK, nb > 1);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t x = (uint32_t)(x0 + src);
      const uint32_t nbx = __shfl_sync(FULL_MASK, nb, src);
      const int64_t slot0 = partial_slot(p, x, 0);
      const int n_c = (int)nbx * k;
      uint64_t best = 0;
      int best_i = 0;
      for (int i = lane; i < n_c; i += 32) {
        const int j = i / k, r = i % k;
        const uint64_t kk = r < p.p_len[slot0 + j] ? p.p_key[(slot0 + j) * k + r] : 0;
        if (kk > best) { best = kk; best_i = i; }
      }
      int found = 0;
      const int64_t row = (int64_t)x * k;
      for (; found < k; ++found) {
        const uint64_t m = warp_max_u64(best);
        if (m == 0) break;
        if (best == m) {
          const int64_t at = (slot0 + best_i / k) * k + best_i % k;
          p.out_y[row + found] = (int32_t)(~(uint32_t)m);
          p.out_w[row + found] = __uint_as_float((uint32_t)(m >> 32));
          if (p.out_cnt) p.out_cnt[row + found] = p.p_cnt[at];
          if (p.out_tsum) p.out_tsum[row + found] = p.p_sum[at];
          p.p_key[at] = 0;
          best = 0;
          for (int i = lane; i < n_c; i += 32) {
            const int j = i / k, r = i % k;
            const uint64_t kk = r < p.p_len[slot0 + j] ? p.p_key[(slot0 + j) * k + r] : 0;
            if (kk > best) { best = kk; best_i = i; }
          }
        }
      }
      for (int r = found + (int)lane; r < k; r += 32) {
        p.out_y[row + r] = -1;
        p.out_w[row + r] = 0.f;
        if (p.out_cnt) p.out_cnt[row + r] = 0;
        if (p.out_tsum) p.out_tsum[row + r] = 0;
      }
      if (lane == 0) p.out_len[x] = found;
      __syncwarp();
    }
  }
}

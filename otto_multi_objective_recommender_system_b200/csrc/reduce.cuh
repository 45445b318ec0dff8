// Accumulate + per-aid top-K: bins of 8-byte pair records -> rows of the top-K table.
//
// Replaces builder steps 7-8 of SURVEY.md Appendix A: groupby(['aid_x','aid_y']).wgt.sum() followed by
// the stable sort (aid_x asc, wgt desc) and cumcount() < K, whose tie-break is aid_y ascending.
//
// A bin is either a whole aid_x row or one aid_y-hash slice of a hot row.  Its records are streamed once
// from HBM into an open-addressing hash table in shared memory keyed by aid_y (atomicCAS claim, 32-bit
// atomicAdd of the integer payloads), so the accumulated weights are exact integers (count, sum of
// ts_x - ts_min, or sum of type weights) regardless of arrival order; the float weight is formed once per
// candidate:
//   time  wgt = float(count + 3 * tsum / (ts_max - ts_min))   (fp64 -> fp32)
//   type / unit  wgt = float(sum)                               (exact below 2^24)
// Final order key = (float bits of wgt) << 32 | ~aid_y, so a plain max implements (wgt desc, aid_y asc).
//
// What the round-1 profiles taught (profiles/r01_*): the kernels are instruction-issue bound, not HBM
// bound, so the work per record is what matters:
//   * claimed slots are appended to an "occupied" list, so the sweeps and the clearing touch the d
//     distinct entries instead of all table slots;
//   * 64-bit shared atomicAdd compiles to a CAS spin loop (ATOMS.CAST.SPIN.64); the time sum is kept as a
//     32-bit low word plus a packed word (count in bits 0-23, carries of the low word in bits 24-31);
//   * top-K: sweep 1 takes, per lane, the max of an integer order key I = count * R + 3 * tsum (exactly
//     proportional to the weight; no fp64); the K-th largest of the 32 lane-group maxima is a lower bound
//     T on the K-th largest entry; sweep 2 pushes the entries with I >= T - T * 2^-22 (about 1.5 K of
//     them; the margin covers entries that tie with T after fp32 rounding) into a candidate list, for
//     which the exact float keys are formed and rank-sorted by one warp.  If an adversarial layout
//     overflows the list, an exact K-round selection over the table runs instead.
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
  uint32_t range;            // ts_max - ts_min (time mode)
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
  uint32_t* list_x;
  uint32_t* counters;        // [0] n_m  [1] n_l  [2] next_m  [3] next_l  [4] n_x  [5] next_x
  unsigned long long* stats; // [0] distinct  [1] checksum  [2] overflow  [3] slow-path selections  [4..7] records per tier
};

constexpr uint32_t TINY_MAX = 32;      // records: one step of one warp, no table at all
constexpr uint32_t SMALL_MAX = 256;    // records: warp kernel, 512-slot table
constexpr uint32_t MEDIUM_MAX = 1024;  // records: 128-thread kernel, 2048-slot table
constexpr uint32_t LARGE_MAX = 3072;   // records: 256-thread kernel, 4096-slot table
// beyond: 512-thread kernel, 8192-slot table; single pass up to 3/4 of the slots, else aid_y-hash passes

__device__ __forceinline__ int64_t partial_slot(const ReduceParams& p, uint32_t x, uint32_t j) {
  const int64_t extra = ((int64_t)p.bin_base[x] - x) - (p.bin_lo - p.aid_lo);
  return 2 * extra + j;
}

__device__ __forceinline__ uint32_t bin_records(const ReduceParams& p, int64_t b) {
  uint32_t n = 0;
  for (int s = 0; s < p.n_seg; ++s) n += (uint32_t)(p.seg[s].offsets[b - p.bin_lo + 1] - p.seg[s].offsets[b - p.bin_lo]);
  return n;
}

// single-segment fast path: the bin's run (NULL when the bin is spread over several segments)
__device__ __forceinline__ const uint2* bin_run(const ReduceParams& p, int64_t b) {
  if (p.n_seg != 1) return nullptr;
  return (const uint2*)p.seg[0].records + (p.seg[0].offsets[b - p.bin_lo] - p.seg[0].offsets[0]);
}

// record i of bin b in the concatenation of the segments' runs (i < bin_records)
__device__ __forceinline__ uint2 bin_record(const ReduceParams& p, int64_t b, uint32_t i) {
  for (int s = 0; s < p.n_seg; ++s) {
    const uint64_t beg = p.seg[s].offsets[b - p.bin_lo], end = p.seg[s].offsets[b - p.bin_lo + 1];
    const uint32_t len = (uint32_t)(end - beg);
    if (i < len) return ld_stream_u2((const uint2*)p.seg[s].records + (beg - p.seg[s].offsets[0]) + i);
    i -= len;
  }
  return make_uint2(KEY_EMPTY, 0);
}

// Open-addressing table over SLOTS = 2^LOG slots.  TIME: hc = count | carries << 24, lo = low 32 bits of
// the time sum.  !TIME: lo = integer weight sum (hc unused).
template <bool TIME, int LOG>
struct Table {
  static constexpr uint32_t SLOTS = 1u << LOG;
  uint32_t* keys;
  uint32_t* lo;
  uint32_t* hc;
  uint16_t* occ;     // claimed slots, in claim order
  uint32_t* n_occ;

  __device__ __forceinline__ void clear_all(uint32_t tid, uint32_t nthreads) {
    for (uint32_t h = tid; h < SLOTS; h += nthreads) {
      keys[h] = KEY_EMPTY;
      lo[h] = 0;
      if (TIME) hc[h] = 0;
    }
    if (tid == 0) *n_occ = 0;
  }
  // resets exactly the claimed slots (n = *n_occ read by the caller after a barrier)
  __device__ __forceinline__ void clear_dirty(uint32_t tid, uint32_t nthreads, uint32_t n) {
    for (uint32_t i = tid; i < n; i += nthreads) {
      const uint32_t h = occ[i];
      keys[h] = KEY_EMPTY;
      lo[h] = 0;
      if (TIME) hc[h] = 0;
    }
  }
  // One record per lane (has = this lane holds one); EVERY lane of the warp must call.  Returns false on
  // overflow.  The probe loop only finds / claims the slot.  Lanes leave it after different numbers of
  // probes, and without the __syncwarp() below the compiler keeps them diverged: profile r01_reduce_v3b
  // shows the payload atomics and the occupied-list append executing with 6.7 of 32 lanes per issue
  // (45 % of all instructions of the kernel).  After reconvergence the append is one atomic per warp.
  __device__ __forceinline__ bool insert(bool has, uint32_t y, uint32_t v) {
    // double hashing: slot from the top bits of a Fibonacci hash, odd stride from a second multiplier.  The
    // profile of the linear-probing version (r01_reduce_v3c) showed ~14 probe iterations per 32-record step
    // (the slowest lane decides), at 16 SASS instructions each: the probe loop was 3/4 of the kernel.
    const uint32_t hm = y * 0x9E3779B1u;
    uint32_t h = hm >> (32 - LOG);
    const uint32_t step = ((y * 0x85EBCA6Bu) >> (32 - LOG)) | 1u;
    uint32_t prev = 0x80000000u;   // neither EMPTY nor an aid (aids are < 2^30)
    if (has) {
#pragma unroll 1
      for (uint32_t probe = SLOTS; probe; --probe) {
        prev = atomicCAS(&keys[h], KEY_EMPTY, y);
        if (prev == KEY_EMPTY || prev == y) break;
        h = (h + step) & (SLOTS - 1);
      }
    }
    const int state = !has ? 0 : prev == KEY_EMPTY ? 2 : (prev == y ? 1 : 0);  // 2 = claimed a fresh slot, 1 = found
    __syncwarp();
    const uint32_t fresh = __ballot_sync(FULL_MASK, state == 2);
    if (fresh) {
      const int leader = __ffs(fresh) - 1;
      uint32_t base = 0;
      if ((int)lane_id() == leader) base = atomicAdd(n_occ, (uint32_t)__popc(fresh));
      base = __shfl_sync(FULL_MASK, base, leader);
      if (state == 2) occ[base + __popc(fresh & lanemask_lt())] = (uint16_t)h;
    }
    if (state != 0) {
      const uint32_t old = atomicAdd(&lo[h], v);
      if (TIME) atomicAdd(&hc[h], 1u + ((old + v < old) ? (1u << 24) : 0u));
    }
    return state != 0 || !has;
  }
  // Two records per lane with their probe sequences interleaved: both CAS are in flight together, and the loop
  // control, the reconvergence, the occupied-list append (one atomic for both) are paid once per 64 records.
  // A lane's two records may carry the same aid_y: the second CAS then finds the first one's claim.
  __device__ __forceinline__ bool insert2(bool has0, uint32_t y0, uint32_t v0, bool has1, uint32_t y1, uint32_t v1) {
    uint32_t h0 = (y0 * 0x9E3779B1u) >> (32 - LOG), h1 = (y1 * 0x9E3779B1u) >> (32 - LOG);
    const uint32_t step0 = ((y0 * 0x85EBCA6Bu) >> (32 - LOG)) | 1u, step1 = ((y1 * 0x85EBCA6Bu) >> (32 - LOG)) | 1u;
    uint32_t prev0 = 0x80000000u, prev1 = 0x80000000u;
    bool p0 = has0, p1 = has1;
#pragma unroll 1
    for (uint32_t probe = SLOTS; probe && (p0 || p1); --probe) {
      if (p0) prev0 = atomicCAS(&keys[h0], KEY_EMPTY, y0);
      if (p1) prev1 = atomicCAS(&keys[h1], KEY_EMPTY, y1);
      if (p0) {
        if (prev0 == KEY_EMPTY || prev0 == y0) p0 = false;
        else h0 = (h0 + step0) & (SLOTS - 1);
      }
      if (p1) {
        if (prev1 == KEY_EMPTY || prev1 == y1) p1 = false;
        else h1 = (h1 + step1) & (SLOTS - 1);
      }
    }
    const int state0 = (!has0 || p0) ? 0 : prev0 == KEY_EMPTY ? 2 : 1;
    const int state1 = (!has1 || p1) ? 0 : prev1 == KEY_EMPTY ? 2 : 1;
    __syncwarp();
    const uint32_t fresh0 = __ballot_sync(FULL_MASK, state0 == 2), fresh1 = __ballot_sync(FULL_MASK, state1 == 2);
    if (fresh0 | fresh1) {
      uint32_t base = 0;
      if (lane_id() == 0) base = atomicAdd(n_occ, (uint32_t)(__popc(fresh0) + __popc(fresh1)));
      base = __shfl_sync(FULL_MASK, base, 0);
      const uint32_t lt = lanemask_lt();
      if (state0 == 2) occ[base + __popc(fresh0 & lt)] = (uint16_t)h0;
      if (state1 == 2) occ[base + __popc(fresh0) + __popc(fresh1 & lt)] = (uint16_t)h1;
    }
    if (state0 != 0) {
      const uint32_t old = atomicAdd(&lo[h0], v0);
      if (TIME) atomicAdd(&hc[h0], 1u + ((old + v0 < old) ? (1u << 24) : 0u));
    }
    if (state1 != 0) {
      const uint32_t old = atomicAdd(&lo[h1], v1);
      if (TIME) atomicAdd(&hc[h1], 1u + ((old + v1 < old) ? (1u << 24) : 0u));
    }
    return (state0 != 0 || !has0) && (state1 != 0 || !has1);
  }
  __device__ __forceinline__ uint32_t count(uint32_t h) const { return TIME ? (hc[h] & 0xffffffu) : 0u; }
  __device__ __forceinline__ uint64_t sum(uint32_t h) const {
    return TIME ? (((uint64_t)(hc[h] >> 24) << 32) | lo[h]) : (uint64_t)lo[h];
  }
  // integer order key, exactly proportional to the weight
  __device__ __forceinline__ uint64_t ikey(uint32_t h, uint32_t range) const {
    return TIME ? (uint64_t)count(h) * range + 3ull * sum(h) : (uint64_t)lo[h];
  }
};

__device__ __forceinline__ uint64_t float_key(bool time_mode, uint32_t y, uint32_t cnt, uint64_t sum, double w_scale) {
  const float w = time_mode ? (float)((double)cnt + w_scale * (double)sum) : (float)sum;
  return ((uint64_t)__float_as_uint(w) << 32) | (uint32_t)(~y);
}

// Candidate lists in shared memory: key (0 = none), cnt, sum.
struct Cands {
  uint64_t* key;
  uint64_t* sum;
  uint32_t* cnt;
};

// Bitonic sort of one u64 per lane, descending (lane 0 ends with the largest).
__device__ __forceinline__ uint64_t warp_sort_desc(uint64_t v) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint32_t olo = __shfl_xor_sync(FULL_MASK, (uint32_t)v, j);
      const uint32_t ohi = __shfl_xor_sync(FULL_MASK, (uint32_t)(v >> 32), j);
      const uint64_t o = ((uint64_t)ohi << 32) | olo;
      const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);
      v = keep_max ? (o > v ? o : v) : (o < v ? o : v);
    }
  }
  return v;
}
// K-th largest of the 32 lane values (0 if fewer than k lanes are non-zero).
__device__ __forceinline__ uint64_t warp_kth_largest(uint64_t v, int k) { return shfl_u64(warp_sort_desc(v), k - 1); }

// candidates must reach T minus the fp32 tie margin
__device__ __forceinline__ uint64_t with_margin(uint64_t thr) { return thr > (thr >> 22) + 1 ? thr - (thr >> 22) - 1 : 0; }

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

// Exact K-round selection over occupied entries [lo, hi) of the occ list (slow path for candidate-list
// overflow).  One warp; results (best first) land in dst[dst_base ..); returns how many.
template <bool TIME, int LOG>
__device__ __forceinline__ int warp_select_slow(Table<TIME, LOG>& t, uint32_t lo, uint32_t hi, int k, double w_scale,
                                                Cands dst, int dst_base) {
  const uint32_t lane = lane_id();
  auto scan = [&](uint64_t& best, uint32_t& best_h) {
    best = 0;
    for (uint32_t i = lo + lane; i < hi; i += 32) {
      const uint32_t h = t.occ[i];
      const uint32_t y = t.keys[h];
      if (y & KEY_TAKEN) continue;
      const uint64_t kk = float_key(TIME, y, t.count(h), t.sum(h), w_scale);
      if (kk > best) { best = kk; best_h = h; }
    }
  };
  uint64_t best;
  uint32_t best_h = 0;
  scan(best, best_h);
  int found = 0;
  for (; found < k; ++found) {
    const uint64_t m = warp_max_u64(best);
    if (m == 0) break;
    if (best == m) {
      dst.key[dst_base + found] = m;
      dst.sum[dst_base + found] = t.sum(best_h);
      dst.cnt[dst_base + found] = t.count(best_h);
      t.keys[best_h] |= KEY_TAKEN;
      scan(best, best_h);
    }
  }
  __syncwarp();
  return found;
}

// Where a finished bin goes: a table row (ordinary bin) or a partial list (slice of a split row).
struct BinOut {
  bool whole;
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

// warp-aggregated append of candidate (y, cnt, sum) with its exact float key; q = this lane has one
template <int CAP>
__device__ __forceinline__ void push_candidate(bool q, uint32_t* n_cand, const Cands& c, bool time_mode, uint32_t y,
                                               uint32_t cnt, uint64_t sum, double w_scale) {
  const uint32_t m = __ballot_sync(FULL_MASK, q);
  if (m == 0) return;
  uint32_t base = 0;
  if (lane_id() == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(n_cand, (uint32_t)__popc(m));
  base = __shfl_sync(FULL_MASK, base, __ffs(m) - 1);
  if (q) {
    const uint32_t at = base + __popc(m & lanemask_lt());
    if (at < (uint32_t)CAP) {
      c.key[at] = float_key(time_mode, y, cnt, sum, w_scale);
      c.sum[at] = sum;
      c.cnt[at] = cnt;
    }
  }
}

// ---- small bins: one warp per bin ----
constexpr int SMALL_WARPS = 4;
constexpr int SMALL_LOG = 9;           // 512 slots
constexpr int SMALL_CANDS = 64;
constexpr int SMALL_DIRECT = 32;       // up to this many distinct entries: skip the threshold
// per warp: cand key[64] sum[64] (u64) | keys[512] lo[512] hc[512] cand cnt[64] counters[4] (u32) | occ[256] (u16)
constexpr uint32_t SMALL_PER_WARP = SMALL_CANDS * 16 + (1u << SMALL_LOG) * 12 + SMALL_CANDS * 4 + 16 + SMALL_MAX * 2;

template <bool TIME>
__global__ void __launch_bounds__(SMALL_WARPS * 32) reduce_small_kernel(const ReduceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned char* base = smem_raw + warp * SMALL_PER_WARP;
  constexpr uint32_t SLOTS = 1u << SMALL_LOG;
  Cands c;
  c.key = (uint64_t*)base;
  c.sum = c.key + SMALL_CANDS;
  Table<TIME, SMALL_LOG> t;
  t.keys = (uint32_t*)(c.sum + SMALL_CANDS);
  t.lo = t.keys + SLOTS;
  t.hc = t.lo + SLOTS;
  c.cnt = t.hc + SLOTS;
  uint32_t* n_cand = c.cnt + SMALL_CANDS;
  t.n_occ = n_cand + 1;
  t.occ = (uint16_t*)(n_cand + 4);
  t.clear_all(lane, 32);
  __syncwarp();

  uint64_t st_occ = 0, st_pay = 0, st_rec = 0;
  uint32_t st_slow = 0;
  bool overflow = false;
  const int64_t n_warps = (int64_t)gridDim.x * SMALL_WARPS;
  // A warp takes 32 consecutive bins at a time: lane l walks the dependent metadata loads of bin c0 + l (record
  // offsets, bin -> aid_x, first bin of the row), so their latency is paid once per 32 bins instead of once per
  // bin (r01: with one chain per bin the kernels were bound by exactly these round trips), then the bins are
  // processed one by one with the metadata coming from registers.
  for (int64_t c0 = p.bin_lo + ((int64_t)blockIdx.x * SMALL_WARPS + warp) * 32; c0 < p.bin_hi; c0 += n_warps * 32) {
    const int64_t bl = c0 + lane;
    const bool inb = bl < p.bin_hi;
    const uint32_t n_l = inb ? bin_records(p, bl) : 0u;
    if (inb && n_l > SMALL_MAX) {  // hand over to a block kernel
      if (n_l <= MEDIUM_MAX) p.list_m[atomicAdd(&p.counters[0], 1u)] = (uint32_t)(bl - p.bin_lo);
      else if (n_l <= LARGE_MAX) p.list_l[atomicAdd(&p.counters[1], 1u)] = (uint32_t)(bl - p.bin_lo);
      else p.list_x[atomicAdd(&p.counters[4], 1u)] = (uint32_t)(bl - p.bin_lo);
    }
    BinOut o_l;
    o_l.whole = true;
    o_l.row = 0;
    o_l.x = 0;
    const uint2* run_l = nullptr;
    if (inb && n_l <= SMALL_MAX) {
      o_l = bin_out(p, bl);
      run_l = bin_run(p, bl);
    }
    uint32_t todo = __ballot_sync(FULL_MASK, inb && n_l <= SMALL_MAX);
    while (todo) {
    const int srcl = __ffs(todo) - 1;
    todo &= todo - 1;
    const int64_t b = c0 + srcl;
    const uint32_t n = __shfl_sync(FULL_MASK, n_l, srcl);
    BinOut o;
    o.whole = __shfl_sync(FULL_MASK, (int)o_l.whole, srcl) != 0;
    o.row = (int64_t)shfl_u64((uint64_t)o_l.row, srcl);
    o.x = __shfl_sync(FULL_MASK, o_l.x, srcl);
    const uint2* run = (const uint2*)shfl_u64((uint64_t)(uintptr_t)run_l, srcl);
    if (n == 0) {
      emit_finish(p, o, 0);
      continue;
    }
    if (lane == 0) st_rec += n;
    if (n <= TINY_MAX) {
      // the whole bin is one step: fold duplicates with match_any, rank the group leaders, done
      const bool has = lane < n;
      uint2 r = make_uint2(0x80000000u | lane, 0);
      if (has) r = run ? ld_stream_u2(run + lane) : bin_record(p, b, lane);
      const uint32_t lt = lanemask_lt();
      const uint32_t peers = __match_any_sync(FULL_MASK, r.x);
      const bool lead = has && (peers & lt) == 0;
      const uint32_t cnt = TIME ? (uint32_t)__popc(peers) : 0u;
      uint64_t sum = r.y;
      uint32_t rest = lead ? (peers & (peers - 1)) : 0u;
      while (__any_sync(FULL_MASK, rest != 0)) {
        const int src = rest ? __ffs(rest) - 1 : (int)lane;
        const uint32_t vv = __shfl_sync(FULL_MASK, r.y, src);
        if (rest) {
          sum += vv;
          rest &= rest - 1;
        }
      }
      const uint64_t key = lead ? float_key(TIME, r.x, cnt, sum, p.w_scale) : 0ull;
      const uint32_t lm = __ballot_sync(FULL_MASK, lead);
      int rank = 0;
      for (uint32_t m = lm; m; m &= m - 1) rank += shfl_u64(key, __ffs(m) - 1) > key;
      if (lead && rank < p.k) emit_entry(p, o, rank, key, cnt, sum);
      const int nl = __popc(lm);
      emit_finish(p, o, nl < p.k ? nl : p.k);
      if (lane == 0) st_occ += nl;
      if (lead) st_pay += TIME ? (uint64_t)cnt : sum;
      continue;
    }
    if (lane == 0) *n_cand = 0;
    for (int s = 0; s < (run ? 1 : p.n_seg); ++s) {
      uint64_t beg = 0, end = n;
      const uint2* rec = run;
      if (!run) {
        const uint64_t o0 = p.seg[s].offsets[0];
        beg = p.seg[s].offsets[b - p.bin_lo] - o0;
        end = p.seg[s].offsets[b - p.bin_lo + 1] - o0;
        rec = (const uint2*)p.seg[s].records;
      }
      for (uint64_t i0 = beg; i0 < end; i0 += 64) {
        const bool has0 = i0 + lane < end, has1 = i0 + 32 + lane < end;
        uint2 r0 = make_uint2(0, 0), r1 = make_uint2(0, 0);
        if (has0) r0 = ld_stream_u2(rec + i0 + lane);
        if (has1) r1 = ld_stream_u2(rec + i0 + 32 + lane);
        if (!t.insert2(has0, r0.x, r0.y, has1, r1.x, r1.y)) overflow = true;
      }
    }
    __syncwarp();
    const uint32_t d = *t.n_occ;
    // sweep 1: lane maxima of the integer key -> threshold
    uint64_t thr = 0;
    if (d > SMALL_DIRECT) {
      uint64_t best = 0;
      for (uint32_t i = lane; i < d; i += 32) {
        const uint64_t ik = t.ikey(t.occ[i], p.range);
        best = ik > best ? ik : best;
      }
      thr = with_margin(warp_kth_largest(best, p.k));
    }
    // sweep 2: candidates (+ stats)
    for (uint32_t i0 = 0; i0 < d; i0 += 32) {
      const uint32_t i = i0 + lane;
      bool q = false;
      uint32_t y = 0, cnt = 0;
      uint64_t sum = 0;
      if (i < d) {
        const uint32_t h = t.occ[i];
        y = t.keys[h];
        cnt = t.count(h);
        sum = t.sum(h);
        st_pay += TIME ? (uint64_t)cnt : sum;
        q = t.ikey(h, p.range) >= thr;
      }
      push_candidate<SMALL_CANDS>(q, n_cand, c, TIME, y, cnt, sum, p.w_scale);
    }
    st_occ += (lane == 0) ? d : 0;
    __syncwarp();
    int n_c = (int)*n_cand;
    if (n_c > SMALL_CANDS) {  // adversarial layout: exact K-round selection straight from the table
      ++st_slow;
      n_c = warp_select_slow<TIME, SMALL_LOG>(t, 0, d, p.k, p.w_scale, c, 0);
    }
    const int found = warp_rank_emit(c, n_c, p.k, [&](int r, uint64_t kk, uint32_t cnt, uint64_t sum) {
      emit_entry(p, o, r, kk, cnt, sum);
    });
    emit_finish(p, o, found);
    t.clear_dirty(lane, 32, d);
    if (lane == 0) *t.n_occ = 0;
    __syncwarp();
    }
  }
  // one stats update per warp
  for (int off = 16; off > 0; off >>= 1) {
    st_occ += shfl_u64(st_occ, lane ^ off);
    st_slow += __shfl_xor_sync(FULL_MASK, st_slow, off);
    st_pay += shfl_u64(st_pay, lane ^ off);
  }
  if (lane == 0 && (st_occ || st_pay)) {
    atomicAdd(&p.stats[0], (unsigned long long)st_occ);
    atomicAdd(&p.stats[1], (unsigned long long)st_pay);
    if (st_slow) atomicAdd(&p.stats[3], (unsigned long long)(st_slow / 32));
  }
  if (lane == 0 && st_rec) atomicAdd(&p.stats[4], (unsigned long long)st_rec);
  if (overflow) atomicOr(&p.stats[2], 1ull);
}

// ---- medium / large bins: one block per bin, work taken from a list through an atomic cursor ----
constexpr int BLOCK_CANDS = 128;

// TIER 0: medium list, 1: large list, 2: extra-large list (single pass up to 3/4 of the slots, otherwise
// aid_y-hash passes whose records are first compacted per warp so that every insert step runs 32 wide)
template <bool TIME, int THREADS, int LOG, int TIER>
__global__ void __launch_bounds__(THREADS) reduce_block_kernel(const ReduceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int WARPS = THREADS / 32;
  constexpr uint32_t SLOTS = 1u << LOG;
  constexpr int NC = BLOCK_CANDS + OTTO_MAX_K;     // candidates + room for the carried best list
  __shared__ uint32_t s_ncand, s_nocc;
  __shared__ uint64_t s_gmax[WARPS][32];
  __shared__ uint64_t s_thr;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  // carve: cand key[NC'] sum[NC'] | best key[K] sum[K] (u64) | keys lo hc | cand cnt[NC'] best cnt[K] (u32) | occ (u16)
  constexpr int NCX = (NC > WARPS * OTTO_MAX_K ? NC : WARPS * OTTO_MAX_K) + OTTO_MAX_K;
  Cands c, best;
  c.key = (uint64_t*)smem_raw;
  c.sum = c.key + NCX;
  best.key = c.sum + NCX;
  best.sum = best.key + OTTO_MAX_K;
  Table<TIME, LOG> t;
  t.keys = (uint32_t*)(best.sum + OTTO_MAX_K);
  t.lo = t.keys + SLOTS;
  t.hc = t.lo + SLOTS;
  c.cnt = t.hc + (TIME ? SLOTS : 0);
  best.cnt = c.cnt + NCX;
  t.occ = (uint16_t*)(best.cnt + OTTO_MAX_K);
  t.n_occ = &s_nocc;
  uint2* stage = (uint2*)(t.occ + SLOTS) + warp * 64;    // TIER 2 only (see reduce_block_smem)
  constexpr int CUR = TIER == 0 ? 2 : TIER == 1 ? 3 : 5;
  constexpr uint32_t SINGLE_CAP = TIER == 2 ? SLOTS / 4 * 3 : 0xffffffffu;
  const uint32_t lt = lanemask_lt();

  const uint32_t* list = TIER == 0 ? p.list_m : TIER == 1 ? p.list_l : p.list_x;
  const uint32_t n_items = p.counters[TIER == 0 ? 0 : TIER == 1 ? 1 : 4];
  uint64_t st_occ = 0, st_pay = 0, st_rec = 0;
  uint32_t st_slow = 0;
  bool overflow = false;
  t.clear_all(threadIdx.x, THREADS);
  // Work items are taken BATCH at a time and their metadata (dependent loads: list -> record offsets -> bin ->
  // aid_x -> first bin of the row) is fetched by BATCH threads in parallel into shared memory: one chain of round
  // trips per batch instead of one per bin (the ablation in profiles/README.md: with empty tables and no inserts
  // the block kernels still took half their time, all of it in these chains).
  constexpr int BATCH = 8;
  __shared__ uint32_t s_first;
  __shared__ uint32_t s_mn[BATCH], s_mx[BATCH], s_mwhole[BATCH];
  __shared__ int64_t s_mb[BATCH], s_mrow[BATCH];
  __shared__ const uint2* s_mrun[BATCH];
  if (threadIdx.x == 0) s_first = atomicAdd(&p.counters[CUR], (uint32_t)BATCH);
  __syncthreads();
  while (true) {
    const uint32_t first = s_first;
    if (first >= n_items) break;
    const uint32_t n_batch = min((uint32_t)BATCH, n_items - first);
    if (threadIdx.x < n_batch) {
      const int64_t bb = p.bin_lo + list[first + threadIdx.x];
      const BinOut ob = bin_out(p, bb);
      s_mb[threadIdx.x] = bb;
      s_mn[threadIdx.x] = bin_records(p, bb);
      s_mx[threadIdx.x] = ob.x;
      s_mwhole[threadIdx.x] = ob.whole ? 1u : 0u;
      s_mrow[threadIdx.x] = ob.row;
      s_mrun[threadIdx.x] = bin_run(p, bb);
    }
    __syncthreads();
    for (uint32_t k = 0; k < n_batch; ++k) {
    const int64_t b = s_mb[k];
    const uint32_t n = s_mn[k];
    const uint32_t n_pass = n > SINGLE_CAP ? (n + SLOTS / 2 - 1) / (SLOTS / 2) : 1;
    BinOut o;
    o.x = s_mx[k];
    o.whole = s_mwhole[k] != 0;
    o.row = s_mrow[k];
    const uint2* run = s_mrun[k];
    if (threadIdx.x == 0) st_rec += n;
    int n_best = 0;  // meaningful in warp 0
    for (uint32_t pass = 0; pass < n_pass; ++pass) {
      const bool last = pass + 1 == n_pass;
      if (threadIdx.x == 0) s_ncand = 0;
      uint32_t n_st = 0;   // records staged by this warp (multi-pass only)
      for (int s = 0; s < (run ? 1 : p.n_seg); ++s) {
        uint64_t beg = 0, end = n;
        const uint2* rec = run;
        if (!run) {
          const uint64_t o0 = p.seg[s].offsets[0];
          beg = p.seg[s].offsets[b - p.bin_lo] - o0;
          end = p.seg[s].offsets[b - p.bin_lo + 1] - o0;
          rec = (const uint2*)p.seg[s].records;
        }
        if (n_pass == 1) {
          // two records per thread and step, the next two in flight while the current ones are inserted
          auto fetch = [&](uint64_t i0, bool& h0, uint2& q0, bool& h1, uint2& q1) {
            h0 = i0 + threadIdx.x < end;
            h1 = i0 + THREADS + threadIdx.x < end;
            q0 = q1 = make_uint2(0, 0);
            if (h0) q0 = ld_stream_u2(rec + i0 + threadIdx.x);
            if (h1) q1 = ld_stream_u2(rec + i0 + THREADS + threadIdx.x);
          };
          bool n0, n1;
          uint2 q0, q1;
          fetch(beg, n0, q0, n1, q1);
          for (uint64_t i0 = beg; i0 < end; i0 += 2 * THREADS) {
            const bool h0 = n0, h1 = n1;
            const uint2 r0 = q0, r1 = q1;
            fetch(i0 + 2 * THREADS, n0, q0, n1, q1);
            if (!t.insert2(h0, r0.x, r0.y, h1, r1.x, r1.y)) overflow = true;
          }
        } else {
          // hash passes: a pass takes 1 / n_pass of the records; compact them per warp so that every insert is 32 wide
          bool has_n = beg + threadIdx.x < end;
          uint2 r_n = make_uint2(0, 0);
          if (has_n) r_n = ld_stream_u2(rec + beg + threadIdx.x);
          for (uint64_t i0 = beg; i0 < end; i0 += THREADS) {
            bool has = has_n;
            const uint2 r = r_n;
            has_n = i0 + THREADS + threadIdx.x < end;
            if (has_n) r_n = ld_stream_u2(rec + i0 + THREADS + threadIdx.x);
            has = has && __umulhi(hash32b(r.x), n_pass) == pass;
            const uint32_t m = __ballot_sync(FULL_MASK, has);
            if (has) stage[n_st + __popc(m & lt)] = r;
            n_st += __popc(m);
            __syncwarp();
            if (n_st >= 32) {
              n_st -= 32;
              const uint2 q = stage[n_st + lane];
              __syncwarp();
              if (!t.insert(true, q.x, q.y)) overflow = true;
            }
          }
        }
      }
      if (n_pass > 1 && n_st) {
        uint2 q = make_uint2(0, 0);
        if (lane < n_st) q = stage[lane];
        __syncwarp();
        if (!t.insert(lane < n_st, q.x, q.y)) overflow = true;
        n_st = 0;
      }
      __syncthreads();
      const uint32_t d = s_nocc;
      // sweep 1: lane-group maxima of the integer key (group = lane index, across warps)
      uint64_t tbest = 0;
      for (uint32_t i = threadIdx.x; i < d; i += THREADS) {
        const uint64_t ik = t.ikey(t.occ[i], p.range);
        tbest = ik > tbest ? ik : tbest;
      }
      s_gmax[warp][lane] = tbest;
      __syncthreads();
      if (warp == 0) {
        uint64_t g = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) g = s_gmax[w][lane] > g ? s_gmax[w][lane] : g;
        const uint64_t thr = with_margin(warp_kth_largest(g, p.k));
        if (lane == 0) s_thr = thr;
      }
      __syncthreads();
      const uint64_t thr = s_thr;
      // sweep 2: candidates (+ stats)
      for (uint32_t i0 = 0; i0 < d; i0 += THREADS) {
        const uint32_t i = i0 + threadIdx.x;
        bool q = false;
        uint32_t y = 0, cnt = 0;
        uint64_t sum = 0;
        if (i < d) {
          const uint32_t h = t.occ[i];
          y = t.keys[h];
          cnt = t.count(h);
          sum = t.sum(h);
          st_pay += TIME ? (uint64_t)cnt : sum;
          q = t.ikey(h, p.range) >= thr;
        }
        push_candidate<BLOCK_CANDS>(q, &s_ncand, c, TIME, y, cnt, sum, p.w_scale);
      }
      if (threadIdx.x == 0) st_occ += d;
      __syncthreads();
      int n_c = (int)s_ncand;
      if (n_c > BLOCK_CANDS) {
        // adversarial layout: exact K-round selection per warp slice of the occ list, lists land at warp * K
        if (threadIdx.x == 0) ++st_slow;
        const uint32_t per = (d + WARPS - 1) / WARPS;
        const uint32_t lo = min(d, warp * per), hi = min(d, (warp + 1) * per);
        const int f = warp_select_slow<TIME, LOG>(t, lo, hi, p.k, p.w_scale, c, warp * OTTO_MAX_K);
        for (int r = f + (int)lane; r < OTTO_MAX_K; r += 32) c.key[warp * OTTO_MAX_K + r] = 0;
        __syncthreads();
        // compact the non-zero keys to the front (one warp)
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
        // the slow path marks taken keys: restore them so that clear_dirty sees plain keys (it only stores)
      }
      // all warps: reset the table for the next pass / bin while warp 0 ranks the candidates
      t.clear_dirty(threadIdx.x, THREADS, d);
      if (warp == 0) {
        n_c = __shfl_sync(FULL_MASK, n_c, 0);
        // entries carried from earlier passes join the candidates (n_c + n_best <= NCX by construction)
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
      if (threadIdx.x == THREADS - 1) s_nocc = 0;
      __syncthreads();
    }
    }
    if (threadIdx.x == 0) s_first = atomicAdd(&p.counters[CUR], (uint32_t)BATCH);
    __syncthreads();
  }
  for (int off = 16; off > 0; off >>= 1) {
    st_occ += shfl_u64(st_occ, lane ^ off);
    st_slow += __shfl_xor_sync(FULL_MASK, st_slow, off);
    st_pay += shfl_u64(st_pay, lane ^ off);
  }
  if (lane == 0 && (st_occ || st_pay)) {
    atomicAdd(&p.stats[0], (unsigned long long)st_occ);
    atomicAdd(&p.stats[1], (unsigned long long)st_pay);
    if (st_slow) atomicAdd(&p.stats[3], (unsigned long long)st_slow);
  }
  if (threadIdx.x == 0 && st_rec) atomicAdd(&p.stats[5 + TIER], (unsigned long long)st_rec);
  if (overflow) atomicOr(&p.stats[2], 1ull);
}

template <bool TIME, int THREADS, int LOG, int TIER>
constexpr size_t reduce_block_smem() {
  constexpr int NC = BLOCK_CANDS + OTTO_MAX_K;
  constexpr int WARPS = THREADS / 32;
  constexpr int NCX = (NC > WARPS * OTTO_MAX_K ? NC : WARPS * OTTO_MAX_K) + OTTO_MAX_K;
  constexpr size_t SLOTS = (size_t)1 << LOG;
  return (size_t)NCX * 16 + OTTO_MAX_K * 16 + SLOTS * 8 + (TIME ? SLOTS * 4 : 0) + NCX * 4 + OTTO_MAX_K * 4 + SLOTS * 2 +
         (TIER == 2 ? (size_t)WARPS * 64 * 8 : 0);
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

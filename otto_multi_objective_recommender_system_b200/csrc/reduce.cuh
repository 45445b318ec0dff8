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
// Round-2 design (v5).  The round-1 kernels issued 9.6 warp-instructions per record where the insert itself
// needs about one; the rest was paid per bin and per warp (profiles/README.md).  What changed:
//   * selection works on a 32-bit order key: time mode the top 32 bits of count * R + 3 * tsum (shift from the
//     bin's record count), type / unit mode (sum << s) | top bits of ~aid_y, i.e. the final order itself as far as
//     32 bits reach.  The K-th largest of the 32 lane-group maxima (a 32-bit bitonic network, ~75 instructions
//     instead of ~530 for 64-bit keys) is a lower bound of the K-th largest entry; entries at or above it
//     (minus the fp32 rounding margin in time mode) are the ~25-35 candidates whose exact keys one warp ranks;
//   * sweep 1 keeps every entry's key in registers, so sweep 2 is a compare per entry plus the reset of the entry's
//     slot; only candidates read their payload again.  Nothing else touches the table, so one pass resets it;
//   * bins of up to 1024 records are processed by ONE warp each (no block barriers at all, the occupied-list
//     counter lives in a register), in two table sizes; larger bins by a block with three barriers per bin:
//     the ranking of bin b's candidates by warp 0 overlaps the inserts of bin b + 1 (candidate lists and the
//     block's counters are double buffered, record chunks are handed out dynamically);
//   * a classify kernel builds one work list per tier; bin metadata is fetched 32 (warp tiers) or 8 (block tiers)
//     bins at a time, for the block tiers one batch ahead;
//   * if the candidate list overflows (adversarial ties), the bin is re-inserted and an exact K-round selection
//     runs instead.
// Slices of split rows write partial top-K lists; merge_split_rows() picks the final K (slices hold disjoint
// aid_y, so the merge is a pure selection).
#pragma once
#include "common.cuh"

struct ReduceParams {
  const uint2* records;      // the one segment of this call
  const uint64_t* offsets;   // offsets[b - bin_lo] .. offsets[b - bin_lo + 1] bound bin b; offsets[0] is subtracted
  const uint32_t* bin_x;     // [B] bin -> aid_x (global bin ids)
  const uint32_t* bin_base;  // [A + 1]
  int64_t bin_lo, bin_hi;    // this call's bins
  int32_t aid_lo, aid_hi;
  int32_t k;
  int32_t time_mode;
  uint32_t range;            // ts_max - ts_min (time mode)
  double w_scale;            // 3 / (ts_max - ts_min)
  uint32_t y_bits;           // bits of n_aids - 1
  uint32_t max_v;            // largest record value outside time mode (largest type weight, or 1)
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
  uint32_t* list[5];         // work list of every tier: bin - bin_lo; [4] = bins the owner-table tiers hand to the hash-table kernel
  uint32_t* counters;        // [0..4] items per tier, [8..12] next item per tier
  unsigned long long* stats; // [0] distinct  [1] checksum  [2] overflow  [3] slow-path selections  [4..7] records per tier
};

constexpr uint32_t TINY_MAX = 32;     // records: one step of one warp, no table at all
constexpr uint32_t TIER0_MAX = 384, TIER1_MAX = 1536, TIER2_MAX = 3072, TIER3_MAX = 6144;   // records per bin of tiers 0..3; tier 4 takes the rest
constexpr int N_TIERS = 5, NEXT_ITEM = 8;   // counters[NEXT_ITEM + t] = next work item of tier t
constexpr int N_CAND = 64;            // candidates a fast selection may produce
constexpr int N_CAND_BUF = N_CAND + OTTO_MAX_K;   // block tiers: + the best list carried between hash passes
constexpr uint32_t KEY_NONE = 0u;     // table keys are aid_y + 1

__device__ __forceinline__ int64_t partial_slot(const ReduceParams& p, uint32_t x, uint32_t j) {
  const int64_t extra = ((int64_t)p.bin_base[x] - x) - (p.bin_lo - p.aid_lo);
  return 2 * extra + j;
}

__device__ __forceinline__ uint64_t bin_start(const ReduceParams& p, int64_t b) { return p.offsets[b - p.bin_lo] - p.offsets[0]; }

// ---- work lists: one thread per bin ----
__global__ void __launch_bounds__(256) reduce_classify_kernel(const ReduceParams p) {
  const int64_t b = p.bin_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int tier = -1;
  if (b < p.bin_hi) {
    const uint64_t n = p.offsets[b - p.bin_lo + 1] - p.offsets[b - p.bin_lo];
    tier = n <= TIER0_MAX ? 0 : n <= TIER1_MAX ? 1 : n <= TIER2_MAX ? 2 : n <= TIER3_MAX ? 3 : 4;
  }
  const uint32_t lt = lanemask_lt();
#pragma unroll
  for (int t = 0; t < N_TIERS; ++t) {
    const uint32_t m = __ballot_sync(FULL_MASK, tier == t);
    if (m == 0) continue;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(&p.counters[t], (uint32_t)__popc(m));
    base = __shfl_sync(FULL_MASK, base, leader);
    if (tier == t) p.list[t][base + __popc(m & lt)] = (uint32_t)(b - p.bin_lo);
  }
}

// Shared memory is addressed through 32-bit shared-space addresses and inline PTX: with generic pointers carved from
// the dynamic allocation the compiler rebuilt the shared window base (S2UR SR_CgaCtaId + 3 uniform ops) in front of
// every access and wrapped single-lane atomics into its own aggregation sequence - the insert path of v5a issued 380
// instructions per 64 records (profiles/r02_reduce_v5a_*), three quarters of the kernel.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t atoms_cas(uint32_t addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void sts_zero16(uint32_t addr) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}

// Open-addressing table over SLOTS = 2^LOG slots, keys = aid_y + 1 (0 = free), double hashing.
// Payload of slot h: TIME an 8-byte pair {lo, hc} - lo = low 32 bits of the time sum, hc = (count - 1) | carries of
// lo << 24, so that the record that claims a slot (three records in four here) only adds to lo; !TIME a 4-byte integer
// weight sum.  Keys and payloads are contiguous, so a full reset is one run of 16-byte stores.
template <bool TIME, int LOG>
struct Table {
  static constexpr uint32_t SLOTS = 1u << LOG;
  static constexpr uint32_t BYTES = SLOTS * (TIME ? 12u : 8u);
  static constexpr uint32_t PSHIFT = TIME ? 3 : 2;   // log2 of the payload stride
  uint32_t keys_s, pay_s;        // shared-space byte addresses
  uint32_t occ_s;                // claimed slots (u16), in claim order

  __device__ __forceinline__ void carve(const void* base, const void* occ) {
    keys_s = smem_u32(base);
    pay_s = keys_s + SLOTS * 4;
    occ_s = smem_u32(occ);
  }
  __device__ __forceinline__ uint32_t occ(uint32_t i) const { return lds_u16(occ_s + i * 2); }
  __device__ __forceinline__ uint32_t key(uint32_t h) const { return lds_u32(keys_s + h * 4); }
  // {lo, hc} of slot h (hc = 0 outside time mode)
  __device__ __forceinline__ uint2 payload(uint32_t h) const {
    return TIME ? lds_u64(pay_s + (h << 3)) : make_uint2(lds_u32(pay_s + (h << 2)), 0u);
  }
  __device__ __forceinline__ void set_key(uint32_t h, uint32_t k) { sts_u32(keys_s + h * 4, k); }
  __device__ __forceinline__ void clear_all(uint32_t tid, uint32_t nthreads) {
    for (uint32_t i = tid; i < BYTES / 16; i += nthreads) sts_zero16(keys_s + i * 16);
  }
  __device__ __forceinline__ void clear_slot(uint32_t h) {
    sts_u32(keys_s + h * 4, KEY_NONE);
    if (TIME) sts_u64(pay_s + (h << 3), 0u, 0u);
    else sts_u32(pay_s + (h << 2), 0u);
  }
  // Two records per lane: both first probes are in flight together (at the load factors here most records settle on
  // the first probe); the few that collide walk their double-hashing sequence in a short divergent loop each.  EVERY
  // lane of the warp must call.  A lane's two records may carry the same aid_y: the second CAS then finds the first
  // one's claim.  The __syncwarp() reconverges the lanes before the warp-wide append to the occupied list.
  // LOCAL: n_occ_reg is a warp-uniform register counter (one warp owns the table); else n_occ_s is the shared-space
  // address of the block's counter.  Returns false when the table is full.
  template <bool LOCAL>
  __device__ __forceinline__ bool insert2(bool has0, uint32_t y0, uint32_t v0, bool has1, uint32_t y1, uint32_t v1,
                                          uint32_t& n_occ_reg, uint32_t n_occ_s) {
    const uint32_t k0 = y0 + 1u, k1 = y1 + 1u;
    uint32_t a0 = keys_s + (((y0 * 0x9E3779B1u) >> (32 - LOG)) << 2), a1 = keys_s + (((y1 * 0x9E3779B1u) >> (32 - LOG)) << 2);
    uint32_t prev0 = k0, prev1 = k1;
    if (has0) prev0 = atoms_cas(a0, KEY_NONE, k0);
    if (has1) prev1 = atoms_cas(a1, KEY_NONE, k1);
    if (prev0 != KEY_NONE && prev0 != k0) {
      const uint32_t step = (((y0 * 0x85EBCA6Bu) >> (32 - LOG)) | 1u) << 2;
      uint32_t left = SLOTS;
      do {
        a0 = keys_s + ((a0 - keys_s + step) & (SLOTS * 4 - 1));
        prev0 = atoms_cas(a0, KEY_NONE, k0);
      } while (prev0 != KEY_NONE && prev0 != k0 && --left);
    }
    if (prev1 != KEY_NONE && prev1 != k1) {
      const uint32_t step = (((y1 * 0x85EBCA6Bu) >> (32 - LOG)) | 1u) << 2;
      uint32_t left = SLOTS;
      do {
        a1 = keys_s + ((a1 - keys_s + step) & (SLOTS * 4 - 1));
        prev1 = atoms_cas(a1, KEY_NONE, k1);
      } while (prev1 != KEY_NONE && prev1 != k1 && --left);
    }
    __syncwarp();
    const bool ok0 = prev0 == KEY_NONE || prev0 == k0, ok1 = prev1 == KEY_NONE || prev1 == k1;
    const uint32_t fresh0 = __ballot_sync(FULL_MASK, has0 && prev0 == KEY_NONE);
    const uint32_t fresh1 = __ballot_sync(FULL_MASK, has1 && prev1 == KEY_NONE);
    if (fresh0 | fresh1) {
      const uint32_t n0 = (uint32_t)__popc(fresh0), add = n0 + (uint32_t)__popc(fresh1);
      uint32_t base;
      if (LOCAL) {
        base = n_occ_reg;
        n_occ_reg = base + add;
      } else {
        base = 0;
        if (lane_id() == 0) base = atoms_add(n_occ_s, add);
        base = __shfl_sync(FULL_MASK, base, 0);
      }
      const uint32_t lt = lanemask_lt();
      if (has0 && prev0 == KEY_NONE) sts_u16(occ_s + (base + __popc(fresh0 & lt)) * 2, (a0 - keys_s) >> 2);
      if (has1 && prev1 == KEY_NONE) sts_u16(occ_s + (base + n0 + __popc(fresh1 & lt)) * 2, (a1 - keys_s) >> 2);
    }
    if (has0 && ok0) {
      const uint32_t pa = pay_s + ((a0 - keys_s) << (PSHIFT - 2));
      const uint32_t old = atoms_add(pa, v0);
      if (TIME) {
        const uint32_t inc = (prev0 == k0 ? 1u : 0u) + ((old + v0 < old) ? (1u << 24) : 0u);
        if (inc) reds_add(pa + 4, inc);
      }
    }
    if (has1 && ok1) {
      const uint32_t pa = pay_s + ((a1 - keys_s) << (PSHIFT - 2));
      const uint32_t old = atoms_add(pa, v1);
      if (TIME) {
        const uint32_t inc = (prev1 == k1 ? 1u : 0u) + ((old + v1 < old) ? (1u << 24) : 0u);
        if (inc) reds_add(pa + 4, inc);
      }
    }
    return ok0 && ok1;
  }
  static __device__ __forceinline__ uint32_t count_of(uint2 pl) { return TIME ? (pl.y & 0xffffffu) + 1u : 0u; }
  static __device__ __forceinline__ uint64_t sum_of(uint2 pl) { return TIME ? (((uint64_t)(pl.y >> 24) << 32) | pl.x) : (uint64_t)pl.x; }
  __device__ __forceinline__ uint32_t count(uint32_t h) const { return count_of(payload(h)); }
  __device__ __forceinline__ uint64_t sum(uint32_t h) const { return sum_of(payload(h)); }
};

// 32-bit order key of an entry, monotone (non-strictly) in the final order of the bin
struct KeyCfg {
  uint32_t shift;    // time: key = (count * R + 3 * tsum) >> shift
  uint32_t s;        // type / unit: key = (sum << s) | ((~y & ymask) >> yshift)
  uint32_t yshift, ymask;
  uint32_t range;
};
template <bool TIME>
__device__ __forceinline__ KeyCfg make_cfg(const ReduceParams& p, uint32_t n) {
  KeyCfg c;
  c.range = p.range;
  c.shift = c.s = c.yshift = c.ymask = 0;
  if (TIME) {
    const int bits = 64 - __clzll((long long)(4ull * n * p.range));   // count * R + 3 * tsum <= 4 n R
    c.shift = bits > 32 ? (uint32_t)(bits - 32) : 0u;
  } else {
    const unsigned long long top = (unsigned long long)n * p.max_v;
    const int wbits = top ? 64 - __clzll((long long)top) : 1;
    int s = wbits >= 32 ? 0 : 32 - wbits;
    if (s > (int)p.y_bits) s = (int)p.y_bits;
    c.s = (uint32_t)s;
    c.yshift = p.y_bits - (uint32_t)s;
    c.ymask = p.y_bits >= 32 ? 0xffffffffu : ((1u << p.y_bits) - 1u);
  }
  return c;
}
template <bool TIME>
__device__ __forceinline__ uint32_t key32(const KeyCfg& c, uint32_t key, uint32_t lo, uint32_t hc) {
  if (TIME) {
    const uint64_t ik = (uint64_t)((hc & 0xffffffu) + 1u) * c.range + 3ull * (((uint64_t)(hc >> 24) << 32) | lo);
    return (uint32_t)(ik >> c.shift);
  }
  const uint32_t y = key - 1u;
  return c.s ? ((lo << c.s) | ((~y & c.ymask) >> c.yshift)) : lo;
}
// time mode: entries whose fp32 weight can tie with the threshold entry's lie within 2^-23 of it
template <bool TIME>
__device__ __forceinline__ uint32_t cand_threshold(uint32_t thr) {
  if (!TIME) return thr;
  const uint32_t m = (thr >> 22) + 1u;
  return thr > m ? thr - m : 0u;
}

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

// K-th largest of the 32 lane values (0 if fewer than k lanes are non-zero): descending bitonic network
__device__ __forceinline__ uint32_t warp_kth_largest32(uint32_t v, int k) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      const uint32_t o = __shfl_xor_sync(FULL_MASK, v, j);
      const bool keep_max = ((lane & j) == 0) == ((lane & kk) == 0);
      v = keep_max ? max(v, o) : min(v, o);
    }
  }
  return __shfl_sync(FULL_MASK, v, k - 1);
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

// Exact K-round selection over occupied entries [lo, hi) of the occ list (slow path: candidate-list overflow).
// One warp; results (best first) land in dst[dst_base ..); returns how many.  Marks taken keys with bit 31.
template <bool TIME, int LOG>
__device__ __forceinline__ int warp_select_slow(Table<TIME, LOG>& t, uint32_t lo, uint32_t hi, int k, double w_scale,
                                                Cands dst, int dst_base) {
  const uint32_t lane = lane_id();
  auto scan = [&](uint64_t& best, uint32_t& best_h) {
    best = 0;
    for (uint32_t i = lo + lane; i < hi; i += 32) {
      const uint32_t h = t.occ(i);
      const uint32_t key = t.key(h);
      if (key & KEY_TAKEN) continue;
      const uint64_t kk = float_key(TIME, key - 1u, t.count(h), t.sum(h), w_scale);
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
      t.set_key(best_h, t.key(best_h) | KEY_TAKEN);
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

struct BinStats {
  uint64_t occ = 0, pay = 0, rec = 0;
  uint32_t slow = 0;
  bool overflow = false;
};


// the whole bin is one step: fold duplicates with match_any, rank the group leaders
template <bool TIME>
__device__ __forceinline__ void tiny_bin(const ReduceParams& p, const BinOut& o, const uint2* run, uint32_t n, BinStats& st) {
  const uint32_t lane = lane_id();
  const bool has = lane < n;
  uint2 r = make_uint2(0x80000000u | lane, 0);
  if (has) r = ld_stream_u2(run + lane);
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
  if (lane == 0) st.occ += nl;
  if (lead) st.pay += TIME ? (uint64_t)cnt : sum;
}

// =====================================================================================================
// block tiers: one block per bin, three barriers per bin
// =====================================================================================================
// MULTI: bins beyond SINGLE_CAP records are accumulated in several aid_y-hash passes (records compacted per warp
// before inserting so that every insert step runs 32 wide); the best K of every pass are carried into the next.
template <bool TIME, int THREADS, int LOG, bool MULTI>
constexpr size_t reduce_block_smem() {
  // cand key / sum x2 buffers, best key / sum (u64) | table | cand cnt x2, best cnt (u32) | occ (u16) | stage (u64)
  constexpr size_t SLOTS = (size_t)1 << LOG;
  constexpr size_t NCB = MULTI ? N_CAND_BUF : N_CAND;
  return 2 * NCB * 16 + OTTO_MAX_K * 16 + Table<TIME, LOG>::BYTES + 2 * NCB * 4 + OTTO_MAX_K * 4 + SLOTS * 2 +
         (MULTI ? (size_t)(THREADS / 32) * 64 * 8 : 0);
}

template <bool TIME, int THREADS, int LOG, int TIER, bool MULTI>
__global__ void __launch_bounds__(THREADS, THREADS == 128 ? 7 : THREADS == 256 ? 3 : 1) reduce_block_kernel(const ReduceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int WARPS = THREADS / 32;
  constexpr uint32_t SLOTS = 1u << LOG;
  constexpr int EPT = SLOTS / THREADS;              // entries per thread: d <= SLOTS
  constexpr uint32_t SINGLE_CAP = SLOTS / 4 * 3;    // records a single pass takes
  constexpr int BATCH = 8;
  constexpr int NCB = MULTI ? N_CAND_BUF : N_CAND;   // candidate buffer: + the carried best list of hash passes
  static_assert(WARPS >= 2, "the metadata prefetch runs on warp 1");
  __shared__ uint32_t s_ncand[2], s_nocc[2], s_chunk[2];
  __shared__ uint32_t s_gmax[WARPS][32];
  __shared__ uint32_t s_first[2], s_slow;
  __shared__ uint32_t s_mn[2][BATCH], s_mx[2][BATCH], s_mwhole[2][BATCH];
  __shared__ int64_t s_mrow[2][BATCH];
  __shared__ const uint2* s_mrun[2][BATCH];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, lt = lanemask_lt();
  // candidate lists of the two parities are slices of one array each (no pointer arrays: they would live in local memory)
  Cands c_all, best;
  c_all.key = (uint64_t*)smem_raw;
  c_all.sum = c_all.key + 2 * NCB;
  best.key = c_all.sum + 2 * NCB;
  best.sum = best.key + OTTO_MAX_K;
  unsigned char* tb = (unsigned char*)(best.sum + OTTO_MAX_K);
  c_all.cnt = (uint32_t*)(tb + Table<TIME, LOG>::BYTES);
  best.cnt = c_all.cnt + 2 * NCB;
  auto cands = [&](int parity) {
    Cands r;
    r.key = c_all.key + parity * NCB;
    r.sum = c_all.sum + parity * NCB;
    r.cnt = c_all.cnt + parity * NCB;
    return r;
  };
  Table<TIME, LOG> t;
  t.carve(tb, best.cnt + OTTO_MAX_K);
  uint2* stage = (uint2*)((uint16_t*)(best.cnt + OTTO_MAX_K) + SLOTS) + warp * 64;    // MULTI only

  const uint32_t* list = p.list[TIER];
  const uint32_t n_items = p.counters[TIER];
  if (n_items == 0) return;   // the usual case of the hand-over tier: nothing to clear, nothing to do
  BinStats st;
  t.clear_all(threadIdx.x, THREADS);
  if (threadIdx.x < 2) {
    s_ncand[threadIdx.x] = 0;
    s_nocc[threadIdx.x] = 0;
    s_chunk[threadIdx.x] = 0;
  }
  if (threadIdx.x == 0) s_slow = 0;
  // metadata of a batch of work items: dependent loads list -> record offsets -> bin -> aid_x -> first bin of the row,
  // fetched by BATCH lanes of warp 1 into buffer `buf`; one chain of round trips per batch
  auto fetch_batch = [&](int buf) {   // called by warp 1 only
    uint32_t first = 0;
    if (lane == 0) first = atomicAdd(&p.counters[NEXT_ITEM + TIER], (uint32_t)BATCH);
    first = __shfl_sync(FULL_MASK, first, 0);
    if (lane == 0) s_first[buf] = first;
    if (lane < BATCH && first + lane < n_items) {
      const int64_t bb = p.bin_lo + list[first + lane];
      const uint64_t beg = p.offsets[bb - p.bin_lo], end = p.offsets[bb - p.bin_lo + 1];
      const BinOut ob = bin_out(p, bb);
      s_mn[buf][lane] = (uint32_t)(end - beg);
      s_mx[buf][lane] = ob.x;
      s_mwhole[buf][lane] = ob.whole ? 1u : 0u;
      s_mrow[buf][lane] = ob.row;
      s_mrun[buf][lane] = p.records + (beg - p.offsets[0]);
    }
  };
  if (warp == 1) fetch_batch(0);
  __syncthreads();
  int par = 0;      // parity of the running bin: candidate list / counters in use
  // what warp 0 still owes for the previous bin (ranking of its candidates overlaps this bin's inserts)
  bool owe = false;
  BinOut owe_o;
  int owe_par = 0;
  auto settle = [&]() {   // warp 0 only
    if (!owe) return;
    const int n_c = (int)s_ncand[owe_par];
    const int found = warp_rank_emit(cands(owe_par), n_c, p.k, [&](int r, uint64_t kk, uint32_t cnt, uint64_t sum) { emit_entry(p, owe_o, r, kk, cnt, sum); });
    emit_finish(p, owe_o, found);
    __syncwarp();
    owe = false;
  };
  for (int buf = 0;; buf ^= 1) {
    const uint32_t first = s_first[buf];
    if (first >= n_items) break;
    const uint32_t n_batch = min((uint32_t)BATCH, n_items - first);
    bool fetched = false;
    for (uint32_t kb = 0; kb < n_batch; ++kb) {
      const uint32_t n = s_mn[buf][kb];
      BinOut o;
      o.x = s_mx[buf][kb];
      o.whole = s_mwhole[buf][kb] != 0;
      o.row = s_mrow[buf][kb];
      const uint2* run = s_mrun[buf][kb];
      const uint32_t n_pass = (MULTI && n > SINGLE_CAP) ? (n + SLOTS / 2 - 1) / (SLOTS / 2) : 1;
      const KeyCfg cfg = make_cfg<TIME>(p, n);
      if (threadIdx.x == 0) st.rec += n;
      int n_best = 0;   // meaningful in warp 0 (multi-pass)
      bool slow = false;
      for (uint32_t pass = 0; pass < n_pass; ++pass) {
        const bool last = pass + 1 == n_pass;
        const Cands cp = cands(par);
        const uint32_t nocc_s = smem_u32(&s_nocc[par]), chunk_s = smem_u32(&s_chunk[par]);
        uint32_t dummy = 0;
        // ---- insert: 64-record chunks handed out dynamically (warp 0 joins late while it ranks the previous bin)
        if (warp == 0) settle();
        if (n_pass == 1) {
          auto grab = [&]() {
            uint32_t cidx = 0;
            if (lane == 0) cidx = atoms_add(chunk_s, 1u);
            return __shfl_sync(FULL_MASK, cidx, 0);
          };
          uint32_t i0 = grab() * 64;
          bool h0 = i0 + lane < n, h1 = i0 + 32 + lane < n;
          uint2 q0 = make_uint2(0, 0), q1 = make_uint2(0, 0);
          if (h0) q0 = ld_stream_u2(run + i0 + lane);
          if (h1) q1 = ld_stream_u2(run + i0 + 32 + lane);
          while (i0 < n) {
            const bool a0 = h0, a1 = h1;
            const uint2 r0 = q0, r1 = q1;
            i0 = grab() * 64;
            h0 = i0 + lane < n;
            h1 = i0 + 32 + lane < n;
            if (h0) q0 = ld_stream_u2(run + i0 + lane);
            if (h1) q1 = ld_stream_u2(run + i0 + 32 + lane);
            if (!t.template insert2<false>(a0, r0.x, r0.y, a1, r1.x, r1.y, dummy, nocc_s)) st.overflow = true;
          }
        } else {
          // hash passes: a pass takes 1 / n_pass of the records
          uint32_t n_st = 0;
          bool has_n = threadIdx.x < n;
          uint2 r_n = make_uint2(0, 0);
          if (has_n) r_n = ld_stream_u2(run + threadIdx.x);
          for (uint32_t i0 = 0; i0 < n; i0 += THREADS) {
            bool has = has_n;
            const uint2 r = r_n;
            has_n = i0 + THREADS + threadIdx.x < n;
            if (has_n) r_n = ld_stream_u2(run + i0 + THREADS + threadIdx.x);
            has = has && __umulhi(hash32b(r.x), n_pass) == pass;
            const uint32_t m = __ballot_sync(FULL_MASK, has);
            if (has) stage[n_st + __popc(m & lt)] = r;
            n_st += __popc(m);
            __syncwarp();
            if (n_st >= 32) {
              n_st -= 32;
              const uint2 q = stage[n_st + lane];
              __syncwarp();
              if (!t.template insert2<false>(true, q.x, q.y, false, 0, 0, dummy, nocc_s)) st.overflow = true;
            }
          }
          if (n_st) {
            uint2 q = make_uint2(0, 0);
            if (lane < n_st) q = stage[lane];
            __syncwarp();
            if (!t.template insert2<false>(lane < n_st, q.x, q.y, false, 0, 0, dummy, nocc_s)) st.overflow = true;
          }
        }
        __syncthreads();   // #1: all records of the pass are in the table
        const uint32_t d = s_nocc[par];
        if (threadIdx.x == 0) {   // the other parity's counters are idle now (warp 0 has settled what it owed)
          s_nocc[par ^ 1] = 0;
          s_chunk[par ^ 1] = 0;
          s_ncand[par ^ 1] = 0;
        }
        if (!fetched && warp == 1) fetch_batch(buf ^ 1);   // next batch's metadata, one batch ahead
        fetched = true;
        if (slow) {
          // exact K-round selection per warp slice of the occ list, lists land at warp * K of the candidate buffer
          // (N_CAND >= WARPS * K is not guaranteed: only warps 0 and 1 select, over half of the list each)
          for (uint32_t i = threadIdx.x; i < d; i += THREADS) st.pay += TIME ? (uint64_t)t.count(t.occ(i)) : t.sum(t.occ(i));
          __syncthreads();
          if (warp < 2) {
            const uint32_t half = (d + 1) / 2;
            const uint32_t lo = warp == 0 ? 0 : half, hi = warp == 0 ? half : d;
            const int f = warp_select_slow<TIME, LOG>(t, lo, hi, p.k, p.w_scale, cp, warp * OTTO_MAX_K);
            for (int r = f + (int)lane; r < OTTO_MAX_K; r += 32) cp.key[warp * OTTO_MAX_K + r] = 0;
          }
          __syncthreads();
          if (warp == 0) {
            // compact the non-zero keys of the two lists (+ the carried best) to the front
            int w = 0;
            for (int i0 = 0; i0 < 2 * OTTO_MAX_K; i0 += 32) {
              const int i = i0 + lane;
              const uint64_t kk = cp.key[i];
              const uint64_t ss = cp.sum[i];
              const uint32_t cc = cp.cnt[i];
              const uint32_t m = __ballot_sync(FULL_MASK, kk != 0);
              __syncwarp();
              if (kk != 0) {
                const int at = w + __popc(m & lt);
                cp.key[at] = kk; cp.sum[at] = ss; cp.cnt[at] = cc;
              }
              w += __popc(m);
              __syncwarp();
            }
            for (int r = lane; r < n_best; r += 32) {
              cp.key[w + r] = best.key[r]; cp.sum[w + r] = best.sum[r]; cp.cnt[w + r] = best.cnt[r];
            }
            __syncwarp();
            w += n_best;
            if (last) {
              const int found = warp_rank_emit(cp, w, p.k, [&](int r, uint64_t kk, uint32_t cnt, uint64_t sum) { emit_entry(p, o, r, kk, cnt, sum); });
              emit_finish(p, o, found);
            } else {
              n_best = warp_rank_emit(cp, w, p.k, [&](int r, uint64_t kk, uint32_t cnt, uint64_t sum) { best.key[r] = kk; best.sum[r] = sum; best.cnt[r] = cnt; });
            }
            __syncwarp();
            if (lane == 0) { st.occ += d; ++st.slow; }
          }
          for (uint32_t i = threadIdx.x; i < d; i += THREADS) t.clear_slot(t.occ(i));
          __syncthreads();
          slow = false;
          par ^= 1;
          continue;
        }
        // ---- sweep 1: 32-bit keys into registers, lane-group maxima (group = lane index, across warps)
        uint32_t k32[EPT];
        uint32_t tbest = 0, pay = 0;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
          k32[j] = 0;
          if ((uint32_t)j * THREADS >= d) break;
          const uint32_t i = j * THREADS + threadIdx.x;
          if (i < d) {
            const uint32_t h = t.occ(i);
            const uint2 pl = t.payload(h);
            k32[j] = key32<TIME>(cfg, TIME ? 0u : t.key(h), pl.x, pl.y) + 1u;
            if (k32[j] == 0) k32[j] = 0xffffffffu;
            pay += TIME ? (pl.y & 0xffffffu) + 1u : pl.x;
            tbest = max(tbest, k32[j]);
          }
        }
        s_gmax[warp][lane] = tbest;
        __syncthreads();   // #2
        uint32_t g = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) g = max(g, s_gmax[w][lane]);
        const uint32_t thr = cand_threshold<TIME>(warp_kth_largest32(g, p.k));
        // ---- sweep 2: candidates read their payload again; every entry resets its slot
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
          if ((uint32_t)j * THREADS >= d) break;
          const uint32_t i = j * THREADS + threadIdx.x;
          const bool mine = i < d;
          const bool q = mine && k32[j] >= thr;
          const uint32_t m = __ballot_sync(FULL_MASK, q);
          const uint32_t h = mine ? t.occ(i) : 0u;
          if (m) {
            const int leader = __ffs(m) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(&s_ncand[par], (uint32_t)__popc(m));
            base = __shfl_sync(FULL_MASK, base, leader);
            if (q) {
              const uint32_t at = base + __popc(m & lt);
              if (at < (uint32_t)N_CAND) {   // the carried best list goes behind them
                const uint2 pl = t.payload(h);
                const uint32_t cnt = t.count_of(pl);
                const uint64_t sum = t.sum_of(pl);
                cp.key[at] = float_key(TIME, t.key(h) - 1u, cnt, sum, p.w_scale);
                cp.sum[at] = sum;
                cp.cnt[at] = cnt;
              }
            }
          }
          if (mine) t.clear_slot(h);
        }
        __syncthreads();   // #3: candidate list complete, table clean
        const uint32_t n_c = s_ncand[par];
        if (n_c > (uint32_t)N_CAND) {   // adversarial ties: this pass again, selected exactly
          slow = true;
          --pass;
          par ^= 1;      // fresh counters (the other parity was reset behind barrier #1)
          continue;
        }
        st.pay += pay;
        if (threadIdx.x == 0) st.occ += d;
        if (warp == 0) {
          if (n_pass == 1) {
            // ranking is deferred: it runs while the other warps insert the next bin
            owe = true;
            owe_o = o;
            owe_par = par;
          } else {
            // entries carried from earlier passes join the candidates
            for (int r = lane; r < n_best; r += 32) {
              cp.key[n_c + r] = best.key[r]; cp.sum[n_c + r] = best.sum[r]; cp.cnt[n_c + r] = best.cnt[r];
            }
            __syncwarp();
            const int n_all = (int)n_c + n_best;
            if (last) {
              const int found = warp_rank_emit(cp, n_all, p.k, [&](int r, uint64_t kk, uint32_t cnt, uint64_t sum) { emit_entry(p, o, r, kk, cnt, sum); });
              emit_finish(p, o, found);
            } else {
              n_best = warp_rank_emit(cp, n_all, p.k, [&](int r, uint64_t kk, uint32_t cnt, uint64_t sum) { best.key[r] = kk; best.sum[r] = sum; best.cnt[r] = cnt; });
            }
            __syncwarp();
          }
        }
        par ^= 1;
      }
    }
    // the next batch's metadata was written by warp 1 behind barrier #1 of this batch's first bin; every later
    // barrier orders it.  A batch always has at least one bin, so at least three barriers have passed.
  }
  if (warp == 0) settle();
  for (int off = 16; off > 0; off >>= 1) {
    st.occ += shfl_u64(st.occ, lane ^ off);
    st.slow += __shfl_xor_sync(FULL_MASK, st.slow, off);
    st.pay += shfl_u64(st.pay, lane ^ off);
  }
  if (lane == 0 && (st.occ || st.pay)) {
    atomicAdd(&p.stats[0], (unsigned long long)st.occ);
    atomicAdd(&p.stats[1], (unsigned long long)st.pay);
    if (st.slow) atomicAdd(&p.stats[3], (unsigned long long)st.slow);
  }
  if (threadIdx.x == 0 && st.rec) atomicAdd(&p.stats[4 + (TIER < 3 ? TIER : 3)], (unsigned long long)st.rec);
  if (st.overflow) atomicOr(&p.stats[2], 1ull);
}

// ---- split rows: merge the slices' partial lists (disjoint aid_y) into the final row ----
// Every slice list is sorted best first, so the K best of the row are prefixes of the lists: the K-th largest of the
// lists' HEADS (weight bits, one pass over the slices) bounds the K-th best entry from below, and only entries at or
// above it - about K .. 2K of up to 4096 * K - go through the exact K-round selection.  Round 1 ran the K rounds over
// all nb * K entries (0.86 ms for 10 k split rows, the hottest rows serialising 20 scans of 80 k entries on one warp).
constexpr int MERGE_CANDS = 512, MERGE_WARPS = 4;

__global__ void __launch_bounds__(MERGE_WARPS * 32) merge_split_rows_kernel(const ReduceParams p) {
  __shared__ uint64_t s_key[MERGE_WARPS][MERGE_CANDS];
  __shared__ uint32_t s_at[MERGE_WARPS][MERGE_CANDS];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, lt = lanemask_lt();
  const int64_t n_warps = (int64_t)gridDim.x * MERGE_WARPS;
  const int k = p.k;
  for (int64_t x0 = p.aid_lo + ((int64_t)blockIdx.x * MERGE_WARPS + warp) * 32; x0 < p.aid_hi; x0 += n_warps * 32) {
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
      // threshold: K-th largest head (weight bits) over the slices, by lane maxima
      uint32_t best = 0;
      for (uint32_t j = lane; j < nbx; j += 32)
        if (p.p_len[slot0 + j] > 0) best = max(best, (uint32_t)(p.p_key[(slot0 + j) * k] >> 32) + 1u);
      const uint32_t thr = warp_kth_largest32(best, k);   // 0 (everything is a candidate) with fewer than K non-empty slices
      // candidates: the prefix of every list whose weight bits reach the threshold
      uint32_t n_c = 0;
      for (uint32_t j0 = 0; j0 < nbx; j0 += 32) {
        const uint32_t j = j0 + lane;
        const int len = j < nbx ? p.p_len[slot0 + j] : 0;
        for (int r = 0; __any_sync(FULL_MASK, r < len); ++r) {
          uint64_t kk = 0;
          if (r < len) kk = p.p_key[(slot0 + j) * k + r];
          const bool q = r < len && (uint32_t)(kk >> 32) + 1u >= thr;
          const uint32_t m = __ballot_sync(FULL_MASK, q);
          if (m == 0) break;                      // lists are sorted: nobody has a further candidate at this depth
          const uint32_t at = n_c + __popc(m & lt);
          if (q && at < (uint32_t)MERGE_CANDS) {
            s_key[warp][at] = kk;
            s_at[warp][at] = (uint32_t)((slot0 + j) * k + r - slot0 * k);
          }
          n_c += __popc(m);
        }
      }
      __syncwarp();
      const int64_t row = (int64_t)x * k;
      int found = 0;
      if (n_c <= (uint32_t)MERGE_CANDS) {
        // exact K rounds over the candidates
        for (; found < k; ++found) {
          uint64_t bk = 0;
          uint32_t bi = 0;
          for (uint32_t i = lane; i < n_c; i += 32) {
            const uint64_t kk = s_key[warp][i];
            if (kk > bk) { bk = kk; bi = i; }
          }
          const uint64_t m = warp_max_u64(bk);
          if (m == 0) break;
          if (bk == m) {
            const int64_t at = slot0 * k + s_at[warp][bi];
            p.out_y[row + found] = (int32_t)(~(uint32_t)m);
            p.out_w[row + found] = __uint_as_float((uint32_t)(m >> 32));
            if (p.out_cnt) p.out_cnt[row + found] = p.p_cnt[at];
            if (p.out_tsum) p.out_tsum[row + found] = p.p_sum[at];
            s_key[warp][bi] = 0;
          }
          __syncwarp();
        }
      } else {
        // more ties than the list holds: K rounds over all entries of the row (round-1 path)
        const int n_all = (int)nbx * k;
        uint64_t bestk = 0;
        int best_i = 0;
        auto scan = [&]() {
          bestk = 0;
          for (int i = lane; i < n_all; i += 32) {
            const int j = i / k, r = i % k;
            const uint64_t kk = r < p.p_len[slot0 + j] ? p.p_key[(slot0 + j) * k + r] : 0;
            if (kk > bestk) { bestk = kk; best_i = i; }
          }
        };
        scan();
        for (; found < k; ++found) {
          const uint64_t m = warp_max_u64(bestk);
          if (m == 0) break;
          if (bestk == m) {
            const int64_t at = (slot0 + best_i / k) * k + best_i % k;
            p.out_y[row + found] = (int32_t)(~(uint32_t)m);
            p.out_w[row + found] = __uint_as_float((uint32_t)(m >> 32));
            if (p.out_cnt) p.out_cnt[row + found] = p.p_cnt[at];
            if (p.out_tsum) p.out_tsum[row + found] = p.p_sum[at];
            p.p_key[at] = 0;
            scan();
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

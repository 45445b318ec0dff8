// Accumulate + per-aid top-K with an OWNER TABLE in shared memory (round 2, v6 of the reduce phase).
//
// Same contract as reduce.cuh (builder steps 7-8 of SURVEY.md Appendix A: groupby(['aid_x','aid_y']).wgt.sum(),
// stable sort (aid_x asc, wgt desc), cumcount() < K, ties by aid_y ascending).  The v5 kernels of reduce.cuh insert
// every record into an open-addressing table and then sweep an occupied list twice; profiles/r02_reduce256_lines.txt
// shows where their 418 warp-instructions per 64 records go: 172 in the insert (62 in the probe loops, 25 in the
// occupied-list append, 30 in the payload atomics and the carry of the 40-bit time sum), 57 in the dynamic chunk
// hand-out, 120 in the two sweeps through the occupied list.  Three quarters of all pair records are the ONLY record
// of their (aid_x, aid_y), and v5 pays a payload atomic, a list append and two indirect sweeps for each of them.
//
// Here the record that claims a slot (atomicCAS, double hashing as before) OWNS the entry: its thread remembers the
// slot in a register and the record's own value never enters the table.  Only the later records of the same aid_y
// (one in four) touch the payload: one red (type / unit weights) or two (time: the low 18 bits of ts_x - ts_min into
// one word, count | high bits << 13 into the other; no carries).  Records are assigned to threads statically
// (12 per thread, a bin fills at most three quarters of its table), and the selection is one pass over the thread's
// own records: owners read their slot, form the 32-bit order key of v5 and keep it in a register; the K-th largest of
// the 32 lane-group maxima bounds the K-th best entry from below; the ~25-35 owners at or above it write their exact
// 64-bit keys to the candidate list that one warp ranks.  The table is cleared with 16-byte stores.
//
// A first version of this file bucketed the records by counting sort instead (no keys, one atomic per record, group
// folding by the first record of an aid_y): bit-exact but 3x slower than v5 - the most frequent aid_y of a row repeat
// 50-200 times, and one lane folded each of those groups serially (55 % of all instructions at 1.2-2 active lanes,
// profiles/r02_experiments.md).
//
// Tiers by records per bin: <= 384 one warp per bin (records stay in registers, __syncwarp only); <= 1536 / 3072 /
// 6144 a block of 128 / 256 / 512 threads, four barriers per bin.  Larger bins and bins whose candidate list
// overflows (more than 64 entries tie around the K-th weight) go to the v5 hash-table kernel of reduce.cuh
// (multi-pass, exact K-round selection): list[4].
#pragma once
#include "reduce.cuh"

constexpr int OT_RPT = 12;              // records per thread (block tiers) / per lane (warp tier); 16 slots per thread
constexpr int OTW_WARPS = 4;            // warps per block of the warp tier
constexpr int OTW_LOG = 9;              // 512 slots per warp

__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// default-cached 8-byte load (the block tiers read a bin twice; the second pass finds it in L1 / L2)
__device__ __forceinline__ uint2 ld_rec(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// Table over SLOTS = 2^LOG slots at shared address keys_s: keys (aid_y + 1, 0 = free) | word A | word B (TIME only).
//   TIME   A = sum of (v & 0x3ffff) of the followers, B = followers | sum of (v >> 18) << 13
//          (v = ts_x - ts_min < 2^24, at most 6143 followers: A < 2^31, followers < 2^13, high sum < 2^19)
//   !TIME  A = sum of the followers' weights
// The owner's own record is not in the table.
constexpr uint32_t OT_NOT_OWNER = 0xffffu;

template <bool TIME, int LOG>
struct OTable {
  static constexpr uint32_t SLOTS = 1u << LOG;
  static constexpr uint32_t BYTES = SLOTS * (TIME ? 12u : 8u);
  uint32_t keys_s;

  // Two records per call, both first probes in flight together.  own0 / own1 = slot the record claimed, or
  // OT_NOT_OWNER.  A lane's two records may carry the same aid_y: the second CAS then finds the first one's claim.
  __device__ __forceinline__ void insert2(bool has0, uint32_t y0, uint32_t v0, bool has1, uint32_t y1, uint32_t v1,
                                          uint32_t& own0, uint32_t& own1, bool& full) const {
    constexpr uint32_t MASK = SLOTS * 4 - 1;
    const uint32_t k0 = y0 + 1u, k1 = y1 + 1u;
    uint32_t a0 = ((y0 * 0x9E3779B1u) >> (32 - LOG)) << 2, a1 = ((y1 * 0x9E3779B1u) >> (32 - LOG)) << 2;
    uint32_t prev0 = 1u, prev1 = 1u;      // placeholders of absent records: every use below is guarded by has0 / has1
    if (has0) prev0 = atoms_cas(keys_s + a0, KEY_NONE, k0);
    if (has1) prev1 = atoms_cas(keys_s + a1, KEY_NONE, k1);
    if (has0 && prev0 != KEY_NONE && prev0 != k0) {
      const uint32_t step = (((y0 * 0x85EBCA6Bu) >> (32 - LOG)) | 1u) << 2;
      uint32_t left = SLOTS;
      do {
        a0 = (a0 + step) & MASK;
        prev0 = atoms_cas(keys_s + a0, KEY_NONE, k0);
      } while (prev0 != KEY_NONE && prev0 != k0 && --left);
      if (prev0 != KEY_NONE && prev0 != k0) full = true;
    }
    if (has1 && prev1 != KEY_NONE && prev1 != k1) {
      const uint32_t step = (((y1 * 0x85EBCA6Bu) >> (32 - LOG)) | 1u) << 2;
      uint32_t left = SLOTS;
      do {
        a1 = (a1 + step) & MASK;
        prev1 = atoms_cas(keys_s + a1, KEY_NONE, k1);
      } while (prev1 != KEY_NONE && prev1 != k1 && --left);
      if (prev1 != KEY_NONE && prev1 != k1) full = true;
    }
    own0 = (has0 && prev0 == KEY_NONE) ? (a0 >> 2) : OT_NOT_OWNER;
    own1 = (has1 && prev1 == KEY_NONE) ? (a1 >> 2) : OT_NOT_OWNER;
    if (has0 && prev0 == k0) {
      if (TIME) {
        reds_add(keys_s + SLOTS * 4 + a0, v0 & 0x3ffffu);
        reds_add(keys_s + SLOTS * 8 + a0, 1u | ((v0 >> 18) << 13));
      } else {
        reds_add(keys_s + SLOTS * 4 + a0, v0);
      }
    }
    if (has1 && prev1 == k1) {
      if (TIME) {
        reds_add(keys_s + SLOTS * 4 + a1, v1 & 0x3ffffu);
        reds_add(keys_s + SLOTS * 8 + a1, 1u | ((v1 >> 18) << 13));
      } else {
        reds_add(keys_s + SLOTS * 4 + a1, v1);
      }
    }
  }
  // entry of the owner of slot h whose own record carries v: count (TIME only) and sum over all records of the aid_y
  __device__ __forceinline__ void entry(uint32_t h, uint32_t v, uint32_t& cnt, uint64_t& sum) const {
    const uint32_t a = lds_u32(keys_s + SLOTS * 4 + h * 4);
    if (TIME) {
      const uint32_t b = lds_u32(keys_s + SLOTS * 8 + h * 4);
      cnt = 1u + (b & 0x1fffu);
      sum = (uint64_t)v + a + ((uint64_t)(b >> 13) << 18);
    } else {
      cnt = 0u;
      sum = (uint64_t)v + a;
    }
  }
  __device__ __forceinline__ void clear_all(uint32_t tid, uint32_t nthreads) const {
#pragma unroll
    for (uint32_t i = 0; i < BYTES / 16 / nthreads; ++i) sts_zero16(keys_s + (i * nthreads + tid) * 16);
  }
};

// 32-bit order key of an entry held as (y, count, sum): the same order as key32() of reduce.cuh, + 1 (0 = no entry)
template <bool TIME>
__device__ __forceinline__ uint32_t ot_key32(const KeyCfg& c, uint32_t y, uint32_t cnt, uint64_t sum) {
  uint32_t k;
  if (TIME) {
    k = (uint32_t)(((uint64_t)cnt * c.range + 3ull * sum) >> c.shift);
  } else {
    const uint32_t lo = (uint32_t)sum;
    k = c.s ? ((lo << c.s) | ((~y & c.ymask) >> c.yshift)) : lo;
  }
  k += 1u;
  return k ? k : 0xffffffffu;
}

// a bin the owner-table tiers cannot finish (candidate list overflow): the hash-table kernel takes it afterwards
__device__ __forceinline__ void ot_hand_over(const ReduceParams& p, uint32_t item) {
  const uint32_t at = atomicAdd(&p.counters[4], 1u);
  p.list[4][at] = item;
}

// =====================================================================================================
// warp tier: one warp per bin of up to 384 records
// =====================================================================================================
template <bool TIME>
struct __align__(16) OtwShared {
  unsigned char table[OTable<TIME, OTW_LOG>::BYTES];
  uint64_t ckey[N_CAND];
  uint64_t csum[N_CAND];
  uint32_t ccnt[N_CAND];
};

template <bool TIME>
__global__ void __launch_bounds__(OTW_WARPS * 32, 6) otable_warp_kernel(const ReduceParams p) {
  __shared__ OtwShared<TIME> sm_all[OTW_WARPS];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, lt = lanemask_lt();
  OtwShared<TIME>& sm = sm_all[warp];
  OTable<TIME, OTW_LOG> t;
  t.keys_s = smem_u32(sm.table);
  Cands c;
  c.key = sm.ckey;
  c.sum = sm.csum;
  c.cnt = sm.ccnt;
  t.clear_all(lane, 32);
  __syncwarp();

  BinStats st;
  bool full = false;
  const uint32_t n_items = p.counters[0];
  const uint32_t* list = p.list[0];
  // 32 consecutive list items per grab: lane l walks the dependent metadata loads of item c0 + l (list -> record
  // offsets -> bin -> aid_x -> first bin of the row), so their latency is paid once per 32 bins
  while (true) {
    uint32_t c0 = 0;
    if (lane == 0) c0 = atomicAdd(&p.counters[NEXT_ITEM + 0], 32u);
    c0 = __shfl_sync(FULL_MASK, c0, 0);
    if (c0 >= n_items) break;
    const bool inb = c0 + lane < n_items;
    uint32_t n_l = 0, item_l = 0;
    BinOut o_l;
    o_l.whole = true;
    o_l.row = 0;
    o_l.x = 0;
    const uint2* run_l = nullptr;
    if (inb) {
      item_l = list[c0 + lane];
      const int64_t bl = p.bin_lo + item_l;
      const uint64_t beg = p.offsets[bl - p.bin_lo], end = p.offsets[bl - p.bin_lo + 1];
      n_l = (uint32_t)(end - beg);
      o_l = bin_out(p, bl);
      run_l = p.records + (beg - p.offsets[0]);
    }
    const int n_here = min(32u, n_items - c0);
    for (int srcl = 0; srcl < n_here; ++srcl) {
      const uint32_t n = __shfl_sync(FULL_MASK, n_l, srcl);
      BinOut o;
      o.whole = __shfl_sync(FULL_MASK, (int)o_l.whole, srcl) != 0;
      o.row = (int64_t)shfl_u64((uint64_t)o_l.row, srcl);
      o.x = __shfl_sync(FULL_MASK, o_l.x, srcl);
      const uint2* run = (const uint2*)shfl_u64((uint64_t)(uintptr_t)run_l, srcl);
      const uint32_t item = __shfl_sync(FULL_MASK, item_l, srcl);
      if (n == 0) {
        emit_finish(p, o, 0);
        continue;
      }
      if (n <= TINY_MAX) {
        if (lane == 0) st.rec += n;
        tiny_bin<TIME>(p, o, run, n, st);
        continue;
      }
      const KeyCfg cfg = make_cfg<TIME>(p, n);
      // ---- insert: all loads in flight, then two records per step
      uint2 rec[OT_RPT];
      uint32_t own[OT_RPT];
#pragma unroll
      for (int r = 0; r < OT_RPT; ++r) {
        rec[r] = make_uint2(0, 0);
        if (r * 32 + lane < n) rec[r] = ld_stream_u2(run + r * 32 + lane);
      }
#pragma unroll
      for (int r = 0; r < OT_RPT; r += 2) {
        own[r] = own[r + 1] = OT_NOT_OWNER;
        if ((uint32_t)r * 32 < n)
          t.insert2(r * 32 + lane < n, rec[r].x, rec[r].y, (r + 1) * 32 + lane < n, rec[r + 1].x, rec[r + 1].y, own[r], own[r + 1], full);
      }
      __syncwarp();
      // ---- owners: entry, 32-bit key (kept in place of the slot), lane maxima -> threshold
      uint32_t best = 0, pay = 0, nd = 0;
      uint32_t k32[OT_RPT];
#pragma unroll
      for (int r = 0; r < OT_RPT; ++r) {
        k32[r] = 0;
        if (own[r] != OT_NOT_OWNER) {
          uint32_t cnt;
          uint64_t sum;
          t.entry(own[r], rec[r].y, cnt, sum);
          k32[r] = ot_key32<TIME>(cfg, rec[r].x, cnt, sum);
          best = max(best, k32[r]);
          pay += TIME ? cnt : (uint32_t)sum;
          ++nd;
        }
      }
      const uint32_t thr = cand_threshold<TIME>(warp_kth_largest32(best, p.k));   // 0 with fewer than K non-empty lanes
      // ---- the owners at or above the threshold, with their exact keys
      uint32_t n_c = 0;
#pragma unroll
      for (int r = 0; r < OT_RPT; ++r) {
        if ((uint32_t)r * 32 >= n) break;
        const bool q = k32[r] != 0 && k32[r] >= thr;
        const uint32_t m = __ballot_sync(FULL_MASK, q);
        if (q) {
          const uint32_t at = n_c + __popc(m & lt);
          if (at < (uint32_t)N_CAND) {
            uint32_t cnt;
            uint64_t sum;
            t.entry(own[r], rec[r].y, cnt, sum);
            c.key[at] = float_key(TIME, rec[r].x, cnt, sum, p.w_scale);
            c.sum[at] = sum;
            c.cnt[at] = cnt;
          }
        }
        n_c += __popc(m);
      }
      __syncwarp();
      t.clear_all(lane, 32);
      if (n_c > (uint32_t)N_CAND) {   // more ties around the K-th weight than the list holds
        if (lane == 0) ot_hand_over(p, item);
        __syncwarp();
        continue;
      }
      const int found = warp_rank_emit(c, (int)n_c, p.k, [&](int rr, uint64_t kk, uint32_t cnt, uint64_t sum) { emit_entry(p, o, rr, kk, cnt, sum); });
      emit_finish(p, o, found);
      st.pay += pay;
      st.occ += nd;
      if (lane == 0) st.rec += n;
      __syncwarp();
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    st.occ += shfl_u64(st.occ, lane ^ off);
    st.pay += shfl_u64(st.pay, lane ^ off);
    st.rec += shfl_u64(st.rec, lane ^ off);
  }
  if (lane == 0 && (st.occ || st.pay)) {
    atomicAdd(&p.stats[0], (unsigned long long)st.occ);
    atomicAdd(&p.stats[1], (unsigned long long)st.pay);
  }
  if (lane == 0 && st.rec) atomicAdd(&p.stats[4], (unsigned long long)st.rec);
  if (full) atomicOr(&p.stats[2], 1ull);
}

// =====================================================================================================
// block tiers: one block per bin of up to 12 * THREADS records, four barriers per bin
// =====================================================================================================
template <bool TIME, int LOG>
constexpr size_t otable_block_smem() {   // table | candidate key / sum (u64) | candidate cnt
  return (size_t)OTable<TIME, LOG>::BYTES + (size_t)N_CAND * 20;
}

template <bool TIME, int THREADS, int LOG, int TIER>
__global__ void __launch_bounds__(THREADS, THREADS == 128 ? 8 : THREADS == 256 ? 4 : 2) otable_block_kernel(const ReduceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int WARPS = THREADS / 32;
  static_assert((1u << LOG) == (uint32_t)THREADS * 16u, "16 slots per thread");
  static_assert(WARPS >= 2, "the metadata prefetch runs on warp 1");
  constexpr int BATCH = 8;
  __shared__ uint32_t s_gmax[WARPS][32], s_ncand[2], s_first[2];
  __shared__ uint32_t s_mn[2][BATCH], s_mx[2][BATCH], s_mwhole[2][BATCH], s_mitem[2][BATCH];
  __shared__ int64_t s_mrow[2][BATCH];
  __shared__ const uint2* s_mrun[2][BATCH];
  const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5, lt = lanemask_lt();
  OTable<TIME, LOG> t;
  t.keys_s = smem_u32(smem_raw);
  Cands c;
  c.key = (uint64_t*)(smem_raw + OTable<TIME, LOG>::BYTES);
  c.sum = c.key + N_CAND;
  c.cnt = (uint32_t*)(c.sum + N_CAND);

  const uint32_t* list = p.list[TIER];
  const uint32_t n_items = p.counters[TIER];
  BinStats st;
  bool full = false;
  t.clear_all(tid, THREADS);
  if (tid < 2) s_ncand[tid] = 0;
  // metadata of a batch of work items (dependent loads list -> record offsets -> bin -> aid_x -> first bin of the
  // row), fetched by BATCH lanes of warp 1 one batch ahead of the bins being processed
  auto fetch_batch = [&](int buf) {   // warp 1 only
    uint32_t first = 0;
    if (lane == 0) first = atomicAdd(&p.counters[NEXT_ITEM + TIER], (uint32_t)BATCH);
    first = __shfl_sync(FULL_MASK, first, 0);
    if (lane == 0) s_first[buf] = first;
    if (lane < BATCH && first + lane < n_items) {
      const uint32_t item = list[first + lane];
      const int64_t bb = p.bin_lo + item;
      const uint64_t beg = p.offsets[bb - p.bin_lo], end = p.offsets[bb - p.bin_lo + 1];
      const BinOut ob = bin_out(p, bb);
      s_mn[buf][lane] = (uint32_t)(end - beg);
      s_mx[buf][lane] = ob.x;
      s_mwhole[buf][lane] = ob.whole ? 1u : 0u;
      s_mrow[buf][lane] = ob.row;
      s_mrun[buf][lane] = p.records + (beg - p.offsets[0]);
      s_mitem[buf][lane] = item;
    }
  };
  if (warp == 1) fetch_batch(0);
  __syncthreads();
  int par = 0;   // parity of the running bin: which candidate counter is in use
  for (int buf = 0;; buf ^= 1) {
    const uint32_t first = s_first[buf];
    if (first >= n_items) break;
    const uint32_t n_batch = min((uint32_t)BATCH, n_items - first);
    bool fetched = false;
    for (uint32_t kb = 0; kb < n_batch; ++kb, par ^= 1) {
      const uint32_t n = s_mn[buf][kb];
      const uint2* run = s_mrun[buf][kb];
      const KeyCfg cfg = make_cfg<TIME>(p, n);
      // ---- insert: the loads are issued in front of the barrier that waits for the cleared table
      uint32_t own[OT_RPT / 2];   // two 16-bit slots per word
      {
        uint2 rec[OT_RPT];
#pragma unroll
        for (int r = 0; r < OT_RPT; ++r) {
          rec[r] = make_uint2(0, 0);
          if (r * THREADS + tid < n) rec[r] = ld_rec(run + r * THREADS + tid);
        }
        __syncthreads();   // #0: table clear, the previous bin's candidates ranked
#pragma unroll
        for (int r = 0; r < OT_RPT; r += 2) {
          uint32_t o0 = OT_NOT_OWNER, o1 = OT_NOT_OWNER;
          if ((uint32_t)r * THREADS < n)
            t.insert2(r * THREADS + tid < n, rec[r].x, rec[r].y, (r + 1) * THREADS + tid < n, rec[r + 1].x, rec[r + 1].y, o0, o1, full);
          own[r >> 1] = o0 | (o1 << 16);
        }
      }
      __syncthreads();   // #1: every record is in the table
      if (tid == 0) s_ncand[par ^ 1] = 0;   // everybody has read the previous bin's candidate count
      if (!fetched && warp == 1) fetch_batch(buf ^ 1);
      fetched = true;
      // ---- owners: entry (own record read again: L1 / L2), 32-bit key, lane-group maxima
      uint32_t k32[OT_RPT];
      uint32_t tbest = 0, pay = 0, nd = 0;
#pragma unroll
      for (int r = 0; r < OT_RPT; ++r) {
        k32[r] = 0;
        const uint32_t h = (own[r >> 1] >> ((r & 1) * 16)) & 0xffffu;
        if (h != OT_NOT_OWNER) {
          const uint2 e = ld_rec(run + r * THREADS + tid);
          uint32_t cnt;
          uint64_t sum;
          t.entry(h, e.y, cnt, sum);
          k32[r] = ot_key32<TIME>(cfg, e.x, cnt, sum);
          tbest = max(tbest, k32[r]);
          pay += TIME ? cnt : (uint32_t)sum;
          ++nd;
        }
      }
      s_gmax[warp][lane] = tbest;
      __syncthreads();   // #2
      uint32_t g = 0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) g = max(g, s_gmax[w][lane]);
      const uint32_t thr = cand_threshold<TIME>(warp_kth_largest32(g, p.k));
      // ---- the owners at or above the threshold, with their exact keys
#pragma unroll
      for (int r = 0; r < OT_RPT; ++r) {
        if ((uint32_t)r * THREADS >= n) break;
        const bool qq = k32[r] != 0 && k32[r] >= thr;
        const uint32_t m = __ballot_sync(FULL_MASK, qq);
        if (m) {
          const int leader = __ffs(m) - 1;
          uint32_t base = 0;
          if ((int)lane == leader) base = atomicAdd(&s_ncand[par], (uint32_t)__popc(m));
          base = __shfl_sync(FULL_MASK, base, leader);
          if (qq) {
            const uint32_t at = base + __popc(m & lt);
            if (at < (uint32_t)N_CAND) {
              const uint2 e = ld_rec(run + r * THREADS + tid);
              uint32_t cnt;
              uint64_t sum;
              t.entry((own[r >> 1] >> ((r & 1) * 16)) & 0xffffu, e.y, cnt, sum);
              c.key[at] = float_key(TIME, e.x, cnt, sum, p.w_scale);
              c.sum[at] = sum;
              c.cnt[at] = cnt;
            }
          }
        }
      }
      __syncthreads();   // #3: candidate list complete, nobody reads the table any more
      t.clear_all(tid, THREADS);
      const uint32_t n_c = s_ncand[par];
      if (n_c <= (uint32_t)N_CAND) {
        st.pay += pay;
        st.occ += nd;
        if (tid == 0) st.rec += n;
      }
      if (warp == 0) {
        if (n_c > (uint32_t)N_CAND) {
          if (lane == 0) ot_hand_over(p, s_mitem[buf][kb]);
        } else {
          BinOut o;
          o.x = s_mx[buf][kb];
          o.whole = s_mwhole[buf][kb] != 0;
          o.row = s_mrow[buf][kb];
          const int found = warp_rank_emit(c, (int)n_c, p.k, [&](int rr, uint64_t kk, uint32_t cnt, uint64_t sum) { emit_entry(p, o, rr, kk, cnt, sum); });
          emit_finish(p, o, found);
        }
        __syncwarp();
      }
      // barrier #0 of the next bin orders the cleared table and the free candidate list; the last bin needs none
    }
    // the next batch's metadata was written by warp 1 behind barrier #1 of this batch's first bin; barriers #2 and #3
    // of that bin order it
  }
  for (int off = 16; off > 0; off >>= 1) {
    st.occ += shfl_u64(st.occ, lane ^ off);
    st.pay += shfl_u64(st.pay, lane ^ off);
  }
  if (lane == 0 && (st.occ || st.pay)) {
    atomicAdd(&p.stats[0], (unsigned long long)st.occ);
    atomicAdd(&p.stats[1], (unsigned long long)st.pay);
  }
  if (tid == 0 && st.rec) atomicAdd(&p.stats[4 + (TIER < 3 ? TIER : 3)], (unsigned long long)st.rec);
  if (full) atomicOr(&p.stats[2], 1ull);
}

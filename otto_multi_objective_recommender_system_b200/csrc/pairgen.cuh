// Pair generation: tail CSR -> in-session deduplicated (aid_x, aid_y) pairs, grouped by bin.
//
// Replaces builder steps 3-6 of SURVEY.md Appendix A (the cuDF self-merge that is not in the reference
// repo; idiom witness: src/matrix_factorization/torch_trainer.py:198-223).
//
// Layout.  The tail CSR holds, per session, its <= tail_n most recent (type-filtered) events in
// (ts desc, row asc) order as two 4-byte columns: aw = aid | type << 30 and ts.  Because the tails of
// consecutive sessions are contiguous, a warp takes 32 consecutive sessions, then repeatedly packs as
// many whole sessions as fit into its 32 lanes (one event per lane).  One coalesced load brings the
// batch in; every later step is register / shuffle work.
//
// Pass 1 (count) finds, for every lane j (an event in its role as aid_y), the 32-bit COLUMN mask of the
// rows i (events of the same session in their role as aid_x) for which (i, j) is the row pandas keeps in
//   merge(on='session') -> |ts_x - ts_y| < W, aid_x != aid_y -> drop_duplicates(['session','aid_x','aid_y'])
// (first row in i-major, j-minor order).  Because a session is sorted by ts, the window of an event is a
// contiguous lane range [lo, hi] (found by a 5-step binary descent, skipped when the whole session spans
// less than W), which makes the dedupe O(1) per lane instead of a loop over rows (v1-v3 of this kernel):
//   V_j   = range(lo_j, hi_j) & ~same_j                      rows in window with another aid
//   y side  a row keeps the first in-window occurrence of aid_y: rows i <= hi_prev(j) already saw the
//           previous occurrence of my aid, so  V_j &= ~lowmask(hi_prev + 1)
//   x side  of several rows with the same aid_x only the first that sees ANY occurrence of aid_y keeps
//           the pair: U_j = union of V over the occurrences of my aid; row i is dropped when an earlier
//           row with the same aid is in U_j.  Only lanes that repeat an aid can be dropped, so this is a
//           loop over the (few) repeat lanes of the batch, not over rows.
// A 5-step butterfly transposes the column masks into ROW masks (bit j of row i), which are stored
// (4 B per tail event) and whose popcounts feed the per-aid_x pair histogram: one RED per (session, row).
// Pass 2 (scatter) re-reads events + row masks, reserves the row's slots with one atomic per
// (session, row) and writes 8-byte records {aid_y, v} as one contiguous run; aid_x is implicit in the position.
// v = ts_x - ts_min (time), type_weight[type_y] (type) or 1 (unit).
// Neither pass knows about sub-bins: the runs of a hot aid_x land in a staging area and are partitioned by
// aid_y hash afterwards (build.cu, partition_*_kernel).  v1-v4 of this kernel took one returning global atomic
// and one lone 8-byte store per PAIR of a hot row (30 % of all pairs); that path held 35 % of the scatter's
// stall samples and most of its partial-sector DRAM traffic (profiles/r01_final_scatter_lines.txt).
#pragma once
#include "common.cuh"

struct PairGenParams {
  const uint32_t* tail_off;   // [S + 1]
  const uint32_t* tail_aw;    // [E30]
  const int32_t* tail_ts;     // [E30]
  uint32_t* winmask;          // [E30]
  uint32_t* row_hist;         // [A]      pass 1: pairs per aid_x
  uint32_t* cursor;           // [A]      pass 2: next free record slot of the row (staging area for hot rows)
  uint2* records;             //          pass 2
  int64_t n_sessions;
  uint32_t window;
  uint32_t x_type_mask, y_type_mask;
  int32_t weight_mode;
  int32_t ts_min;
  uint32_t type_weight[3];
  // MODE 2 (owner-direct scatter, multi-GPU): row x belongs to owner o with cuts[o] <= x < cuts[o + 1]; its records
  // go to owner_rec[o] (the owner's record buffer mapped into this process, NVLink stores) at the slot the cursor
  // holds, which the host derived from the all-gathered row counts
  int32_t n_owners;
  uint32_t cuts[OTTO_MAX_OWNERS + 1];
  uint2* owner_rec[OTTO_MAX_OWNERS];
  // MODE 3 (staged scatter): the records of row x go to coarse bucket x >> STAGE_LOGA of this rank's staging buffer,
  // packed into 64 bits as aid_y | v << y_bits | (x & (2^STAGE_LOGA - 1)) << (y_bits + v_bits); bcur[b] = next free slot
  uint32_t* bcur;              // [buckets * STAGE_CUR_STRIDE]
  uint32_t y_bits, v_bits;
};

constexpr int STAGE_LOGA = 11;      // staged scatter: 2048 aid_x rows per coarse bucket
// every bucket cursor in its own 128-byte line: 125 M atomics on ~900 adjacent words serialise in ~30 L2 lines (11.5 ms
// for pass A), spread over one line each they do not (profiles/r02_experiments.md)
constexpr int STAGE_CUR_STRIDE = 32;
constexpr int PAIRGEN_WARPS = 8;

__device__ __forceinline__ uint32_t lowmask(int k) { return k >= 32 ? FULL_MASK : ((1u << k) - 1u); }

// 32 x 32 bit-matrix transpose across the lanes of a warp: on return lane i holds bit i of every lane's x
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t m = s == 16 ? 0x0000ffffu : s == 8 ? 0x00ff00ffu : s == 4 ? 0x0f0f0f0fu : s == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t o = __shfl_xor_sync(FULL_MASK, x, s);
    x = (lane & s) ? ((x & ~m) | ((o >> s) & m)) : ((x & m) | ((o << s) & ~m));
  }
  return x;
}

// MODE 0: count pass, 1: scatter into p.records, 2: scatter into the owners' buffers, 3: staged scatter into coarse buckets
template <int MODE>
__global__ void __launch_bounds__(PAIRGEN_WARPS * 32) pairgen_kernel(const PairGenParams p) {
  constexpr bool SCATTER = MODE != 0;
  const uint32_t lane = lane_id();
  const uint32_t lt = lanemask_lt();
  const int64_t warp = (int64_t)blockIdx.x * PAIRGEN_WARPS + (threadIdx.x >> 5);
  const int64_t s0 = warp * 32;
  if (s0 >= p.n_sessions) return;
  const int ns = (int)min((int64_t)32, p.n_sessions - s0);
  const bool masks_on = ((p.x_type_mask & p.y_type_mask & 7u) != 7u);

  // lane l: tail offsets of session s0 + l (clamped so every lane holds a valid pair)
  const uint32_t o = p.tail_off[s0 + min((int)lane, ns)];
  const uint32_t o_next = p.tail_off[s0 + min((int)lane + 1, ns)];

  int a = 0;  // first session (slab-relative) of the next batch; warp-uniform
  while (a < ns) {
    const uint32_t e0 = __shfl_sync(FULL_MASK, o, a);
    const uint32_t fit = __ballot_sync(FULL_MASK, (int)lane >= a && (int)lane < ns && (o_next - e0) <= 32u);
    const int b = a + __popc(fit);  // sessions [a, b) form the batch; tail_n <= 32 guarantees b > a
    const uint32_t total = __shfl_sync(FULL_MASK, o_next, b - 1) - e0;
    if (total == 0) { a = b; continue; }
    const uint32_t heads = __reduce_or_sync(
        FULL_MASK, ((int)lane >= a && (int)lane < b && o_next > o) ? (1u << (o - e0)) : 0u);
    a = b;

    const bool active = lane < total;
    // my session's lanes: [base, gend)
    const uint32_t le = heads & (FULL_MASK >> (31 - lane));
    const uint32_t gt = heads & ~(FULL_MASK >> (31 - lane));
    const int base = active ? 31 - __clz(le) : (int)lane;
    const int gend = active ? (gt ? __ffs(gt) - 1 : (int)total) : (int)lane + 1;
    const int n = gend - base;
    const int maxn = __reduce_max_sync(FULL_MASK, active ? n : 0);

    const uint32_t aw = active ? p.tail_aw[e0 + lane] : (0x80000000u | lane);
    const int32_t t = active ? p.tail_ts[e0 + lane] : 0;
    const uint32_t aid = aw & AID_MASK;
    const uint32_t ty = aw >> 30;

    if (!SCATTER) {
      const uint32_t grpmask = active ? (lowmask(n) << base) : 0u;
      const uint32_t same = __match_any_sync(FULL_MASK, active ? aid : (0x80000000u | lane)) & grpmask;
      bool y_ok = active, x_ok = active;
      uint32_t xrows = FULL_MASK, yok_m = FULL_MASK;
      if (masks_on) {
        y_ok = active && ((p.y_type_mask >> ty) & 1u);
        x_ok = active && ((p.x_type_mask >> ty) & 1u);
        xrows = __ballot_sync(FULL_MASK, x_ok);
        yok_m = __ballot_sync(FULL_MASK, y_ok);
      }
      // window [lo, hi] of my event (lane indices); a session is sorted by ts descending
      int lo = base, hi = gend - 1;
      const int32_t t_first = __shfl_sync(FULL_MASK, t, base);
      const int32_t t_last = __shfl_sync(FULL_MASK, t, gend - 1);
      if (__any_sync(FULL_MASK, active && (uint32_t)(t_first - t_last) >= p.window)) {
        lo = hi = (int)lane;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
          const int cl = lo - step, ch = hi + step;
          const bool okl = active && cl >= base, okh = active && ch < gend;
          const int32_t tl = __shfl_sync(FULL_MASK, t, okl ? cl : (int)lane);
          const int32_t th = __shfl_sync(FULL_MASK, t, okh ? ch : (int)lane);
          if (okl && (uint32_t)(tl - t) < p.window) lo = cl;
          if (okh && (uint32_t)(t - th) < p.window) hi = ch;
        }
      }
      uint32_t V = y_ok ? ((lowmask(hi - lo + 1) << lo) & ~same & xrows) : 0u;
      // y side: rows up to hi(previous occurrence of my aid that may act as aid_y) keep that occurrence
      const uint32_t prevs = same & lt & yok_m;
      const int prev = prevs ? 31 - __clz(prevs) : (int)lane;
      const int hi_prev = __shfl_sync(FULL_MASK, hi, prev);
      if (prevs) V &= ~lowmask(hi_prev + 1);
      // x side: only lanes that repeat an aid can lose a pair to an earlier row with the same aid
      const uint32_t rep = __ballot_sync(FULL_MASK, active && (same & lt) != 0);
      if (rep) {
        const int first = same ? __ffs(same) - 1 : (int)lane;
        uint32_t U = V | __shfl_sync(FULL_MASK, V, first);
        for (uint32_t m = rep; m; m &= m - 1) {
          const int d = __ffs(m) - 1;
          const uint32_t vd = __shfl_sync(FULL_MASK, V, d);
          if ((same >> d) & 1u) U |= vd;
        }
        const uint32_t earlier_x = same & lt & xrows;     // earlier rows with my aid that may act as aid_x
        for (uint32_t m = rep; m; m &= m - 1) {
          const int i = __ffs(m) - 1;
          const uint32_t ei = __shfl_sync(FULL_MASK, earlier_x, i);
          if (((V >> i) & 1u) && (ei & U)) V &= ~(1u << i);
        }
      }
      // column masks -> row masks; lane i now owns the winners of row i
      const uint32_t R = warp_transpose32(V);
      if (active) p.winmask[e0 + lane] = R >> base;
      const uint32_t cnt = __popc(R);
      if (active && cnt) atomicAdd(&p.row_hist[aid], cnt);
    } else {
      const uint32_t mywm = active ? (p.winmask[e0 + lane] << base) : 0u;
      const uint32_t cnt = __popc(mywm);
      uint32_t slot = 0, xl = 0;
      if (MODE == 3) {
        xl = aid & ((1u << STAGE_LOGA) - 1u);
        if (active && cnt) slot = atomicAdd(&p.bcur[(aid >> STAGE_LOGA) * STAGE_CUR_STRIDE], cnt);
      } else if (active && cnt) {
        slot = atomicAdd(&p.cursor[aid], cnt);
      }
      uint2* dst = p.records;
      if (MODE == 2) {
        dst = p.owner_rec[0];
#pragma unroll
        for (int g = 1; g < OTTO_MAX_OWNERS; ++g)
          if (g < p.n_owners && aid >= p.cuts[g]) dst = p.owner_rec[g];
      }
      uint32_t v = 1;
      if (p.weight_mode == OTTO_WEIGHT_TYPE) v = ty == 0 ? p.type_weight[0] : (ty == 1 ? p.type_weight[1] : p.type_weight[2]);
      const bool time_mode = p.weight_mode == OTTO_WEIGHT_TIME;
      const int32_t tv = t - p.ts_min;
      // row i's winners write one contiguous run
      for (int i = 0; i < maxn; ++i) {
        const int src = (active && i < n) ? base + i : (int)lane;   // own row mask never holds the own lane
        const uint32_t wmi = __shfl_sync(FULL_MASK, mywm, src);
        const uint32_t sloti = __shfl_sync(FULL_MASK, slot, src);
        const uint32_t val = time_mode ? (uint32_t)__shfl_sync(FULL_MASK, tv, src) : v;
        uint2* dsti = p.records;
        if (MODE == 2) dsti = (uint2*)__shfl_sync(FULL_MASK, (unsigned long long)dst, src);
        if (MODE == 3) {
          const uint32_t xli = __shfl_sync(FULL_MASK, xl, src);
          const unsigned long long rec = (unsigned long long)aid | ((unsigned long long)val << p.y_bits) |
                                         ((unsigned long long)xli << (p.y_bits + p.v_bits));
          if ((wmi >> lane) & 1u) st_stream_u2(dsti + (sloti + __popc(wmi & lt)), make_uint2((uint32_t)rec, (uint32_t)(rec >> 32)));
        } else if ((wmi >> lane) & 1u) st_stream_u2(dsti + (sloti + __popc(wmi & lt)), make_uint2(aid, val));
      }
    }
  }
}

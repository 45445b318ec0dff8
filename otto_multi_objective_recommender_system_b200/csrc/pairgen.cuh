// Pair generation: tail CSR -> in-session deduplicated (aid_x, aid_y) pairs, grouped by bin.
//
// Replaces builder steps 3-6 of SURVEY.md Appendix A (the cuDF self-merge that is not in the reference
// repo; idiom witness: src/matrix_factorization/torch_trainer.py:198-223).
//
// Layout.  The tail CSR holds, per session, its <= tail_n most recent (type-filtered) events in
// (ts desc, row asc) order as two 4-byte columns: aw = aid | type << 30 and ts.  Because the tails of
// consecutive sessions are contiguous, a warp takes 32 consecutive sessions, then repeatedly packs as
// many whole sessions as fit into its 32 lanes (one event per lane).  One coalesced load brings the
// batch in; every later step is register / shuffle work:
//   row loop i = 0..max_n-1: each session group broadcasts its i-th event (aid_x, ts_x, type_x) and every
//   lane j of the group decides "pair (i, j) valid" (window, aid_x != aid_y, type masks).  The dedupe
//   winner of pandas' drop_duplicates(['session','aid_x','aid_y']) (first row in i-major, j-minor
//   order) is found with three 32-bit masks per lane:
//     same   lanes of my session holding my aid          (one __match_any_sync per batch)
//     v      ballot of valid lanes of this row
//     cov    rows (by owner lane) in which my aid was valid as aid_y
//   lane j wins row i  <=>  valid  &&  no earlier lane of the row has my aid (v & same & lt == 0)
//                           &&  no earlier row with the same aid_x covered my aid (cov & samelt_x == 0).
//   The winner mask of row i is kept by the lane that owns event i ("row owner").
// Pass 1 (count) stores the group-relative winner masks (4 B per tail event) and adds the per-row pair
// counts into the bin histogram: one RED per (session, row) for ordinary rows, one per pair for rows that
// are split into aid_y-hash sub-bins (hot aid_x).  Pass 2 (scatter) re-reads events + masks, reserves the
// row's slots with one atomic per (session, row) and writes 8-byte records {aid_y, v}; aid_x is implicit
// in the bin.  v = ts_x - ts_min (time), type_weight[type_y] (type) or 1 (unit).
#pragma once
#include "common.cuh"

struct PairGenParams {
  const uint32_t* tail_off;   // [S + 1]
  const uint32_t* tail_aw;    // [E30]
  const int32_t* tail_ts;     // [E30]
  uint32_t* winmask;          // [E30]
  const uint32_t* bin_base;   // [A + 1]
  uint32_t* hist;             // [B]      pass 1
  unsigned long long* cursor; // [B]      pass 2 (starts as the exclusive scan of hist)
  uint2* records;             //          pass 2
  int64_t n_sessions;
  uint32_t window;
  uint32_t x_type_mask, y_type_mask;
  int32_t weight_mode;
  int32_t ts_min;
  uint32_t type_weight[3];
};

__device__ __forceinline__ uint32_t abs_diff_i32(int32_t a, int32_t b) {
  return a > b ? (uint32_t)a - (uint32_t)b : (uint32_t)b - (uint32_t)a;
}

constexpr int PAIRGEN_WARPS = 8;

template <bool SCATTER>
__global__ void __launch_bounds__(PAIRGEN_WARPS * 32) pairgen_kernel(const PairGenParams p) {
  const uint32_t lane = lane_id();
  const uint32_t lt = lanemask_lt();
  const int64_t warp = (int64_t)blockIdx.x * PAIRGEN_WARPS + (threadIdx.x >> 5);
  const int64_t s0 = warp * 32;
  if (s0 >= p.n_sessions) return;
  const int ns = (int)min((int64_t)32, p.n_sessions - s0);

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
    // my session's lanes: [base, base + n)
    const uint32_t le = heads & (FULL_MASK >> (31 - lane));
    const uint32_t gt = heads & ~(FULL_MASK >> (31 - lane));
    const int base = active ? 31 - __clz(le) : (int)lane;
    const int gend = active ? (gt ? __ffs(gt) - 1 : (int)total) : (int)lane + 1;
    const int n = gend - base;
    const uint32_t grpmask = active ? ((n >= 32 ? FULL_MASK : ((1u << n) - 1u)) << base) : 0u;
    const int maxn = __reduce_max_sync(FULL_MASK, active ? n : 0);

    const uint32_t aw = active ? p.tail_aw[e0 + lane] : (0x80000000u | lane);
    const int32_t t = active ? p.tail_ts[e0 + lane] : 0;
    const uint32_t aid = aw & AID_MASK;
    const uint32_t ty = aw >> 30;

    // row-owner view: bins of my aid_x
    uint32_t bb0 = 0, nbx = 1;
    if (active) {
      bb0 = p.bin_base[aid];
      nbx = p.bin_base[aid + 1] - bb0;
    }

    uint32_t mywm = 0;  // winner lanes of the row I own (absolute lane bits)
    if (!SCATTER) {
      const uint32_t same = __match_any_sync(FULL_MASK, active ? aid : (0x80000000u | lane)) & grpmask;
      const uint32_t samelt = same & lt;
      const bool y_ok = active && ((p.y_type_mask >> ty) & 1u);
      uint32_t cov = 0;
      for (int i = 0; i < maxn; ++i) {
        const bool rowact = active && i < n;
        const int src = rowact ? base + i : (int)lane;
        const uint32_t axw = __shfl_sync(FULL_MASK, aw, src);
        const int32_t tx = __shfl_sync(FULL_MASK, t, src);
        const uint32_t sx = __shfl_sync(FULL_MASK, samelt, src);
        const bool valid = rowact && y_ok && ((axw & AID_MASK) != aid) && (abs_diff_i32(tx, t) < p.window) &&
                           ((p.x_type_mask >> (axw >> 30)) & 1u);
        const uint32_t v = __ballot_sync(FULL_MASK, valid) & grpmask;
        const bool win = valid && !(v & samelt) && !(cov & sx);
        const uint32_t wmi = __ballot_sync(FULL_MASK, win) & grpmask;
        if (rowact && (int)lane == src) mywm = wmi;
        if (v & same) cov |= 1u << src;
      }
      if (active) p.winmask[e0 + lane] = mywm >> base;
      const uint32_t cnt = __popc(mywm);
      if (active && cnt && nbx == 1) atomicAdd(&p.hist[bb0], cnt);
      // split rows: one RED per pair into the aid_y-hash sub-bin
      if (__ballot_sync(FULL_MASK, active && cnt && nbx > 1)) {
        for (int i = 0; i < maxn; ++i) {
          const bool rowact = active && i < n;
          const int src = rowact ? base + i : (int)lane;
          const uint32_t wmi = __shfl_sync(FULL_MASK, mywm, src);
          const uint32_t nbi = __shfl_sync(FULL_MASK, nbx, src);
          const uint32_t bbi = __shfl_sync(FULL_MASK, bb0, src);
          if (rowact && nbi > 1 && ((wmi >> lane) & 1u)) atomicAdd(&p.hist[bbi + sub_bin(aid, nbi)], 1u);
        }
      }
    } else {
      mywm = active ? (p.winmask[e0 + lane] << base) : 0u;
      const uint32_t cnt = __popc(mywm);
      unsigned long long slot = 0;
      if (active && cnt && nbx == 1) slot = atomicAdd(&p.cursor[bb0], (unsigned long long)cnt);
      uint32_t v = 1;
      if (p.weight_mode == OTTO_WEIGHT_TYPE) v = ty == 0 ? p.type_weight[0] : (ty == 1 ? p.type_weight[1] : p.type_weight[2]);
      if (__ballot_sync(FULL_MASK, cnt != 0)) {
        for (int i = 0; i < maxn; ++i) {
          const bool rowact = active && i < n;
          const int src = rowact ? base + i : (int)lane;
          const uint32_t wmi = __shfl_sync(FULL_MASK, mywm, src);
          const unsigned long long sloti = shfl_u64(slot, src);
          const uint32_t nbi = __shfl_sync(FULL_MASK, nbx, src);
          const uint32_t bbi = __shfl_sync(FULL_MASK, bb0, src);
          const int32_t tx = __shfl_sync(FULL_MASK, t, src);
          if (rowact && ((wmi >> lane) & 1u)) {
            unsigned long long pos;
            if (nbi == 1) pos = sloti + __popc(wmi & lt);
            else pos = atomicAdd(&p.cursor[bbi + sub_bin(aid, nbi)], 1ull);
            const uint32_t val = p.weight_mode == OTTO_WEIGHT_TIME ? (uint32_t)(tx - p.ts_min) : v;
            st_stream_u2(p.records + pos, make_uint2(aid, val));
          }
        }
      }
    }
  }
}

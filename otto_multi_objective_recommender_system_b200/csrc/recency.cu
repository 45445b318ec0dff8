// Long-session branch of the standalone covisitation model: sessions with >= 20 unique aids are not served by
// the vote count but by recency-weighted event scores plus small covisitation bonuses
// (src/covisitation/inference.py:142-199 validation, :336-392 submission):
//   w_click[i] = 2 ** linspace(0.1, 1, L)[i] - 1,  w_cart = w_order = 2 ** linspace(0.5, 1, L) - 1     (file order)
//   Counter[aid] += w_t[i] * {0: 1, 1: 9, 2: 6}[type_i]                         for every event, in file order
//   clicks  += 0.05 per occurrence in time_weighted[a]  for a in np.unique(aids[type == 0])
//   carts   += 0.05 per occurrence in cart_weighted[a]  for a in np.unique(aids[type <= 1])
//   orders  += 0.15 per occurrence in cart_order[a]     for a in np.unique(aids[type >= 1])
//   prediction = [aid for aid, w in Counter.most_common(20)]      (weight desc, first insertion asc)
// The fastText / Annoy neighbour bonus (:165-170) is not on this path (no model offline).
// Everything is fp64 and bit-exact against the Python loop: the recency weights come from the host (numpy, the
// same libm pow), every `+=` is one __dadd_rn in the reference's order (an aid's events in file order, then its
// bonuses one by one), products are __dmul_rn (no FMA contraction).
// Long sessions are rare (about one test session in fifty) and at most a few hundred events: one 128-thread
// block per session over a global scratch slab; this kernel is not on the throughput path.
#include <string.h>

#include "common.cuh"

struct RecencyParams {
  const int32_t* off;
  const int32_t* aid;
  const uint8_t* type;
  const int32_t* list;        // session indices to process
  int32_t n_list;
  int64_t n_sessions;
  int32_t n_aids;
  // per target: table of the bonus, history selector (OTTO_HIST_*), bonus value
  const int32_t* table_aid_y[3];
  const int32_t* table_len[3];
  int32_t table_k[3];
  int32_t hist[3];
  double bonus[3];
  double coeff[3];
  const double* w_click;      // concatenated per session length: w[offset[L] + i]
  const double* w_cart;
  const int64_t* w_offset;    // [max_len + 1]
  int32_t n;                  // predictions per target (20)
  int32_t* pred;              // [3][rows][n]; row = session index, or position in `list` when rows_by_list
  double* score;              // optional [3][rows][n]: the weights of pred
  int32_t* out_len;           // optional [3][rows]: entries written (the rest is -1 / 0.0)
  int64_t rows;
  int32_t rows_by_list;
  // scratch slab per block
  uint64_t* slab;
  int64_t slab_words;
  int32_t lcap, hs;           // events capacity, hash slots (power of two)
  int32_t max_k;              // largest table_k among the present tables
};

constexpr int REC_THREADS = 128;   // threads per session (see RECENCY_BLOCKS)

struct RecWork {
  int32_t* uaid;      // [lcap] unique aids in file order of first occurrence
  int32_t* utm;       // [lcap] type mask per unique aid
  double* ubase;      // [lcap] recency score per unique aid
  uint64_t* ord;      // [pow2 lcap] subset sort
  int32_t* sel;       // [lcap] sorted subset
  uint32_t* keys;     // [hs]
  uint32_t* cnt;      // [hs] bonus occurrences
  uint32_t* first;    // [hs] first position among the gathered neighbours
  int32_t* sidx;      // [hs] index among the session's unique aids, or -1
  uint32_t* occ;      // [hs]
  double* val;        // [hs] final weight per occupied entry (indexed like occ)
  uint32_t* pos;      // [hs] insertion order per occupied entry
};

__device__ __forceinline__ uint32_t rec_find_or_claim(const RecWork& w, uint32_t hmask, int hshift, uint32_t y, int32_t* n_occ) {
  uint32_t h = (y * 0x9E3779B1u) >> hshift;
  while (true) {
    const uint32_t prev = atomicCAS(&w.keys[h], KEY_EMPTY, y);
    if (prev == KEY_EMPTY) {
      w.occ[atomicAdd(n_occ, 1)] = h;
      return h;
    }
    if (prev == y) return h;
    h = (h + 1) & hmask;
  }
}

__global__ void __launch_bounds__(REC_THREADS) recency_long_kernel(const RecencyParams p) {
  constexpr int T = REC_THREADS;
  __shared__ int32_t s_n_occ, s_U, s_ne;
  __shared__ unsigned long long s_best_v[T / 32];
  __shared__ uint32_t s_best_p[T / 32], s_best_i[T / 32];
  constexpr int REC_CANDS = 96;
  __shared__ unsigned long long s_cv[REC_CANDS];
  __shared__ uint32_t s_cq[REC_CANDS], s_ci[REC_CANDS];
  __shared__ int s_selected;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* base = (unsigned char*)(p.slab + (int64_t)blockIdx.x * p.slab_words);
  RecWork w;
  const int64_t lcap = p.lcap, hs = p.hs;
  int64_t lp = 1;
  while (lp < lcap) lp <<= 1;
  w.ubase = (double*)base;
  w.val = w.ubase + lcap;
  w.ord = (uint64_t*)(w.val + hs);
  w.uaid = (int32_t*)(w.ord + lp);
  w.utm = w.uaid + lcap;
  w.sel = w.utm + lcap;
  w.keys = (uint32_t*)(w.sel + lcap);
  w.cnt = w.keys + hs;
  w.first = w.cnt + hs;
  w.sidx = (int32_t*)(w.first + hs);
  w.occ = (uint32_t*)(w.sidx + hs);
  w.pos = w.occ + hs;

  // the slab is sized for the longest session (tens of thousands of slots): clear it once, afterwards only the
  // claimed slots are reset (a full reset per session and target wrote 40 GB for 27 k sessions)
  // ... and only as far as the longest session this block will see needs it
  __shared__ int s_lmax;
  if (tid == 0) s_lmax = 0;
  __syncthreads();
  {
    int lmax = 0;
    for (int64_t item = blockIdx.x + (int64_t)tid * gridDim.x; item < p.n_list; item += (int64_t)T * gridDim.x) {
      const int64_t s = p.list[item];
      lmax = max(lmax, p.off[s + 1] - p.off[s]);
    }
    if (lmax) atomicMax(&s_lmax, lmax);
  }
  __syncthreads();
  int64_t hs_clear = 64;
  while (hs_clear < 2 * (int64_t)s_lmax * (1 + p.max_k) && hs_clear < hs) hs_clear <<= 1;
  for (int64_t h = tid; h < hs_clear; h += T) {
    w.keys[h] = KEY_EMPTY;
    w.cnt[h] = 0;
    w.first[h] = 0xffffffffu;
    w.sidx[h] = -1;
  }
  __syncthreads();
  for (int item = blockIdx.x; item < p.n_list; item += gridDim.x) {
    const int64_t s = p.list[item];
    const int32_t beg = p.off[s], end = p.off[s + 1];
    const int L = end - beg;
    const double* wc = p.w_click + p.w_offset[L];
    const double* wk = p.w_cart + p.w_offset[L];
    // this session's share of the table: the slab is sized for the longest session, but a typical long session has
    // ~50 events, and 4 k slots stay in L2 where 32 k slots per block (1.3 GB over all blocks) went to DRAM
    int64_t hs_s = 64;
    while (hs_s < 2 * (int64_t)L * (1 + p.max_k) && hs_s < hs) hs_s <<= 1;
    const uint32_t hmask = (uint32_t)(hs_s - 1);
    const int hshift = 32 - (63 - __clzll((long long)hs_s));
    // unique aids in FILE order of first occurrence (Counter insertion order) + type masks
    if (tid == 0) s_U = 0;
    __syncthreads();
    for (int i = tid; i < L; i += T) {
      const int32_t a = p.aid[beg + i];
      bool first = true;
      for (int j = 0; j < i; ++j)
        if (p.aid[beg + j] == a) { first = false; break; }
      w.sel[i] = first ? 1 : 0;      // temporary flag
    }
    __syncthreads();
    for (int i = tid; i < L; i += T) {
      if (w.sel[i]) {
        int u = 0;
        for (int j = 0; j < i; ++j) u += w.sel[j];
        w.uaid[u] = p.aid[beg + i];
        atomicAdd(&s_U, 1);
      }
    }
    __syncthreads();
    const int U = s_U;
    for (int u = tid; u < U; u += T) {
      const int32_t a = w.uaid[u];
      int m = 0;
      for (int i = 0; i < L; ++i)
        if (p.aid[beg + i] == a) m |= 1 << p.type[beg + i];
      w.utm[u] = m;
    }
    __syncthreads();

    for (int tg = 0; tg < 3; ++tg) {
      const double* wt = tg == 0 ? wc : wk;
      // the session's aids enter first, in insertion order (the table is clean: see the reset after the selection)
      if (tid == 0) s_n_occ = 0;
      __syncthreads();
      for (int u = tid; u < U; u += T) {
        const int32_t a = w.uaid[u];
        // recency score: this aid's events in file order, one rounded add each
        double acc = 0.0;
        for (int i = 0; i < L; ++i)
          if (p.aid[beg + i] == a) acc = __dadd_rn(acc, __dmul_rn(wt[i], p.coeff[p.type[beg + i]]));
        w.ubase[u] = acc;
        const uint32_t h = rec_find_or_claim(w, hmask, hshift, (uint32_t)a, &s_n_occ);
        w.sidx[h] = u;
      }
      __syncthreads();
      // bonus source: np.unique(aids[type filter]) ascending
      const int sel = p.hist[tg];
      const uint32_t mask = sel == OTTO_HIST_TYPE_LE1 ? 3u : sel == OTTO_HIST_TYPE_GE1 ? 6u : sel == OTTO_HIST_TYPE_EQ0 ? 1u : 7u;
      int np2 = 1;
      while (np2 < U) np2 <<= 1;
      for (int u = tid; u < np2; u += T) w.ord[u] = (u < U && (w.utm[u] & mask)) ? (uint64_t)(uint32_t)w.uaid[u] : ~0ull;
      __syncthreads();
      for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int t = tid; t < (np2 >> 1); t += T) {
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
            const uint64_t x = w.ord[i], y = w.ord[l];
            if ((x > y) == ((i & k) == 0)) { w.ord[i] = y; w.ord[l] = x; }
          }
          __syncthreads();
        }
      if (tid == 0) {
        int c = 0;
        while (c < np2 && w.ord[c] != ~0ull) ++c;
        s_ne = c;
      }
      __syncthreads();
      const int ne = s_ne;
      const int K = p.table_k[tg];
      const int32_t* tlen = p.table_len[tg];
      const int32_t* ty = p.table_aid_y[tg];
      if (ty != nullptr) {
        // positions in the concatenation: element e contributes rows [e * K, e * K + len) - any order-preserving
        // numbering does, only the order of first appearance matters
        for (int idx = tid; idx < ne * K; idx += T) {
          const int e = idx / K, r = idx - e * K;
          const int32_t a = (int32_t)w.ord[e];
          if (a >= 0 && a < p.n_aids && r < tlen[a]) {
            const uint32_t y = (uint32_t)ty[(int64_t)a * K + r];
            const uint32_t h = rec_find_or_claim(w, hmask, hshift, y, &s_n_occ);
            atomicAdd(&w.cnt[h], 1u);
            atomicMin(&w.first[h], (uint32_t)idx);
          }
        }
      }
      __syncthreads();
      const int d = s_n_occ;
      // final weights and insertion order of every Counter entry
      for (int i = tid; i < d; i += T) {
        const uint32_t h = w.occ[i];
        const int u = w.sidx[h];
        double v = u >= 0 ? w.ubase[u] : 0.0;
        for (uint32_t c = 0; c < w.cnt[h]; ++c) v = __dadd_rn(v, p.bonus[tg]);
        w.val[i] = v;
        w.pos[i] = u >= 0 ? (uint32_t)u : (uint32_t)U + w.first[h];
      }
      __syncthreads();
      // most_common(n): min(n, entries) rounds of (weight desc, insertion asc) selection; weights are positive doubles
      const int64_t row = p.rows_by_list ? (int64_t)item : s;
      int32_t* out = p.pred + ((int64_t)tg * p.rows + row) * p.n;
      double* out_score = p.score ? p.score + ((int64_t)tg * p.rows + row) * p.n : nullptr;
      const int rounds = d < p.n ? d : p.n;
      for (int r = rounds + tid; r < p.n; r += T) {
        out[r] = -1;
        if (out_score) out_score[r] = 0.0;
      }
      if (tid == 0 && p.out_len) p.out_len[(int64_t)tg * p.rows + row] = rounds;
      // Fast selection (n <= 32, the standalone model's 20): one warp takes the n-th largest of its 32 lane maxima in
      // the full order (weight desc, insertion asc) as a lower bound of the n-th best entry, collects the entries at or
      // above it (the order is strict, so there are about n .. 2n of them) and ranks those.  The n rounds of block-wide
      // arg-max below cost 60 rounds x 8 warps x ~150 instructions per session (two thirds of this kernel's 112 k
      // warp-instructions per session, profiles/r02_cand_full_launch_table.txt); they stay as the path for n > 32
      // (otto_recency_scored keeps every aid) and for a candidate list that overflows.
      bool selected = false;
      if (p.n <= 32 && rounds > 0) {
        if (warp == 0) {
          auto gt = [](unsigned long long av, uint32_t aq, unsigned long long bv, uint32_t bq) { return av > bv || (av == bv && aq > bq); };
          unsigned long long bv = 0;
          uint32_t bq = 0;
          for (int i = lane; i < d; i += 32) {
            const unsigned long long v = (unsigned long long)__double_as_longlong(w.val[i]);
            const uint32_t q = ~w.pos[i];
            if (gt(v, q, bv, bq)) { bv = v; bq = q; }
          }
          unsigned long long sv = bv;
          uint32_t sq = bq;
#pragma unroll
          for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
              const unsigned long long ov = shfl_u64(sv, lane ^ j);
              const uint32_t oq = __shfl_xor_sync(FULL_MASK, sq, j);
              const bool keep_max = ((lane & j) == 0) == ((lane & kk) == 0);
              const bool take = keep_max ? gt(ov, oq, sv, sq) : gt(sv, sq, ov, oq);
              if (take) { sv = ov; sq = oq; }
            }
          }
          const unsigned long long tv = shfl_u64(sv, rounds - 1);   // (0, 0) when fewer than `rounds` lanes hold an entry: d < 32
          const uint32_t tq = __shfl_sync(FULL_MASK, sq, rounds - 1);
          int n_c = 0;
          for (int i0 = 0; i0 < d; i0 += 32) {
            const int i = i0 + lane;
            unsigned long long v = 0;
            uint32_t q = 0;
            if (i < d) {
              v = (unsigned long long)__double_as_longlong(w.val[i]);
              q = ~w.pos[i];
            }
            const bool cand = i < d && !gt(tv, tq, v, q);
            const uint32_t m = __ballot_sync(FULL_MASK, cand);
            if (cand) {
              const int at = n_c + __popc(m & ((1u << lane) - 1u));
              if (at < REC_CANDS) { s_cv[at] = v; s_cq[at] = q; s_ci[at] = (uint32_t)i; }
            }
            n_c += __popc(m);
          }
          __syncwarp();
          if (n_c <= REC_CANDS) {
            for (int c = lane; c < n_c; c += 32) {
              const unsigned long long v = s_cv[c];
              const uint32_t q = s_cq[c];
              int rank = 0;
              for (int j = 0; j < n_c; ++j) rank += gt(s_cv[j], s_cq[j], v, q) ? 1 : 0;
              if (rank < rounds) {
                out[rank] = (int32_t)w.keys[w.occ[s_ci[c]]];
                if (out_score) out_score[rank] = __longlong_as_double((long long)v);
              }
            }
          }
          if (lane == 0) s_selected = n_c <= REC_CANDS ? 1 : 0;
        }
        __syncthreads();
        selected = s_selected != 0;
      }
      for (int r = 0; r < rounds && !selected; ++r) {
        unsigned long long bv = 0;
        uint32_t bp = 0xffffffffu, bi = 0xffffffffu;
        for (int i = tid; i < d; i += T) {
          const unsigned long long v = (unsigned long long)__double_as_longlong(w.val[i]);
          const uint32_t ps = w.pos[i];
          if (ps != 0xffffffffu && (bi == 0xffffffffu || v > bv || (v == bv && ps < bp))) { bv = v; bp = ps; bi = (uint32_t)i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long ov = shfl_u64(bv, lane ^ o);
          const uint32_t op = __shfl_xor_sync(FULL_MASK, bp, o), oi = __shfl_xor_sync(FULL_MASK, bi, o);
          if (oi != 0xffffffffu && (bi == 0xffffffffu || ov > bv || (ov == bv && op < bp))) { bv = ov; bp = op; bi = oi; }
        }
        if (lane == 0) { s_best_v[warp] = bv; s_best_p[warp] = bp; s_best_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
          unsigned long long v = s_best_v[0];
          uint32_t ps = s_best_p[0], ix = s_best_i[0];
          for (int k = 1; k < T / 32; ++k)
            if (s_best_i[k] != 0xffffffffu && (ix == 0xffffffffu || s_best_v[k] > v || (s_best_v[k] == v && s_best_p[k] < ps))) {
              v = s_best_v[k]; ps = s_best_p[k]; ix = s_best_i[k];
            }
          out[r] = ix == 0xffffffffu ? -1 : (int32_t)w.keys[w.occ[ix]];
          if (out_score) out_score[r] = ix == 0xffffffffu ? 0.0 : __longlong_as_double((long long)v);
          if (ix != 0xffffffffu) w.pos[ix] = 0xffffffffu;     // taken
        }
        __syncthreads();
      }
      for (int i = tid; i < d; i += T) {
        const uint32_t h = w.occ[i];
        w.keys[h] = KEY_EMPTY;
        w.cnt[h] = 0;
        w.first[h] = 0xffffffffu;
        w.sidx[h] = -1;
      }
      __syncthreads();
    }
  }
}

static int64_t recency_slab_words(int64_t lcap, int64_t hs) {
  int64_t lp = 1;
  while (lp < lcap) lp <<= 1;
  return (lcap * 8 + hs * 8 + lp * 8 + lcap * 4 * 3 + hs * 4 * 6 + 7) / 8 + 2;
}
static void recency_caps(int32_t max_len, int32_t max_k, int64_t* lcap, int64_t* hs) {
  *lcap = max_len > 1 ? max_len : 1;
  int64_t need = 2 * ((int64_t)max_len * (1 + max_k));
  int64_t h = 64;
  while (h < need) h <<= 1;
  *hs = h;
}
constexpr int RECENCY_BLOCKS = 2368;   // 16 per SM: the kernel is bound by latency (global-memory table, barriers), not by throughput -
                                       // sessions in flight are what counts (256 threads x 8 per SM: 4.5 ms; 128 x 16: see profiles/)

extern "C" int64_t otto_recency_scratch_bytes(int32_t max_session_len, int32_t max_table_k) {
  int64_t lcap, hs;
  recency_caps(max_session_len, max_table_k, &lcap, &hs);
  return RECENCY_BLOCKS * recency_slab_words(lcap, hs) * 8 + 256;
}

extern "C" int otto_recency_scored(const OttoSessions* sessions, const int32_t* session_list, int32_t n_list,
                                   int32_t max_session_len, const OttoRecencySpec* spec, void* scratch, int64_t scratch_bytes,
                                   int32_t rows_by_list, int32_t* pred, double* score, int32_t* len, void* stream) {
  if (!sessions || !spec || !pred) { otto_set_error("NULL argument"); return OTTO_EINVAL; }
  if (n_list <= 0) return OTTO_OK;
  if (!session_list) { otto_set_error("session_list is NULL"); return OTTO_EINVAL; }
  if (spec->n < 1 || spec->n > 4096) { otto_set_error("n must be in [1, 4096]"); return OTTO_EINVAL; }
  int max_k = 1;
  for (int t = 0; t < 3; ++t) {
    if (spec->table_aid_y[t] && (spec->table_k[t] < 1 || spec->table_k[t] > OTTO_MAX_K)) { otto_set_error("table_k must be in [1, 32]"); return OTTO_EINVAL; }
    if (spec->table_aid_y[t] && spec->table_k[t] > max_k) max_k = spec->table_k[t];
  }
  const int64_t need = otto_recency_scratch_bytes(max_session_len, max_k);
  if (!scratch || scratch_bytes < need) { otto_set_error("recency scratch too small: need %lld bytes", (long long)need); return OTTO_ENOSPC; }
  RecencyParams p;
  memset(&p, 0, sizeof(p));
  p.off = sessions->session_offsets;
  p.aid = sessions->aid;
  p.type = sessions->type;
  p.list = session_list;
  p.n_list = n_list;
  p.n_sessions = sessions->n_sessions;
  p.n_aids = spec->n_aids;
  for (int t = 0; t < 3; ++t) {
    p.table_aid_y[t] = spec->table_aid_y[t];
    p.table_len[t] = spec->table_len[t];
    p.table_k[t] = spec->table_aid_y[t] ? spec->table_k[t] : 1;
    p.hist[t] = spec->hist[t];
    p.bonus[t] = spec->bonus[t];
    p.coeff[t] = spec->type_coefficient[t];
  }
  p.w_click = spec->w_click;
  p.w_cart = spec->w_cart;
  p.w_offset = spec->w_offset;
  p.n = spec->n;
  p.pred = pred;
  p.score = score;
  p.out_len = len;
  p.rows_by_list = rows_by_list ? 1 : 0;
  p.rows = rows_by_list ? (int64_t)n_list : sessions->n_sessions;
  int64_t lcap, hs;
  recency_caps(max_session_len, max_k, &lcap, &hs);
  p.lcap = (int32_t)lcap;
  p.hs = (int32_t)hs;
  p.max_k = max_k;
  p.slab = (uint64_t*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  p.slab_words = recency_slab_words(lcap, hs);
  const int blocks = n_list < RECENCY_BLOCKS ? n_list : RECENCY_BLOCKS;
  recency_long_kernel<<<blocks, REC_THREADS, 0, (cudaStream_t)stream>>>(p);
  LAUNCH_CHECK();
  return OTTO_OK;
}

extern "C" int otto_recency_long(const OttoSessions* sessions, const int32_t* session_list, int32_t n_list,
                                 int32_t max_session_len, const OttoRecencySpec* spec, void* scratch, int64_t scratch_bytes,
                                 int32_t* pred, void* stream) {
  return otto_recency_scored(sessions, session_list, n_list, max_session_len, spec, scratch, scratch_bytes, 0, pred, nullptr,
                             nullptr, stream);
}

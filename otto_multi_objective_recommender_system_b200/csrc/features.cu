// Interaction features over a candidate frame: the immediate consumer of the hot path's output.
//
// Replaces src/ranker/interaction_feature_engineering.py:47-113 (a polars script): per candidate row the occurrences
// of the candidate in its session (all / per event type, position of the last one), then mean / std / min / max of
// candidate_scores and mean / sum / max of the two counts per session and per candidate aid, joined back onto every
// row.  Three HBM-bound passes:
//   rows      one thread per row: binary search of the row's session in the event CSR, one walk over the session's
//             events (:50-60 as counters), row features out; the row's contributions are reduced over the runs of
//             equal session inside a warp (a frame sorted by session has ~100 rows per session) and added to the
//             per-session accumulators with one atomic set per run; per-aid accumulators take one atomic set per row
//   finalize  accumulators -> the ten / nine float / integer features of every session and aid
//   gather    one thread per row copies the features of its session and its aid into the output columns
// Sums are exact integers (scores in fixed point with 8 fractional bits, squares in 128 bits), so the features do not
// depend on the order of the atomics; mean and std (ddof 1) are formed once per key in fp64 and cast to fp32.
#include <string.h>

#include "common.cuh"

struct KeyAcc {                     // 64 bytes
  unsigned long long sum;           // sum of round(score * 256), two's complement
  unsigned long long sq_lo, sq_hi;  // sum of squares of the same, 128 bits
  unsigned long long occ_sum;
  unsigned long long last_sum;
  uint32_t n, last_n;
  uint32_t max_key, min_key;        // order-preserving integer image of the float score
  uint32_t occ_max, last_max;
};
static_assert(sizeof(KeyAcc) == 64, "accumulator layout");

struct KeyFeat {                    // finalized features of a session / an aid
  float score_mean, score_std, score_min, score_max, occ_mean, last_mean;
  uint32_t occ_sum, last_sum;
  uint16_t occ_max, last_max;
};

__device__ __forceinline__ uint32_t float_order_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct RowContribution {
  long long x;            // round(score * 256)
  unsigned long long sq;  // x * x (fits: |x| < 2^31 is checked by the caller's score range; larger scores saturate)
  uint32_t occ, last, key;
};

__device__ __forceinline__ void acc_add(KeyAcc* a, unsigned long long sum, unsigned long long sq_lo, unsigned long long sq_hi,
                                        uint32_t n, unsigned long long occ_sum, uint32_t occ_max, unsigned long long last_sum,
                                        uint32_t last_max, uint32_t last_n, uint32_t max_key, uint32_t min_key) {
  atomicAdd(&a->sum, sum);
  const unsigned long long old = atomicAdd(&a->sq_lo, sq_lo);
  const unsigned long long carry = (old + sq_lo < old) ? 1ull : 0ull;
  if (sq_hi + carry) atomicAdd(&a->sq_hi, sq_hi + carry);
  atomicAdd(&a->n, n);
  atomicMax(&a->max_key, max_key);
  atomicMin(&a->min_key, min_key);
  if (occ_sum) {
    atomicAdd(&a->occ_sum, occ_sum);
    atomicMax(&a->occ_max, occ_max);
    atomicAdd(&a->last_sum, last_sum);
    atomicMax(&a->last_max, last_max);
    atomicAdd(&a->last_n, last_n);
  }
}

struct FeatParams {
  const int32_t* off;
  const int32_t* ev_aid;
  const uint8_t* ev_type;
  const int32_t* session_ids;
  int64_t n_sessions;
  int64_t n_rows;
  const int32_t* row_session;
  const uint64_t* row_cand;
  const float* row_score;
  int32_t n_aids;
  int32_t* row_si;        // [n_rows] index of the row's session in the CSR (-1: not there)
  KeyAcc* sess_acc;       // [n_sessions]
  KeyAcc* aid_acc;        // [n_aids]
  OttoInteractionFeatures out;
};

__global__ void init_acc_kernel(KeyAcc* acc, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  KeyAcc a;
  memset(&a, 0, sizeof(a));
  a.min_key = 0xffffffffu;
  acc[i] = a;
}

__global__ void __launch_bounds__(256) feature_rows_kernel(const FeatParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = lane_id();
  const bool has = i < p.n_rows;
  int32_t si = -1;
  uint32_t occ = 0, last = 0, tc[3] = {0, 0, 0};
  long long x = 0;
  uint32_t key = 0;
  int64_t cand = -1;
  if (has) {
    const int32_t sid = p.row_session[i];
    int64_t lo = 0, hi = p.n_sessions;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (p.session_ids[mid] < sid) lo = mid + 1;
      else hi = mid;
    }
    if (lo < p.n_sessions && p.session_ids[lo] == sid) si = (int32_t)lo;
    cand = (int64_t)p.row_cand[i];
    if (si >= 0) {
      const int32_t beg = p.off[si], end = p.off[si + 1];
      for (int32_t e = beg; e < end; ++e) {
        if ((int64_t)p.ev_aid[e] == cand) {
          ++occ;
          last = (uint32_t)(e - beg + 1);                 // :50-57: 1-based position, last occurrence wins
          const uint32_t t = p.ev_type[e];
          if (t < 3) ++tc[t];
        }
      }
    }
    const float sc = p.row_score[i];
    x = llrintf(sc * 256.0f);
    key = float_order_key(sc);
    p.row_si[i] = si;
    const auto sat16 = [](uint32_t v) { return (uint16_t)(v > 0xffffu ? 0xffffu : v); };
    if (p.out.occurrence_count) p.out.occurrence_count[i] = sat16(occ);
    if (p.out.cumcount_last) p.out.cumcount_last[i] = sat16(last);
    if (p.out.click_occurrence_count) p.out.click_occurrence_count[i] = sat16(tc[0]);
    if (p.out.cart_occurrence_count) p.out.cart_occurrence_count[i] = sat16(tc[1]);
    if (p.out.order_occurrence_count) p.out.order_occurrence_count[i] = sat16(tc[2]);
  }
  const unsigned long long ax = (unsigned long long)(x < 0 ? -x : x);
  const unsigned __int128 sq128 = (unsigned __int128)ax * ax;
  // per aid: one atomic set per row
  if (has && cand >= 0 && cand < p.n_aids)
    acc_add(&p.aid_acc[cand], (unsigned long long)x, (unsigned long long)sq128, (unsigned long long)(sq128 >> 64), 1u, occ, occ, last, last,
            occ ? 1u : 0u, key, key);
  // per session: reduce the run of equal session inside the warp first (rows of a session are adjacent in a sorted frame)
  const uint32_t peers = __match_any_sync(FULL_MASK, has ? si : (int32_t)(0x40000000u | lane));
  const bool lead = has && si >= 0 && (peers & lanemask_lt()) == 0;
  unsigned long long s_sum = 0, s_sqlo = 0, s_sqhi = 0, s_occ = 0, s_last = 0;
  uint32_t s_n = 0, s_lastn = 0, s_max = 0, s_min = 0xffffffffu, s_occmax = 0, s_lastmax = 0;
  for (uint32_t m = peers; __any_sync(FULL_MASK, m != 0); m &= m - 1) {
    const int src = m ? __ffs(m) - 1 : (int)lane;
    const unsigned long long vx = shfl_u64((unsigned long long)x, src);
    const unsigned long long vlo = shfl_u64((unsigned long long)sq128, src), vhi = shfl_u64((unsigned long long)(sq128 >> 64), src);
    const uint32_t vocc = __shfl_sync(FULL_MASK, occ, src), vlast = __shfl_sync(FULL_MASK, last, src), vkey = __shfl_sync(FULL_MASK, key, src);
    if (m) {
      s_sum += vx;
      const unsigned long long before = s_sqlo;
      s_sqlo += vlo;
      s_sqhi += vhi + (s_sqlo < before ? 1ull : 0ull);
      ++s_n;
      s_occ += vocc;
      s_occmax = max(s_occmax, vocc);
      s_last += vlast;
      s_lastmax = max(s_lastmax, vlast);
      s_lastn += vocc ? 1u : 0u;
      s_max = max(s_max, vkey);
      s_min = min(s_min, vkey);
    }
  }
  if (lead) acc_add(&p.sess_acc[si], s_sum, s_sqlo, s_sqhi, s_n, s_occ, s_occmax, s_last, s_lastmax, s_lastn, s_max, s_min);
}

__global__ void finalize_kernel(const KeyAcc* __restrict__ acc, int64_t n, KeyFeat* __restrict__ feat) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const KeyAcc a = acc[i];
  KeyFeat f;
  memset(&f, 0, sizeof(f));
  if (a.n) {
    const double nn = (double)a.n;
    const double sum = (double)(long long)a.sum / 256.0;
    const double sq = ((double)a.sq_hi * 18446744073709551616.0 + (double)a.sq_lo) / 65536.0;
    f.score_mean = (float)(sum / nn);
    // sample standard deviation (ddof 1); a single row has none (NaN, the script's null)
    f.score_std = a.n > 1 ? (float)sqrt(fmax(0.0, (sq - sum * sum / nn) / (nn - 1.0))) : __int_as_float(0x7fc00000);
    f.score_min = float_from_order_key(a.min_key);
    f.score_max = float_from_order_key(a.max_key);
    f.occ_mean = (float)((double)a.occ_sum / nn);
    f.occ_sum = (uint32_t)a.occ_sum;
    f.occ_max = (uint16_t)(a.occ_max > 0xffffu ? 0xffffu : a.occ_max);
    f.last_mean = a.last_n ? (float)((double)a.last_sum / (double)a.last_n) : __int_as_float(0x7fc00000);
    f.last_sum = (uint32_t)a.last_sum;
    f.last_max = (uint16_t)(a.last_max > 0xffffu ? 0xffffu : a.last_max);
  }
  feat[i] = f;
}

__global__ void __launch_bounds__(256)
    feature_gather_kernel(const FeatParams p, const KeyFeat* __restrict__ sess_feat, const KeyFeat* __restrict__ aid_feat) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_rows) return;
  const OttoInteractionFeatures& o = p.out;
  const int32_t si = p.row_si[i];
  KeyFeat f;
  memset(&f, 0, sizeof(f));
  if (si >= 0) f = sess_feat[si];
  if (o.session_score_mean) o.session_score_mean[i] = f.score_mean;
  if (o.session_score_std) o.session_score_std[i] = f.score_std;
  if (o.session_score_min) o.session_score_min[i] = f.score_min;
  if (o.session_score_max) o.session_score_max[i] = f.score_max;
  if (o.session_occurrence_count_mean) o.session_occurrence_count_mean[i] = f.occ_mean;
  if (o.session_occurrence_count_sum) o.session_occurrence_count_sum[i] = f.occ_sum;
  if (o.session_occurrence_count_max) o.session_occurrence_count_max[i] = f.occ_max;
  if (o.session_cumcount_last_mean) o.session_cumcount_last_mean[i] = f.last_mean;
  if (o.session_cumcount_last_sum) o.session_cumcount_last_sum[i] = f.last_sum;
  if (o.session_cumcount_last_max) o.session_cumcount_last_max[i] = f.last_max;
  const int64_t cand = (int64_t)p.row_cand[i];
  memset(&f, 0, sizeof(f));
  if (cand >= 0 && cand < p.n_aids) f = aid_feat[cand];
  if (o.aid_score_mean) o.aid_score_mean[i] = f.score_mean;
  if (o.aid_score_std) o.aid_score_std[i] = f.score_std;
  if (o.aid_score_max) o.aid_score_max[i] = f.score_max;
  if (o.aid_occurrence_count_mean) o.aid_occurrence_count_mean[i] = f.occ_mean;
  if (o.aid_occurrence_count_sum) o.aid_occurrence_count_sum[i] = f.occ_sum;
  if (o.aid_occurrence_count_max) o.aid_occurrence_count_max[i] = f.occ_max;
  if (o.aid_cumcount_last_mean) o.aid_cumcount_last_mean[i] = f.last_mean;
  if (o.aid_cumcount_last_sum) o.aid_cumcount_last_sum[i] = f.last_sum;
  if (o.aid_cumcount_last_max) o.aid_cumcount_last_max[i] = f.last_max;
}

static int64_t feat_part(int64_t n, int64_t bytes) { return align_up((n > 0 ? n : 1) * bytes, 256); }

extern "C" int64_t otto_interaction_scratch_bytes(int64_t n_sessions, int64_t n_rows, int32_t n_aids) {
  return feat_part(n_rows, 4) + feat_part(n_sessions, sizeof(KeyAcc)) + feat_part(n_aids, sizeof(KeyAcc)) +
         feat_part(n_sessions, sizeof(KeyFeat)) + feat_part(n_aids, sizeof(KeyFeat));
}

extern "C" int otto_interaction_features(const OttoSessions* sessions, const int32_t* session_ids, const OttoCandidateFrame* frame,
                                         int32_t n_aids, const OttoInteractionFeatures* out, void* scratch, int64_t scratch_bytes,
                                         void* stream) {
  if (!sessions || !session_ids || !frame || !out || n_aids <= 0) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  if (frame->n_rows < 0 || (frame->n_rows > 0 && (!frame->session || !frame->candidates || !frame->candidate_scores))) {
    otto_set_error("candidate frame columns missing");
    return OTTO_EINVAL;
  }
  const int64_t S = sessions->n_sessions, R = frame->n_rows;
  if (!scratch || scratch_bytes < otto_interaction_scratch_bytes(S, R, n_aids)) { otto_set_error("interaction scratch too small"); return OTTO_ENOSPC; }
  if (R == 0) return OTTO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  char* sc = (char*)scratch;
  FeatParams p;
  memset(&p, 0, sizeof(p));
  p.off = sessions->session_offsets;
  p.ev_aid = sessions->aid;
  p.ev_type = sessions->type;
  p.session_ids = session_ids;
  p.n_sessions = S;
  p.n_rows = R;
  p.row_session = frame->session;
  p.row_cand = frame->candidates;
  p.row_score = frame->candidate_scores;
  p.n_aids = n_aids;
  p.row_si = (int32_t*)sc;
  sc += feat_part(R, 4);
  p.sess_acc = (KeyAcc*)sc;
  sc += feat_part(S, sizeof(KeyAcc));
  p.aid_acc = (KeyAcc*)sc;
  sc += feat_part(n_aids, sizeof(KeyAcc));
  KeyFeat* sess_feat = (KeyFeat*)sc;
  sc += feat_part(S, sizeof(KeyFeat));
  KeyFeat* aid_feat = (KeyFeat*)sc;
  p.out = *out;
  if (S > 0) {
    init_acc_kernel<<<(unsigned)ceil_div(S, 256), 256, 0, st>>>(p.sess_acc, S);
    LAUNCH_CHECK();
  }
  init_acc_kernel<<<(unsigned)ceil_div(n_aids, 256), 256, 0, st>>>(p.aid_acc, n_aids);
  LAUNCH_CHECK();
  feature_rows_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, st>>>(p);
  LAUNCH_CHECK();
  if (S > 0) {
    finalize_kernel<<<(unsigned)ceil_div(S, 256), 256, 0, st>>>(p.sess_acc, S, sess_feat);
    LAUNCH_CHECK();
  }
  finalize_kernel<<<(unsigned)ceil_div(n_aids, 256), 256, 0, st>>>(p.aid_acc, n_aids, aid_feat);
  LAUNCH_CHECK();
  feature_gather_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, st>>>(p, sess_feat, aid_feat);
  LAUNCH_CHECK();
  return OTTO_OK;
}

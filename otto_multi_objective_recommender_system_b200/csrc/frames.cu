// Candidate frames on the device: the flat columns the ranker scripts pickle, their labels and recall@k.
//
// Replaces, per event type, the host tail of src/ranker/covisitation_candidate_generation.py:
//   :151-153  candidate_labels = [int(aid in labels) for aid in sorted_aids]           -> label marking
//   :159-165  hits = len(set(candidates) & set(labels)); recall = sum(hits) / sum(min(len(labels), 20))
//   :177-197  df.explode([candidates, candidate_scores, candidate_labels]) + dtype casts -> flat columns
// and the same steps of src/ranker/regular_candidate_generation.py:160-180,225-257 (history aids prepended with
// scores |H| .. 1) and src/covisitation/inference.py:251-257 (recall@20 of the assembled predictions).
// Round 1 did these with per-row Python loops over ~3 * 100 * 1.67 M rows (missing #2 / #6).
// Labels come as a CSR over the frame's sessions with the aids of a session sorted ascending (host: one lexsort).
#include <string.h>

#include "common.cuh"
#include "scan.cuh"

__global__ void i32_to_u64_kernel(const int32_t* __restrict__ len, int64_t n, unsigned long long* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (unsigned long long)(len[i] > 0 ? len[i] : 0);
}

extern "C" int64_t otto_row_offsets_scratch_bytes(int64_t n) { return scan_scratch_elems(n) * 8 + 256; }

extern "C" int otto_row_offsets(const int32_t* len, int64_t n, int64_t* offsets, int64_t* total_host, void* scratch,
                                int64_t scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n < 0 || !offsets) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  if (scratch_bytes < otto_row_offsets_scratch_bytes(n)) { otto_set_error("scan scratch too small"); return OTTO_ENOSPC; }
  if (n > 0) {
    i32_to_u64_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(len, n, (unsigned long long*)offsets);
    LAUNCH_CHECK();
  }
  int rc = exclusive_scan<unsigned long long, unsigned long long>((unsigned long long*)offsets, n, (unsigned long long*)offsets,
                                                                  (unsigned long long*)scratch, st);
  if (rc) return rc;
  if (total_host) {
    CUDA_TRY(cudaMemcpyAsync(total_host, offsets + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return OTTO_OK;
}

// aid in the (ascending) labels of session s
__device__ __forceinline__ bool in_labels(const OttoLabels& lab, int64_t s, int32_t a) {
  int64_t lo = lab.offsets[s], hi = lab.offsets[s + 1];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int32_t v = lab.aid[mid];
    if (v == a) return true;
    if (v < a) lo = mid + 1;
    else hi = mid;
  }
  return false;
}

struct ExplodeParams {
  const int32_t* aid;
  const int32_t* score;
  const int32_t* len;
  int64_t n_sessions;
  int32_t top_n;
  const int64_t* row_offsets;
  const int32_t* session_ids;
  OttoLabels labels;
  int has_labels;
  int32_t* session_out;
  uint64_t* candidates_out;
  float* scores_out;
  uint8_t* labels_out;
};

// one thread per slot (s, r) of the fixed-stride lists: consecutive threads read consecutive slots and, inside a
// session, write consecutive rows
__global__ void __launch_bounds__(256) explode_kernel(const ExplodeParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_sessions * p.top_n) return;
  const int64_t s = i / p.top_n;
  const int r = (int)(i - s * p.top_n);
  if (r >= p.len[s]) return;
  const int64_t at = p.row_offsets[s] + r;
  const int32_t a = p.aid[i];
  p.session_out[at] = p.session_ids[s];
  p.candidates_out[at] = (uint64_t)(uint32_t)a;
  p.scores_out[at] = (float)p.score[i];
  if (p.has_labels) p.labels_out[at] = in_labels(p.labels, s, a) ? 1 : 0;
}

extern "C" int otto_explode_candidates(const int32_t* aid, const int32_t* score, const int32_t* len, int64_t n_sessions,
                                       int32_t top_n, const int64_t* row_offsets, const int32_t* session_ids,
                                       const OttoLabels* labels, int32_t* session_out, uint64_t* candidates_out,
                                       float* scores_out, uint8_t* labels_out, void* stream) {
  if (!aid || !score || !len || !row_offsets || !session_ids || !session_out || !candidates_out || !scores_out || top_n < 1 ||
      n_sessions < 0) {
    otto_set_error("bad argument");
    return OTTO_EINVAL;
  }
  if (labels && !labels_out) { otto_set_error("labels given but labels_out is NULL"); return OTTO_EINVAL; }
  if (n_sessions == 0) return OTTO_OK;
  ExplodeParams p;
  p.aid = aid; p.score = score; p.len = len;
  p.n_sessions = n_sessions; p.top_n = top_n;
  p.row_offsets = row_offsets; p.session_ids = session_ids;
  p.has_labels = labels != nullptr;
  if (labels) p.labels = *labels;
  else { p.labels.offsets = nullptr; p.labels.aid = nullptr; }
  p.session_out = session_out; p.candidates_out = candidates_out; p.scores_out = scores_out; p.labels_out = labels_out;
  explode_kernel<<<(unsigned)ceil_div(n_sessions * top_n, 256), 256, 0, (cudaStream_t)stream>>>(p);
  LAUNCH_CHECK();
  return OTTO_OK;
}

// ---- recall@k: hits = |set(pred) & set(labels)| summed over sessions, denominator = sum of min(|labels|, k) ----
// One warp per session: lane l takes label l, l + 32, ... (labels are unique) and scans the <= n predictions.
__global__ void __launch_bounds__(256)
    recall_kernel(const int32_t* __restrict__ pred, int64_t S, int n, const OttoLabels lab, int k_clip, unsigned long long* out) {
  const uint32_t lane = lane_id();
  const int64_t n_warps = (int64_t)gridDim.x * 8;
  unsigned long long hits = 0, denom = 0;
  for (int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); s < S; s += n_warps) {
    const int64_t lo = lab.offsets[s], hi = lab.offsets[s + 1];
    if (hi == lo) continue;
    if (lane == 0) denom += (unsigned long long)((hi - lo) < k_clip ? (hi - lo) : k_clip);
    const int32_t pl = (int)lane < n ? pred[s * n + lane] : -1;     // n <= 32: one prediction per lane
    for (int64_t j0 = lo; j0 < hi; j0 += 32) {                      // warp-uniform trip count
      const int64_t j = j0 + lane;
      bool has = j < hi;
      const int32_t a = has ? lab.aid[j] : -1;
      if (has && j > lo && lab.aid[j - 1] == a) has = false;        // defensive: duplicate label
      bool hit = false;
      if (n <= 32) {
        for (int r = 0; r < n; ++r) hit |= __shfl_sync(FULL_MASK, pl, r) == a;
      } else {
        for (int r = 0; r < n; ++r) hit |= pred[s * n + r] == a;
      }
      hits += (has && a >= 0 && hit) ? 1 : 0;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    hits += shfl_u64(hits, lane ^ o);
    denom += shfl_u64(denom, lane ^ o);
  }
  if (lane == 0 && (hits || denom)) {
    atomicAdd(&out[0], hits);
    atomicAdd(&out[1], denom);
  }
}

extern "C" int otto_recall_counts(const int32_t* pred, int64_t n_sessions, int32_t n, const OttoLabels* labels, int32_t k_clip,
                                  uint64_t* out_dev, void* stream) {
  if (!pred || !labels || !out_dev || n < 1 || k_clip < 1) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(out_dev, 0, 16, st));
  if (n_sessions == 0) return OTTO_OK;
  int64_t blocks = ceil_div(n_sessions, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  recall_kernel<<<(unsigned)blocks, 256, 0, st>>>(pred, n_sessions, n, *labels, k_clip, (unsigned long long*)out_dev);
  LAUNCH_CHECK();
  return OTTO_OK;
}

// ---- regular candidate form (ranker/regular_candidate_generation.py:139-180): history rows then vote rows ----
// Per session: its unique aids, most recent first, with scores |H| .. 1, followed by the ranker-form votes (aid, count).
// One warp per session; the unique aids are found chunk by chunk (a chunk's events against all earlier ones), so
// a 458-event session costs ~3 k compares per lane.  COUNT: rows per session; else write at row_offsets.
struct RegularParams {
  const int32_t* off;
  const int32_t* ev_aid;
  int64_t n_sessions;
  const int32_t* cand_aid;     // [S][N] one target
  const int32_t* cand_score;
  const int32_t* cand_len;
  int32_t top_n;
  int32_t* rows_out;           // COUNT: [S]
  const int64_t* row_offsets;
  const int32_t* session_ids;
  OttoLabels labels;
  int has_labels;
  int32_t* session_out;
  uint64_t* candidates_out;
  float* scores_out;
  uint8_t* labels_out;
};

template <bool COUNT>
__global__ void __launch_bounds__(256) regular_rows_kernel(const RegularParams p) {
  const uint32_t lane = lane_id(), lt = lanemask_lt();
  const int64_t n_warps = (int64_t)gridDim.x * 8;
  for (int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); s < p.n_sessions; s += n_warps) {
    const int32_t beg = p.off[s], end = p.off[s + 1];
    const int L = end - beg;
    // pass 1: number of unique aids
    int U = 0;
    for (int i0 = 0; i0 < L; i0 += 32) {
      const int i = i0 + (int)lane;                     // recency index: event end - 1 - i
      bool first = i < L;
      const int32_t a = first ? p.ev_aid[end - 1 - i] : -1;
      for (int j = 0; j < i && first; ++j) first = p.ev_aid[end - 1 - j] != a;
      U += __popc(__ballot_sync(FULL_MASK, first));
    }
    const int cl = p.cand_len[s];
    if (COUNT) {
      if (lane == 0) p.rows_out[s] = U + cl;
      continue;
    }
    const int64_t at0 = p.row_offsets[s];
    const int32_t sid = p.session_ids[s];
    int u = 0;
    for (int i0 = 0; i0 < L; i0 += 32) {
      const int i = i0 + (int)lane;
      bool first = i < L;
      const int32_t a = first ? p.ev_aid[end - 1 - i] : -1;
      for (int j = 0; j < i && first; ++j) first = p.ev_aid[end - 1 - j] != a;
      const uint32_t m = __ballot_sync(FULL_MASK, first);
      if (first) {
        const int r = u + __popc(m & lt);
        p.session_out[at0 + r] = sid;
        p.candidates_out[at0 + r] = (uint64_t)(uint32_t)a;
        p.scores_out[at0 + r] = (float)(U - r);           // :163  scores len(H) .. 1
        if (p.has_labels) p.labels_out[at0 + r] = in_labels(p.labels, s, a) ? 1 : 0;
      }
      u += __popc(m);
    }
    for (int r = (int)lane; r < cl; r += 32) {
      const int32_t a = p.cand_aid[s * p.top_n + r];
      p.session_out[at0 + U + r] = sid;
      p.candidates_out[at0 + U + r] = (uint64_t)(uint32_t)a;
      p.scores_out[at0 + U + r] = (float)p.cand_score[s * p.top_n + r];
      if (p.has_labels) p.labels_out[at0 + U + r] = in_labels(p.labels, s, a) ? 1 : 0;
    }
  }
}

static RegularParams regular_params(const OttoSessions* sessions, const int32_t* aid, const int32_t* score, const int32_t* len,
                                    int32_t top_n) {
  RegularParams p;
  memset(&p, 0, sizeof(p));
  p.off = sessions->session_offsets;
  p.ev_aid = sessions->aid;
  p.n_sessions = sessions->n_sessions;
  p.cand_aid = aid;
  p.cand_score = score;
  p.cand_len = len;
  p.top_n = top_n;
  return p;
}

extern "C" int otto_regular_row_counts(const OttoSessions* sessions, const int32_t* aid, const int32_t* score, const int32_t* len,
                                       int32_t top_n, int32_t* rows_out, void* stream) {
  if (!sessions || !len || !rows_out || top_n < 1) { otto_set_error("bad argument"); return OTTO_EINVAL; }
  if (sessions->n_sessions == 0) return OTTO_OK;
  RegularParams p = regular_params(sessions, aid, score, len, top_n);
  p.rows_out = rows_out;
  int64_t blocks = ceil_div(sessions->n_sessions, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  regular_rows_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  LAUNCH_CHECK();
  return OTTO_OK;
}

extern "C" int otto_regular_rows(const OttoSessions* sessions, const int32_t* aid, const int32_t* score, const int32_t* len,
                                 int32_t top_n, const int64_t* row_offsets, const int32_t* session_ids, const OttoLabels* labels,
                                 int32_t* session_out, uint64_t* candidates_out, float* scores_out, uint8_t* labels_out,
                                 void* stream) {
  if (!sessions || !aid || !score || !len || !row_offsets || !session_ids || !session_out || !candidates_out || !scores_out ||
      top_n < 1) {
    otto_set_error("bad argument");
    return OTTO_EINVAL;
  }
  if (labels && !labels_out) { otto_set_error("labels given but labels_out is NULL"); return OTTO_EINVAL; }
  if (sessions->n_sessions == 0) return OTTO_OK;
  RegularParams p = regular_params(sessions, aid, score, len, top_n);
  p.row_offsets = row_offsets;
  p.session_ids = session_ids;
  p.has_labels = labels != nullptr;
  if (labels) p.labels = *labels;
  p.session_out = session_out;
  p.candidates_out = candidates_out;
  p.scores_out = scores_out;
  p.labels_out = labels_out;
  int64_t blocks = ceil_div(sessions->n_sessions, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  regular_rows_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  LAUNCH_CHECK();
  return OTTO_OK;
}

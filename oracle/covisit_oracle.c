/* CPU oracle for the covisitation-matrix BUILD half of the hot path, in plain C.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product package never does.  Build: gcc -O2 -fPIC -shared -o oracle/_ref/libcovisit_oracle.so oracle/covisit_oracle.c
 * (__graft_entry__.build() and oracle/covisit_oracle_c.py do that).
 *
 * PARITY UNPINNED: /root/reference contains no covisitation builder - every script under
 * src/covisitation and src/ranker only reads pre-built top_15_<stem>_<part>.pqt files
 * (src/covisitation/inference.py:87-111, src/ranker/covisitation_candidate_generation.py:49-73); the matrices were
 * made out of tree with cuDF 22.10 (requirements.txt:23).  This file restates the north_star recipe (SURVEY.md
 * Appendix A, steps 1-7) a third time - beside the pandas oracle (oracle/covisit_oracle.py) and the plain-Python
 * witness (tests/test_oracle_bruteforce.py) - with different machinery (per-session nested loops over the tail, a
 * small open-addressing set for the in-session dedupe, one radix sort of the pair records for the group-by), and it
 * returns the EXACT integer accumulators of every distinct pair: cnt, tsum = sum(ts_x - ts_min), wsum = sum of the
 * integer type weights.  The column idiom it follows is the reference's src/matrix_factorization/torch_trainer.py:198-223
 * (`df.merge(df, on='session')`, `aid_x != aid_y`, `groupby(['aid_x', 'aid_y'])`).
 *
 * Appendix A numbering:
 *   1 type pre-filter   2 order (session asc, ts desc, ties in row order)   3 the tail_n most recent per session
 *   4 all ordered (i, j) of the tail, i-major then j; keep |ts_x - ts_y| < W (strict), aid_x != aid_y, type masks
 *   5 the first row of every (session, aid_x, aid_y) wins   6 weight from the winner row   7 group-by sum
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int64_t n_events;
  const int32_t* session;
  const int32_t* aid;
  const int32_t* ts;
  const uint8_t* type;
} OracleFrame;

typedef struct {
  uint32_t event_type_mask; /* bit t set: events of type t survive step 1 */
  uint32_t x_type_mask;     /* pair-level filters of step 4 */
  uint32_t y_type_mask;
  int32_t window_s;
  int32_t tail_n;
  int32_t ts_min;
  int32_t type_weight[3];   /* integer weights of type_y (wsum) */
  int32_t x_lo, x_hi;       /* only pairs with x_lo <= aid_x < x_hi are produced: a row of the matrix depends on nothing
                               else, so a large frame can be accumulated one aid_x range at a time */
} OracleRecipe;

typedef struct {
  uint64_t key; /* aid_x << 32 | aid_y */
  uint32_t tv;  /* ts_x - ts_min of the winner row */
  uint32_t ty;  /* type_y of the winner row */
} PairRec;

typedef struct {
  int64_t n;      /* distinct pairs */
  int64_t pairs;  /* pair rows after the in-session dedupe */
  int32_t* aid_x;
  int32_t* aid_y;
  int64_t* cnt;
  int64_t* tsum;
  int64_t* wsum;
} OracleResult;

/* step 2 orders ROW NUMBERS; the comparator reads the columns through these (one oracle call at a time per process) */
static const int32_t* g_session;
static const int32_t* g_ts;

static int order_cmp(const void* a, const void* b) {
  const int32_t p = *(const int32_t*)a, q = *(const int32_t*)b;
  if (g_session[p] != g_session[q]) return g_session[p] < g_session[q] ? -1 : 1;
  if (g_ts[p] != g_ts[q]) return g_ts[p] > g_ts[q] ? -1 : 1;      /* ts descending */
  return p < q ? -1 : (p > q ? 1 : 0);                            /* stable: ties keep the row order */
}

/* LSD radix sort of the pair records by key, 16 bits per pass, only as many passes as the largest key needs */
static int radix_sort(PairRec* a, int64_t n, uint64_t max_key) {
  if (n < 2) return 0;
  PairRec* b = (PairRec*)malloc((size_t)n * sizeof(PairRec));
  int64_t* count = (int64_t*)malloc(65537 * sizeof(int64_t));
  if (!b || !count) { free(b); free(count); return -1; }
  PairRec* src = a;
  PairRec* dst = b;
  for (int shift = 0; shift < 64 && (max_key >> shift) != 0; shift += 16) {
    memset(count, 0, 65537 * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) ++count[((src[i].key >> shift) & 0xffffu) + 1];
    for (int d = 0; d < 65536; ++d) count[d + 1] += count[d];
    for (int64_t i = 0; i < n; ++i) dst[count[(src[i].key >> shift) & 0xffffu]++] = src[i];
    PairRec* t = src; src = dst; dst = t;
  }
  if (src != a) memcpy(a, src, (size_t)n * sizeof(PairRec));
  free(b);
  free(count);
  return 0;
}

void covisit_oracle_free(OracleResult* r) {
  if (!r) return;
  free(r->aid_x); free(r->aid_y); free(r->cnt); free(r->tsum); free(r->wsum);
  free(r);
}

/* steps 1-7; NULL when memory runs out or an argument is unusable */
OracleResult* covisit_oracle_accumulate(const OracleFrame* f, const OracleRecipe* rc) {
  if (!f || !rc || f->n_events < 0 || rc->tail_n < 1 || rc->tail_n > 512 || rc->x_hi < rc->x_lo) return NULL;
  const int64_t n = f->n_events;
  if (n >= (1ll << 31)) return NULL;
  int32_t* order = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
  if (!order) return NULL;
  int64_t m = 0;
  int grouped = 1;                                                   /* rows of a session adjacent, sessions ascending */
  for (int64_t i = 0; i < n; ++i) {                                  /* step 1 */
    if (f->type[i] > 2 || !((rc->event_type_mask >> f->type[i]) & 1u)) continue;
    if (m > 0 && f->session[i] < f->session[order[m - 1]]) grouped = 0;
    order[m++] = (int32_t)i;
  }
  g_session = f->session;
  g_ts = f->ts;
  if (!grouped) {
    qsort(order, (size_t)m, sizeof(int32_t), order_cmp);            /* step 2 */
  } else {                                                           /* step 2, session by session (frames in file order) */
    for (int64_t s0 = 0; s0 < m;) {
      int64_t s1 = s0;
      while (s1 < m && f->session[order[s1]] == f->session[order[s0]]) ++s1;
      qsort(order + s0, (size_t)(s1 - s0), sizeof(int32_t), order_cmp);
      s0 = s1;
    }
  }

  int64_t cap = 1 << 16, np = 0;
  PairRec* pairs = (PairRec*)malloc((size_t)cap * sizeof(PairRec));
  /* in-session set of (aid_x, aid_y): at most tail_n * (tail_n - 1) members, table at least twice that */
  int64_t set_size = 64;
  while (set_size < 4 * (int64_t)rc->tail_n * rc->tail_n) set_size <<= 1;
  uint64_t* set = (uint64_t*)malloc((size_t)set_size * sizeof(uint64_t));
  int64_t* used = (int64_t*)malloc((size_t)rc->tail_n * rc->tail_n * sizeof(int64_t));
  if (!pairs || !set || !used) { free(order); free(pairs); free(set); free(used); return NULL; }
  memset(set, 0xff, (size_t)set_size * sizeof(uint64_t));
  uint64_t max_key = 0;
  int fail = 0;

  for (int64_t s0 = 0; s0 < m && !fail;) {
    int64_t s1 = s0;
    while (s1 < m && f->session[order[s1]] == f->session[order[s0]]) ++s1;
    const int64_t t = s1 - s0 < rc->tail_n ? s1 - s0 : rc->tail_n;   /* step 3 */
    int64_t n_used = 0;
    for (int64_t i = 0; i < t; ++i) {                                /* step 4: i-major ... */
      const int64_t ri = order[s0 + i];
      const int32_t ax = f->aid[ri], tx = f->ts[ri];
      if (!((rc->x_type_mask >> f->type[ri]) & 1u) || ax < rc->x_lo || ax >= rc->x_hi) continue;
      for (int64_t j = 0; j < t; ++j) {                              /* ... then j */
        const int64_t rj = order[s0 + j];
        const int32_t ay = f->aid[rj];
        int64_t dt = (int64_t)tx - (int64_t)f->ts[rj];
        if (dt < 0) dt = -dt;
        if (dt >= rc->window_s || ax == ay) continue;
        if (!((rc->y_type_mask >> f->type[rj]) & 1u)) continue;
        const uint64_t key = ((uint64_t)(uint32_t)ax << 32) | (uint32_t)ay;
        uint64_t h = (key * 0x9E3779B97F4A7C15ull) >> 20;            /* step 5: first row wins */
        int64_t slot = (int64_t)(h & (uint64_t)(set_size - 1));
        while (set[slot] != ~0ull && set[slot] != key) slot = (slot + 1) & (set_size - 1);
        if (set[slot] == key) continue;
        set[slot] = key;
        used[n_used++] = slot;
        if (np == cap) {
          cap *= 2;
          PairRec* grown = (PairRec*)realloc(pairs, (size_t)cap * sizeof(PairRec));
          if (!grown) { fail = 1; break; }
          pairs = grown;
        }
        pairs[np].key = key;                                         /* step 6: what the weight is formed from */
        pairs[np].tv = (uint32_t)((int64_t)tx - rc->ts_min);
        pairs[np].ty = f->type[rj];
        if (key > max_key) max_key = key;
        ++np;
      }
      if (fail) break;
    }
    for (int64_t u = 0; u < n_used; ++u) set[used[u]] = ~0ull;
    s0 = s1;
  }
  free(order); free(set); free(used);
  if (fail || radix_sort(pairs, np, max_key)) { free(pairs); return NULL; }

  int64_t d = 0;                                                     /* step 7 */
  for (int64_t i = 0; i < np; ++i) d += (i == 0 || pairs[i].key != pairs[i - 1].key);
  OracleResult* r = (OracleResult*)calloc(1, sizeof(OracleResult));
  if (!r) { free(pairs); return NULL; }
  const size_t dn = (size_t)(d > 0 ? d : 1);
  r->n = d;
  r->pairs = np;
  r->aid_x = (int32_t*)malloc(dn * 4); r->aid_y = (int32_t*)malloc(dn * 4);
  r->cnt = (int64_t*)malloc(dn * 8); r->tsum = (int64_t*)malloc(dn * 8); r->wsum = (int64_t*)malloc(dn * 8);
  if (!r->aid_x || !r->aid_y || !r->cnt || !r->tsum || !r->wsum) { free(pairs); covisit_oracle_free(r); return NULL; }
  int64_t o = -1;
  for (int64_t i = 0; i < np; ++i) {
    if (i == 0 || pairs[i].key != pairs[i - 1].key) {
      ++o;
      r->aid_x[o] = (int32_t)(pairs[i].key >> 32);
      r->aid_y[o] = (int32_t)(pairs[i].key & 0xffffffffu);
      r->cnt[o] = 0; r->tsum[o] = 0; r->wsum[o] = 0;
    }
    r->cnt[o] += 1;
    r->tsum[o] += pairs[i].tv;
    r->wsum[o] += rc->type_weight[pairs[i].ty];
  }
  free(pairs);
  return r;
}

int64_t covisit_oracle_count(const OracleResult* r) { return r ? r->n : -1; }
int64_t covisit_oracle_pairs(const OracleResult* r) { return r ? r->pairs : -1; }

/* copies the rows (sorted by aid_x, then aid_y) into caller arrays of covisit_oracle_count() elements */
void covisit_oracle_fetch(const OracleResult* r, int32_t* aid_x, int32_t* aid_y, int64_t* cnt, int64_t* tsum, int64_t* wsum) {
  if (!r || r->n == 0) return;
  memcpy(aid_x, r->aid_x, (size_t)r->n * 4);
  memcpy(aid_y, r->aid_y, (size_t)r->n * 4);
  memcpy(cnt, r->cnt, (size_t)r->n * 8);
  memcpy(tsum, r->tsum, (size_t)r->n * 8);
  memcpy(wsum, r->wsum, (size_t)r->n * 8);
}

/* Steps 6-8 over the accumulators: one float32 weight per distinct pair, formed ONCE from the exact integers
 * (weight_mode 2, time: cnt + w_scale * tsum evaluated in double, w_scale = 3 / (ts_max - ts_min); 1, type: wsum;
 * 0, unit: cnt), then per aid_x the k best by (weight descending, aid_y ascending) - the stable sort of Appendix A step 8
 * on rows that arrive in aid_y order.  Writes at most k rows per aid_x into the caller's arrays (capacity
 * covisit_oracle_count()), rows of an aid_x best first; returns the number of rows written. */
int64_t covisit_oracle_topk(const OracleResult* r, int32_t weight_mode, int32_t k, double w_scale, int32_t* aid_x, int32_t* aid_y,
                            float* wgt, int64_t* cnt, int64_t* tsum) {
  if (!r || k < 1 || k > 64) return -1;
  int64_t out = 0;
  for (int64_t g0 = 0; g0 < r->n;) {
    int64_t g1 = g0;
    while (g1 < r->n && r->aid_x[g1] == r->aid_x[g0]) ++g1;
    int64_t best[64];
    float bw[64];
    int kept = 0;
    for (int64_t i = g0; i < g1; ++i) {
      float w;
      if (weight_mode == 2) {
        volatile double prod = w_scale * (double)r->tsum[i];      /* no fused multiply-add: numpy rounds the product first */
        w = (float)((double)r->cnt[i] + prod);
      } else if (weight_mode == 1) {
        w = (float)(double)r->wsum[i];
      } else {
        w = (float)(double)r->cnt[i];
      }
      /* insertion into the k best; an equal weight stays behind the earlier (smaller aid_y) entry */
      int pos = kept;
      while (pos > 0 && bw[pos - 1] < w) --pos;
      if (pos >= k) continue;
      const int last = kept < k ? kept : k - 1;
      for (int q = last; q > pos; --q) { best[q] = best[q - 1]; bw[q] = bw[q - 1]; }
      best[pos] = i;
      bw[pos] = w;
      if (kept < k) ++kept;
    }
    for (int q = 0; q < kept; ++q, ++out) {
      aid_x[out] = r->aid_x[best[q]];
      aid_y[out] = r->aid_y[best[q]];
      wgt[out] = bw[q];
      cnt[out] = r->cnt[best[q]];
      tsum[out] = r->tsum[best[q]];
    }
    g0 = g1;
  }
  return out;
}

/*
 * otto_covisit.h - C ABI of the B200-native covisitation hot path (libotto_covisit.so).
 *
 * The reference (gunesevitan/otto-multi-objective-recommender-system) has no FFI, plugin or operator
 * table: its boundary for this path is files + one helper + a CLI.  Each entry point
 * below names the reference code whose work it replaces; the Python mirror of the reference scripts
 * lives in otto_multi_objective_recommender_system_b200/ and binds these symbols with ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all buffers
 *     (torch.Tensor.data_ptr() in our host code); the library allocates nothing on the device
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*); functions that
 *     return a count through a *_host pointer synchronise that stream before returning
 *   - return value 0 = ok, negative = error (OTTO_E*), message via otto_last_error(); no exceptions
 *     cross the ABI; calls are thread-compatible (one stream per call, no global mutable state except
 *     the thread-local error string)
 *   - aids are int32 in [0, n_aids), n_aids <= 2^30; ts are int32 seconds; type is uint8 in {0,1,2}
 *     (dtypes of utilities/dataset_writer_pickle.py:57-60 after the /1000 the consumers apply)
 */
#ifndef OTTO_COVISIT_H_
#define OTTO_COVISIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OTTO_OK 0
#define OTTO_EINVAL (-22)      /* bad argument */
#define OTTO_ENOSPC (-28)      /* caller buffer / workspace too small */
#define OTTO_ECUDA (-5)        /* CUDA runtime error */
#define OTTO_EOVERFLOW (-75)   /* an in-SM hash table overflowed: lower split_ub and rebuild */
#define OTTO_EUNSORTED (-71)   /* frame is not sorted by (session, ts) */

#define OTTO_WEIGHT_UNIT 0     /* buy2buy: wgt = number of sessions holding the pair            */
#define OTTO_WEIGHT_TYPE 1     /* carts-orders: wgt = sum of type_weight[type_y]                 */
#define OTTO_WEIGHT_TIME 2     /* clicks: wgt = sum of 1 + 3 (ts_x - ts_min) / (ts_max - ts_min) */

#define OTTO_MAX_TAIL 32       /* tail_n <= 32: one event per lane */
#define OTTO_MAX_K 32
#define OTTO_MAX_SEGMENTS 8    /* pair segments per bin = GPUs of one box */
#define OTTO_MAX_OWNERS 8      /* ranks of one box that can own aid_x ranges (owner-direct scatter) */
#define OTTO_MAX_TABLES 8
#define OTTO_MAX_SOURCES 8
#define OTTO_MAX_TARGETS 4

/* Session-sorted CSR of events, most recent first inside a session:
 * order (session asc, ts desc, original row asc) - exactly the row order after the builder's
 * sort_values(['session','ts'], ascending=[True, False]).
 * Produced by otto_ingest_desc(). */
typedef struct {
  int64_t n_sessions;
  int64_t n_events;
  const int32_t* session_offsets; /* [n_sessions + 1] */
  const int32_t* aid;             /* [n_events] */
  const int32_t* ts;              /* [n_events] */
  const uint8_t* type;            /* [n_events] */
} OttoEvents;

/* The matrix recipe (generic form §8a "CovisitSpec"). */
typedef struct {
  int32_t n_aids;
  int32_t weight_mode;       /* OTTO_WEIGHT_* */
  int32_t type_weight[3];    /* OTTO_WEIGHT_TYPE: {1, 6, 3} (baseline/aid_weight.py:34) */
  uint32_t event_type_mask;  /* bit t set = events of type t enter the tail (buy2buy: 0b110) */
  uint32_t x_type_mask;      /* pair-level filters on type_x / type_y (0b111 for the graded variants) */
  uint32_t y_type_mask;
  int32_t window_s;          /* keep |ts_x - ts_y| <  window_s (strict) */
  int32_t tail_n;            /* most recent events per session that enter the self-join (30) */
  int32_t k;                 /* rows kept per aid_x (15 / 20) */
  int32_t ts_min;            /* OTTO_WEIGHT_TIME constants: 1659304800, 1662328791 */
  int32_t ts_max;
  int32_t split_ub;          /* rows with more pairs than this are split into aid_y-hash sub-bins of ~split_ub / 3; 0 = default (6144) */
  int64_t global_events;     /* multi-GPU: events of ALL ranks (bins are formed from the all-reduced bounds, so the
                                bin arrays must be sized for the global frame); 0 = this rank's frame is the whole frame */
} OttoCovisitSpec;

/* Sizes the caller needs to allocate the build workspace for a given input shape. */
typedef struct {
  int64_t tail_capacity;     /* events of the tail CSR: min(n_events, n_sessions * tail_n) */
  int64_t max_bins;          /* upper bound on bins = n_aids + sub-bins */
  int64_t workspace_bytes;   /* bytes of the fixed workspace (everything except the pair records) */
} OttoBuildSizes;

/* Counters a build publishes (host struct, filled by otto_covisit_count / otto_covisit_reduce). */
typedef struct {
  int64_t tail_events;       /* E30: events that entered the self-join */
  int64_t pairs;             /* P: pairs after in-session dedupe (records this rank emits) */
  int64_t bins;              /* B: aid_x rows + sub-bins */
  int64_t split_rows;        /* aid_x rows that were split into sub-bins */
  int64_t distinct;          /* D: distinct (aid_x, aid_y) accumulated by this rank (after reduce) */
  int64_t pair_checksum;     /* sum of counts over all accumulated entries (== pairs received) */
  int64_t table_overflow;    /* != 0: a hash table overflowed (result invalid) */
  int64_t tier_records[4];   /* records accumulated by the warp / 128- / 256- / 512-thread (and hash-table) reduce kernels */
  int64_t hot_pairs;         /* pairs of hot (split) rows: they pass through a staging area behind the P final records,
                                so the record buffer must hold pairs + hot_pairs records */
  /* otto_covisit_count_finish_owned only (identical on every rank: the layout of all owners is computed everywhere) */
  int64_t owner_records_max; /* largest record buffer any owner needs (final + staged records) */
  int64_t owner_bin_cuts[9]; /* [OTTO_MAX_OWNERS + 1] first bin of every owner's aid range; [n_owners] = bins */
} OttoBuildStats;

/* Pair records of one producer for a contiguous range of bins (multi-GPU: one segment per sender). */
typedef struct {
  const void* records;       /* 8-byte records {uint32 aid_y, uint32 v} */
  const uint64_t* offsets;   /* offsets[b - bin_lo] .. offsets[b - bin_lo + 1] bound bin b's records */
} OttoPairSegment;

/* Per-aid top-K table, fixed stride: row aid_x holds len[aid_x] <= k entries, best first
 * (wgt desc, aid_y asc); the file form top_<k>_<stem>_<part>.pqt (covisitation/inference.py:87-111)
 * is these rows compacted. */
typedef struct {
  int32_t n_aids;
  int32_t k;
  int32_t* aid_y;            /* [n_aids * k], -1 padded */
  float* wgt;                /* [n_aids * k] */
  int32_t* len;              /* [n_aids] */
  uint32_t* cnt;             /* optional [n_aids * k]: exact pair count (NULL to skip) */
  uint64_t* tsum;            /* optional [n_aids * k]: exact sum(ts_x - ts_min), time mode (NULL to skip) */
} OttoTopK;

const char* otto_last_error(void);
int otto_version(void);
/* Kernels this library has launched so far in this process (bench.py reports the per-step delta). */
uint64_t otto_launch_count(void);

/* Measurement aid for bench.py: with profiling on, otto_covisit_reduce brackets each of its kernels with CUDA
 * events on the caller's stream; otto_profile_reduce_ms synchronises the last one and returns the five
 * durations of the most recent call in milliseconds: classify + owner-table warp tier [bins <= 384 records],
 * 512-thread tier [<= 6144], 256-thread tier [<= 3072], 128-thread tier [<= 1536], hash-table kernel [hand-overs
 * and larger bins] + merge_split_rows.  Profiled calls run the tier kernels back to back on the caller's stream;
 * unprofiled calls run them concurrently on three internal side streams (forked from and joined to the caller's
 * stream). */
int otto_profile_enable(int on);
int otto_profile_reduce_ms(float* ms_host /* [5] */);
/* Same for otto_covisit_scatter: pairgen scatter kernel; bin counts + partition count + offset scan; partition move. */
int otto_profile_scatter_ms(float* ms_host /* [3] */);

/* ---- ingest: frame columns -> CSR (replaces the sort + chunk writers of
 *      utilities/split_dataset_writer_parquet.py:13-33 and builder step 2) ---- */

/* Checks that (session, ts) is non-decreasing; *sorted_host = 1/0. Synchronises. */
int otto_frame_is_sorted(const int32_t* session, const int32_t* ts, int64_t n_events, int32_t* flag_dev,
                         int32_t* sorted_host, void* stream);

/* Event contents the kernels index with: counts the rows whose aid is outside [0, n_aids) or whose type is above 2
 * (the dtypes of utilities/dataset_writer_pickle.py:57-60 allow both).  count_dev: 8 bytes of device scratch.
 * Returns OTTO_EINVAL (and the count in *n_bad_host) when any row offends.  Synchronises. */
int otto_frame_check(const int32_t* aid, const uint8_t* type, int64_t n_events, int32_t n_aids, void* count_dev,
                     int64_t* n_bad_host, void* stream);

/* Frame -> session CSR in two passes over the session column (no per-event temporaries).  Pass 1 counts the session
 * starts per tile and checks, in the same read, the (session, ts) order and the event contents: info_host [3] =
 * {n_sessions, sorted (1 / 0), events with an aid outside [0, n_aids) or a type above 2} (aid == NULL skips the content
 * check); synchronises.  Pass 2 (same scratch, untouched in between) writes the session ids [n_sessions], the offsets
 * int32 [n_sessions + 1] and the longest session.  Replaces groupby('session') of covisitation/inference.py:117 and
 * the chunk writer's sort check (utilities/split_dataset_writer_parquet.py:17). */
int64_t otto_ingest_scratch_bytes(int64_t n_events);
int otto_ingest_scan(const int32_t* session, const int32_t* aid, const int32_t* ts, const uint8_t* type, int64_t n_events,
                     int32_t n_aids, void* scratch, int64_t scratch_bytes, int64_t* info_host /* [3] */, void* stream);
int otto_ingest_offsets(const int32_t* session, int64_t n_events, int64_t n_sessions, void* scratch, int64_t scratch_bytes,
                        int32_t* session_ids, int32_t* offsets, int32_t* max_len_dev, void* stream);

/* From a frame sorted by (session, ts) ascending with run-length offsets (ascending CSR), writes the
 * most-recent-first CSR columns: inside each session ts descending, ties in original row order. */
int otto_ingest_desc(const int32_t* session_offsets, int64_t n_sessions, const int32_t* aid, const int32_t* ts,
                     const uint8_t* type, int64_t n_events, int32_t* aid_out, int32_t* ts_out, uint8_t* type_out,
                     void* stream);

/* ---- build (the missing builder) ----
 *
 * Phases (all on one stream, one rank):
 *   count_begin   tail CSR (steps 1-3), in-session dedupe (steps 4-5, row masks), pairs per aid_x row
 *   [multi-GPU: all-reduce the row totals so that every rank forms identical bins]
 *   count_finish  bins (a row whose total exceeds split_ub is split into aid_y-hash sub-bins), record offsets
 *   scatter       8-byte pair records {aid_y, v} (step 6 in integer form): ordinary rows at their final position,
 *                 hot rows into a staging area and from there, partitioned by aid_y hash, into their sub-bins;
 *                 bin_offsets are valid after this phase
 *   [multi-GPU: owners read the per-owner bin ranges of every rank (otto_covisit_merge_segments)]
 *   reduce        accumulate per (aid_x, aid_y) + top-k per aid_x (steps 7-8) -> OttoTopK
 * otto_covisit_build runs all of them for one GPU. */

int otto_covisit_sizes(int64_t n_sessions, int64_t n_events, const OttoCovisitSpec* spec, OttoBuildSizes* out_host);

int otto_covisit_count_begin(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                             int64_t workspace_bytes, void* stream);
/* The same for a CSR in FILE order (ts ascending inside a session, the order the reference's frames are written in:
 * utilities/split_dataset_writer_parquet.py:17): the tail kernels apply the builder's stable ts-descending sort
 * (Appendix A step 2) while they copy, so otto_ingest_desc is not needed.  Only the <= tail_n most recent events of
 * a session are read: ev->aid and ev->type may point to pinned HOST memory (cudaHostAlloc; valid on the device under
 * UVA), in which case 5 bytes per tail event cross PCIe instead of 5 bytes per event.  ev->session_offsets and ev->ts
 * must be device memory.  Every later phase takes the same ev. */
int otto_covisit_count_begin_asc(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                                 int64_t workspace_bytes, void* stream);
/* Fills stats_host->{tail_events,pairs,bins,split_rows,hot_pairs}; synchronises.  OTTO_EINVAL when a tail event carried
 * an aid outside [0, n_aids), a type above 2 or (time mode) a ts outside [ts_min, ts_max]: count_begin replaces such
 * events by a harmless one and flags them, so nothing is written out of bounds (in owner-direct mode the pair kernels
 * store into a peer's memory at positions derived from the aid). */
int otto_covisit_count_finish(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                              int64_t workspace_bytes, OttoBuildStats* stats_host, void* stream);
/* count_begin + count_finish */
int otto_covisit_count(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                       OttoBuildStats* stats_host, void* stream);

/* Device views into the workspace (valid after the phase that writes them):
 *   row_total   uint32 [n_aids]      pairs per aid_x row, after count_begin (all-reduce target for multi-GPU)
 *   bin_base    uint32 [n_aids + 1]  first bin of each aid_x row, after count_finish
 *   bin_x       uint32 [bins]        bin -> aid_x, after count_finish
 *   bin_offsets uint64 [bins + 1]    record offsets of this rank's bins, after scatter */
int otto_covisit_views(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                       uint64_t** bin_offsets, uint32_t** bin_base, uint32_t** bin_x, uint32_t** row_total);

/* Writes the pair records, grouped by bin, into `records`: capacity >= (stats.pairs + stats.hot_pairs) records of
 * 8 bytes; the first stats.pairs of them are the result, the rest is the staging area of the hot rows. */
int otto_covisit_scatter(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                         void* records, int64_t records_capacity, void* stream);

/* For bins [bin_lo, bin_hi) = rows [aid_lo, aid_hi): accumulate the records of ONE bin-contiguous segment per
 * (aid_x, aid_y), select the top k per aid_x and write those rows of `out` (n_segments must be 1: a multi-GPU owner
 * merges what it received with otto_covisit_merge_segments first; OTTO_EINVAL otherwise).  The segment's offsets are
 * indexed by (bin - bin_lo) and may carry any base (offsets[0] is subtracted).  Fills
 * stats_host->{distinct,pair_checksum,table_overflow}; synchronises when stats_host != NULL. */
int64_t otto_covisit_reduce_scratch_bytes(const OttoCovisitSpec* spec, int64_t n_bins, int64_t n_aids_range);
int otto_covisit_reduce(const OttoCovisitSpec* spec, const uint32_t* bin_base, const uint32_t* bin_x, int64_t bin_lo,
                        int64_t bin_hi, int32_t aid_lo, int32_t aid_hi, const OttoPairSegment* segments_host,
                        int32_t n_segments, void* scratch, int64_t scratch_bytes, const OttoTopK* out,
                        OttoBuildStats* stats_host, void* stream);

/* Multi-GPU owner side: gathers the G received segments of bins [0, n_bins) into ONE bin-contiguous record
 * array (merged_records, capacity >= sum of all segment lengths) with offsets merged_offsets [n_bins + 1],
 * so that the reduce kernels stream each bin as a single run (with G runs per bin every small bin costs G
 * partial warp steps).  scratch: otto_covisit_merge_scratch_bytes(n_bins).  *n_records_host (optional)
 * receives the total and synchronises. */
int64_t otto_covisit_merge_scratch_bytes(int64_t n_bins);
int otto_covisit_merge_segments(const OttoPairSegment* segments_host, int32_t n_segments, int64_t n_bins,
                                void* merged_records, int64_t merged_capacity, uint64_t* merged_offsets, void* scratch,
                                int64_t scratch_bytes, int64_t* n_records_host, void* stream);

/* Peer memory over NVLink (one process per GPU of one box).  The pair records of a rank live in a buffer
 * that every other rank maps (CUDA IPC); the owner's merge kernel then READS the senders' slabs straight
 * from their HBM over NVLink / NVSwitch - the all-to-all of SURVEY.md §8e is fused into the merge, no
 * staging copy and no NCCL bulk transfer.  Handles are 64 opaque bytes, exchanged by the host (all_gather). */
#define OTTO_PEER_HANDLE_BYTES 64
int otto_peer_alloc(int64_t bytes, void** ptr_host);
int otto_peer_free(void* ptr);
int otto_peer_get_handle(void* ptr, uint8_t* handle_host /* [64] */);
int otto_peer_open(const uint8_t* handle_host /* [64] */, void** ptr_host);
int otto_peer_close(void* ptr);

/* ---- owner-direct scatter (multi-GPU, the exchange of SURVEY.md §8e fused into the scatter kernel) ----
 *
 * Every rank writes each pair record straight into the record buffer of the rank that OWNS the aid_x row, through
 * the peer mapping of that buffer (NVLink stores, fire and forget), at its final position: after the scatter the
 * owner holds its rows exactly as a single-GPU build of the whole frame would (ordinary rows final, hot rows in its
 * staging area), so partition and reduce run locally on one segment and nothing is read remotely afterwards.
 * Host protocol (distributed.py):
 *   count_begin -> all-gather the per-row pair counts (otto_covisit_views: row_total holds this rank's counts) ->
 *   row_total := sum over ranks, row_before := sum over lower ranks, aid cuts balanced on the totals ->
 *   otto_covisit_count_finish_owned (bins; layout of MY rows; scatter cursors = position inside the owner's buffer;
 *   stats_host->pairs / hot_pairs = what MY buffer must hold) -> peers map each other's buffers ->
 *   otto_covisit_scatter_owned -> any collective (orders "all scatters done") -> otto_covisit_partition ->
 *   otto_covisit_reduce over bins [bin_base[aid_cuts[rank]], bin_base[aid_cuts[rank + 1]]) with ONE local segment. */
typedef struct {
  int32_t n_owners;                          /* G <= OTTO_MAX_OWNERS */
  int32_t rank;                              /* this process' owner index */
  int32_t aid_cuts[OTTO_MAX_OWNERS + 1];     /* owner o holds rows [aid_cuts[o], aid_cuts[o + 1]); [0] = 0, [G] = n_aids */
  void* owner_records[OTTO_MAX_OWNERS];      /* record buffer of every owner as mapped on THIS device (scatter only) */
} OttoOwnerPlan;

/* The plan on the device, from the all-gathered per-row counts counts_all [n_owners][n_aids] (uint32, rank-major):
 * row_total [n_aids] := sum over ranks (pass the workspace view of otto_covisit_views), row_before [n_aids] := sum over
 * the ranks below `rank`, aid_cuts_host [OTTO_MAX_OWNERS + 1] := contiguous aid ranges with (nearly) equal pair counts
 * (cut g = first row with at least total * g / n_owners pairs in the rows before it).  One synchronisation.
 * OTTO_EINVAL when a row holds 2^32 or more pairs. */
int64_t otto_covisit_plan_scratch_bytes(int32_t n_aids);
int otto_covisit_plan_owners(const uint32_t* counts_all, int32_t n_owners, int32_t rank, int32_t n_aids, uint32_t* row_total,
                             uint32_t* row_before, void* scratch, int64_t scratch_bytes, int32_t* aid_cuts_host, void* stream);

/* row_before: device uint32 [n_aids], pairs of each row held by ranks below this one.  Synchronises; returns
 * OTTO_EINVAL on every rank alike when any owner would exceed 2^32 - 1 records. */
int otto_covisit_count_finish_owned(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                                    int64_t workspace_bytes, const OttoOwnerPlan* plan, const uint32_t* row_before,
                                    OttoBuildStats* stats_host, void* stream);
int otto_covisit_scatter_owned(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace,
                               int64_t workspace_bytes, const OttoOwnerPlan* plan, void* stream);
/* Second half of otto_covisit_scatter: records per bin, hot rows from the staging area into their sub-bins, bin
 * offsets.  `records` is this rank's own buffer. */
int otto_covisit_partition(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                           void* records, int64_t records_capacity, void* stream);

/* Staged scatter (multi-GPU, sender-side combining): the alternative to otto_covisit_scatter_owned for links that
 * punish short stores.  The direct scatter crosses NVLink as ~63-byte runs (one run = the pairs of one event);
 * here a rank first appends its runs to COARSE buckets of its OWN staging buffer (bucket = 2^20 records of the
 * global row order, each record packed into 64 bits as aid_y | v | row - first row of the bucket), and the owner of
 * a row range then streams the segments of its buckets out of every rank's staging buffer (large contiguous reads
 * through the peer mapping) and places the records (final position of an ordinary row, staging area of a hot row).
 *   ... otto_covisit_count_finish_owned -> otto_covisit_stage_plan (one synchronisation; staged_records_host[g] =
 *   records rank g stages, identical on every rank; or, to save the synchronisation: stage_plan with a NULL
 *   staged_records_host BEFORE count_finish_owned and otto_covisit_stage_totals after it) -> peers map each other's
 *   staging buffers ->
 *   otto_covisit_scatter_staged -> any collective (orders "every rank has staged") -> otto_covisit_place_staged
 *   into MY record buffer -> otto_covisit_partition -> otto_covisit_reduce as before.
 * counts_all [n_ranks][n_aids]: the all-gathered per-row counts (NULL with one rank = the workspace's own).
 * OTTO_EOVERFLOW: the 64-bit staged record cannot carry aid_y, v and the row (use the direct scatter).
 * Like the bin arrays of the workspace, the plan scratch (its tile list) is sized from spec->global_events: with more
 * than one rank it must be the event count of ALL ranks (a smaller value under-sizes the list).
 * The reference has no distributed code; this replaces nothing of it. */
int64_t otto_covisit_stage_plan_bytes(const OttoCovisitSpec* spec, int64_t n_sessions, int64_t n_events, int32_t n_ranks);
int otto_covisit_stage_plan(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                            const uint32_t* counts_all, int32_t n_ranks, int32_t rank, void* plan, int64_t plan_bytes,
                            int64_t* staged_records_host, void* stream);
int otto_covisit_stage_totals(const OttoCovisitSpec* spec, int64_t n_sessions, int64_t n_events, void* plan, int32_t n_ranks,
                              int64_t* staged_records_host, void* stream);
int otto_covisit_scatter_staged(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                                void* plan, int32_t n_ranks, void* staged, void* stream);
int otto_covisit_place_staged(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                              void* plan, int32_t n_ranks, const void* const* staged_host, int32_t aid_lo, int32_t aid_hi,
                              void* records, int64_t records_capacity, void* stream);

/* One-shot single-GPU build; the pair records and the reduce scratch are carved from the workspace after
 * the fixed part.  Returns OTTO_ENOSPC (stats_host->pairs / bins set) when the workspace cannot hold
 * them; otto_covisit_build_bytes (pairs = stats.pairs + stats.hot_pairs) gives the size to retry with. */
int otto_covisit_build(const OttoEvents* ev, const OttoCovisitSpec* spec, void* workspace, int64_t workspace_bytes,
                       const OttoTopK* out, OttoBuildStats* stats_host, void* stream);
int64_t otto_covisit_build_bytes(int64_t n_sessions, int64_t n_events, const OttoCovisitSpec* spec, int64_t pairs,
                                 int64_t bins);

/* Compacts a fixed-stride table into file rows (aid_x, aid_y, wgt) sorted aid_x asc, best first.
 * row_offsets [n_aids + 1] = exclusive scan of len (int64). */
int otto_topk_row_offsets(const OttoTopK* table, int64_t* row_offsets, int64_t* n_rows_host, void* scratch,
                          int64_t scratch_bytes, void* stream);
int otto_topk_to_rows(const OttoTopK* table, const int64_t* row_offsets, int32_t* aid_x, int32_t* aid_y, float* wgt,
                      void* stream);
/* Inverse: file rows (grouped by aid_x, ranked) -> fixed-stride table; what covisitation_df_to_dict
 * (covisitation/inference.py:19-35) builds as a dict of lists. */
int otto_rows_to_topk(const int32_t* aid_x, const int32_t* aid_y, const float* wgt, int64_t n_rows,
                      const OttoTopK* table, void* stream);

/* ---- candidate generation (ranker/covisitation_candidate_generation.py:108-141,
 *      covisitation/inference.py:204-247) ---- */

/* Test / validation sessions in FILE order (ts ascending), as groupby('session').agg(list) sees them. */
typedef struct {
  int64_t n_sessions;
  int64_t n_events;
  const int32_t* session_offsets; /* [n_sessions + 1] */
  const int32_t* aid;
  const uint8_t* type;
} OttoSessions;

#define OTTO_HIST_RECENCY 0   /* unique aids, most recent first: list(dict.fromkeys(aids[::-1])) */
#define OTTO_HIST_TYPE_LE1 1  /* np.unique(aids[types <= 1]) ascending */
#define OTTO_HIST_TYPE_GE1 2  /* np.unique(aids[types >= 1]) ascending */
#define OTTO_HIST_TYPE_EQ0 3  /* np.unique(aids[types == 0]) ascending */

typedef struct {
  int32_t n_tables;
  const int32_t* table_aid_y[OTTO_MAX_TABLES]; /* fixed-stride tables [n_aids * table_k[i]] */
  const int32_t* table_len[OTTO_MAX_TABLES];   /* [n_aids] */
  int32_t table_k[OTTO_MAX_TABLES];
  int32_t n_aids;
  int32_t n_sources;                           /* a source = one table gathered over one history set */
  int32_t source_table[OTTO_MAX_SOURCES];
  int32_t source_hist[OTTO_MAX_SOURCES];       /* OTTO_HIST_* */
  int32_t n_targets;                           /* clicks, carts, orders */
  int32_t target_n_sources[OTTO_MAX_TARGETS];
  int32_t target_sources[OTTO_MAX_TARGETS][OTTO_MAX_SOURCES]; /* concatenation order */
  int32_t top_n;                               /* Counter.most_common(top_n): 100 (ranker) / 20 (standalone) */
  int32_t drop_history;                        /* 1: drop aids that are in the session (after truncation) */
} OttoCandidateSpec;

typedef struct {
  int32_t* aid;    /* [n_targets][n_sessions][top_n], -1 padded, count desc then first-seen asc */
  int32_t* score;  /* [n_targets][n_sessions][top_n] vote counts */
  int32_t* len;    /* [n_targets][n_sessions] */
} OttoCandidates;

/* max_session_len = longest session of the frame (events); positions in a concatenation are 16-bit, so
 * max_session_len * (largest sum of table_k over one target's sources) must stay below 65535. */
int64_t otto_candidates_scratch_bytes(int64_t n_sessions, int32_t max_session_len, const OttoCandidateSpec* spec);
int otto_candidates(const OttoSessions* sessions, int32_t max_session_len, const OttoCandidateSpec* spec, void* scratch,
                    int64_t scratch_bytes, const OttoCandidates* out, void* stream);

/* covisitation/inference.py:238-243: history + votes[:n - |H|] + popular[:n - len], -1 padded.  Sessions
 * with >= n unique aids keep their n most recent unique aids and are flagged in long_session (the
 * reference routes them to its recency branch, :128-131).  popular is [n_targets][n_popular]. */
int otto_assemble_predictions(const OttoSessions* sessions, const OttoCandidates* cand, int32_t n_targets, int32_t top_n,
                              const int32_t* popular, int32_t n_popular, int32_t n, int32_t* pred /* [n_targets][n_sessions][n] */,
                              uint8_t* long_session /* [n_sessions] or NULL */, void* stream);

/* ---- candidate frames on the device (ranker/covisitation_candidate_generation.py:151-165,177-197;
 *      ranker/regular_candidate_generation.py:160-180,225-257; covisitation/inference.py:251-257) ---- */

/* Ground truth of a frame's sessions as a CSR: the aids of session s are aid[offsets[s] .. offsets[s + 1]), unique and
 * sorted ascending (splits/val_labels.parquet rows of one event type, covisitation/inference.py:116-122). */
typedef struct {
  const int64_t* offsets;   /* [n_sessions + 1] */
  const int32_t* aid;
} OttoLabels;

/* Exclusive scan of int32 lengths (negative = 0) into int64 offsets [n + 1]; *total_host (optional) synchronises. */
int64_t otto_row_offsets_scratch_bytes(int64_t n);
int otto_row_offsets(const int32_t* len, int64_t n, int64_t* offsets, int64_t* total_host, void* scratch, int64_t scratch_bytes,
                     void* stream);

/* df.explode of one target's lists (aid / score [n_sessions][top_n], len [n_sessions]) into the flat columns the ranker
 * scripts pickle: session (repeated), candidates uint64, candidate_scores float32 and, when labels != NULL,
 * candidate_labels uint8 = int(aid in labels of the session).  Session s fills rows row_offsets[s] .. + len[s]. */
int otto_explode_candidates(const int32_t* aid, const int32_t* score, const int32_t* len, int64_t n_sessions, int32_t top_n,
                            const int64_t* row_offsets, const int32_t* session_ids, const OttoLabels* labels,
                            int32_t* session_out, uint64_t* candidates_out, float* scores_out, uint8_t* labels_out, void* stream);

/* recall@k parts of pred [n_sessions][n] (-1 = empty): out_dev[0] = sum over sessions of |set(pred) & set(labels)|,
 * out_dev[1] = sum of min(|labels|, k_clip); recall = out[0] / out[1] (covisitation/inference.py:251-252). */
int otto_recall_counts(const int32_t* pred, int64_t n_sessions, int32_t n, const OttoLabels* labels, int32_t k_clip,
                       uint64_t* out_dev /* [2] */, void* stream);

/* Regular candidate form (ranker/regular_candidate_generation.py:139-180): per session its unique aids, most recent
 * first, with scores |H| .. 1 (:163), followed by one target's ranker-form votes (aid, count).  row_counts gives the rows
 * per session (-> otto_row_offsets), regular_rows writes the flat columns (labels optional, as above). */
int otto_regular_row_counts(const OttoSessions* sessions, const int32_t* aid, const int32_t* score, const int32_t* len,
                            int32_t top_n, int32_t* rows_out, void* stream);
int otto_regular_rows(const OttoSessions* sessions, const int32_t* aid, const int32_t* score, const int32_t* len, int32_t top_n,
                      const int64_t* row_offsets, const int32_t* session_ids, const OttoLabels* labels, int32_t* session_out,
                      uint64_t* candidates_out, float* scores_out, uint8_t* labels_out, void* stream);

/* ---- interaction features over a candidate frame (ranker/interaction_feature_engineering.py:47-113) ---- */

/* One event type's candidate frame as flat device columns (what otto_explode_candidates / otto_regular_rows write and
 * candidate/{event}_{validation,test}.pkl holds), sorted by session. */
typedef struct {
  int64_t n_rows;
  const int32_t* session;
  const uint64_t* candidates;
  const float* candidate_scores;
} OttoCandidateFrame;

/* Output columns, each [n_rows]; a NULL pointer skips the column.  Names follow the script's columns with the
 * "session_candidate_" / "aid_candidate_" / "aid_session_candidate_" prefixes shortened.  cumcount_last = 0 and the
 * *_cumcount_last_mean = NaN stand for the script's nulls (candidate absent from the session / from every session of
 * the group); *_score_std of a single row is NaN (ddof 1). */
typedef struct {
  uint16_t* occurrence_count;          /* events of the session with aid == candidate (:59, :70) */
  uint16_t* cumcount_last;             /* 1-based position of the last such event (:55-57) */
  uint16_t* click_occurrence_count;    /* the same per event type (:60, :72-83) */
  uint16_t* cart_occurrence_count;
  uint16_t* order_occurrence_count;
  float* session_score_mean;           /* per session over its candidate rows (:86-97) */
  float* session_score_std;
  float* session_score_min;
  float* session_score_max;
  float* session_occurrence_count_mean;
  uint32_t* session_occurrence_count_sum;
  uint16_t* session_occurrence_count_max;
  float* session_cumcount_last_mean;
  uint32_t* session_cumcount_last_sum;
  uint16_t* session_cumcount_last_max;
  float* aid_score_mean;               /* per candidate aid over all its rows (:101-111) */
  float* aid_score_std;
  float* aid_score_max;
  float* aid_occurrence_count_mean;
  uint32_t* aid_occurrence_count_sum;
  uint16_t* aid_occurrence_count_max;
  float* aid_cumcount_last_mean;
  uint32_t* aid_cumcount_last_sum;
  uint16_t* aid_cumcount_last_max;
} OttoInteractionFeatures;

/* sessions: the event CSR in file order (ts ascending) of the frame the candidates were generated from; session_ids
 * [n_sessions] ascending.  Scores are summed in fixed point with 8 fractional bits (exact for vote counts and history
 * ranks), squares in 128 bits: the features do not depend on the order of the device atomics.  Every candidate session
 * is expected in the event CSR (the script filters the events BY the candidate sessions, :47); rows of an unknown
 * session get zero counts and zero session aggregates. */
int64_t otto_interaction_scratch_bytes(int64_t n_sessions, int64_t n_rows, int32_t n_aids);
int otto_interaction_features(const OttoSessions* sessions, const int32_t* session_ids, const OttoCandidateFrame* frame,
                              int32_t n_aids, const OttoInteractionFeatures* out, void* scratch, int64_t scratch_bytes,
                              void* stream);

/* ---- long-session branch of the standalone model (covisitation/inference.py:142-199, :336-392) ----
 * Sessions with >= n unique aids (flagged by otto_assemble_predictions) are ranked by recency-weighted event
 * scores plus covisitation bonuses instead of votes.  Targets are clicks, carts, orders (index 0, 1, 2):
 *   score[aid] += w_t[i] * type_coefficient[type_i]          for every event i in file order
 *   score[y]   += bonus[t]   per occurrence of y in table_t[a], a over the target's history set (hist[t])
 *   prediction  = the n best (score desc, first insertion asc)
 * w_click / w_cart hold, for every session length L, np.logspace(0.1 | 0.5, 1, L, base=2) - 1 at
 * w_offset[L] .. w_offset[L] + L (computed by the host with numpy so that the fp64 values are the reference's). */
typedef struct {
  int32_t n_aids;
  int32_t n;                          /* predictions per target: 20 */
  const int32_t* table_aid_y[3];      /* time_weighted, cart_weighted, cart_order (NULL = table absent) */
  const int32_t* table_len[3];
  int32_t table_k[3];
  int32_t hist[3];                    /* OTTO_HIST_TYPE_EQ0, OTTO_HIST_TYPE_LE1, OTTO_HIST_TYPE_GE1 */
  double bonus[3];                    /* 0.05, 0.05, 0.15 */
  double type_coefficient[3];         /* {0: 1, 1: 9, 2: 6} (covisitation/inference.py:72) */
  const double* w_click;
  const double* w_cart;
  const int64_t* w_offset;            /* [max_session_len + 1] */
} OttoRecencySpec;

int64_t otto_recency_scratch_bytes(int32_t max_session_len, int32_t max_table_k);
/* Overwrites rows session_list[0 .. n_list) of pred [3][n_sessions][n]. */
int otto_recency_long(const OttoSessions* sessions, const int32_t* session_list, int32_t n_list, int32_t max_session_len,
                      const OttoRecencySpec* spec, void* scratch, int64_t scratch_bytes, int32_t* pred, void* stream);

/* The same ranking with its weights, for the recency-weighted candidate generator
 * (ranker/recency_weighted_candidate_generator.py:61-105: type coefficients {0: 1, 1: 6, 2: 1}, no covisitation bonus
 * (all tables NULL), every unique aid of the session kept: n >= the longest session of the list).
 * rows_by_list != 0: row i of the outputs belongs to session_list[i] (outputs are [3][n_list][n]); else rows are
 * session indices ([3][n_sessions][n]) and only the listed rows are written.  score (fp64, bit-exact against the
 * Python Counter) and len ([3][rows]) are optional; entries beyond len are -1 / 0.0. */
int otto_recency_scored(const OttoSessions* sessions, const int32_t* session_list, int32_t n_list, int32_t max_session_len,
                        const OttoRecencySpec* spec, void* scratch, int64_t scratch_bytes, int32_t rows_by_list, int32_t* pred,
                        double* score, int32_t* len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OTTO_COVISIT_H_ */

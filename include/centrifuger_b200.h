/*
 * centrifuger_b200.h -- C ABI of the B200-native classification path.
 *
 * This is the drop-in boundary for the reference's
 *     Classifier<Sequence_RunBlock>::Init(char *idxPrefix, _classifierParam)   Classifier.hpp:902
 *     Classifier<Sequence_RunBlock>::Query(char *r1, char *r2, _classifierResult&) Classifier.hpp:950
 * as they are driven, one batch at a time, from ClassifyReads_Thread
 * (CentrifugerClass.cpp:240-340, incl. the DUST masking at :276-316), plus the
 * name/taxonomy look-ups ResultWriter::Output needs (ResultWriter.hpp:199-236).
 *
 * Plain C, plain pointers and sizes; no torch / C++ types.  One handle per GPU;
 * calls on one handle are stream-ordered and must not overlap; different
 * handles are independent.  All functions return CFR_OK (0) or a negative
 * cfr_status; cfr_last_error() gives the message.  There is NO CPU fallback:
 * without a usable CUDA device every entry point fails with CFR_ERR_CUDA.
 */
#ifndef CENTRIFUGER_B200_H
#define CENTRIFUGER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFR_B200_ABI_VERSION 1

typedef enum {
  CFR_OK = 0,
  CFR_ERR_ARG = -1,          /* bad argument */
  CFR_ERR_IO = -2,           /* index file missing / short */
  CFR_ERR_FORMAT = -3,       /* .cfr grammar mismatch */
  CFR_ERR_UNSUPPORTED = -4,  /* protein index, non-ACGT alphabet ... */
  CFR_ERR_CUDA = -5,         /* no device / CUDA runtime error */
  CFR_ERR_NOMEM = -6,
  CFR_ERR_OVERFLOW = -7      /* a per-read device work area was exceeded; the condition stays raised on the
                                handle (later batches report it too) until cfr_reset_counters() */
} cfr_status;

/* In-HBM layout of the BWT.  The *.cfr files are always read unchanged. */
typedef enum {
  CFR_LAYOUT_AUTO = 0,
  CFR_LAYOUT_RUNBLOCK = 1, /* the run-block arrays exactly as stored (rank9 + wavelet trees) */
  CFR_LAYOUT_OCCLINE = 2   /* transcoded on the GPU at load into 32-byte occ sectors (64 BWT rows each) */
} cfr_layout;

/* replaces _classifierParam (Classifier.hpp:17-38) + the --no-dust switch
 * (CentrifugerClass.cpp:367) */
typedef struct {
  int32_t max_result;                /* -k            [1]; <= 0: every best-scoring sequence, never reduced by rank
                                        (Classifier.hpp:620-623, :784-785) */
  int32_t min_hit_len;               /* --min-hitlen  [0 = infer, Classifier.hpp:113-129] */
  int32_t max_result_per_hit_factor; /* --hitk-factor [40] */
  int32_t dust;                      /* 1 = SDUST-mask reads first (reference default) */
  uint64_t consider_secondary_hit_len;    /* [2000]  */
  double consider_secondary_score_factor; /* [0.995] */
  int32_t layout;                    /* cfr_layout */
  int32_t max_batch_reads;           /* device chunk size, 0 = default */
  uint64_t arena_rows;               /* locate work-area rows per chunk, 0 = default */
  int32_t expand_taxid;              /* --expand-taxid (_classifierParam.outputExpandedResult, Classifier.hpp:22):
                                        keep the ids that were promoted into each reported id; read them
                                        with cfr_fetch_expanded / cfr_batch_fetch_expanded  [0] */
  int32_t unlimited_cap;             /* id slots per read when max_result <= 0 (report every best-scoring sequence) [0 = 64] */
} cfr_params;

/* One batch of reads, structure-of-arrays, HOST memory (pinned recommended).
 * Read i of mate m is seq[m][off[m][i] .. off[m][i+1]).  seq2/off2 NULL for
 * single-end.  Bytes are used as given (no upper-casing; any byte outside
 * "ACGT" stops a match, FMIndex.hpp:396,500). */
typedef struct {
  uint64_t n_reads;
  const char *seq1;
  const uint64_t *off1; /* n_reads + 1 entries */
  const char *seq2;
  const uint64_t *off2; /* n_reads + 1 entries, or NULL */
} cfr_read_batch;

/* replaces _classifierResult (Classifier.hpp:41-59).  The assignment ids of
 * read i are ids[i*S .. i*S + n_assign), S = the id stride: max_result, or with max_result <= 0
 * unlimited_cap (cfr_index_info(h, 24) tells). */
typedef struct {
  uint64_t score;
  uint64_t secondary_score;
  int32_t hit_length;
  int32_t query_length;
  int32_t n_assign; /* 0 = unclassified */
  int32_t by_rank;  /* 0: ids are sequence ids; 1: ids are compact taxonomy ids
                       produced by Taxonomy::ReduceTaxIds (Classifier.hpp:798-841) */
} cfr_result;

/* operation counters, accumulated since the last reset; they feed the
 * algorithmic-bytes formula of SURVEY.md 8(d) */
typedef struct {
  uint64_t n_rank;    /* Sequence_RunBlock::Rank-equivalent queries */
  uint64_t n_access;  /* Sequence_RunBlock::Access-equivalent queries */
  uint64_t n_search;  /* lookup-table probes (BackwardSearch calls) */
  uint64_t n_locate;  /* resolved BWT rows */
  uint64_t n_lf;      /* LF steps in locate walks */
  uint64_t n_extend;  /* range BackwardExtend steps */
  uint64_t n_bases;   /* read bases consumed */
  uint64_t n_reads;   /* reads (pairs) classified */
  uint64_t n_launches;/* kernels launched by this library */
} cfr_counters;

typedef struct cfr_handle cfr_handle;
typedef struct cfr_device_batch cfr_device_batch;

void cfr_default_params(cfr_params *p);

/* Classifier::Init: parse <prefix>.1.cfr/.2.cfr (and .4.cfr for the sequence
 * type), upload to `device`'s HBM. */
int cfr_open(const char *idx_prefix, const cfr_params *p, int device, cfr_handle **out);
void cfr_close(cfr_handle *h);
const char *cfr_last_error(void);

/* ClassifyReads_Thread for one batch: H2D, [DUST], search, locate, score,
 * LCA, D2H.  `stream` is a cudaStream_t (NULL = the handle's own stream); the
 * call returns after the results are in `results`/`ids` (host memory). */
int cfr_classify_batch(cfr_handle *h, const cfr_read_batch *in, cfr_result *results,
                       uint64_t *ids, void *stream);

/* Streaming form of cfr_classify_batch for callers that feed batch after batch (the CLI, bench.py's
 * end-to-end loop): submit returns as soon as the copies and kernels are enqueued, so the upload of
 * batch i+1 overlaps the kernels of batch i and the download of batch i-1.  At most three batches are
 * in flight; submitting a fourth first completes the oldest one.  `in`, `results` and `ids` must stay
 * valid (and should be pinned) until cfr_wait_batch(ticket) returns.  n_reads must not exceed
 * cfr_params.max_batch_reads (default 2^20).  The device hit tables of a batch hold
 * n_reads x (longest read / (min_hit_len + 1)) entries per strand, so a caller whose reads differ wildly in
 * length should cut batches by bases and keep very long reads apart (the CLI does: cfr_main.cpp, slot
 * budget); a batch that does not fit returns CFR_ERR_NOMEM. */
int cfr_submit_batch(cfr_handle *h, const cfr_read_batch *in, cfr_result *results, uint64_t *ids, void *stream,
                     int *ticket);
int cfr_wait_batch(cfr_handle *h, int ticket);

/* The same batch with the bases already packed by the producer: 2-bit codes plus one "not ACGT" bit per base,
 * 2.25 bits per base over the host link instead of 8 (what the device works on anyway: on this path a byte
 * outside "ACGT" only stops a match, complements to N and is DUST's fifth symbol -- FMIndex.hpp:396,500,
 * Classifier.hpp:846-856, Dustmasker.hpp:298-302 --, so nothing is lost).  The batch buffer holds mate 1's reads
 * from position 0 and mate 2's from the next multiple of 32 after them; position j lives in word j / 32:
 *   codes[j / 32] bits 2(j % 32) ..   A = 0, C = 1, G = 2, T = 3 (0 for any other byte)
 *   nmask[j / 32] bit  j % 32         1 = the byte was not one of "ACGT" (set for every padding position too)
 * off1 / off2 are positions in that buffer (off1[0] = 0, off2[0] = (off1[n_reads] + 31) & ~31).
 * cfr_pack_reads produces all of it from a cfr_read_batch with `threads` host threads (the CLI's ingest stage
 * calls it per parsed batch); n_words = cfr_packed_words(in) sizes codes / nmask, off1 / off2 hold n_reads + 1. */
typedef struct {
  uint64_t n_reads;
  const uint64_t *codes;
  const uint32_t *nmask;
  uint64_t n_words;
  const uint64_t *off1;
  const uint64_t *off2; /* NULL for single-end */
} cfr_packed_batch;
uint64_t cfr_packed_words(const cfr_read_batch *in);
int cfr_pack_reads(const cfr_read_batch *in, uint64_t *codes, uint32_t *nmask, uint64_t *off1, uint64_t *off2, int threads,
                   cfr_packed_batch *out);
/* cfr_submit_batch for a packed batch (same tickets, same cfr_wait_batch / cfr_fetch_expanded) */
int cfr_submit_packed(cfr_handle *h, const cfr_packed_batch *in, cfr_result *results, uint64_t *ids, void *stream, int *ticket);

/* cfr_submit_batch that also returns the reads as Classifier::Query saw them: the bytes of seq1 / seq2
 * with the DUST-masked intervals replaced by 'N' (CentrifugerClass.cpp:276-316 masks the reads in
 * place, and ResultWriter::Output writes those to the --un / --cl files, ResultWriter.hpp:244-262).
 * masked1 / masked2 (host, as large as this batch's seq1 / seq2 bytes; masked2 may be NULL without
 * mates) are filled when cfr_wait_batch(ticket) returns. */
int cfr_submit_batch_masked(cfr_handle *h, const cfr_read_batch *in, cfr_result *results, uint64_t *ids,
                            char *masked1, char *masked2, void *stream, int *ticket);

/* The expandedTaxIDs column of --expand-taxid (Classifier.hpp:807-838 fills
 * _classifierResult::expandedTaxIdStrings from the child lists of Taxonomy::ReduceTaxIds / LCA,
 * Taxonomy.hpp:767-831, :871-878, :938-971).  For a handle opened with cfr_params.expand_taxid, after
 * cfr_wait_batch(ticket) and before three more batches are submitted:
 *   exp_cnt[i*max_result + j] = number of ids promoted into assignment j of read i (0 when the read was
 *                               not reduced by rank, or the reference prints an empty string);
 *   exp_off[i]                = where read i's lists start in exp_ids (list j follows list j-1);
 *   exp_ids                   = compact taxonomy ids (print with cfr_orig_taxid), *exp_n of them.
 * exp_cap = entries exp_ids can hold; CFR_ERR_OVERFLOW (and *exp_n = the number needed) if too small.
 * cfr_classify_batch cuts its input into chunks that share device slots and keeps no lists: use the
 * streaming form or the resident form (cfr_batch_fetch_expanded) when the lists are wanted. */
int cfr_fetch_expanded(cfr_handle *h, int ticket, uint32_t *exp_cnt, uint64_t *exp_off, uint64_t *exp_ids,
                       uint64_t exp_cap, uint64_t *exp_n);

/* Same work with the batch already resident in HBM (kernel-only timing;
 * multi-GPU shards).  upload = H2D + layout; classify_resident = kernels
 * only, asynchronous on `stream`; fetch = D2H + synchronize. */
int cfr_batch_upload(cfr_handle *h, const cfr_read_batch *in, void *stream, cfr_device_batch **out);
int cfr_classify_resident(cfr_handle *h, cfr_device_batch *b, void *stream);
int cfr_batch_fetch(cfr_handle *h, cfr_device_batch *b, cfr_result *results, uint64_t *ids, void *stream);
/* cfr_fetch_expanded for a resident batch, after cfr_batch_fetch */
int cfr_batch_fetch_expanded(cfr_handle *h, cfr_device_batch *b, uint32_t *exp_cnt, uint64_t *exp_off,
                             uint64_t *exp_ids, uint64_t exp_cap, uint64_t *exp_n, void *stream);
void cfr_batch_free(cfr_handle *h, cfr_device_batch *b);

/* Page-locked host memory for read / result buffers (cudaHostAlloc), so callers that do
 * not link the CUDA runtime can still get asynchronous, overlapped copies. */
void *cfr_host_alloc(size_t bytes);
void cfr_host_free(void *p);

/* index facts: 0 n, 1 b, 2 blockCnt, 3 firstISA, 4 min_hit_len in effect,
 * 5 nodeCnt, 6 seqCnt(+extra), 7 root ctid, 8 layout in use, 9 HBM bytes held,
 * 10 sampleRate, 11 precomputeWidth, 12 max_result,
 * 13 / 14 bytes the batch calls have moved host->device / device->host so far,
 * 15 occ-sector bytes, 16 wide-lookup-table bytes, 17 dense-locate-table bytes,
 * 18 microseconds cfr_open took, 19 run-block bytes released after the transcode,
 * 20 dense-locate spacing (log2; 255 = none), 21 wide-lookup width (0 = none),
 * 22 width of the BWT positions the kernels walk with (32 / 64), 23 pair-line bytes,
 * 24 id slots per read in the `ids` arrays (the stride) */
uint64_t cfr_index_info(const cfr_handle *h, int which);

/* Taxonomy look-ups used by ResultWriter (host tables) */
const char *cfr_seq_name(const cfr_handle *h, uint64_t seq_id);   /* Taxonomy::SeqIdToName   :704 */
const char *cfr_rank_name(const cfr_handle *h, uint64_t ctid);    /* GetTaxRankString(GetTaxIdRank) :497/:659 */
uint64_t cfr_orig_taxid(const cfr_handle *h, uint64_t ctid);      /* Taxonomy::GetOrigTaxId   :633 */
uint64_t cfr_seq_taxid(const cfr_handle *h, uint64_t seq_id);     /* Taxonomy::SeqIdToTaxId   :718 */

/* ResultWriter::Output (ResultWriter.hpp:199-236): TSV row(s) of one read.
 * Returns bytes written or CFR_ERR_ARG if cap is too small. */
int cfr_format_tsv(const cfr_handle *h, const char *read_id, const cfr_result *r,
                   const uint64_t *ids, char *buf, size_t cap);

/* Per-taxon assignment counters kept in HBM (uint64[nodeCnt + 3]: reads
 * assigned per compact taxId, slot nodeCnt = unknown/root, then {reads,
 * classified}).  The device pointer is handed out so the caller can
 * all-reduce it over NCCL; cfr_taxon_counts_read copies it to the host. */
int cfr_taxon_counts_device(cfr_handle *h, void **dev_ptr, uint64_t *n_entries);
int cfr_taxon_counts_read(cfr_handle *h, uint64_t *out, uint64_t n_entries, void *stream);
int cfr_taxon_counts_reset(cfr_handle *h, void *stream);

/* The path's one collective: SUM all-reduce of the counter vector over NCCL, once, after the
 * last batch.  A snapshot of the live counters is reduced, so the handle keeps its own cumulative counts.
 * libnccl.so.2 is loaded at run time; CFR_ERR_UNSUPPORTED if it is not there.
 *   cfr_counts_allreduce        one handle per process: `nccl_comm` is the caller's ncclComm_t (MPI / torchrun style);
 *   cfr_counts_allreduce_local  one process, one handle per GPU (the CLI's --gpus): communicators are made with
 *                               ncclCommInitAll for the call and the reduction runs grouped over NVLink.
 * `out` (host, n_entries, may be NULL) receives the global sum. */
int cfr_counts_allreduce(cfr_handle *h, void *nccl_comm, uint64_t *out, uint64_t n_entries);
int cfr_counts_allreduce_local(cfr_handle **handles, int n_handles, uint64_t *out, uint64_t n_entries);

/* Quantification (replaces `centrifuger-quant`, Quantifier.hpp: LoadReadAssignments :515, Quantification :640,
 * EstimateAbundanceWithEM :236, Output :746) without the round trip through a classification file.  After
 * cfr_quant_enable every batch a handle finishes is coalesced ON THE DEVICE into distinct (targets, weight,
 * unique) records with their multiplicities; cfr_quant_report merges the records of the given handles (one per
 * GPU), runs the reference's abundance estimation over them and writes the report -- byte for byte what
 * centrifuger-quant prints for the TSV of the same reads.  format: 0 centrifuge, 1 metaphlan, 2 CAMI,
 * 3 kraken-report (--output-format); path NULL or "-" = stdout.  min_score / min_hit_length = --min-score / --min-length. */
int cfr_quant_enable(cfr_handle *h, uint64_t min_score, uint64_t min_hit_length);
int cfr_quant_reset(cfr_handle *h);
int cfr_quant_stats(cfr_handle *h, uint64_t *distinct_records, uint64_t *batches, uint64_t *records_moved);
int cfr_quant_report(cfr_handle **handles, int n_handles, const char *idx_prefix, int format, const char *path);

int cfr_get_counters(cfr_handle *h, cfr_counters *c, void *stream);
int cfr_reset_counters(cfr_handle *h, void *stream);

/* Per-stage profile.  Stages: 0 dust, 1 search, 2 select (boundary adjust +
 * strand pick + row plan), 3 locate, 4 score (+LCA), 5 other (memsets / D2D).
 * With profiling on, every kernel is bracketed by CUDA events on its launch
 * stream; cfr_get_stage_times synchronises, folds the finished events into
 * `out` (milliseconds and launch counts since the last reset) and optionally
 * resets.  cfr_get_stage_counters gives the operation counters of one stage
 * (search / select / locate are the stages that touch the index). */
typedef struct {
  double ms[6];
  uint64_t launches[6];
} cfr_stage_times;
int cfr_set_profiling(cfr_handle *h, int on);
int cfr_get_stage_times(cfr_handle *h, cfr_stage_times *out, int reset);
int cfr_get_stage_counters(cfr_handle *h, int stage, cfr_counters *c, void *stream);

/* Diagnostics used by the parity tests (device results copied to host). */
int cfr_debug_bwt_rank(cfr_handle *h, const uint8_t *codes, const uint64_t *pos, const int32_t *inclusive,
                       uint64_t n, uint64_t *out);   /* Sequence_RunBlock::Rank  Sequence_RunBlock.hpp:378 */
int cfr_debug_bwt_access(cfr_handle *h, const uint64_t *pos, uint64_t n, uint8_t *out); /* ::Access :360 */
int cfr_debug_locate(cfr_handle *h, const uint64_t *rows, uint64_t n, uint64_t *seq_ids); /* FMIndex::BackwardToSampledSA FMIndex.hpp:514 */
int cfr_debug_dust(cfr_handle *h, const cfr_read_batch *in, char *masked1, char *masked2); /* Dustmasker.hpp:357 */

#ifdef __cplusplus
}
#endif
#endif

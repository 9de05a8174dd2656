/*
 * cfr_oracle.h -- TEST INFRASTRUCTURE ONLY (parity oracle).
 *
 * A plain-C, single-threaded CPU restatement of the reference classifier's
 * hot path (mourisl/centrifuger @ v1.1.3-r347): *.cfr loading, rank9 /
 * wavelet-tree / run-block BWT rank+access, FM-index backward search, the
 * sampled-SA locate walk, hit scoring, taxonomy LCA / rank promotion and the
 * SDUST read masker.  Every function cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * this library, and only as the checker.  The product (centrifuger_b200/) never
 * links, loads or calls it.
 *
 * Pinning (tests/test_oracle_golden.py, tests/golden/, tests/fuzz/): checked against
 * (1) /root/reference/example/example_class.out, (2) ~100 committed outputs of the
 * unmodified reference binary built into oracle/_ref/ (option grid, --expand-taxid,
 * long reads, --consider-secondary), (3) the reference's own headers compiled into
 * small drivers under oracle/_ref/ and queried at random: fm_ref (rank, access,
 * FM rank, backward search, locate walk), classifier_ref (hit lists of the search
 * stage), taxonomy_ref (ReduceTaxIds with its child lists), dust_ref (SDUST), and
 * (4) a differential fuzzer against the reference binary on generated reads.
 */
#ifndef CFR_ORACLE_H
#define CFR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cfr_oracle cfr_oracle;

/* Classifier.hpp:17-38 (_classifierParam) */
typedef struct {
  int32_t max_result;            /* -k, default 1 */
  int32_t min_hit_len;           /* <=0: infer (Classifier.hpp:113-129) */
  int32_t max_result_per_hit_factor; /* --hitk-factor, default 40 */
  int32_t pad_;
  uint64_t consider_secondary_hit_len;   /* default 2000 */
  double consider_secondary_score_factor; /* default 0.995 */
} cfr_oracle_param;

/* Classifier.hpp:70-85 (_BWTHit) */
typedef struct {
  uint64_t sp, ep;
  int32_t l;
  int32_t strand;
  int32_t offset;
  int32_t pad_;
} cfr_oracle_hit;

/* Classifier.hpp:41-59 (_classifierResult); names are resolved lazily through
 * cfr_oracle_seq_name / cfr_oracle_rank_name so the struct stays POD. */
typedef struct {
  uint64_t score;
  uint64_t secondary_score;
  int32_t hit_length;
  int32_t query_length;
  int32_t n;        /* number of assignments (may exceed cap; only cap stored) */
  int32_t by_rank;  /* 0: ids[] are seqIds (name = sequence name);
                       1: ids[] are compact taxIds after ReduceTaxIds
                          (name = rank string) */
  uint64_t ids[64];
  uint64_t tax_ids[64]; /* original taxonomy ids */
} cfr_oracle_result;

/* operation counters (SURVEY.md 8(d) algorithmic-bytes formula) */
typedef struct {
  uint64_t n_rank;    /* Sequence_RunBlock::Rank calls */
  uint64_t n_access;  /* Sequence_RunBlock::Access calls */
  uint64_t n_search;  /* FMIndex::BackwardSearch calls (lookup-table probes) */
  uint64_t n_locate;  /* sampled-SA reads that ended a locate walk */
  uint64_t n_lf;      /* LF steps inside locate walks */
  uint64_t n_extend;  /* range BackwardExtend calls */
} cfr_oracle_counters;

cfr_oracle *cfr_oracle_open(const char *idx_prefix);
void cfr_oracle_close(cfr_oracle *o);
void cfr_oracle_default_param(cfr_oracle_param *p);

/* scalars: 0 n, 1 b, 2 blockCnt, 3 firstISA, 4 lastChr, 5 sampleRate,
 * 6 sampledSA bits, 7 sampleSize, 8 precomputeWidth, 9 selectedSA count,
 * 10 nodeCnt, 11 seqCnt, 12 extraSeqCnt, 13 rootCTaxId, 14 inferred minHitLen,
 * 15 plain-tree n, 16 run-tree n, 17 adjustedSA0, 18..22 C[0..4] */
uint64_t cfr_oracle_scalar(const cfr_oracle *o, int which);

/* compactds layer */
uint64_t cfr_oracle_bwt_rank(cfr_oracle *o, char c, uint64_t i, int inclusive);
char cfr_oracle_bwt_access(cfr_oracle *o, uint64_t i);
uint64_t cfr_oracle_fm_rank(cfr_oracle *o, char c, uint64_t p, int inclusive);
uint64_t cfr_oracle_backward_search(cfr_oracle *o, const char *s, uint64_t m,
                                    uint64_t *sp, uint64_t *ep);
uint64_t cfr_oracle_locate(cfr_oracle *o, uint64_t row, uint64_t *steps);

/* classifier layer */
int cfr_oracle_infer_min_hit_len(const cfr_oracle *o);
/* SearchForwardAndReverse: returns the number of hits (stores <= cap) */
int cfr_oracle_search(cfr_oracle *o, const cfr_oracle_param *p, const char *r1,
                      const char *r2, cfr_oracle_hit *hits, int cap);
void cfr_oracle_query(cfr_oracle *o, const cfr_oracle_param *p, const char *r1,
                      const char *r2, cfr_oracle_result *res);

/* taxonomy layer */
uint64_t cfr_oracle_seqid_to_taxid(const cfr_oracle *o, uint64_t seq_id);
uint64_t cfr_oracle_orig_taxid(const cfr_oracle *o, uint64_t ctid);
const char *cfr_oracle_seq_name(const cfr_oracle *o, uint64_t seq_id);
const char *cfr_oracle_rank_name(const cfr_oracle *o, uint64_t ctid);
/* ReduceTaxIds (Taxonomy.hpp:839-973); returns count written to out */
int cfr_oracle_reduce_taxids(const cfr_oracle *o, const uint64_t *tax_ids,
                             int n, int k, uint64_t *out, int cap);

/* the same with promotedChildTaxIds: *n_lists lists (child_cnt[i] entries each, concatenated in
 * child); the classifier prints them only when *n_lists equals the returned count */
int cfr_oracle_reduce_taxids_expanded(const cfr_oracle *o, const uint64_t *tax_ids, int n, int k,
                                      uint64_t *out, int cap, uint64_t *child, int child_cap,
                                      int32_t *child_cnt, int *n_lists);
/* Query with outputExpandedResult (--expand-taxid, Classifier.hpp:22, :792-838): child_cnt[64]
 * = entries per reported id, child = the compact tax ids list after list; returns their total */
int cfr_oracle_query_expanded(cfr_oracle *o, const cfr_oracle_param *p, const char *r1, const char *r2,
                              cfr_oracle_result *res, uint64_t *child, int child_cap, int32_t *child_cnt);
int cfr_oracle_format_tsv_expanded(const cfr_oracle *o, const char *read_id, const cfr_oracle_result *res,
                                   const uint64_t *child, const int32_t *child_cnt, char *buf, size_t cap);

/* SDUST masker (Dustmasker.hpp:357-421 + CentrifugerClass.cpp:276-316):
 * masks seq[0..n) in place with 'N'; returns the number of masked intervals */
int cfr_oracle_dust_mask(char *seq, size_t n);

/* ResultWriter.hpp:199-236: appends the TSV row(s) of one read; returns bytes
 * written (excluding the NUL), or -1 if cap is too small */
int cfr_oracle_format_tsv(const cfr_oracle *o, const char *read_id,
                          const cfr_oracle_result *res, char *buf, size_t cap);

void cfr_oracle_get_counters(const cfr_oracle *o, cfr_oracle_counters *c);
void cfr_oracle_reset_counters(cfr_oracle *o);

#ifdef __cplusplus
}
#endif
#endif

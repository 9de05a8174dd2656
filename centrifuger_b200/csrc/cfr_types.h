// cfr_types.h -- PODs shared by the host runtime and the device kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CFR_HD __host__ __device__ __forceinline__
#define CFR_D __device__ __forceinline__
#else
#define CFR_HD inline
#define CFR_D inline
#endif

namespace cfrb200 {

typedef unsigned long long u64;
typedef unsigned int u32;

struct u64x2 {
  u64 x, y;
};

// rank9 bitvector as stored by the reference (Bitvector_Plain + DS_Rank9)
struct DevBV {
  const u64 *B;
  const u64 *R;
  u64 n;
};

// 3-node wavelet tree over {A,C,G,T} = codes {00,01,10,11}
struct DevWT {
  DevBV node[3];
  int child[3][2];
  u64 n;
};

// One 32-byte occ sector of the transcoded layout: 64 BWT symbols and the number
// of A, C and G before them; the number of T follows because every BWT row holds
// one of the four symbols (the text has no '$', FMIndex.hpp:355-358):
//     #T before the sector = 64 * sector - (#A + #C + #G).
//   lo, hi  bit planes of the 64 two-bit symbol codes
//   w2      low 32 bits of #A | low 32 bits of #C << 32
//   w3      low 32 bits of #G | (bits 32..39 of #A, #C, #G as three bytes) << 32   (40-bit counters: n < 2^40)
// 32 bytes = one L2 / HBM3e sector: a rank or an LF step reads exactly one sector.
struct alignas(32) OccLine {
  u64 lo, hi, w2, w3;
};

// One 128-byte line of the PAIR layout (two BackwardExtend steps per DRAM line): 64 BWT rows.
// code1 = the row's own symbol B[i]; code2 = B[img(i)], the symbol of the row that i maps to when it
// is extended by its own symbol: img(i) = C[code1] + occ(code1, i) + FMIndex::Rank's correction.
//   w[0..3]    bit planes: code1 low, code1 high, code2 low, code2 high
//   w[4..11]   P[c1*4+c2]: rows before this line whose (code1, code2) is that pair; 32 bits each (pair 2k in
//              the low half of w[4+k], pair 2k+1 in the high half), relative to the line's superblock
//   w[12..13]  S[c1]: rows before this line with code1 = c1, stored the same way
//   w[14..15]  unused
// Four adjacent lanes fetch the line as ONE coalesced 128-byte request (32 bytes each): lane 0 the planes,
// lanes 1 and 2 the pair counters of c1 = 0,1 / 2,3, lane 3 the single counters.
struct alignas(128) PairLine {
  u64 w[16];
};
#define CFR_PAIR_SB_SHIFT 26  // lines per superblock (2^32 rows): the 32-bit counters are relative to it

struct DevIndex {
  // FM-index scalars (FMIndex.hpp:191-199)
  u64 n;
  u64 first_isa;
  int last_code;  // code of _lastChr
  u64 C[5];       // _plainAlphabetPartialSum
  // run-block BWT exactly as stored (Sequence_RunBlock.hpp:15-20)
  u64 b, block_cnt;
  DevBV block_type;
  DevWT plain, run;
  // transcoded layout (nullptr when not built)
  const OccLine *occ;
  // locate (FMIndex.hpp:13-41)
  int sample_rate;
  int sample_shift;  // log2(sample_rate) or -1
  int sa_bits;
  const u64 *sampled_sa;
  u64 adjusted_sa0;
  const u64x2 *sel;  // {row, seqId}, ascending row
  u64 sel_cnt;
  const u64 *sel_filter;
  int sel_filter_rate;
  int filter_shift;  // log2(sel_filter_rate) or -1
  // 10-mer (precomputeWidth-mer) lookup table
  int pre_width;
  const u64x2 *lookup;  // {start, len}
  // optional wide table, built at load from `lookup` + BackwardExtend for indexes that live in HBM:
  // entry of a WW-mer = where FMIndex::BackwardSearch stands after its last WW bases
  // {sp, ep | l << 56}: l < WW means the search ended inside them (l = W - 1: empty W-mer)
  int wide_width;  // 0 = none
  const u64x2 *wide;
  // optional pair layout (nullptr when not built): k_search then does two BackwardExtend steps per line
  const PairLine *pairs;
  const u64 *pair_sb;  // [superblock][20]: absolute P[16] then S[4] at the superblock's first line
  u64 pair_D[16];      // D[c1*4+c2] = # of c2 among the rows below C[c1]
  int pair_E;          // B[C[last_code]]: the symbol at the row the virtual '$' row maps to
  int pair_F;          // code2 of row first_isa
  // optional dense locate table, built at load by running FMIndex::BackwardToSampledSA once from
  // every row that is a multiple of 2^dense_shift: dense[row >> dense_shift] = its sequence id.  A walk
  // that reaches such a row ends there with the answer the reference's longer walk would find.
  int dense_shift;  // -1 = none
  int dense_idx_shift;  // the entry of dense row i is dense[i >> dense_idx_shift] (= dense_shift once the table is built; while
                        // it is being built level by level the rows of the coarser levels are the dense ones)
  int dense16;      // entries are 16 bits wide (every sequence id of the index is below 2^16)
  const u32 *dense;
  // taxonomy (Taxonomy.hpp)
  u64 node_cnt, seq_cnt, root;
  const u32 *parent;
  const unsigned char *rank;
  const u32 *seq_to_tax;  // value node_cnt = unknown
  unsigned char rank_num[32];
};

struct DevParams {
  int max_result;  // -k; <= 0 = every best-scoring sequence is reported (Classifier.hpp:620-623, :784-785)
  int ids_stride;  // id slots per read in out_ids: max_result, or the cap of the unlimited form
  int min_hit_len;
  int hitk_factor;
  u64 secondary_len;
  double secondary_factor;
  int quorum;  // tasks that must be waiting before a warp runs its transition block
};

// Classifier.hpp:70-85 (_BWTHit); strand is kept in the low bits of `meta`
struct Hit {
  u64 sp, ep;
  int l;
  int offset;
};

struct FinalHit {
  u64 sp, ep;
  int l;
  int offset;
  int strand;   // -1 / +1 (template strand, Classifier.hpp:560)
  u32 row_cnt;  // rows expanded for locate
};

// per-read bookkeeping between the pipeline stages
struct ReadWork {
  u64 arena_base;  // first arena row of this read
  u32 arena_rows;  // rows reserved
  u32 n_hits;      // final hits
  int status;      // 0 ok, 1 deferred (arena full)
};

struct SeqRec {  // value of std::map<size_t,_seqHitRecord> (Classifier.hpp:62-67,590)
  u64 score;
  u32 seq_id;
  int hit_length;
};

struct DevResult {  // mirrors cfr_result
  u64 score;
  u64 secondary_score;
  int hit_length;
  int query_length;
  int n_assign;
  int by_rank;
};

struct DevCounters {
  u64 n_rank, n_access, n_search, n_locate, n_lf, n_extend, n_bases, n_reads;
  u64 error_flags;  // bit 0: taxonomy path deeper than the device cap; bit 1: --expand-taxid lists exceed the batch's list area
                    // (sticky until cfr_reset_counters: batches overlap on the device, so a flag cannot be attributed and cleared per batch); bit 1: --expand-taxid lists exceed the batch's list area
};

enum { CFR_TAX_PATH_CAP = 128 };

// pipeline stages (profiling slots; DevCounters is an array indexed by stage)
enum { CFR_STAGE_DUST = 0, CFR_STAGE_SEARCH = 1, CFR_STAGE_SELECT = 2, CFR_STAGE_LOCATE = 3, CFR_STAGE_SCORE = 4,
       CFR_STAGE_OTHER = 5, CFR_N_STAGES = 6 };

}  // namespace cfrb200

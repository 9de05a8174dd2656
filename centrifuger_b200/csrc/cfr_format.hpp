// cfr_format.hpp -- zero-copy parser for the reference's *.cfr index files.
//
// Grammar (little-endian, unpadded; every array after the 1-byte lastChr at
// offset 24 is unaligned, so the views below are BYTE pointers into the mmap):
//   .1.cfr  FMIndex::Save                FMIndex.hpp:571-586
//           Sequence_RunBlock::Save      compactds/Sequence_RunBlock.hpp:468-476
//           Sequence::Save / Alphabet    Sequence.hpp:24-29, Alphabet.hpp:194-205
//           Bitvector_Plain::Save        Bitvector_Plain.hpp:182-196
//           DS_Rank9::Save               DS_Rank.hpp:275-282
//           DS_Select::Save (speed 0)    DS_Select.hpp:679-686
//           Sequence_WaveletTree::Save   Sequence_WaveletTree.hpp:303-311 (+ node :21-27)
//           _FMIndexAuxData::Save        FMIndex.hpp:100-134
//           FixedSizeElemArray::Save     FixedSizeElemArray.hpp:388-394
//   .2.cfr  Taxonomy::Save               Taxonomy.hpp:1238-1257, MapID.hpp:76-81
//   .4.cfr  text key/value               Classifier.hpp:867-895 (sequence_type)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace cfrb200 {

struct MappedFile {
  const uint8_t *data = nullptr;
  size_t size = 0;
  int fd = -1;
  bool open(const std::string &path, std::string &err);
  void close();
  ~MappedFile() { close(); }
  MappedFile() = default;
  MappedFile(const MappedFile &) = delete;
  MappedFile &operator=(const MappedFile &) = delete;
};

// one rank9 bitvector: bit words B and the interleaved rank9 counters R
struct BvView {
  uint64_t nbits = 0;
  const uint8_t *B = nullptr;  // ceil(nbits/64) u64
  uint64_t words = 0;
  const uint8_t *R = nullptr;  // 2*ceil(wordCnt/8) u64
  uint64_t rwords = 0;
};

// 3-node wavelet tree over ACGT (root, A/C leaf, G/T leaf); children are read
// from the file, not assumed
struct WtView {
  uint64_t n = 0;
  int32_t node_cnt = 0;  // 0 = empty tree (no run blocks / no plain blocks)
  BvView node[3];
  int32_t child[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}};
};

struct TaxonomyHost {
  uint64_t node_cnt = 0, seq_cnt = 0, extra_seq_cnt = 0, root = 0;
  std::vector<uint64_t> parent;     // TaxonomyNode::parentTid
  std::vector<uint8_t> rank;        // TaxonomyNode::rank
  std::vector<uint8_t> leaf;        // TaxonomyNode::leaf
  std::vector<uint64_t> orig_taxid; // MapID<uint64_t>::_toOrigElem
  std::vector<std::string> tax_name;
  std::vector<uint64_t> seq_to_tax;
  std::vector<std::string> seq_name; // seq_cnt + extra_seq_cnt
};

struct CfrIndexFile {
  MappedFile map1;
  // FMIndex
  uint64_t n = 0, alphabet_bits = 0, first_isa = 0;
  char last_chr = 0;
  // Sequence_RunBlock
  uint64_t rb_n = 0, b = 0, block_cnt = 0;
  BvView block_type;
  WtView plain, run;  // _waveletSeq, _runBlockSeq
  uint64_t C[5] = {0, 0, 0, 0, 0};
  // aux
  int32_t sample_strategy = 0, sample_rate = 0;
  uint64_t sample_size = 0, precompute_width = 0, precompute_size = 0, adjusted_sa0 = 0;
  int32_t sa_bits = 0;
  uint64_t sa_n = 0, sa_words = 0;
  const uint8_t *sa_w = nullptr;
  const uint8_t *lookup = nullptr;  // precompute_size x {start u64, len u64}
  uint64_t sel_cnt = 0;
  int32_t sel_filter_rate = 1024;
  const uint8_t *sel = nullptr;     // sel_cnt x {row u64, seqId u64}, ascending row
  bool has_end_marker = false;
  bool protein = false;
  TaxonomyHost tax;

  // returns a cfr_status-compatible code (0 ok)
  int load(const std::string &prefix, std::string &err);
};

// Classifier::InferMinHitLen (Classifier.hpp:113-129), nucleotide branch
int infer_min_hit_len(uint64_t n);

// Taxonomy::InitTaxRankNum (Taxonomy.hpp:100-144): rank id -> promotion level
void init_tax_rank_num(uint8_t out[32]);
// Taxonomy::GetTaxRankString (Taxonomy.hpp:497-532)
const char *tax_rank_string(uint8_t rank);
// rank numbers of Taxonomy.hpp:25-58 that the quantifier's reports name
enum { TAX_RANK_STRAIN = 1, TAX_RANK_SPECIES = 2, TAX_RANK_GENUS = 3, TAX_RANK_FAMILY = 4, TAX_RANK_ORDER = 5, TAX_RANK_CLASS = 6,
       TAX_RANK_PHYLUM = 7, TAX_RANK_KINGDOM = 8, TAX_RANK_DOMAIN = 9, TAX_RANK_SUPER_KINGDOM = 24, TAX_RANK_ACELLULAR_ROOT = 30 };
// Taxonomy::IsCanonicalRankNum (Taxonomy.hpp:435-443)
inline bool tax_rank_is_canonical(uint8_t r) {
  return r == TAX_RANK_STRAIN || r == TAX_RANK_SPECIES || r == TAX_RANK_GENUS || r == TAX_RANK_FAMILY || r == TAX_RANK_ORDER ||
         r == TAX_RANK_CLASS || r == TAX_RANK_PHYLUM || r == TAX_RANK_KINGDOM || r == TAX_RANK_SUPER_KINGDOM ||
         r == TAX_RANK_DOMAIN || r == TAX_RANK_ACELLULAR_ROOT;
}
// <prefix>.2.cfr alone (Taxonomy::Load, Taxonomy.hpp:1259-1290): 0 or a negative status with `err` set
int load_taxonomy_file(const std::string &path, TaxonomyHost &t, std::string &err);

inline uint64_t load_u64(const uint8_t *p) {
  uint64_t v;
  __builtin_memcpy(&v, p, 8);
  return v;
}

}  // namespace cfrb200

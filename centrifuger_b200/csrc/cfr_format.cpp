// cfr_format.cpp -- see cfr_format.hpp for the grammar citations.
#include "cfr_format.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>

namespace cfrb200 {

enum { ST_OK = 0, ST_IO = -2, ST_FORMAT = -3, ST_UNSUPPORTED = -4 };

bool MappedFile::open(const std::string &path, std::string &err) {
  close();
  fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) {
    err = "cannot open " + path;
    return false;
  }
  struct stat st;
  if (fstat(fd, &st) != 0) {
    err = "cannot stat " + path;
    return false;
  }
  size = (size_t)st.st_size;
  if (size == 0) {
    err = path + " is empty";
    return false;
  }
  void *p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  if (p == MAP_FAILED) {
    err = "mmap failed for " + path;
    data = nullptr;
    return false;
  }
  data = (const uint8_t *)p;
  madvise(p, size, MADV_SEQUENTIAL);
  return true;
}

void MappedFile::close() {
  if (data) munmap((void *)data, size);
  if (fd >= 0) ::close(fd);
  data = nullptr;
  fd = -1;
  size = 0;
}

namespace {

struct Cursor {
  const uint8_t *p;
  const uint8_t *end;
  bool ok = true;
  Cursor(const uint8_t *b, size_t n) : p(b), end(b + n) {}
  bool need(uint64_t n) {
    if (!ok || (uint64_t)(end - p) < n) ok = false;
    return ok;
  }
  template <class T>
  T get() {
    T v{};
    if (need(sizeof(T))) {
      memcpy(&v, p, sizeof(T));
      p += sizeof(T);
    }
    return v;
  }
  const uint8_t *bytes(uint64_t n) {
    if (!need(n)) return nullptr;
    const uint8_t *r = p;
    p += n;
    return r;
  }
};

uint64_t div_ceil(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

struct AlphabetInfo {
  uint64_t n = 0;
  int32_t method = 0;
  std::string list;
  int32_t code[256];
  int16_t code_len[256];
};

void parse_alphabet(Cursor &c, AlphabetInfo &a) {
  c.get<uint64_t>();  // _space
  a.method = c.get<int32_t>();
  a.n = c.get<uint64_t>();
  if (a.n != 0) {
    const uint8_t *l = c.bytes(a.n);
    if (l) a.list.assign((const char *)l, a.n);
    const uint8_t *cd = c.bytes(256 * 4);
    const uint8_t *cl = c.bytes(256 * 2);
    if (cd && cl) {
      memcpy(a.code, cd, 256 * 4);
      memcpy(a.code_len, cl, 256 * 2);
    }
  }
}

// The device code hard-wires A,C,G,T -> 0,1,2,3 with 2-bit plain codes.
bool alphabet_is_plain_acgt(const AlphabetInfo &a) {
  if (a.n != 4 || a.method != 1 || a.list != "ACGT") return false;
  const char *s = "ACGT";
  for (int i = 0; i < 4; ++i)
    if (a.code[(int)s[i]] != i || a.code_len[(int)s[i]] != 2) return false;
  return true;
}

int parse_bv(Cursor &c, BvView &v, std::string &err) {
  c.get<uint64_t>();  // Bitvector::_space
  v.nbits = c.get<uint64_t>();
  c.get<int32_t>();  // _rb
  c.get<int32_t>();  // _sb
  c.get<int32_t>();  // _selectSpeed
  c.get<int32_t>();  // _selectTypeSupport
  if (v.nbits > 0) {
    v.words = div_ceil(v.nbits, 64);
    v.B = c.bytes(v.words * 8);
    c.get<uint64_t>();  // DS_Rank9::_space
    uint64_t word_cnt = c.get<uint64_t>();
    v.rwords = div_ceil(word_cnt, 8) * 2;
    v.R = c.bytes(v.rwords * 8);
    c.get<uint64_t>();  // DS_Select::_space
    uint64_t sn = c.get<uint64_t>();
    int32_t speed = c.get<int32_t>();
    if (c.ok && !(speed == 0 || sn == 0)) {
      err = "bitvector carries select tables (speed != 0): not a centrifuger index";
      return ST_UNSUPPORTED;
    }
    if (c.ok && word_cnt != v.words) {
      err = "rank9 word count disagrees with the bitvector length";
      return ST_FORMAT;
    }
  }
  return c.ok ? ST_OK : ST_FORMAT;
}

int parse_wt(Cursor &c, WtView &t, std::string &err) {
  c.get<uint64_t>();  // Sequence::_space
  t.n = c.get<uint64_t>();
  AlphabetInfo a;
  parse_alphabet(c, a);
  int32_t node_cnt = c.get<int32_t>();
  c.get<int32_t>();  // _selectSpeed
  if (!c.ok) return ST_FORMAT;
  if (a.n == 0) {  // empty tree: header only (Sequence_WaveletTree.hpp:320-321)
    t.node_cnt = 0;
    return ST_OK;
  }
  if (!alphabet_is_plain_acgt(a)) {
    err = "wavelet tree alphabet is not plain ACGT (protein / custom alphabets are unsupported)";
    return ST_UNSUPPORTED;
  }
  if (node_cnt < 1 || node_cnt > 3) {
    err = "unexpected wavelet tree node count";
    return ST_FORMAT;
  }
  t.node_cnt = node_cnt;
  for (int i = 0; i < node_cnt; ++i) {
    c.get<uint64_t>();  // prefix
    c.get<int32_t>();   // prefixLen
    t.child[i][0] = c.get<int32_t>();
    t.child[i][1] = c.get<int32_t>();
    int st = parse_bv(c, t.node[i], err);
    if (st != ST_OK) return st;
    for (int k = 0; k < 2; ++k)
      if (t.child[i][k] < -1 || t.child[i][k] >= node_cnt) {
        err = "wavelet tree child index out of range";
        return ST_FORMAT;
      }
  }
  return c.ok ? ST_OK : ST_FORMAT;
}

std::string get_string(Cursor &c) {
  uint64_t len = c.get<uint64_t>();
  const uint8_t *b = c.bytes(len);
  return b ? std::string((const char *)b, len) : std::string();
}

int parse_taxonomy(const std::string &path, TaxonomyHost &t, std::string &err) {
  MappedFile m;
  if (!m.open(path, err)) return ST_IO;
  Cursor c(m.data, m.size);
  t.node_cnt = c.get<uint64_t>();
  t.seq_cnt = c.get<uint64_t>();
  t.extra_seq_cnt = c.get<uint64_t>();
  if (!c.ok || t.node_cnt > (m.size / 16) || t.seq_cnt > m.size) {
    err = path + ": bad taxonomy header";
    return ST_FORMAT;
  }
  t.parent.resize(t.node_cnt);
  t.rank.resize(t.node_cnt);
  t.leaf.resize(t.node_cnt);
  for (uint64_t i = 0; i < t.node_cnt; ++i) {  // TaxonomyNode: u64 parent, u8 rank, u8 leaf, pad[6]
    t.parent[i] = c.get<uint64_t>();
    t.rank[i] = c.get<uint8_t>();
    t.leaf[i] = c.get<uint8_t>();
    c.bytes(6);
  }
  uint64_t map_n = c.get<uint64_t>();
  if (!c.need(map_n * 8)) {
    err = path + ": bad taxid map";
    return ST_FORMAT;
  }
  t.orig_taxid.resize(map_n);
  for (uint64_t i = 0; i < map_n; ++i) t.orig_taxid[i] = c.get<uint64_t>();
  t.tax_name.resize(t.node_cnt);
  for (uint64_t i = 0; i < t.node_cnt && c.ok; ++i) t.tax_name[i] = get_string(c);
  if (!c.need(t.seq_cnt * 8)) {
    err = path + ": bad seqid table";
    return ST_FORMAT;
  }
  t.seq_to_tax.resize(t.seq_cnt);
  for (uint64_t i = 0; i < t.seq_cnt; ++i) t.seq_to_tax[i] = c.get<uint64_t>();
  t.seq_name.resize(t.seq_cnt + t.extra_seq_cnt);
  for (uint64_t i = 0; i < t.seq_cnt + t.extra_seq_cnt && c.ok; ++i) t.seq_name[i] = get_string(c);
  if (!c.ok || c.p != c.end) {
    err = path + ": taxonomy grammar mismatch";
    return ST_FORMAT;
  }
  // Taxonomy::FindRoot (Taxonomy.hpp:426-433)
  t.root = t.node_cnt;
  for (uint64_t i = 0; i < t.node_cnt; ++i)
    if (t.parent[i] == i) {
      t.root = i;
      break;
    }
  for (uint64_t i = 0; i < t.node_cnt; ++i)
    if (t.parent[i] >= t.node_cnt) {
      err = path + ": taxonomy parent id out of range";
      return ST_FORMAT;
    }
  return ST_OK;
}

// Classifier::IsProteinDatabase (Classifier.hpp:867-895)
bool is_protein_database(const std::string &path) {
  FILE *fp = fopen(path.c_str(), "r");
  if (!fp) return false;
  char key[128], val[128];
  bool ret = false;
  while (fscanf(fp, "%127s %127s", key, val) == 2)
    if (!strcmp(key, "sequence_type") && !strcmp(val, "amino_acid")) ret = true;
  fclose(fp);
  return ret;
}

}  // namespace

int CfrIndexFile::load(const std::string &prefix, std::string &err) {
  protein = is_protein_database(prefix + ".4.cfr");
  if (protein) {
    err = "protein (amino_acid) indexes are not supported by the B200 path";
    return ST_UNSUPPORTED;
  }
  if (!map1.open(prefix + ".1.cfr", err)) return ST_IO;
  Cursor c(map1.data, map1.size);
  n = c.get<uint64_t>();
  alphabet_bits = c.get<uint64_t>();
  first_isa = c.get<uint64_t>();
  last_chr = (char)c.get<uint8_t>();
  // Sequence_RunBlock
  c.get<uint64_t>();  // Sequence::_space
  rb_n = c.get<uint64_t>();
  AlphabetInfo rb_alpha;
  parse_alphabet(c, rb_alpha);
  b = c.get<uint64_t>();
  block_cnt = c.get<uint64_t>();
  if (!c.ok) {
    err = ".1.cfr: truncated header";
    return ST_FORMAT;
  }
  if (!alphabet_is_plain_acgt(rb_alpha) || alphabet_bits != 2) {
    err = "index alphabet is not plain ACGT";
    return ST_UNSUPPORTED;
  }
  int st = parse_bv(c, block_type, err);
  if (st != ST_OK) {
    if (err.empty()) err = ".1.cfr: block-type bitvector";
    return st;
  }
  if ((st = parse_wt(c, plain, err)) != ST_OK || (st = parse_wt(c, run, err)) != ST_OK) {
    if (err.empty()) err = ".1.cfr: wavelet trees";
    return st;
  }
  AlphabetInfo a1, a2;
  parse_alphabet(c, a1);  // FMIndex::_alphabets
  parse_alphabet(c, a2);  // FMIndex::_plainAlphabetCoder
  if (!c.ok) {
    err = ".1.cfr: truncated alphabets";
    return ST_FORMAT;
  }
  if (!alphabet_is_plain_acgt(a1) || !alphabet_is_plain_acgt(a2)) {
    err = "FM-index alphabets are not plain ACGT";
    return ST_UNSUPPORTED;
  }
  for (int i = 0; i < 5; ++i) C[i] = c.get<uint64_t>();
  // _FMIndexAuxData
  uint64_t aux_n = c.get<uint64_t>();
  sample_strategy = c.get<int32_t>();
  sample_rate = c.get<int32_t>();
  sample_size = c.get<uint64_t>();
  precompute_width = c.get<uint64_t>();
  precompute_size = c.get<uint64_t>();
  adjusted_sa0 = c.get<uint64_t>();
  c.get<uint64_t>();  // FixedSizeElemArray::_size
  sa_bits = c.get<int32_t>();
  sa_n = c.get<uint64_t>();
  if (!c.ok || sa_bits < 0 || sa_bits > 64) {
    err = ".1.cfr: bad sampled-SA header";
    return ST_FORMAT;
  }
  sa_words = div_ceil(sa_n * (uint64_t)sa_bits, 64);
  sa_w = c.bytes(sa_words * 8);
  lookup = c.bytes(precompute_size * 16);
  uint64_t max_lcp = c.get<uint64_t>();
  if (c.ok && max_lcp > 0) c.bytes(div_ceil(aux_n, 64) * 8 * 2);  // semiLcp arrays, unused here
  sel_cnt = c.get<uint64_t>();
  sel_filter_rate = c.get<int32_t>();
  sel = c.bytes(sel_cnt * 16);
  if (!c.ok) {
    err = ".1.cfr: truncated auxiliary data";
    return ST_FORMAT;
  }
  // hasEndMarker is absent in old indexes (FMIndex.hpp:178-181)
  has_end_marker = false;
  if (c.p != c.end) has_end_marker = c.get<uint8_t>() != 0;
  if (has_end_marker) {
    err = "indexes with end-marker SA are not supported";
    return ST_UNSUPPORTED;
  }
  if (c.p != c.end) {
    err = ".1.cfr: trailing bytes after parse (grammar mismatch)";
    return ST_FORMAT;
  }
  // sanity
  if (n == 0 || rb_n != n || aux_n != n || b == 0 || block_cnt != div_ceil(n, b) ||
      block_type.nbits != block_cnt || sample_rate <= 0 || C[4] != n || first_isa >= n ||
      (precompute_width > 0 && precompute_size != (1ull << (2 * precompute_width))) ||
      precompute_width > 16 || sa_n != sample_size || sel_filter_rate <= 0) {
    err = ".1.cfr: inconsistent header fields";
    return ST_FORMAT;
  }
  if (last_chr != 'A' && last_chr != 'C' && last_chr != 'G' && last_chr != 'T') {
    err = ".1.cfr: lastChr outside ACGT";
    return ST_FORMAT;
  }
  for (uint64_t i = 1; i < sel_cnt; ++i)
    if (load_u64(sel + (i - 1) * 16) >= load_u64(sel + i * 16)) {
      err = ".1.cfr: selectedSA rows not ascending";
      return ST_FORMAT;
    }
  st = parse_taxonomy(prefix + ".2.cfr", tax, err);
  if (st != ST_OK) return st;
  return ST_OK;
}

int infer_min_hit_len(uint64_t n) {
  int mhl = 23;
  uint64_t kmerspace = 1;
  for (int i = 0; i < mhl; ++i) kmerspace *= 4;
  kmerspace /= 2;
  for (; mhl <= 32; ++mhl) {
    if (kmerspace >= 100 * n) break;
    kmerspace *= 4;
  }
  return mhl;
}

void init_tax_rank_num(uint8_t r[32]) {
  // rank ids follow the enum of Taxonomy.hpp:25-59
  enum { UNKNOWN = 0, STRAIN, SPECIES, GENUS, FAMILY, ORDER, CLASS, PHYLUM, KINGDOM, DOMAIN_, FORMA,
         INFRA_CLASS, INFRA_ORDER, PARV_ORDER, SUB_CLASS, SUB_FAMILY, SUB_GENUS, SUB_KINGDOM, SUB_ORDER,
         SUB_PHYLUM, SUB_SPECIES, SUB_TRIBE, SUPER_CLASS, SUPER_FAMILY, SUPER_KINGDOM, SUPER_ORDER,
         SUPER_PHYLUM, TRIBE, VARIETAS, LIFE, ACELLULAR_ROOT };
  memset(r, 0, 32);
  uint8_t k = 0;
  r[SUB_SPECIES] = k; r[STRAIN] = k++;
  r[SPECIES] = k++;
  r[SUB_GENUS] = k; r[GENUS] = k++;
  r[SUB_FAMILY] = k; r[FAMILY] = k; r[SUPER_FAMILY] = k++;
  r[SUB_ORDER] = k; r[INFRA_ORDER] = k; r[PARV_ORDER] = k; r[ORDER] = k; r[SUPER_ORDER] = k++;
  r[INFRA_CLASS] = k; r[SUB_CLASS] = k; r[CLASS] = k; r[SUPER_CLASS] = k++;
  r[SUB_PHYLUM] = k; r[PHYLUM] = k; r[SUPER_PHYLUM] = k++;
  r[SUB_KINGDOM] = k; r[KINGDOM] = k++;
  r[SUPER_KINGDOM] = k; r[ACELLULAR_ROOT] = k; r[DOMAIN_] = k++;
  r[FORMA] = k; r[SUB_TRIBE] = k; r[TRIBE] = k; r[VARIETAS] = k; r[LIFE] = k; r[UNKNOWN] = k;
  r[31] = k;  // slot 31: level of "unknown" for ids beyond the table
}

int load_taxonomy_file(const std::string &path, TaxonomyHost &t, std::string &err) { return parse_taxonomy(path, t, err); }

const char *tax_rank_string(uint8_t rank) {
  static const char *names[] = {
      "no rank", "strain", "species", "genus", "family", "order", "class", "phylum", "kingdom",
      "domain", "forma", "infraclass", "infraorder", "parvorder", "subclass", "subfamily",
      "subgenus", "subkingdom", "suborder", "subphylum", "subspecies", "subtribe", "superclass",
      "superfamily", "superkingdom", "superorder", "superphylum", "tribe", "varietas", "life",
      "acellular root"};
  if (rank >= 1 && rank <= 30) return names[rank];
  return "no rank";
}

}  // namespace cfrb200

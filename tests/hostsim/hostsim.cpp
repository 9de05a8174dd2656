// hostsim.cpp -- TEST INFRASTRUCTURE: runs the product's per-task stage functions
// (centrifuger_b200/csrc/cfr_core.cuh, cfr_pipeline.cuh) sequentially on the host
// so the classification logic can be debugged in a container without a GPU.
// It is compiled only by tests/ (g++ -DCFR_HOSTSIM); the shipped library never
// contains it and has no CPU path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../centrifuger_b200/csrc/cfr_format.hpp"
#include "../../centrifuger_b200/csrc/cfr_pipeline.cuh"
#include "../../include/centrifuger_b200.h"

using namespace cfrb200;

namespace {

struct HostIndex {
  CfrIndexFile file;
  std::vector<std::vector<u64>> bufs;
  std::vector<u32> parent, seq_to_tax;
  std::vector<unsigned char> rank;
  std::vector<u64> sel_filter;
  std::vector<OccLine> occ;
  std::vector<PairLine> pairs;  // layout 4: the search stage walks pair lines (two extends per line)
  std::vector<u64> pair_sb;
  bool use_pairs = false;
  std::vector<u64x2> wide;
  std::vector<u32> dense;
  bool pos32 = false;
  DevIndex ix;
  DevParams P;
  int layout;
  std::string err;

  const u64 *copy_words(const uint8_t *p, u64 words) {
    bufs.emplace_back(words + 2, 0);
    if (p && words) memcpy(bufs.back().data(), p, words * 8);
    return bufs.back().data();
  }
  DevBV mk_bv(const BvView &v) {
    DevBV d;
    d.n = v.nbits;
    d.B = v.nbits ? copy_words(v.B, v.words) : nullptr;
    d.R = v.nbits ? copy_words(v.R, v.rwords) : nullptr;
    return d;
  }
  DevWT mk_wt(const WtView &t) {
    DevWT d;
    memset(&d, 0, sizeof(d));
    d.n = t.n;
    for (int i = 0; i < 3; ++i) {
      d.child[i][0] = t.child[i][0];
      d.child[i][1] = t.child[i][1];
      if (i < t.node_cnt) d.node[i] = mk_bv(t.node[i]);
    }
    return d;
  }
};

}  // namespace

extern "C" {

void *hostsim_open(const char *prefix, const cfr_params *p) {
  HostIndex *h = new HostIndex();
  int st = h->file.load(prefix, h->err);
  if (st != 0) {
    fprintf(stderr, "hostsim: %s\n", h->err.c_str());
    delete h;
    return nullptr;
  }
  CfrIndexFile &f = h->file;
  DevIndex &ix = h->ix;
  memset(&ix, 0, sizeof(ix));
  ix.n = f.n;
  ix.first_isa = f.first_isa;
  ix.last_code = base_code((unsigned char)f.last_chr);
  for (int i = 0; i < 5; ++i) ix.C[i] = f.C[i];
  ix.b = f.b;
  ix.block_cnt = f.block_cnt;
  ix.block_type = h->mk_bv(f.block_type);
  ix.plain = h->mk_wt(f.plain);
  ix.run = h->mk_wt(f.run);
  ix.sample_rate = f.sample_rate;
  ix.sample_shift = (f.sample_rate & (f.sample_rate - 1)) == 0 ? __builtin_ctz((unsigned)f.sample_rate) : -1;
  ix.sa_bits = f.sa_bits;
  ix.sampled_sa = h->copy_words(f.sa_w, f.sa_words);
  ix.adjusted_sa0 = f.adjusted_sa0;
  ix.sel = (const u64x2 *)h->copy_words(f.sel, f.sel_cnt * 2);
  ix.sel_cnt = f.sel_cnt;
  ix.sel_filter_rate = f.sel_filter_rate;
  ix.filter_shift = (f.sel_filter_rate & (f.sel_filter_rate - 1)) == 0 ? __builtin_ctz((unsigned)f.sel_filter_rate) : -1;
  if (f.sel_cnt > 0) {
    u64 fbits = (f.n + f.sel_filter_rate - 1) / f.sel_filter_rate;
    h->sel_filter.assign(fbits / 64 + 2, 0);
    for (u64 i = 0; i < f.sel_cnt; ++i) {
      u64 fb = load_u64(f.sel + i * 16) / (u64)f.sel_filter_rate;
      h->sel_filter[fb >> 6] |= 1ull << (fb & 63);
    }
    ix.sel_filter = h->sel_filter.data();
  }
  ix.pre_width = (int)f.precompute_width;
  ix.lookup = (const u64x2 *)h->copy_words(f.lookup, f.precompute_size * 2);
  ix.node_cnt = f.tax.node_cnt;
  ix.seq_cnt = f.tax.seq_cnt;
  ix.root = f.tax.root;
  h->parent.resize(f.tax.node_cnt + 1);
  h->rank.resize(f.tax.node_cnt + 1);
  for (u64 i = 0; i < f.tax.node_cnt; ++i) {
    h->parent[i] = (u32)f.tax.parent[i];
    h->rank[i] = f.tax.rank[i];
  }
  h->seq_to_tax.resize(f.tax.seq_cnt + 1);
  for (u64 i = 0; i < f.tax.seq_cnt; ++i)
    h->seq_to_tax[i] = f.tax.seq_to_tax[i] >= f.tax.node_cnt ? (u32)f.tax.node_cnt : (u32)f.tax.seq_to_tax[i];
  ix.parent = h->parent.data();
  ix.rank = h->rank.data();
  ix.seq_to_tax = h->seq_to_tax.data();
  init_tax_rank_num(ix.rank_num);
  // layout 2 = occ sectors (32-bit positions when the index allows, as the library does), 3 = occ
  // sectors with 64-bit positions forced, anything else = the run-block arrays
  // 4 = occ sectors + pair lines for the search stage
  h->use_pairs = p->layout == 4;
  h->layout = (p->layout == CFR_LAYOUT_OCCLINE || p->layout == 4) ? 2 : (p->layout == 3 ? 3 : 1);
  h->pos32 = h->layout == 2 && f.n < CFR_POS32_MAX_N;
  if (h->layout == 3) h->layout = 2;
  if (h->layout == 2) {  // same construction the transcode kernel performs
    const u64 lines = f.n / 64 + 1;
    h->occ.resize(lines);
    for (u64 L = 0; L < lines; ++L) {
      u64 cnt[3], lo = 0, hi = 0;
      for (int c = 0; c < 3; ++c) cnt[c] = rb_rank(ix, c, L * 64, 0);
      for (int w = 0; w < 64; ++w) {
        const u64 pos = L * 64 + (u64)w;
        if (pos >= f.n) break;
        const int s = rb_access(ix, pos);
        lo |= (u64)(s & 1) << w;
        hi |= (u64)(s >> 1) << w;
      }
      h->occ[L] = occ_pack(lo, hi, cnt[0], cnt[1], cnt[2]);
    }
    ix.occ = h->occ.data();
  }
  if (h->use_pairs) {  // the passes of the library's load-time kernels, run as loops
    const u64 n_lines = f.n / 64 + 1, n_chunk = (n_lines + CFR_PAIR_CHUNK - 1) / CFR_PAIR_CHUNK;
    h->pairs.resize(n_lines);
    std::vector<u64> tot(20 * n_chunk);
    for (u64 c = 0; c < n_chunk; ++c) pair_chunk_planes(ix, h->pairs.data(), n_lines, c, tot.data(), n_chunk);
    for (int k = 0; k < 20; ++k) {
      u64 run = 0;
      for (u64 c = 0; c < n_chunk; ++c) {
        const u64 x = tot[(u64)k * n_chunk + c];
        tot[(u64)k * n_chunk + c] = run;
        run += x;
      }
    }
    const u64 n_sb = ((n_lines - 1) >> CFR_PAIR_SB_SHIFT) + 1;
    h->pair_sb.assign(20 * n_sb, 0);
    for (u64 sb = 0; sb < n_sb; ++sb)
      for (int k = 0; k < 20; ++k) h->pair_sb[sb * 20 + k] = tot[(u64)k * n_chunk + ((sb << CFR_PAIR_SB_SHIFT) / CFR_PAIR_CHUNK)];
    for (u64 c = 0; c < n_chunk; ++c) pair_chunk_counters(ix, h->pairs.data(), n_lines, c, tot.data(), n_chunk, h->pair_sb.data());
    ix.pairs = h->pairs.data();
    ix.pair_sb = h->pair_sb.data();
    u64 k18[18];
    pair_constants(ix, k18);
    for (int i = 0; i < 16; ++i) ix.pair_D[i] = k18[i];
    ix.pair_E = (int)k18[16];
    ix.pair_F = (int)k18[17];
  }
  ix.dense_shift = -1;
  ix.dense_idx_shift = 0;
  ix.dense16 = 0;
  if (const char *e = getenv("HOSTSIM_DENSE_LOCATE")) {  // the library's dense locate table
    const int shift = atoi(e);
    if (shift >= 0 && ix.sample_shift > shift) {
      const u64 n_rows = ((ix.n - 1) >> shift) + 1;
      const char *e16 = getenv("HOSTSIM_DENSE16");  // 16-bit entries (the library picks them when every id fits)
      ix.dense16 = e16 && atoi(e16) != 0 ? 1 : 0;
      h->dense.resize(n_rows + 1);
      OpCount oc{};
      for (u64 j = 0; j < n_rows; ++j) {
        const u32 id = (u32)(h->layout == 2 ? locate_row<BwtOccLine>(ix, j << shift, oc) : locate_row<BwtRunBlock>(ix, j << shift, oc));
        if (ix.dense16) reinterpret_cast<unsigned short *>(h->dense.data())[j] = (unsigned short)id;
        else h->dense[j] = id;
      }
      ix.dense = h->dense.data();
      ix.dense_shift = shift;
      ix.dense_idx_shift = shift;
    }
  }
  if (const char *e = getenv("HOSTSIM_WIDE_LOOKUP")) {  // the library's wide lookup table, any width
    const int ww = atoi(e);
    if (ww > ix.pre_width && ix.pre_width > 0) {
      h->wide.resize(1ull << (2 * ww));
      for (u64 key = 0; key < h->wide.size(); ++key)
        h->wide[key] = h->layout == 2 ? wide_lookup_entry<BwtOccLine>(ix, key, ww) : wide_lookup_entry<BwtRunBlock>(ix, key, ww);
      ix.wide = h->wide.data();
      ix.wide_width = ww;
    }
  }
  h->P.max_result = p->max_result;
  h->P.ids_stride = p->max_result > 0 ? p->max_result : (p->unlimited_cap > 0 ? p->unlimited_cap : 64);
  h->P.min_hit_len = p->min_hit_len > 0 ? p->min_hit_len : infer_min_hit_len(f.n);
  h->P.hitk_factor = p->max_result_per_hit_factor;
  h->P.secondary_len = p->consider_secondary_hit_len;
  h->P.secondary_factor = p->consider_secondary_score_factor;
  h->P.quorum = 8;
  return h;
}

void hostsim_close(void *hh) { delete (HostIndex *)hh; }

// Pair layout against the literal steps: for every boundary x and symbol pair, step(c2, step(c1, x)) from
// the pair line equals two FMIndex::Rank-based steps; and for `n_ranges` pseudo-random ranges (single rows,
// narrow and wide ranges, ranges around firstISA and the ends) extend2 equals two literal
// BackwardExtend calls with their stop tests.  Returns the number of disagreements.
u64 hostsim_pair_check(void *hh, u64 n_ranges, u64 seed) {
  HostIndex *h = (HostIndex *)hh;
  const DevIndex &ix = h->ix;
  if (!ix.pairs) return ~0ull;
  u64 bad = 0;
  OpCount oc{}, oc2{};
  auto literal = [&](int c, u64 sp, u64 ep, u64 &nsp, u64 &nep) {
    BwtOccLine::extend(ix, c, sp, ep, nsp, nep, oc);
    return !(nsp > nep || nep > ix.n);
  };
  auto one = [&](u64 sp, u64 ep) {
    for (int c1 = 0; c1 < 4; ++c1)
      for (int c2 = -1; c2 < 4; ++c2) {
        u64 y1, y2, z1 = 0, z2 = 0;
        int want = 0;
        u64 wsp = sp, wep = ep;
        if (literal(c1, sp, ep, y1, y2)) {
          want = 1;
          wsp = y1;
          wep = y2;
          if (c2 >= 0 && literal(c2, y1, y2, z1, z2)) {
            want = 2;
            wsp = z1;
            wep = z2;
          }
        }
        u64 gsp = sp, gep = ep;
        const int got = BwtPairT<0>::extend2(ix, true, c1, c2, gsp, gep, oc2);
        if (got != want || (want > 0 && (gsp != wsp || gep != wep))) ++bad;
      }
  };
  // every boundary as a single row and as the low end of short ranges
  for (u64 x = 0; x < ix.n; ++x) {
    one(x, x);
    if (x + 1 < ix.n) one(x, x + 1);
    if (x + 5 < ix.n && (x % 7) == 0) one(x, x + 5);
  }
  u64 st = seed * 0x9E3779B97F4A7C15ull + 1;
  auto rnd = [&]() {
    st ^= st << 13;
    st ^= st >> 7;
    st ^= st << 17;
    return st;
  };
  for (u64 i = 0; i < n_ranges; ++i) {
    u64 a = rnd() % ix.n, len = (i & 3) == 0 ? rnd() % ix.n : rnd() % 300;
    if ((i & 15) == 1) a = ix.first_isa > 100 ? ix.first_isa - rnd() % 100 : 0;
    u64 b = a + len;
    if (b >= ix.n) b = ix.n - 1;
    one(a, b);
  }
  one(0, ix.n - 1);
  // the counters advance as the literal calls do
  oc_fold(oc);
  oc_fold(oc2);
  if (oc.rank != oc2.rank || oc.access != oc2.access || oc.extend != oc2.extend) ++bad;
  return bad;
}

// ---- primitives with positions beyond 2^32 (no index of that size exists on the CPU side of the tests; the GPU suite
// loads real ones).  Each function answers ONE query so that the Python test states the expected value itself.
// occ sector: pack the counts of A, C, G before sector `sec` and read the count of symbol c back
uint64_t hostsim_occ_roundtrip(uint64_t a, uint64_t c, uint64_t g, uint64_t sec, int sym) {
  const OccLine o = occ_pack(0x0123456789abcdefull, 0xfedcba9876543210ull, a, c, g);
  return occ_base(o.w2, o.w3, sym, sec);
}
// the 32-bit walker's view of the same sector (valid below 2^32 rows: the high bytes are zero)
uint32_t hostsim_occ_roundtrip32(uint64_t a, uint64_t c, uint64_t g, uint64_t sec, int sym) {
  const OccLine o = occ_pack(0, 0, a, c, g);
  return occ_base32(o.w2, o.w3, sym, (u32)sec);
}
// FixedSizeElemArray::Read at element i of a bit-packed array of `bits`-wide elements (words = the array)
uint64_t hostsim_sa_read(const uint64_t *words, int bits, uint64_t i) {
  DevIndex ix;
  memset(&ix, 0, sizeof(ix));
  ix.sampled_sa = reinterpret_cast<const u64 *>(words);
  ix.sa_bits = bits;
  return sa_read(ix, i);
}
// the row plan of one hit (Classifier.hpp:620-666): out = {step, fwd, total, row(t0), row(t1)}
void hostsim_plan_rows(uint64_t sp, uint64_t ep, int k, int hitk, uint64_t t0, uint64_t t1, uint64_t *out) {
  DevParams P;
  memset(&P, 0, sizeof(P));
  P.max_result = k;
  P.hitk_factor = hitk;
  const RowPlan rp = plan_rows(sp, ep, P);
  out[0] = rp.step;
  out[1] = rp.fwd;
  out[2] = rp.total;
  out[3] = plan_row_at(sp, ep, rp, t0);
  out[4] = plan_row_at(sp, ep, rp, t1);
}
// one occ-sector extend step (FMIndex::BackwardExtend) on a two-sector toy BWT placed at sector `sec0` of a virtual
// index of n rows: the sectors' counts carry the rows before them, so every product of the step has high bits
void hostsim_extend_high(uint64_t sec0, uint64_t n, const uint64_t *C5, uint64_t first_isa, int last_code, uint64_t lo0,
                         uint64_t hi0, uint64_t lo1, uint64_t hi1, const uint64_t *cnt0, int c, uint64_t sp, uint64_t ep,
                         uint64_t *out) {
  // cnt0 = counts of A, C, G before sector sec0; the counts before sec0 + 1 follow from sector sec0's symbols
  std::vector<OccLine> lines(3);
  u64 cnt1[3];
  for (int s = 0; s < 3; ++s) cnt1[s] = cnt0[s] + (u64)popc64(occ_match(lo0, hi0, s));
  lines[0] = occ_pack(lo0, hi0, cnt0[0], cnt0[1], cnt0[2]);
  lines[1] = occ_pack(lo1, hi1, cnt1[0], cnt1[1], cnt1[2]);
  lines[2] = occ_pack(0, 0, 0, 0, 0);
  DevIndex ix;
  memset(&ix, 0, sizeof(ix));
  ix.n = n;
  for (int i = 0; i < 5; ++i) ix.C[i] = C5[i];
  ix.first_isa = first_isa;
  ix.last_code = last_code;
  ix.occ = lines.data() - sec0;  // sector sec0 of the virtual index is lines[0]
  OpCount oc{};
  u64 nsp = 0, nep = 0;
  BwtOccLine::extend(ix, c, sp, ep, nsp, nep, oc);
  out[0] = nsp;
  out[1] = nep;
  BwtOccLineT<4>::extend(ix, c, sp, ep, nsp, nep, oc);
  out[2] = nsp;
  out[3] = nep;
  out[4] = BwtOccLine::lf(ix, sp, oc);
}

int hostsim_min_hit_len(void *hh) { return ((HostIndex *)hh)->P.min_hit_len; }

uint64_t hostsim_bwt_rank(void *hh, int c, uint64_t i, int inclusive) {
  HostIndex *h = (HostIndex *)hh;
  return h->layout == 2 ? BwtOccLine::rank(h->ix, c, i, inclusive) : BwtRunBlock::rank(h->ix, c, i, inclusive);
}
int hostsim_bwt_access(void *hh, uint64_t i) {
  HostIndex *h = (HostIndex *)hh;
  return h->layout == 2 ? BwtOccLine::access(h->ix, i) : BwtRunBlock::access(h->ix, i);
}
uint64_t hostsim_locate(void *hh, uint64_t row) {
  HostIndex *h = (HostIndex *)hh;
  OpCount oc{};
  return h->layout == 2 ? locate_row<BwtOccLine>(h->ix, row, oc) : locate_row<BwtRunBlock>(h->ix, row, oc);
}

// Taxonomy::ReduceTaxIds as the scoring stage runs it (tax_reduce / tax_lca); returns the number of ids
int hostsim_reduce_taxids(void *hh, const uint64_t *tax_ids, int cnt, int k, uint64_t *out) {
  HostIndex *h = (HostIndex *)hh;
  std::vector<u64> scratch(cnt + 1), o(cnt + 1 + (k > 0 ? k : 1));
  u64 err = 0;
  const int n = tax_reduce(h->ix, (const u64 *)tax_ids, cnt, k, scratch.data(), o.data(), &err);
  if (err) return -1;
  for (int i = 0; i < n; ++i) out[i] = o[i];
  return n;
}

// the child lists (tax_expand) for the same ids; child_cnt has max(k,1) slots; returns the total
int hostsim_expand_taxids(void *hh, const uint64_t *tax_ids, int cnt, int k, uint64_t *child, uint32_t *child_cnt) {
  HostIndex *h = (HostIndex *)hh;
  std::vector<u64> scratch(cnt + 1), o(cnt + 1 + (k > 0 ? k : 1)), c(cnt + 1);
  u64 err = 0;
  const int n = tax_reduce(h->ix, (const u64 *)tax_ids, cnt, k, scratch.data(), o.data(), &err);
  const int total = tax_expand(h->ix, (const u64 *)tax_ids, cnt, k, o.data(), n, scratch.data(), c.data(), child_cnt, &err);
  if (err) return -1;
  for (int i = 0; i < total; ++i) child[i] = c[i];
  return total;
}

void hostsim_dust(const char *in, int n, char *out) {
  static DustState d;
  memcpy(out, in, (size_t)n);
  const u64 words = (u64)n / 32 + 2;
  std::vector<unsigned char> raw((size_t)words * 32 + 32, 0);
  memcpy(raw.data(), in, (size_t)n);
  std::vector<u64> codes(words);
  std::vector<u32> mraw(words), mwork(words), dbits(words, 0);
  ChunkDev B;
  memset(&B, 0, sizeof(B));
  B.seq_raw = raw.data();
  B.codes = codes.data();
  B.mask_raw = mraw.data();
  B.mask = mwork.data();
  B.dust_bits = dbits.data();
  for (u64 w = 0; w < words; ++w) encode_stage(B, w, (u64)n);
  DustIn din{B.codes, B.mask_raw, 0};
  const DustOut dout{B.mask, B.dust_bits, 0};
  dust_task(din, n, dout, d);
  for (int i = 0; i < n; ++i)
    if ((dbits[i >> 5] >> (i & 31)) & 1u) out[i] = 'N';
}

// the register-only screen on one read: 1 = the full SDUST must run, 0 = provably nothing is masked
int hostsim_dust_screen(const char *in, int n) {
  const u64 words = (u64)n / 32 + 2;
  std::vector<unsigned char> raw((size_t)words * 32 + 32, 0);
  memcpy(raw.data(), in, (size_t)n);
  std::vector<u64> codes(words);
  std::vector<u32> mraw(words), mwork(words);
  std::vector<u64> off = {0, (u64)n};
  ChunkDev B;
  memset(&B, 0, sizeof(B));
  B.n_reads = 1;
  B.mates = 1;
  B.seq_raw = raw.data();
  B.codes = codes.data();
  B.mask_raw = mraw.data();
  B.mask = mwork.data();
  B.off[0] = off.data();
  for (u64 w = 0; w < words; ++w) encode_stage(B, w, (u64)n);
  return dust_screen_stage(B, 0) ? 1 : 0;
}

// the whole pipeline for one batch; arena_rows small values exercise the deferral loop.
// exp_cnt != NULL asks for the --expand-taxid lists (exp_cnt[n*k], exp_off[n], exp_ids[exp_cap]).
int hostsim_classify_expanded(void *hh, int dust, uint64_t arena_rows, const cfr_read_batch *in, cfr_result *results,
                              uint64_t *ids, cfr_counters *counters, uint32_t *exp_cnt, uint64_t *exp_off,
                              uint64_t *exp_ids, uint64_t exp_cap, uint64_t *exp_n) {
  HostIndex *h = (HostIndex *)hh;
  const DevIndex &ix = h->ix;
  const DevParams &P = h->P;
  const u64 n = in->n_reads;
  const int mates = in->seq2 ? 2 : 1;
  const int S = 2 * mates;
  // pack both mates into one buffer
  const u64 len1 = in->off1[n], len2 = mates == 2 ? in->off2[n] : 0;
  std::vector<unsigned char> raw(len1 + len2 + 64);
  memcpy(raw.data(), in->seq1, len1);
  if (mates == 2) memcpy(raw.data() + len1, in->seq2, len2);
  const u64 n_words = (len1 + len2) / 32 + 2;
  std::vector<u64> codes(n_words);
  std::vector<u32> mask_raw(n_words), mask_work(n_words);
  std::vector<u64> off1(in->off1, in->off1 + n + 1), off2(n + 1, 0);
  if (mates == 2)
    for (u64 i = 0; i <= n; ++i) off2[i] = in->off2[i] + len1;
  int max_len = 0;
  for (u64 i = 0; i < n; ++i) {
    max_len = std::max<int>(max_len, (int)(off1[i + 1] - off1[i]));
    if (mates == 2) max_len = std::max<int>(max_len, (int)(off2[i + 1] - off2[i]));
  }
  ChunkDev B;
  memset(&B, 0, sizeof(B));
  B.n_reads = n;
  B.mates = mates;
  B.cap_h = std::max(1, max_hits_for_len(max_len, P.min_hit_len));
  B.seq_raw = raw.data();
  B.n_words = n_words;
  B.codes = codes.data();
  B.mask_raw = mask_raw.data();
  B.mask = dust ? mask_work.data() : mask_raw.data();
  B.off[0] = off1.data();
  B.off[1] = off2.data();
  std::vector<Hit> strand_hits(n * S * B.cap_h + 1);
  std::vector<int> strand_nhits(n * S + 1);
  std::vector<FinalHit> fhits(n * S * B.cap_h + 1);
  std::vector<ReadWork> work_v(n + 1);
  if (arena_rows == 0) arena_rows = n * 64 + 1024;
  std::vector<u64> rows(arena_rows), best(arena_rows), tmp(arena_rows);
  std::vector<u32> seq_ids(arena_rows);
  std::vector<SeqRec> rec0(arena_rows), rec1(arena_rows);
  std::vector<DevResult> res(n + 1);
  std::vector<u64> out_ids(n * P.ids_stride + 1);
  std::vector<u64> taxon(ix.node_cnt + 3, 0);
  DevCounters cnt;
  memset(&cnt, 0, sizeof(cnt));
  u64 arena_used = 0;
  u32 n_deferred = 0;
  std::vector<u32> deferred(n + 1), list;
  B.strand_hits = strand_hits.data();
  B.strand_nhits = strand_nhits.data();
  B.fhits = fhits.data();
  B.work = work_v.data();
  B.arena_cap = arena_rows;
  B.arena_used = &arena_used;
  B.rows = rows.data();
  B.seq_ids = seq_ids.data();
  B.rec0 = rec0.data();
  B.rec1 = rec1.data();
  B.best = best.data();
  B.tmp = tmp.data();
  B.results = res.data();
  B.out_ids = out_ids.data();
  B.taxon_counts = taxon.data();
  B.counters = &cnt;
  B.deferred = deferred.data();
  B.n_deferred = &n_deferred;
  u64 exp_used = 0;
  if (exp_cnt) {
    B.exp_cnt = exp_cnt;
    B.exp_off = (u64 *)exp_off;
    B.exp_ids = (u64 *)exp_ids;
    B.exp_cap = exp_cap;
    B.exp_used = &exp_used;
  }
  OpCount oc{};
  static DustState ds;
  for (u64 w = 0; w < n_words; ++w) encode_stage(B, w, len1 + len2);
  u64 dust_counter = 0;
  B.dust_counter = &dust_counter;
  std::vector<u32> dust_list(n * mates + 1);
  u32 dust_list_n = 0;
  if (dust && !getenv("HOSTSIM_NO_DUST_SCREEN")) {  // the screen kernel's loop, one lane
    for (u64 t = 0; t < n * mates; ++t)
      if (dust_screen_stage(B, t)) dust_list[dust_list_n++] = (u32)t;
    B.dust_list = dust_list.data();
    B.dust_list_n = &dust_list_n;
  }
  if (dust) dust_tasks(B, B.dust_list ? (u64)dust_list_n : n * mates, ds, P.quorum, true, B.dust_list != nullptr);
  u64 task_counter = 0, row_counter = 0;
  B.task_counter = &task_counter;
  B.row_counter = &row_counter;
  if (h->use_pairs) search_tasks<BwtPairT<0>>(ix, P, B, n * S, oc);
  else if (h->layout == 2 && h->pos32) search_tasks<BwtOccLine32T<0>>(ix, P, B, n * S, oc);
  else if (h->layout == 2) search_tasks<BwtOccLine>(ix, P, B, n * S, oc);
  else search_tasks<BwtRunBlock>(ix, P, B, n * S, oc);
  B.read_list = nullptr;
  B.n_list = n;
  u64 err_flags = 0;
  bool first = true;
  for (;;) {
    arena_used = 0;
    n_deferred = 0;
    u64 arena_valid = ~0ull;
    for (u64 t = 0; t < B.n_list; ++t) {
      const u64 read = chunk_read_id(B, t);
      u32 r;
      if (first) {
        r = h->layout == 2 ? select_plan<BwtOccLine>(ix, P, B, read, oc) : select_plan<BwtRunBlock>(ix, P, B, read, oc);
      } else {
        r = B.work[read].arena_rows;
      }
      const u64 base = arena_used;
      arena_used += r;
      const bool fits = base + r <= B.arena_cap;
      oc.locate += select_write_rows(ix, P, B, read, base, fits);
      if (!fits) {
        deferred[n_deferred++] = (u32)read;
        arena_valid = std::min(arena_valid, base);
      }
    }
    const u64 used = std::min(std::min(arena_used, arena_valid), B.arena_cap);
    row_counter = 0;
    if (locate_in_select(ix)) {
    } else if (h->layout == 2 && h->pos32) locate_rows<BwtOccLine32T<0>>(ix, P, B, used, oc);
    else if (h->layout == 2) locate_rows<BwtOccLine>(ix, P, B, used, oc);
    else locate_rows<BwtRunBlock>(ix, P, B, used, oc);
    for (u64 t = 0; t < B.n_list; ++t) {
      const u64 read = chunk_read_id(B, t);
      if (B.work[read].status != 0) continue;
      score_stage(ix, P, B, read, &err_flags);
    }
    if (n_deferred == 0) break;
    if (n_deferred == B.n_list && B.work[deferred[0]].arena_rows > B.arena_cap) return CFR_ERR_OVERFLOW;
    list.assign(deferred.begin(), deferred.begin() + n_deferred);
    B.read_list = list.data();
    B.n_list = n_deferred;
    first = false;
  }
  if (err_flags) return CFR_ERR_OVERFLOW;
  if (exp_n) *exp_n = exp_used;
  for (u64 i = 0; i < n; ++i) {
    results[i].score = res[i].score;
    results[i].secondary_score = res[i].secondary_score;
    results[i].hit_length = res[i].hit_length;
    results[i].query_length = res[i].query_length;
    results[i].n_assign = res[i].n_assign;
    results[i].by_rank = res[i].by_rank;
    for (int k = 0; k < P.ids_stride; ++k) ids[i * P.ids_stride + k] = out_ids[i * P.ids_stride + k];
  }
  oc_fold(oc);
  if (counters) {
    memset(counters, 0, sizeof(*counters));
    counters->n_rank = oc.rank;
    counters->n_access = oc.access;
    counters->n_search = oc.search;
    counters->n_locate = oc.locate;
    counters->n_lf = oc.lf;
    counters->n_extend = oc.extend;
    counters->n_reads = n;
  }
  return 0;
}

int hostsim_classify(void *hh, int dust, uint64_t arena_rows, const cfr_read_batch *in, cfr_result *results,
                     uint64_t *ids, cfr_counters *counters) {
  return hostsim_classify_expanded(hh, dust, arena_rows, in, results, ids, counters, nullptr, nullptr, nullptr, 0, nullptr);
}

}  // extern "C"

// cfr_core.cuh -- the classification hot path as per-task device functions.
//
// Everything here is written once and compiled twice: by nvcc for sm_100a (the
// product) and by g++ with -DCFR_HOSTSIM for the host-side simulation harness
// under tests/hostsim (a debugging twin; never part of the shipped library).
//
// Reference map (all file:line relative to the reference tree):
//   rank9 query                DS_Rank.hpp:255-273          -> bv_rank1 / bv_rank_bit
//   Bitvector::Rank0/Rank      Bitvector.hpp:45-57          -> bv_rank
//   wavelet Rank/RankAndTest   Sequence_WaveletTree.hpp:235-293 -> wt_rank / wt_rank_test
//   wavelet Access             Sequence_WaveletTree.hpp:215-232 -> wt_access
//   run-block Rank / Access    Sequence_RunBlock.hpp:378-416 / :360-376 -> rb_rank / rb_access
//   FMIndex::Rank              FMIndex.hpp:352-362          -> lastChr fix in *_extend / *_lf
//   BackwardExtend             FMIndex.hpp:364-386          -> Bwt::extend / Bwt::lf
//   BackwardSearch             FMIndex.hpp:388-422,487-510  -> backward_search
//   GetSampledSA / locate      FMIndex.hpp:203-231,514-524  -> locate_row
//   GetHitsFromRead            Classifier.hpp:274-293       -> search_tasks (cfr_pipeline.cuh)
//   AdjustHitBoundary...       Classifier.hpp:303-401       -> adjust_hit_boundary
//   SearchForwardAndReverse    Classifier.hpp:509-583       -> select_task
//   GetClassificationFromHits  Classifier.hpp:585-843       -> select_task (row plan) + score_task
//   Taxonomy::LCA/ReduceTaxIds Taxonomy.hpp:733-973         -> tax_lca / tax_reduce
//   Dustmasker                 Dustmasker.hpp:93-421        -> dust_task
#pragma once
#include "cfr_types.h"

namespace cfrb200 {

// ---------------------------------------------------------------------------
// low-level helpers
// ---------------------------------------------------------------------------
CFR_HD int popc64(u64 x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}

CFR_HD u64 ld64(const u64 *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

CFR_HD u32 ld32(const u32 *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

CFR_HD u32 ld16(const unsigned short *p) {
#if defined(__CUDA_ARCH__)
  return (u32)__ldg(p);
#else
  return (u32)*p;
#endif
}

CFR_HD unsigned char ld8(const unsigned char *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

CFR_HD u64x2 ld128(const u64x2 *p) {
#if defined(__CUDA_ARCH__)
  ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
  u64x2 r;
  r.x = v.x;
  r.y = v.y;
  return r;
#else
  return *p;
#endif
}

// per-thread operation counters (flushed once per task batch)
struct OpCount {
  u32 rank, access, search, locate, lf, extend, bases;
  // occ-sector extends are counted as (steps, single-row steps) in the hot loop and folded into
  // rank / access / extend by oc_fold: a range step is two ranks, a single-row step one rank + one access
  u32 xext, xsingle;
};

CFR_HD void oc_fold(OpCount &oc) {
  oc.extend += oc.xext;
  oc.rank += 2u * oc.xext - oc.xsingle;
  oc.access += oc.xsingle;
  oc.xext = oc.xsingle = 0;
}

// base byte -> code: A,C,G,T -> 0..3, anything else (incl. lowercase) -> 4
CFR_HD int base_code(unsigned char c) {
  return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
}

// ---------------------------------------------------------------------------
// Reads in HBM: 2-bit codes + an "N" bit per base
// ---------------------------------------------------------------------------
// The uploaded bytes are re-coded once per batch (k_encode): base q of the batch
// buffer lives in bits [2(q&31), 2(q&31)+2) of codes[q>>5] and in bit (q&31) of
// nmask[q>>5].  Any byte outside "ACGT" (lower case, IUPAC, N ...) sets the N bit:
// on this path such bytes only ever stop a match (FMIndex.hpp:396,500), complement
// to 'N' (Classifier.hpp:846-856) and form DUST's fifth symbol (Dustmasker.hpp:
// 298-302), so the recoding loses nothing.  DUST masking = setting N bits.

CFR_HD u32 brev32(u32 x) {
#if defined(__CUDA_ARCH__)
  return __brev(x);
#else
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
  return (x >> 16) | (x << 16);
#endif
}

CFR_HD int clz32(u32 x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}

CFR_HD int ctz32(u32 x) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return x ? __builtin_ctz(x) : -1;
#endif
}

// 32 bases -> one code word + one N-mask word (the body of k_encode)
CFR_HD void encode_bases(const unsigned char *b, int n, u64 &codes, u32 &nmask) {
  codes = 0;
  nmask = 0;
  for (int i = 0; i < n; ++i) {
    const int c = base_code(b[i]);
    if (c > 3) nmask |= 1u << i; else codes |= (u64)c << (2 * i);
  }
}

// `cnt` (<= 16) consecutive bases starting at batch position q: their 2-bit codes
// (little endian) and N bits
CFR_HD void packed_field(const u64 *codes, const u32 *nmask, u64 q, int cnt, u32 &field, u32 &nbits) {
  const u64 wi = q >> 5;
  const int sh = (int)(q & 31);
  u64 c = ld64(codes + wi) >> (2 * sh);
  u32 m = ld32(nmask + wi) >> sh;
  if (sh + cnt > 32) {
    c |= ld64(codes + wi + 1) << (64 - 2 * sh);
    m |= ld32(nmask + wi + 1) << (32 - sh);
  }
  field = (u32)c & (cnt >= 16 ? 0xffffffffu : ((1u << (2 * cnt)) - 1u));
  nbits = m & ((1u << cnt) - 1u);
}

// One strand of one read as the backward search sees it.  rc strands are never
// materialised: rc[p] = comp(r[len-1-p]) (Classifier.hpp:99-111,846-856).
struct StrandSeq {
  const u64 *codes;
  const u32 *nmask;
  u64 base;  // batch position of the mate's first base
  int len;
  int rc;
  u64 cw = 0;
  u32 mw = 0;
  u64 widx = ~0ull;
  // base p of the strand: 0..3, or 4 for anything that is not ACGT
  CFR_HD int operator()(int p) {
    const u64 q = base + (u64)(rc ? len - 1 - p : p);
    const u64 wi = q >> 5;
    if (wi != widx) {
      cw = ld64(codes + wi);
      mw = ld32(nmask + wi);
      widx = wi;
    }
    const int sh = (int)(q & 31);
    if ((mw >> sh) & 1u) return 4;
    const int c = (int)((cw >> (2 * sh)) & 3ull);
    return rc ? 3 - c : c;
  }
  // ---- sequential cursor for the extend loop: a backward search reads strand positions
  // p, p-1, p-2, ... ; in batch coordinates that is a walk by -1 (as read) or +1 (reverse complement)
  int csh = 0;  // base index of the cursor inside the cached word
  CFR_HD void seek(int p) {
    const u64 q = base + (u64)(rc ? len - 1 - p : p);
    widx = q >> 5;
    csh = (int)(q & 31);
    cw = ld64(codes + widx);
    mw = ld32(nmask + widx);
  }
  CFR_HD int peek() const {  // 0..3, or 4 for anything that is not ACGT
    const int c = (int)((cw >> (2 * csh)) & 3ull) ^ (rc ? 3 : 0);
    return ((mw >> csh) & 1u) ? 4 : c;
  }
  CFR_HD void advance() {  // to the next lower strand position (which must exist)
    csh += rc ? 1 : -1;
    if (csh & ~31) {
      widx += rc ? 1ull : ~0ull;
      csh &= 31;
      cw = ld64(codes + widx);
      mw = ld32(nmask + widx);
    }
  }
  // GetBackwardSearchInitialRange's loop (FMIndex.hpp:394-403) over the last W
  // bases of strand[0..m): true + the table key, or false + the number of valid
  // characters seen before the first non-ACGT one.
  CFR_HD bool init_key(int m, int W, u64 &key, int &nvalid) const {
    // one code path for both strands (the lanes of a warp mix them): the W bases are the same
    // packed field either way; the reverse complement reverses the 2-bit groups and complements
    u32 field, nbits;
    packed_field(codes, nmask, base + (u64)(rc ? len - m : m - W), W, field, nbits);
    if (nbits) {
      nvalid = rc ? ctz32(nbits) : W - 1 - (31 - clz32(nbits));
      return false;
    }
    u32 r = brev32(field);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    r >>= (32 - 2 * W);
    r ^= (W >= 16 ? 0xffffffffu : ((1u << (2 * W)) - 1u));
    // as read: s[m-1-i] lands in bits 2(W-1-i), exactly the little-endian field
    key = rc ? (u64)r : (u64)field;
    return true;
  }
};

// ---------------------------------------------------------------------------
// Layout 1: the reference's run-block BWT arrays as stored
// ---------------------------------------------------------------------------

// DS_Rank9::Query (inclusive) fused with Bitvector_Plain::Access on the same word
CFR_HD u64 bv_rank_bit(const DevBV &v, u64 i, int &bit) {
  const u64 wb = ld64(v.B + (i >> 6));
  bit = (int)((wb >> (i & 63)) & 1ull);
  const u64 wi = i >> 6;
  const u64x2 r = ld128(reinterpret_cast<const u64x2 *>(v.R) + (wi >> 3));
  const u64 t = (wi & 7) - 1;
  return r.x + ((r.y >> ((t + ((t >> 60) & 8)) * 9)) & 0x1ff) + (u64)popc64(wb & ((2ull << (i & 63)) - 1ull));
}

CFR_HD u64 bv_rank1(const DevBV &v, u64 i) {
  if (i >= v.n) i = v.n - 1;  // DS_Rank.hpp:259-260
  int bit;
  return bv_rank_bit(v, i, bit);
}

// Bitvector::Rank(type, i) inclusive; note Rank0 uses the UNclamped i (Bitvector.hpp:45-49)
CFR_HD u64 bv_rank(const DevBV &v, int type, u64 i) {
  const u64 r1 = bv_rank1(v, i);
  return type ? r1 : i + 1 - r1;
}

CFR_HD int bv_access(const DevBV &v, u64 i) { return (int)((ld64(v.B + (i >> 6)) >> (i & 63)) & 1ull); }

// Sequence_WaveletTree::Rank (inclusive), sigma = 4, code bits MSB first
CFR_HD u64 wt_rank(const DevWT &t, int c, u64 i) {
  const int b1 = c >> 1;
  i = bv_rank(t.node[0], b1, i);
  if (i == 0) return 0;
  --i;
  return bv_rank(t.node[t.child[0][b1]], c & 1, i);
}

// Sequence_WaveletTree::RankAndTest
CFR_HD u64 wt_rank_test(const DevWT &t, int c, u64 i, bool &is_c) {
  const int b1 = c >> 1, b0 = c & 1;
  is_c = true;
  if (b1 != bv_access(t.node[0], i)) is_c = false;
  i = bv_rank(t.node[0], b1, i);
  if (i == 0) return 0;
  --i;
  const DevBV &leaf = t.node[t.child[0][b1]];
  if (is_c && b0 != bv_access(leaf, i)) is_c = false;
  return bv_rank(leaf, b0, i);
}

// Sequence_WaveletTree::Access
CFR_HD int wt_access(const DevWT &t, u64 i) {
  int b1;
  const u64 r1 = bv_rank_bit(t.node[0], i, b1);
  i = (b1 ? r1 : i + 1 - r1) - 1;
  return (b1 << 1) | bv_access(t.node[t.child[0][b1]], i);
}

// Sequence_RunBlock::Rank
CFR_HD u64 rb_rank(const DevIndex &ix, int c, u64 i, int inclusive) {
  if (!inclusive) {
    if (i == 0) return 0;
    --i;
  }
  const u64 bi = i / ix.b;
  const u64 ib = i - bi * ix.b;
  int type;
  u64 ranki;
  if (ix.b < ix.n) {
    if (bi < ix.block_type.n) {
      const u64 r1 = bv_rank_bit(ix.block_type, bi, type);
      ranki = type ? r1 : bi + 1 - r1;
    } else {
      type = bv_access(ix.block_type, bi);
      ranki = bv_rank(ix.block_type, type, bi);
    }
  } else {
    type = bv_access(ix.block_type, bi);
    ranki = 1;
  }
  const u64 other = (bi + 1) - ranki;
  u64 ret;
  if (type == 0) {
    ret = wt_rank(ix.plain, c, (ranki - 1) * ix.b + ib);
  } else {
    bool in_run;
    const u64 rb = wt_rank_test(ix.run, c, ranki - 1, in_run);
    ret = in_run ? (rb - 1) * ix.b + ib + 1 : rb * ix.b;
  }
  if (other == 0) return ret;
  if (type == 0)
    ret += wt_rank(ix.run, c, other - 1) * ix.b;
  else
    ret += wt_rank(ix.plain, c, other * ix.b - 1);
  return ret;
}

// Sequence_RunBlock::Access
CFR_HD int rb_access(const DevIndex &ix, u64 i) {
  const u64 bi = i / ix.b;
  int type;
  const u64 r1 = bv_rank_bit(ix.block_type, bi, type);
  if (type == 0) {
    return wt_access(ix.plain, i - ix.b * r1);  // r1 = Rank(1, bi)
  } else {
    const u64 r0 = bi + 1 - r1;  // Rank(0, bi)
    return wt_access(ix.run, (i - ix.b * r0) / ix.b);
  }
}

// FMIndex::Rank's correction for the missing '$' (FMIndex.hpp:359)
CFR_HD u64 last_chr_fix(const DevIndex &ix, int c, u64 p, int inclusive) {
  return (c == ix.last_code && (p < ix.first_isa || (!inclusive && p == ix.first_isa))) ? 1ull : 0ull;
}

struct BwtRunBlock {
  typedef u64 pos_t;
  // FMIndex::BackwardExtend (range form), FMIndex.hpp:364-379
  static CFR_HD void extend(const DevIndex &ix, int c, u64 sp, u64 ep, u64 &nsp, u64 &nep, OpCount &oc) {
    const u64 off = ix.C[c];
    ++oc.extend;
    ++oc.rank;
    nsp = off + rb_rank(ix, c, sp, 0) + last_chr_fix(ix, c, sp, 0);
    if (sp != ep) {
      ++oc.rank;
      nep = off + rb_rank(ix, c, ep, 1) + last_chr_fix(ix, c, ep, 1) - 1;
    } else {
      ++oc.access;
      nep = nsp + ((rb_access(ix, ep) == c) ? 0ull : ~0ull);
    }
  }
  enum { STEPS_COUNTED_AT_CLOSE = 0 };
  static CFR_HD void extend_step(const DevIndex &ix, int c, u64 sp, u64 ep, u64 &nsp, u64 &nep, OpCount &oc) {
    extend(ix, c, sp, ep, nsp, nep, oc);
  }
  // one LF step: i -> C[c] + Rank(c, i) - 1 with c = BWT[i]  (FMIndex.hpp:382-386,520)
  static CFR_HD u64 lf(const DevIndex &ix, u64 i, OpCount &oc) {
    ++oc.access;
    ++oc.rank;
    const int c = rb_access(ix, i);
    return ix.C[c] + rb_rank(ix, c, i, 1) + last_chr_fix(ix, c, i, 1) - 1;
  }
  static CFR_HD u64 rank(const DevIndex &ix, int c, u64 i, int inclusive) { return rb_rank(ix, c, i, inclusive); }
  static CFR_HD int access(const DevIndex &ix, u64 i) { return rb_access(ix, i); }
  static CFR_HD bool leader() { return true; }
  enum { LANES = 1, PAIR = 0 };
};

// ---------------------------------------------------------------------------
// Layout 2: 32-byte occ sectors (64 symbols + the counts of A, C, G before them)
// ---------------------------------------------------------------------------

// occ(c, x) = # of symbol c in BWT[0..x), x in [0, n]: ONE 32-byte sector.
struct OccRank {
  u64 count;  // occ(c, x)
  u64 lo, hi; // planes of the sector that holds position x
};

CFR_HD u64 occ_match(u64 lo, u64 hi, int c) {
  const u64 ml = (c & 1) ? ~0ull : 0ull, mh = (c & 2) ? ~0ull : 0ull;
  return ~(lo ^ ml) & ~(hi ^ mh);
}

// # of symbol c before sector `sec` from its packed counters (branch-free: the four
// symbols of a warp's lanes differ, a switch on c would serialise them)
CFR_HD u64 occ_base(u64 w2, u64 w3, int c, u64 sec) {
  const u32 x3 = (u32)(w3 >> 32);
  const u64 a = (w2 & 0xffffffffull) | ((u64)(x3 & 0xffu) << 32);
  const u64 cc = (w2 >> 32) | ((u64)((x3 >> 8) & 0xffu) << 32);
  const u64 g = (w3 & 0xffffffffull) | ((u64)((x3 >> 16) & 0xffu) << 32);
  const u64 t = (sec << 6) - (a + cc + g);
  const u64 r01 = (c & 1) ? cc : a;
  const u64 r23 = (c & 1) ? t : g;
  return (c & 2) ? r23 : r01;
}

CFR_HD OccLine occ_pack(u64 lo, u64 hi, u64 a, u64 c, u64 g) {
  OccLine o;
  o.lo = lo;
  o.hi = hi;
  o.w2 = (a & 0xffffffffull) | (c << 32);
  o.w3 = (g & 0xffffffffull) | (((a >> 32) & 0xffull) << 32) | (((c >> 32) & 0xffull) << 40) | (((g >> 32) & 0xffull) << 48);
  return o;
}

// How a sector is fetched (LOAD template parameter of BwtOccLineT):
//   0  two 128-bit loads through the read-only (nc) path
//   1  two 128-bit plain loads (L1 + L2)
//   2  two 128-bit loads that bypass L1 (.cg)
//   3  one 256-bit plain load           4  one 256-bit load through the read-only path
template <int LOAD>
CFR_HD void occ_load(const OccLine *L, u64 &lo, u64 &hi, u64 &w2, u64 &w3) {
#if defined(__CUDA_ARCH__)
  if (LOAD == 1) {
    asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(L));
    asm volatile("ld.global.v2.u64 {%0,%1}, [%2+16];" : "=l"(w2), "=l"(w3) : "l"(L));
    return;
  }
  if (LOAD == 2) {
    asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(L));
    asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2+16];" : "=l"(w2), "=l"(w3) : "l"(L));
    return;
  }
  if (LOAD == 3) {
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(lo), "=l"(hi), "=l"(w2), "=l"(w3) : "l"(L));
    return;
  }
  if (LOAD == 4) {
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(lo), "=l"(hi), "=l"(w2), "=l"(w3) : "l"(L));
    return;
  }
#endif
  const u64x2 p = ld128(reinterpret_cast<const u64x2 *>(L));
  const u64x2 q = ld128(reinterpret_cast<const u64x2 *>(L) + 1);
  lo = p.x;
  hi = p.y;
  w2 = q.x;
  w3 = q.y;
}

// the same fetch, issued only by the lanes with `pred` (the other lanes keep the zeros passed in);
// predicated inside the asm block so that the compiler cannot tie it to an earlier load
template <int LOAD>
CFR_HD void occ_load_if(const OccLine *L, bool pred, u64 &lo, u64 &hi, u64 &w2, u64 &w3) {
#if defined(__CUDA_ARCH__)
  const u32 p = pred ? 1u : 0u;
  if (LOAD == 4) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4]; }"
                 : "+l"(lo), "+l"(hi), "+l"(w2), "+l"(w3) : "l"(L), "r"(p));
  } else {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q ld.global.nc.v2.u64 {%0,%1}, [%4]; @q ld.global.nc.v2.u64 {%2,%3}, [%4+16]; }"
                 : "+l"(lo), "+l"(hi), "+l"(w2), "+l"(w3) : "l"(L), "r"(p));
  }
#else
  if (pred) occ_load<LOAD>(L, lo, hi, w2, w3);
#endif
}

template <int LOAD>
CFR_HD OccRank occ_rank(const DevIndex &ix, int c, u64 x) {
  const u64 sec = x >> 6;
  const int s = (int)(x & 63);
  OccRank r;
  u64 w2, w3;
  occ_load<LOAD>(ix.occ + sec, r.lo, r.hi, w2, w3);
  r.count = occ_base(w2, w3, c, sec) + (u64)popc64(occ_match(r.lo, r.hi, c) & ((1ull << s) - 1ull));
  return r;
}

CFR_HD int occ_symbol(u64 lo, u64 hi, int s) { return (int)(((lo >> s) & 1ull) | (((hi >> s) & 1ull) << 1)); }

template <int LOAD>
struct BwtOccLineT {
  typedef u64 pos_t;
  // straight-line: both ranks and the symbol test are always computed, the range /
  // single-row forms of FMIndex::BackwardExtend are selected at the end.  Returns sp != ep.
  static CFR_HD bool extend_core(const DevIndex &ix, int c, u64 sp, u64 ep, u64 &nsp, u64 &nep) {
    const u64 off = ix.C[c];
    const bool range = sp != ep;
    const u64 xe = ep + 1;
    const u64 sa = sp >> 6, se = xe >> 6;
    u64 alo, ahi, aw2, aw3, elo = 0, ehi = 0, ew2 = 0, ew3 = 0;
    occ_load<LOAD>(ix.occ + sa, alo, ahi, aw2, aw3);  // Rank(c, sp, exclusive)
    occ_load<LOAD>(ix.occ + se, elo, ehi, ew2, ew3);  // Rank(c, ep, inclusive); (predicating it on se != sa, as the
                                                      // 32-bit walker does, measured slower here: 1.32 vs 1.24 ms)
    const u64 a_count = occ_base(aw2, aw3, c, sa) + (u64)popc64(occ_match(alo, ahi, c) & ((1ull << (sp & 63)) - 1ull));
    const u64 e_count = occ_base(ew2, ew3, c, se) + (u64)popc64(occ_match(elo, ehi, c) & ((1ull << (xe & 63)) - 1ull));
    const int sym = occ_symbol(alo, ahi, (int)(ep & 63));  // BWT[ep] when sp == ep (same sector as sp)
    // FMIndex::Rank's correction for the missing '$' (FMIndex.hpp:359), both forms at once
    const u64 is_last = c == ix.last_code ? 1ull : 0ull;
    nsp = off + a_count + (is_last & (sp <= ix.first_isa ? 1ull : 0ull));
    const u64 nep_range = off + e_count + (is_last & (ep < ix.first_isa ? 1ull : 0ull)) - 1;
    const u64 nep_single = nsp + ((sym == c) ? 0ull : ~0ull);
    nep = range ? nep_range : nep_single;
    return range;
  }
  static CFR_HD void extend(const DevIndex &ix, int c, u64 sp, u64 ep, u64 &nsp, u64 &nep, OpCount &oc) {
    ++oc.xext;
    oc.xsingle += extend_core(ix, c, sp, ep, nsp, nep) ? 0u : 1u;
  }
  // the search state machine's form: the number of steps of a search is added when the search
  // closes (search_tasks), only the single-row steps are counted here
  enum { STEPS_COUNTED_AT_CLOSE = 1 };
  static CFR_HD void extend_step(const DevIndex &ix, int c, u64 sp, u64 ep, u64 &nsp, u64 &nep, OpCount &oc) {
    oc.xsingle += extend_core(ix, c, sp, ep, nsp, nep) ? 0u : 1u;
  }
  static CFR_HD u64 lf(const DevIndex &ix, u64 i, OpCount &oc) {
    ++oc.access;
    ++oc.rank;
    const u64 sec = i >> 6;
    const int s = (int)(i & 63);
    u64 lo, hi, w2, w3;
    occ_load<LOAD>(ix.occ + sec, lo, hi, w2, w3);
    const int c = occ_symbol(lo, hi, s);
    const u64 r = occ_base(w2, w3, c, sec) + (u64)popc64(occ_match(lo, hi, c) & ((1ull << s) - 1ull));
    // inclusive rank at i = exclusive count at i, plus the symbol itself
    return ix.C[c] + r + 1 + last_chr_fix(ix, c, i, 1) - 1;
  }
  static CFR_HD u64 rank(const DevIndex &ix, int c, u64 i, int inclusive) {
    return occ_rank<LOAD>(ix, c, inclusive ? i + 1 : i).count;
  }
  static CFR_HD int access(const DevIndex &ix, u64 i) {
    const u64x2 p = ld128(reinterpret_cast<const u64x2 *>(ix.occ + (i >> 6)));
    return occ_symbol(p.x, p.y, (int)(i & 63));
  }
  static CFR_HD bool leader() { return true; }
  enum { LANES = 1, PAIR = 0 };
};

typedef BwtOccLineT<0> BwtOccLine;

// The same sectors walked with 32-bit positions, for collections below 2^32 - 16 BWT rows (the
// high count bytes of every sector are zero then): BackwardExtend and LF are a chain of
// position arithmetic, and half-width positions shorten it and free registers for more
// resident warps.  Wrap-around plays the role it plays in the reference's size_t arithmetic:
// "C + rank - 1" wraps to 2^32 - 1 > n exactly where the 64-bit form wraps past n
// (FMIndex.hpp:371-378,503).
CFR_HD u32 occ_base32(u64 w2, u64 w3, int c, u32 sec) {
  const u32 a = (u32)w2, cc = (u32)(w2 >> 32), g = (u32)w3;
  const u32 t = (sec << 6) - (a + cc + g);
  const u32 r01 = (c & 1) ? cc : a;
  const u32 r23 = (c & 1) ? t : g;
  return (c & 2) ? r23 : r01;
}

template <int LOAD>
struct BwtOccLine32T {
  typedef u32 pos_t;
  static CFR_HD bool extend_core(const DevIndex &ix, int c, u32 sp, u32 ep, u32 &nsp, u32 &nep) {
    const u32 off = (u32)ix.C[c];
    const bool range = sp != ep;
    const u32 xe = ep + 1;
    const u32 sa = sp >> 6, se = xe >> 6;
    u64 alo, ahi, aw2, aw3, elo = 0, ehi = 0, ew2 = 0, ew3 = 0;
    occ_load<LOAD>(ix.occ + sa, alo, ahi, aw2, aw3);  // Rank(c, sp, exclusive)
    // Rank(c, ep, inclusive): ep + 1 mostly falls into the sector just requested; a second request
    // for it right behind the first one is not merged by L1 and goes to L2 again
    const bool other = se != sa;
    occ_load_if<LOAD>(ix.occ + se, other, elo, ehi, ew2, ew3);
    elo = other ? elo : alo;
    ehi = other ? ehi : ahi;
    ew2 = other ? ew2 : aw2;
    ew3 = other ? ew3 : aw3;
    const u32 ca = occ_base32(aw2, aw3, c, sa) + (u32)popc64(occ_match(alo, ahi, c) & ((1ull << (sp & 63)) - 1ull));
    const u32 ce = occ_base32(ew2, ew3, c, se) + (u32)popc64(occ_match(elo, ehi, c) & ((1ull << (xe & 63)) - 1ull));
    const int sym = occ_symbol(alo, ahi, (int)(ep & 63));  // BWT[ep] when sp == ep (same sector as sp)
    const u32 is_last = c == ix.last_code ? 1u : 0u;
    const u32 fisa = (u32)ix.first_isa;
    nsp = off + ca + (is_last & (sp <= fisa ? 1u : 0u));
    const u32 nep_range = off + ce + (is_last & (ep < fisa ? 1u : 0u)) - 1u;
    const u32 nep_single = nsp + ((sym == c) ? 0u : ~0u);
    nep = range ? nep_range : nep_single;
    return range;
  }
  enum { STEPS_COUNTED_AT_CLOSE = 1 };
  static CFR_HD void extend_step(const DevIndex &ix, int c, u32 sp, u32 ep, u32 &nsp, u32 &nep, OpCount &oc) {
    oc.xsingle += extend_core(ix, c, sp, ep, nsp, nep) ? 0u : 1u;
  }
  static CFR_HD u32 lf(const DevIndex &ix, u32 i, OpCount &oc) {
    ++oc.access;
    ++oc.rank;
    const u32 sec = i >> 6;
    const int s = (int)(i & 63);
    u64 lo, hi, w2, w3;
    occ_load<LOAD>(ix.occ + sec, lo, hi, w2, w3);
    const int c = occ_symbol(lo, hi, s);
    const u32 r = occ_base32(w2, w3, c, sec) + (u32)popc64(occ_match(lo, hi, c) & ((1ull << s) - 1ull));
    // inclusive rank at i = exclusive count at i, plus the symbol itself; FMIndex::Rank's lastChr fix
    return (u32)ix.C[c] + r + ((c == ix.last_code && i < (u32)ix.first_isa) ? 1u : 0u);
  }
  static CFR_HD bool leader() { return true; }
  enum { LANES = 1, PAIR = 0 };
};

// ---------------------------------------------------------------------------
// Layout 3 (search kernel only): 128-byte PAIR lines, two BackwardExtend steps per DRAM line
// ---------------------------------------------------------------------------
// For an index that lives in HBM every rank is one DRAM line fill whatever its size (128 bytes are
// filled for a 32-byte sector), and the chip sustains a fixed number of such requests per second
// (DESIGN.md section 3).  A line that four adjacent lanes fetch as ONE coalesced request costs the
// same as a sector, so the line is spent on the second-order FM index: besides the row's own symbol
// it holds the symbol of the row it maps to, and counters of the 16 symbol pairs.  From the line at
// a boundary x both
//     step(c1, x)            = C[c1] + occ(c1, x) + [c1 == lastChr && x <= firstISA]      (FMIndex.hpp:352-379)
//     step(c2, step(c1, x))  = C[c2] + D[c1][c2] + P(c1, c2, x) + corrections            (derivation below)
// follow, i.e. two steps of FMIndex::BackwardSearch's loop with their own stop tests
// (FMIndex.hpp:495-508) for one request per boundary.
//
// Derivation of the second step.  step(c2, y) with y = step(c1, x) needs occ(c2, y) = occ(c2, C[c1]) +
// #{j in [C[c1], y): B[j] = c2}.  The rows [C[c1], y) are the images img(i) = step(c1, i) of the rows
// i < x with B[i] = c1, one each -- except that for c1 == lastChr (i) row C[c1] itself is the image of
// the virtual '$' row, not of a stored row: its symbol E is added; (ii) the stored row firstISA (which
// holds lastChr in place of '$') shares its image with the next c1 row, or maps just past y: once
// x > firstISA its code2 F is taken out again.  tests/test_hostsim.py checks the identity for every
// boundary and pair of the test indexes against the literal two steps.
struct PairQuery {
  u64 s1;    // occ(c1, x), absolute
  u64 p;     // P(c1, c2, x), absolute
  int sym1;  // code1 of row x (meaningful for x < n)
  int sym2;  // code2 of row x
};

CFR_HD u64 pair_match(u64 lo, u64 hi, int c) { return occ_match(lo, hi, c); }

// what one boundary x = 64 L + s contributes, read from the whole line by one lane (host twin, load-time
// kernels, and the reference the cooperative form is checked against)
CFR_HD PairQuery pair_query_scalar(const DevIndex &ix, int c1, int c2, u64 x) {
  const u64 L = x >> 6;
  const int s = (int)(x & 63);
  const int idx = c1 * 4 + c2;
  const u64 *sb = ix.pair_sb + (L >> CFR_PAIR_SB_SHIFT) * 20;
  const u64 *w = ix.pairs[L].w;
  const u64 a = ld64(w), b = ld64(w + 1), c = ld64(w + 2), d = ld64(w + 3);
  const u64 below = (1ull << s) - 1ull;
  const u64 m1 = pair_match(a, b, c1), m12 = m1 & pair_match(c, d, c2);
  const u64 pw = ld64(w + 4 + (idx >> 1)), sw = ld64(w + 12 + (c1 >> 1));
  PairQuery q;
  q.s1 = ld64(sb + 16 + c1) + (u64)(u32)((c1 & 1) ? (sw >> 32) : sw) + (u64)popc64(m1 & below);
  q.p = ld64(sb + idx) + (u64)(u32)((idx & 1) ? (pw >> 32) : pw) + (u64)popc64(m12 & below);
  q.sym1 = (int)(((a >> s) & 1ull) | (((b >> s) & 1ull) << 1));
  q.sym2 = (int)(((c >> s) & 1ull) | (((d >> s) & 1ull) << 1));
  return q;
}

#if defined(__CUDA_ARCH__)
// The cooperative fetch: lines are STAGED IN SHARED MEMORY.  Every lane of the warp owns one search (one lane
// per strand task, as in the sector walkers) and names the line it needs; the lines are fetched by groups
// of four adjacent lanes in four rounds -- in round j the group fetches the line of its member j as ONE
// coalesced 128-byte request (32 bytes per lane) and stores it into member j's slot.  After the four
// rounds every lane reads its own line from its slot and does the arithmetic once, 32 searches per
// instruction.  Slots are 144 bytes apart (128 + 16 of padding): a lane's 16-byte reads of its own slot
// fall on distinct banks within each quarter warp.
#define CFR_PAIR_SLOT_WORDS 36
#define CFR_PAIR_NO_LINE 0xffffffffu
#define CFR_PAIR_FAR 8  // far slots per warp (MODE 3)

// The same staging with asynchronous copies (cp.async, 16 bytes per lane): groups of EIGHT adjacent lanes fetch
// one line per round as one coalesced 128-byte request, eight rounds, and no round waits for the one
// before it -- all 32 lines of the warp are in flight together, without passing through registers.
CFR_D void pair_stage_warp_async(const DevIndex &ix, u32 L, u32 *slots) {
  const int lane = threadIdx.x & 31, sub = lane & 7, gbase = lane & ~7;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const u32 lj = __shfl_sync(0xffffffffu, L, gbase + j);
    if (lj != CFR_PAIR_NO_LINE) {  // uniform over the eight lanes of the group
      const char *src = reinterpret_cast<const char *>(ix.pairs + lj) + 16 * sub;
      const u32 dst = (u32)__cvta_generic_to_shared(slots + (gbase + j) * CFR_PAIR_SLOT_WORDS + sub * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
}

// must be called by all 32 lanes; `slots` = this warp's 32 slots
CFR_D void pair_stage_warp(const DevIndex &ix, u32 L, u32 *slots) {
  const int lane = threadIdx.x & 31, sub = lane & 3, gbase = lane & ~3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const u32 lj = __shfl_sync(0xffffffffu, L, gbase + j);
    if (lj != CFR_PAIR_NO_LINE) {  // uniform over the four lanes of the group
      u64 a, b, c, d;
      asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                   : "l"(reinterpret_cast<const char *>(ix.pairs + lj) + 32 * sub));
      ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(slots + (gbase + j) * CFR_PAIR_SLOT_WORDS + sub * 8);
      dst[0] = make_ulonglong2(a, b);
      dst[1] = make_ulonglong2(c, d);
    }
  }
  __syncwarp();
}

// the boundaries x1 = 64 L + s1 and (if two) x2 = 64 L + s2 from the lane's staged line
CFR_D void pair_query_slot(const DevIndex &ix, const u32 *slot, u64 L, int c1, int c2, int s1, int s2, bool two, PairQuery &q1,
                           PairQuery &q2) {
  const int idx = c1 * 4 + c2;
  const u64 *sb = ix.pair_sb + (L >> CFR_PAIR_SB_SHIFT) * 20;
  const ulonglong2 p01 = *reinterpret_cast<const ulonglong2 *>(slot);
  const ulonglong2 p23 = *reinterpret_cast<const ulonglong2 *>(slot + 4);
  const u64 m1 = pair_match(p01.x, p01.y, c1), m12 = m1 & pair_match(p23.x, p23.y, c2);
  const u64 bs = ld64(sb + 16 + c1) + (u64)slot[24 + c1], bp = ld64(sb + idx) + (u64)slot[8 + idx];
  const u64 b1 = (1ull << s1) - 1ull;
  q1.s1 = bs + (u64)popc64(m1 & b1);
  q1.p = bp + (u64)popc64(m12 & b1);
  q1.sym1 = (int)(((p01.x >> s1) & 1ull) | (((p01.y >> s1) & 1ull) << 1));
  q1.sym2 = (int)(((p23.x >> s1) & 1ull) | (((p23.y >> s1) & 1ull) << 1));
  q2 = q1;
  if (two) {
    const u64 b2 = (1ull << s2) - 1ull;
    q2.s1 = bs + (u64)popc64(m1 & b2);
    q2.p = bp + (u64)popc64(m12 & b2);
  }
}
#endif

// ---- building the pair lines at load time, from the occ sectors (which k_transcode derived from the
// run-block arrays with the literal Sequence_RunBlock::Rank / Access)
// planes of line L: code1 = B[i], code2 = B[img(i)] for the rows i of the line
CFR_HD void pair_line_planes(const DevIndex &ix, u64 L, u64 &a, u64 &b, u64 &c, u64 &d) {
  a = b = c = d = 0;
  const u64 p0 = L * 64;
  if (p0 >= ix.n) return;
  u64 lo, hi, w2, w3;
  occ_load<0>(ix.occ + L, lo, hi, w2, w3);
  for (int w = 0; w < 64; ++w) {
    const u64 pos = p0 + (u64)w;
    if (pos >= ix.n) break;
    const int c1 = occ_symbol(lo, hi, w);
    const u64 occ = occ_base(w2, w3, c1, L) + (u64)popc64(occ_match(lo, hi, c1) & ((1ull << w) - 1ull));
    const u64 img = ix.C[c1] + occ + ((c1 == ix.last_code && pos <= ix.first_isa) ? 1ull : 0ull);
    int c2 = 0;
    if (img < ix.n) {
      const u64x2 pl = ld128(reinterpret_cast<const u64x2 *>(ix.occ + (img >> 6)));
      c2 = occ_symbol(pl.x, pl.y, (int)(img & 63));
    }
    a |= (u64)(c1 & 1) << w;
    b |= (u64)(c1 >> 1) << w;
    c |= (u64)(c2 & 1) << w;
    d |= (u64)(c2 >> 1) << w;
  }
}

// rows of one line per pair (cnt[0..15]) and per code1 (cnt[16..19]); `valid` = mask of the rows below n
CFR_HD void pair_line_counts(u64 a, u64 b, u64 c, u64 d, u64 valid, u32 *cnt) {
  for (int c1 = 0; c1 < 4; ++c1) {
    const u64 m1 = pair_match(a, b, c1) & valid;
    cnt[16 + c1] = (u32)popc64(m1);
    for (int c2 = 0; c2 < 4; ++c2) cnt[c1 * 4 + c2] = (u32)popc64(m1 & pair_match(c, d, c2));
  }
}

CFR_HD u64 pair_valid_mask(const DevIndex &ix, u64 L) {
  const u64 p0 = L * 64;
  if (p0 >= ix.n) return 0;
  const u64 left = ix.n - p0;
  return left >= 64 ? ~0ull : ((1ull << left) - 1ull);
}

enum { CFR_PAIR_CHUNK = 64 };  // lines per chunk of the counter prefix passes

// pass A: planes of every line of a chunk + the chunk's totals, tot[k * n_chunk + chunk]
CFR_HD void pair_chunk_planes(const DevIndex &ix, PairLine *lines, u64 n_lines, u64 chunk, u64 *tot, u64 n_chunk) {
  u64 sum[20];
  for (int k = 0; k < 20; ++k) sum[k] = 0;
  for (u64 L = chunk * CFR_PAIR_CHUNK; L < (chunk + 1) * CFR_PAIR_CHUNK && L < n_lines; ++L) {
    u64 a, b, c, d;
    pair_line_planes(ix, L, a, b, c, d);
    lines[L].w[0] = a;
    lines[L].w[1] = b;
    lines[L].w[2] = c;
    lines[L].w[3] = d;
    u32 cnt[20];
    pair_line_counts(a, b, c, d, pair_valid_mask(ix, L), cnt);
    for (int k = 0; k < 20; ++k) sum[k] += cnt[k];
  }
  for (int k = 0; k < 20; ++k) tot[(u64)k * n_chunk + chunk] = sum[k];
}

// pass B (after the exclusive scan of tot over the chunks and the superblock table): the counters of every line
CFR_HD void pair_chunk_counters(const DevIndex &ix, PairLine *lines, u64 n_lines, u64 chunk, const u64 *tot, u64 n_chunk,
                                const u64 *sb_table) {
  u64 run[20];
  const u64 L0 = chunk * CFR_PAIR_CHUNK;
  const u64 *sb = sb_table + (L0 >> CFR_PAIR_SB_SHIFT) * 20;  // a chunk never straddles a superblock
  for (int k = 0; k < 20; ++k) run[k] = tot[(u64)k * n_chunk + chunk] - sb[k];
  for (u64 L = L0; L < L0 + CFR_PAIR_CHUNK && L < n_lines; ++L) {
    u64 *w = lines[L].w;
    for (int k = 0; k < 10; ++k) w[4 + k] = (run[2 * k] & 0xffffffffull) | (run[2 * k + 1] << 32);
    w[14] = 0;
    w[15] = 0;
    u32 cnt[20];
    pair_line_counts(w[0], w[1], w[2], w[3], pair_valid_mask(ix, L), cnt);
    for (int k = 0; k < 20; ++k) run[k] += cnt[k];
  }
}

// the constants of the second step: out[c1*4+c2] = D, out[16] = E, out[17] = F (see the derivation above)
CFR_HD void pair_constants(const DevIndex &ix, u64 *out) {
  for (int c1 = 0; c1 < 4; ++c1)
    for (int c2 = 0; c2 < 4; ++c2) out[c1 * 4 + c2] = occ_rank<0>(ix, c2, ix.C[c1]).count;
  const int lc = ix.last_code;
  {
    const u64 r = ix.C[lc];
    const u64x2 pl = ld128(reinterpret_cast<const u64x2 *>(ix.occ + (r >> 6)));
    out[16] = (u64)occ_symbol(pl.x, pl.y, (int)(r & 63));
  }
  {
    const u64 img = ix.C[lc] + occ_rank<0>(ix, lc, ix.first_isa).count + 1ull;
    u64 f = 0;
    if (img < ix.n) {
      const u64x2 pl = ld128(reinterpret_cast<const u64x2 *>(ix.occ + (img >> 6)));
      f = (u64)occ_symbol(pl.x, pl.y, (int)(img & 63));
    }
    out[17] = f;
  }
}

// the arithmetic of the two steps once the line contents at both boundaries are known (qa at sp, qe at ep + 1;
// qe is ignored when sp == ep)
CFR_HD int pair_finish(const DevIndex &ix, int c1, int c2, u64 &sp, u64 &ep, const PairQuery &qa, const PairQuery &qe, OpCount &oc) {
  const bool two = c2 >= 0;
  const bool range = sp != ep;
  const u64 xe = ep + 1;
  const u64 fisa = ix.first_isa;
  const bool l1 = c1 == ix.last_code;
  ++oc.xext;
  oc.xsingle += range ? 0u : 1u;
  const u64 y1 = ix.C[c1] + qa.s1 + ((l1 && sp <= fisa) ? 1ull : 0ull);
  const u64 y2 = range ? ix.C[c1] + qe.s1 + ((l1 && xe <= fisa) ? 1ull : 0ull) - 1ull
                       : y1 + ((qa.sym1 == c1) ? 0ull : ~0ull);
  if (y1 > y2 || y2 > ix.n) return 0;
  if (!two) {
    sp = y1;
    ep = y2;
    return 1;
  }
  const bool l2 = c2 == ix.last_code;
  const int idx2 = c1 * 4 + c2;
  const u64 g0 = ix.C[c2] + ix.pair_D[idx2];
  const u64 e_add = (l1 && ix.pair_E == c2) ? 1ull : 0ull;
  const u64 f_sub = (l1 && ix.pair_F == c2) ? 1ull : 0ull;
  const bool range2 = y1 != y2;
  ++oc.xext;
  oc.xsingle += range2 ? 0u : 1u;
  const u64 z1 = g0 + qa.p + e_add - (sp > fisa ? f_sub : 0ull) + ((l2 && y1 <= fisa) ? 1ull : 0ull);
  u64 z2;
  if (range) {
    const u64 Y = y2 + 1;  // = step(c1, ep + 1)
    z2 = g0 + qe.p + e_add - (xe > fisa ? f_sub : 0ull) + ((l2 && Y <= fisa) ? 1ull : 0ull) - 1ull;
    if (!range2 && l2 && y1 == fisa) z2 += 1;  // single row y1: the reference tests B[y1] instead of ranking
  } else {
    z2 = z1 + ((qa.sym2 == c2) ? 0ull : ~0ull);  // code2(sp) = B[y1]
  }
  sp = y1;
  ep = y2;
  if (z1 > z2 || z2 > ix.n) return 1;
  sp = z1;
  ep = z2;
  return 2;
}

// Two steps of FMIndex::BackwardSearch's loop (FMIndex.hpp:495-508) from the range [sp, ep]: first c1,
// then -- if c2 >= 0 -- c2.  Returns how many succeeded (0, 1, 2) and leaves the range after the last
// successful one in (sp, ep).  The operation counters advance as the reference's calls would.
// MODE 1 / 2 (device): one lane per search, lines staged in shared memory by the warp cooperatively
// (pair_stage_warp / pair_stage_warp_async); the call is warp-uniform -- lanes without a step to do pass
// go = false.  MODE 0: one lane reads whole lines (host twin).
template <int MODE>
struct BwtPairT {
  enum { COOP = MODE != 0 };
  typedef u64 pos_t;
  // PAIR = 2: the search loop calls round() (one memory wait per iteration) instead of extend2()
  enum { LANES = 1, PAIR = (MODE >= 3 ? 2 : 1), STEPS_COUNTED_AT_CLOSE = 0 };
  static CFR_HD bool leader() { return true; }
  // single-step form: not used by the search loop of a pair policy (Bwt::PAIR selects extend2)
  static CFR_HD void extend_step(const DevIndex &, int, u64 sp, u64 ep, u64 &nsp, u64 &nep, OpCount &) {
    nsp = sp;
    nep = ep;
  }
  static CFR_HD int extend2(const DevIndex &ix, bool go, int c1, int c2, u64 &sp, u64 &ep, OpCount &oc) {
    const bool two = c2 >= 0;
    const int c2q = two ? c2 : 0;
    const bool range = sp != ep;
    const u64 xe = ep + 1;
    const u64 La = sp >> 6, Le = xe >> 6;
    PairQuery qa, qe;
#if defined(__CUDA_ARCH__)
    if (COOP) {
      __shared__ __align__(16) u32 pair_slots[4][32 * CFR_PAIR_SLOT_WORDS];  // [warp of the block][lane slot]
      u32 *slots = pair_slots[(threadIdx.x >> 5) & 3];
      const u32 *mine = slots + (threadIdx.x & 31) * CFR_PAIR_SLOT_WORDS;
      c1 &= 3;  // lanes with go == false carry anything
      const bool near = range && Le == La;
      if (MODE == 2) pair_stage_warp_async(ix, go ? (u32)La : CFR_PAIR_NO_LINE, slots);
      else pair_stage_warp(ix, go ? (u32)La : CFR_PAIR_NO_LINE, slots);
      if (go) pair_query_slot(ix, mine, La, c1, c2q, (int)(sp & 63), (int)(xe & 63), near, qa, qe);
      __syncwarp();
      const bool far = go && range && Le != La;
      if (__ballot_sync(0xffffffffu, far)) {  // the second boundary lies in another line: a few lanes per step
        if (MODE == 2) pair_stage_warp_async(ix, far ? (u32)Le : CFR_PAIR_NO_LINE, slots);
        else pair_stage_warp(ix, far ? (u32)Le : CFR_PAIR_NO_LINE, slots);
        if (far) {
          PairQuery dummy;
          pair_query_slot(ix, mine, Le, c1, c2q, (int)(xe & 63), 0, false, qe, dummy);
        }
        __syncwarp();
      }
      if (!go) return 0;
    } else
#endif
    {
      if (!go) return 0;
      qa = pair_query_scalar(ix, c1, c2q, sp);
      qe = range ? pair_query_scalar(ix, c1, c2q, xe) : qa;
    }
    return pair_finish(ix, c1, c2, sp, ep, qa, qe, oc);
  }
#if defined(__CUDA_ARCH__)
  // MODE 3: ONE memory round of the warp per iteration of the search loop.  Everything a lane can be waiting for
  // travels in the same group of asynchronous copies and is awaited once:
  //   * the line at sp of a lane that extends (`go`; eight lanes fetch it as one 128-byte request, as in MODE 2),
  //   * the line at ep + 1 when that lies in another line ("far"): up to CFR_PAIR_FAR of them per round go to
  //     extra slots handed out by ballot rank (in MODEs 1 / 2 they cost every iteration a second, dependent
  //     DRAM round trip -- three quarters of all iterations have such a lane); the rare surplus takes a second round,
  //   * the 16-byte wide-lookup-table entry of a lane that starts a search (`probe`), copied by the lane itself.
  // Returns the steps done (as extend2) for go lanes; `entry` = the table entry for probe lanes.
  // MODE 4 state: one mbarrier and its phase parity per warp of the block (function-scope shared memory: one copy per block)
  static CFR_D unsigned long long *tma_bar() {
    __shared__ __align__(8) unsigned long long bar[4];
    return bar;
  }
  static CFR_D unsigned *tma_phase() {
    __shared__ unsigned phase[4];
    return phase;
  }
  // called once by every thread at the start of the search kernel
  static CFR_D void block_init() {
    if (MODE == 4) {
      if ((threadIdx.x & 31) == 0) {
        const int wi = (threadIdx.x >> 5) & 3;
        const u32 mb = (u32)__cvta_generic_to_shared(tma_bar() + wi);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        tma_phase()[wi] = 0;
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      __syncthreads();
    }
  }
  static CFR_D int round(const DevIndex &ix, bool go, int c1, int c2, u64 &sp, u64 &ep, bool probe, u64 key, u64x2 &entry,
                         OpCount &oc) {
    enum { F = CFR_PAIR_FAR, SLOTS = 32 + F };
    __shared__ __align__(16) u32 pair_slots[4][SLOTS * CFR_PAIR_SLOT_WORDS];  // [warp of the block][lane slots, far slots]
    const unsigned full = 0xffffffffu;
    u32 *slots = pair_slots[(threadIdx.x >> 5) & 3];
    const int lane = threadIdx.x & 31, sub = lane & 7, gbase = lane & ~7;
    u32 *mine = slots + lane * CFR_PAIR_SLOT_WORDS;
    const bool two = c2 >= 0;
    const int c2q = two ? c2 : 0;
    const bool range = sp != ep;
    const u64 xe = ep + 1;
    const u64 La = sp >> 6, Le = xe >> 6;
    c1 &= 3;  // lanes with go == false carry anything
    const bool near = range && Le == La;
    const bool far = go && range && Le != La;
    const u32 farmask = __ballot_sync(full, far);
    const int frank = __popc(farmask & ((1u << lane) - 1u));
    const int nfar = min(__popc(farmask), (int)F);
    const bool fslot = far && frank < F;
    u32 *fmine = slots + (32 + (fslot ? frank : 0)) * CFR_PAIR_SLOT_WORDS;
    if (fslot) fmine[32] = (u32)Le;  // a padding word of the far slot names its line
    __syncwarp();
    if (MODE == 4) {
      // the same round with bulk (TMA) copies: every lane copies its own line(s) / table entry with one cp.async.bulk each,
      // completion is counted in bytes on one mbarrier per warp (no cooperation between lanes, no shuffles)
      const int wi = (threadIdx.x >> 5) & 3;
      const u32 mb = (u32)__cvta_generic_to_shared(tma_bar() + wi);
      const unsigned parity = tma_phase()[wi] & 1u;
      const u32 bytes = 128u * (u32)(__popc(__ballot_sync(full, go)) + nfar) + 16u * (u32)__popc(__ballot_sync(full, probe));
      if (bytes) {
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slots were last touched through the generic proxy
        if (go) {
          const u32 dst = (u32)__cvta_generic_to_shared(mine);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(dst),
                       "l"(ix.pairs + La), "r"(mb)
                       : "memory");
        }
        if (fslot) {
          const u32 dst = (u32)__cvta_generic_to_shared(fmine);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(dst),
                       "l"(ix.pairs + Le), "r"(mb)
                       : "memory");
        }
        if (probe) {
          const u32 dst = (u32)__cvta_generic_to_shared(mine);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(dst),
                       "l"(ix.wide + key), "r"(mb)
                       : "memory");
        }
        unsigned done = 0;
        while (!done)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                       : "=r"(done)
                       : "r"(mb), "r"(parity)
                       : "memory");
        __syncwarp();
        if (lane == 0) tma_phase()[wi] = parity ^ 1u;
        __syncwarp();
      }
    } else {
    const u32 L = go ? (u32)La : CFR_PAIR_NO_LINE;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const u32 lj = __shfl_sync(full, L, gbase + j);
      if (lj != CFR_PAIR_NO_LINE) {  // uniform over the eight lanes of the group
        const char *src = reinterpret_cast<const char *>(ix.pairs + lj) + 16 * sub;
        const u32 dst = (u32)__cvta_generic_to_shared(slots + (gbase + j) * CFR_PAIR_SLOT_WORDS + sub * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
#pragma unroll
    for (int j = 0; j < F / 4; ++j) {  // far slot r is fetched by the group r mod 4
      const int r = (lane >> 3) + 4 * j;
      if (r < nfar) {
        u32 *fs = slots + (32 + r) * CFR_PAIR_SLOT_WORDS;
        const char *src = reinterpret_cast<const char *>(ix.pairs + fs[32]) + 16 * sub;
        const u32 dst = (u32)__cvta_generic_to_shared(fs + sub * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
    if (probe) {
      const u32 dst = (u32)__cvta_generic_to_shared(mine);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(ix.wide + key) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    }
    PairQuery qa, qe;
    if (go) pair_query_slot(ix, mine, La, c1, c2q, (int)(sp & 63), (int)(xe & 63), near, qa, qe);
    if (fslot) {
      PairQuery dummy;
      pair_query_slot(ix, fmine, Le, c1, c2q, (int)(xe & 63), 0, false, qe, dummy);
    }
    if (probe) {
      const ulonglong2 e = *reinterpret_cast<const ulonglong2 *>(mine);
      entry.x = e.x;
      entry.y = e.y;
    }
    __syncwarp();
    const bool surplus = far && !fslot;
    if (__ballot_sync(full, surplus)) {  // more than F far lines in one round
      pair_stage_warp_async(ix, surplus ? (u32)Le : CFR_PAIR_NO_LINE, slots);
      if (surplus) {
        PairQuery dummy;
        pair_query_slot(ix, mine, Le, c1, c2q, (int)(xe & 63), 0, false, qe, dummy);
      }
      __syncwarp();
    }
    if (!go) return 0;
    return pair_finish(ix, c1, c2, sp, ep, qa, qe, oc);
  }
#endif
};

// largest row count the 32-bit walkers accept
#define CFR_POS32_MAX_N 0xfffffff0ull

// ---------------------------------------------------------------------------
// FM-index search and locate
// ---------------------------------------------------------------------------

// FMIndex::BackwardSearch with GetBackwardSearchInitialRange inlined (nested-loop
// form; the bulk search uses the flattened loop in cfr_pipeline.cuh)
template <class Bwt>
CFR_HD int backward_search(const DevIndex &ix, StrandSeq &s, int m, u64 &sp, u64 &ep, OpCount &oc) {
  const int W = ix.pre_width;
  if (m < W) return 0;
  ++oc.search;
  int l = 0;
  if (W > 0) {
    u64 key;
    int nvalid;
    if (!s.init_key(m, W, key, nvalid)) {
      sp = 1;
      ep = 0;
      return nvalid;
    }
    const u64x2 e = ld128(ix.lookup + key);
    if (e.y == 0) {
      sp = 1;
      ep = 0;
      return W - 1;
    }
    sp = e.x;
    ep = e.x + e.y - 1;
    l = W;
  } else {
    sp = 0;
    ep = ix.n - 1;
  }
  while (l < m) {
    const int c = s(m - 1 - l);
    if (c > 3) break;
    u64 nsp, nep;
    Bwt::extend(ix, c, sp, ep, nsp, nep, oc);
    if (nsp > nep || nep > ix.n) break;
    sp = nsp;
    ep = nep;
    ++l;
  }
  return l;
}

// One entry of the wide lookup table: FMIndex::BackwardSearch over exactly the WW bases packed in
// `key` (base j of the field = the base WW-1-j steps from the end, as in init_key): the W-mer table
// probe on the last W bases, then BackwardExtend base by base until it fails.  (sp, ep) is the last
// valid range, l the number of bases matched -- exactly where the search stands, or ended, after
// these WW bases whatever precedes them.
template <class Bwt>
CFR_HD u64x2 wide_lookup_entry(const DevIndex &ix, u64 key, int WW) {
  const int W = ix.pre_width;
  u64x2 r;
  const u64x2 e = ld128(ix.lookup + (key >> (2 * (WW - W))));
  if (e.y == 0) {
    r.x = 1;
    r.y = 0 | ((u64)(W - 1) << 56);
    return r;
  }
  u64 sp = e.x, ep = e.x + e.y - 1;
  int l = W;
  OpCount oc{};
  while (l < WW) {
    const int c = (int)((key >> (2 * (WW - 1 - l))) & 3ull);
    u64 nsp, nep;
    Bwt::extend(ix, c, sp, ep, nsp, nep, oc);
    if (nsp > nep || nep > ix.n) break;
    sp = nsp;
    ep = nep;
    ++l;
  }
  r.x = sp;
  r.y = ep | ((u64)l << 56);
  return r;
}

// FixedSizeElemArray::Read (FixedSizeElemArray.hpp:102, Utils.hpp:197-219)
CFR_HD u64 sa_read(const DevIndex &ix, u64 i) {
  const u64 s = i * (u64)ix.sa_bits, e = s + (u64)ix.sa_bits - 1;
  const u64 is = s >> 6, ie = e >> 6;
  const int rs = (int)(s & 63);
  if (is == ie) {
    const u64 m = ix.sa_bits >= 64 ? ~0ull : ((1ull << ix.sa_bits) - 1ull);
    return (ld64(ix.sampled_sa + is) >> rs) & m;
  }
  const int re = (int)(e & 63);
  return (ld64(ix.sampled_sa + is) >> rs) | ((ld64(ix.sampled_sa + ie) & ((1ull << (re + 1)) - 1ull)) << (64 - rs));
}

// i % sampleRate == 0 and i / filterRate; both rates are powers of two in every
// index the reference builder writes (--offrate, 1024), the general form is kept
template <typename Pos>
CFR_HD bool is_sampled_row(const DevIndex &ix, Pos i) {
  return ix.sample_shift >= 0 ? (i & (((Pos)1 << ix.sample_shift) - (Pos)1)) == 0 : (i % (Pos)ix.sample_rate) == 0;
}
template <typename Pos>
CFR_HD Pos filter_bit_index(const DevIndex &ix, Pos i) {
  return ix.filter_shift >= 0 ? (i >> ix.filter_shift) : i / (Pos)ix.sel_filter_rate;
}

// FMIndex::GetSampledSA
CFR_HD bool get_sampled_sa(const DevIndex &ix, u64 i, u64 &sa) {
  if (i == ix.first_isa) {
    sa = ix.adjusted_sa0;
    return true;
  }
  if (is_sampled_row(ix, i)) {
    sa = sa_read(ix, ix.sample_shift >= 0 ? (i >> ix.sample_shift) : i / (u64)ix.sample_rate);
    return true;
  }
  if (ix.sel_filter) {
    const u64 fb = filter_bit_index(ix, i);
    if ((ld64(ix.sel_filter + (fb >> 6)) >> (fb & 63)) & 1ull) {
      u64 lo = 0, hi = ix.sel_cnt;
      while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (ld64(&ix.sel[mid].x) < i) lo = mid + 1; else hi = mid;
      }
      if (lo < ix.sel_cnt && ld64(&ix.sel[lo].x) == i) {
        sa = ld64(&ix.sel[lo].y);
        return true;
      }
    }
  }
  return false;
}

// rows answered by the dense locate table (when built)
template <typename Pos>
CFR_HD bool is_dense_row(const DevIndex &ix, Pos i) {
  return ix.dense_shift >= 0 && (i & (((Pos)1 << ix.dense_shift) - (Pos)1)) == 0;
}

// entry j of the dense locate table: 16-bit entries when every sequence id fits (half the HBM, so twice the density)
CFR_HD u32 dense_read(const DevIndex &ix, u64 j) {
  return ix.dense16 ? ld16(reinterpret_cast<const unsigned short *>(ix.dense) + j) : ld32(ix.dense + j);
}

// GetSampledSA with the dense table in front: a dense row's entry is what the literal procedure
// below returns for the walk that starts there (checks at the row itself included)
CFR_HD bool get_located(const DevIndex &ix, u64 i, u64 &sa) {
  if (is_dense_row(ix, i)) {
    sa = (u64)dense_read(ix, i >> ix.dense_idx_shift);
    return true;
  }
  return get_sampled_sa(ix, i, sa);
}

// FMIndex::BackwardToSampledSA: the sampled SA holds sequence ids (Builder.hpp:27-71)
template <class Bwt>
CFR_HD u64 locate_row(const DevIndex &ix, u64 i, OpCount &oc) {
  u64 sa = 0;
  while (!get_sampled_sa(ix, i, sa)) {
    i = Bwt::lf(ix, i, oc);
    ++oc.lf;
  }
  ++oc.locate;
  return sa;
}

// ---------------------------------------------------------------------------
// Classifier: seed search
// ---------------------------------------------------------------------------

CFR_HD int max_hits_for_len(int len, int mhl) { return len < mhl ? 0 : (len + 1) / (mhl + 1); }

// Classifier::CalculateHitScore (nucleotide, _scoreHitLenAdjust = 15)
CFR_HD u64 hit_score(int l, int mhl) {
  if (l < mhl) return 0;
  return (u64)(long long)(l - 15) * (u64)(long long)(l - 15);
}

// Classifier::AdjustHitBoundaryFromStrandHits.  h1 = strandHits[1] (the mate as
// read), h0 = strandHits[0] (its reverse complement).
template <class Bwt>
CFR_HD void adjust_hit_boundary(const DevIndex &ix, StrandSeq fw, Hit *h0, int n0, Hit *h1, int n1, OpCount &oc) {
  if (!n0 || !n1) return;
  const int len = fw.len;
  StrandSeq rc = fw;
  rc.rc = 1;
  u64 sp = 0, ep = 0;
  int j = n0 - 1;
  bool need_fix0 = false, need_fix1 = false;
  for (int i = 0; i < n1; ++i) {
    const int right = len - h1[i].offset - 1;
    const int left = right - h1[i].l + 1;
    for (; j >= 0; --j) {
      const int rc_left = h0[j].offset;
      const int rc_right = rc_left + h0[j].l - 1;
      if (rc_left >= right) continue;
      if (left >= rc_right) break;
      if (left == rc_left && right == rc_right) break;
      if (left < rc_left && rc_right < right) break;
      if (rc_left < left && right < rc_right) break;
      if (rc_right > right) {
        const int l = backward_search<Bwt>(ix, fw, rc_right + 1, sp, ep, oc);
        if (rc_right - l + 1 == left && sp <= ep) {
          h1[i].sp = sp;
          h1[i].ep = ep;
          h1[i].l = l;
          h1[i].offset = len - rc_right - 1;
          need_fix1 = true;
        }
      }
      if (left < rc_left) {
        const int l = backward_search<Bwt>(ix, rc, len - left, sp, ep, oc);
        if (left + l - 1 == rc_right && sp <= ep) {
          h0[j].sp = sp;
          h0[j].ep = ep;
          h0[j].l = l;
          h0[j].offset = left;
          need_fix0 = true;
        }
      }
    }
  }
  for (int k = 0; k <= 1; ++k) {  // Classifier.hpp:361-400
    Hit *h = k ? h1 : h0;
    const int n = k ? n1 : n0;
    if (!(k ? need_fix1 : need_fix0)) continue;
    for (int i = 0; i < n - 1; ++i) {
      const int starti = h[i].offset;
      const int endi = starti + h[i].l - 1;
      for (int q = i + 1; q < n; ++q) {
        const int startj = h[q].offset;
        if (startj > endi) break;
        const int endj = startj + h[q].l - 1;
        if (h[q].l >= h[i].l) {
          h[i].l = startj - starti;
          break;
        } else {
          if (endj <= endi)
            h[q].l = 0;
          else {
            h[q].offset = endi + 1;
            h[q].l = endj - (endi + 1) + 1;
            break;
          }
        }
      }
    }
  }
}

// number of BWT rows GetClassificationFromHits resolves for one hit and the
// stride it uses (Classifier.hpp:620-666)
struct RowPlan {
  u64 step;   // 0: every row of [sp, ep]
  u64 fwd;    // rows taken walking up from sp
  u64 total;  // fwd + rows taken walking down from ep
};

CFR_HD RowPlan plan_rows(u64 sp, u64 ep, const DevParams &p) {
  RowPlan rp;
  const u64 range = ep - sp + 1;
  const u64 max_entries = (u64)(long long)(p.max_result * p.hitk_factor);
  if (range <= max_entries || p.hitk_factor <= 0 || p.max_result <= 0) {
    rp.step = 0;
    rp.fwd = range;
    rp.total = range;
    return rp;
  }
  const u64 step = (range % max_entries) ? range / max_entries + 1 : range / max_entries;
  rp.step = step;
  rp.fwd = (range - 1) / step + 1;  // j = sp, sp+step, ... <= ep
  // walking down from ep: stops when resolved >= max_entries (checked after each
  // locate) or when j would pass sp
  const u64 down_avail = (range - 1) / step + 1;
  u64 down = rp.fwd >= max_entries ? 1 : max_entries - rp.fwd;
  if (down > down_avail) down = down_avail;
  rp.total = rp.fwd + down;
  return rp;
}

CFR_HD u64 plan_row_at(u64 sp, u64 ep, const RowPlan &rp, u64 t) {
  if (rp.step == 0) return sp + t;
  if (t < rp.fwd) return sp + t * rp.step;
  return ep - (t - rp.fwd) * rp.step;
}

// ---------------------------------------------------------------------------
// Taxonomy
// ---------------------------------------------------------------------------

CFR_HD u32 tax_parent(const DevIndex &ix, u32 t) { return ld32(ix.parent + t); }

// Taxonomy::LCA (lcaChildTaxIds == NULL).  ids are all < node_cnt.
// The backbone is the root path of the first non-root id; for every other id the
// reference counts, per backbone position, the ids whose root-aligned path agrees
// from the top down to it, and returns the lowest position all non-root ids share.
// Equivalent without the count array: the answer index is the maximum over ids of
// (highest mismatching aligned position + 1).
CFR_HD u32 tax_lca(const DevIndex &ix, const u64 *ids, int cnt, u64 *err_flags) {
  const u32 root = (u32)ix.root;
  int k = 0;
  while (k < cnt && (u32)ids[k] == root) ++k;
  if (k >= cnt) return root;
  {
    // Shortcut for the common case (strains of one species tie): no id is a root-level node and all
    // share one parent.  Their root paths have equal length and differ at most in the lowest node,
    // so the general procedure below returns the id itself when all are equal, else the parent.
    // `cnt` independent loads instead of several dependent parent-chain walks.
    const u32 x0 = (u32)ids[0];
    const u32 p0 = tax_parent(ix, x0);
    bool same_parent = p0 != x0, all_equal = true;
    for (int i = 1; i < cnt; ++i) {
      const u32 x = (u32)ids[i];
      const u32 px = tax_parent(ix, x);
      same_parent = same_parent && px == p0 && px != x;
      all_equal = all_equal && x == x0;
    }
    if (same_parent) {
      if (all_equal) return x0;
      return tax_parent(ix, p0) == p0 ? root : p0;  // a root-level parent is reported as the root (the pushed path end)
    }
  }
  u32 path[CFR_TAX_PATH_CAP];
  int plen = 0;
  u32 t = (u32)ids[k];
  do {
    if (plen >= CFR_TAX_PATH_CAP - 1) {
      *err_flags |= 1ull;
      return root;
    }
    path[plen++] = t;
    t = tax_parent(ix, t);
  } while (t != tax_parent(ix, t));
  path[plen++] = root;
  int first_common = 0;  // lowest backbone index shared by everybody seen so far
  for (int i = 0; i < cnt; ++i) {
    if (i == k) continue;
    u32 x = (u32)ids[i];
    if (x == tax_parent(ix, x)) continue;  // a root-level id: ignored (rootCount)
    // length of x's path (nodes below the root-level node, plus the pushed root)
    int tlen = 0;
    t = x;
    do {
      ++tlen;
      t = tax_parent(ix, t);
    } while (t != tax_parent(ix, t));
    ++tlen;
    // align the two paths at their root ends
    int ib, it;  // indices into backbone / tmp path of the lowest aligned pair
    if (tlen >= plen) {
      it = tlen - plen;
      ib = 0;
    } else {
      it = 0;
      ib = plen - tlen;
    }
    t = x;
    for (int s = 0; s < it; ++s) t = tax_parent(ix, t);
    int last_mismatch = ib - 1;  // positions below the aligned window are never counted
    for (int q = ib; q < plen; ++q, ++it) {
      const u32 node = (it == tlen - 1) ? root : t;
      if (node != path[q]) last_mismatch = q;
      if (it < tlen - 1) t = tax_parent(ix, t);
    }
    if (last_mismatch + 1 > first_common) first_common = last_mismatch + 1;
  }
  return first_common >= plen ? root : path[first_common];
}

CFR_HD int tax_level_of(const DevIndex &ix, u32 t) {
  const unsigned char r = ld8(ix.rank + t);
  return ix.rank_num[r < 31 ? r : 0];
}

// the node an id contributes to promotion level `ri` (Taxonomy.hpp:904-930), or
// ~0u when its lineage ends below that level
CFR_HD u32 tax_node_at_level(const DevIndex &ix, u32 id, int ri) {
  if (ri == 0) return id;
  const int unknown = ix.rank_num[0];
  int prev = 0;
  u32 t = id;
  do {
    const int rn = tax_level_of(ix, t);
    if (rn != unknown && rn > prev) {
      if (ri <= rn) return t;  // levels (prev, rn] all receive t
      prev = rn;
    }
    t = tax_parent(ix, t);
  } while (t != tax_parent(ix, t));
  return ~0u;
}

// in-place ascending sort of u64 keys (heapsort above 24 entries: no recursion, O(n log n))
CFR_HD void sort_u64(u64 *a, int n) {
  if (n <= 24) {
    for (int i = 1; i < n; ++i) {
      const u64 x = a[i];
      int j = i - 1;
      while (j >= 0 && a[j] > x) {
        a[j + 1] = a[j];
        --j;
      }
      a[j + 1] = x;
    }
    return;
  }
  for (int start = n / 2 - 1; start >= 0; --start) {
    int root = start;
    for (;;) {
      int child = 2 * root + 1;
      if (child >= n) break;
      if (child + 1 < n && a[child] < a[child + 1]) ++child;
      if (a[root] >= a[child]) break;
      const u64 tmp = a[root]; a[root] = a[child]; a[child] = tmp;
      root = child;
    }
  }
  for (int end = n - 1; end > 0; --end) {
    u64 tmp = a[0]; a[0] = a[end]; a[end] = tmp;
    int root = 0;
    for (;;) {
      int child = 2 * root + 1;
      if (child >= end) break;
      if (child + 1 < end && a[child] < a[child + 1]) ++child;
      if (a[root] >= a[child]) break;
      tmp = a[root]; a[root] = a[child]; a[child] = tmp;
      root = child;
    }
  }
}

CFR_HD int unique_u64(u64 *a, int n) {
  if (n == 0) return 0;
  int m = 1;
  for (int i = 1; i < n; ++i)
    if (a[i] != a[m - 1]) a[m++] = a[i];
  return m;
}

// Taxonomy::ReduceTaxIds (promotedChildTaxIds == NULL).  tax_ids[0..cnt) are the
// compact tax ids of the best sequences (cnt > k guaranteed by the caller);
// `scratch` has room for cnt entries.  Writes <= max(k,1) ids, ascending, to out.
CFR_HD int tax_reduce(const DevIndex &ix, const u64 *tax_ids, int cnt, int k, u64 *scratch, u64 *out,
                      u64 *err_flags) {
  for (int i = 0; i < cnt; ++i)
    if (tax_ids[i] >= ix.node_cnt) {  // Taxonomy.hpp:855-882
      out[0] = ix.node_cnt;
      return 1;
    }
  if (k == 1) {
    out[0] = tax_lca(ix, tax_ids, cnt, err_flags);
    return 1;
  }
  const int unknown = ix.rank_num[0];
  for (int ri = 0; ri < unknown; ++ri) {  // Taxonomy.hpp:933-936
    int m = 0;
    for (int i = 0; i < cnt; ++i) {
      const u32 node = tax_node_at_level(ix, (u32)tax_ids[i], ri);
      if (node != ~0u) scratch[m++] = node;
    }
    sort_u64(scratch, m);
    m = unique_u64(scratch, m);
    if (m <= k) {
      for (int i = 0; i < m; ++i) out[i] = scratch[i];
      if (m == 0) {
        out[0] = ix.root;
        return 1;
      }
      return m;
    }
  }
  out[0] = ix.root;  // level `unknown` is always empty
  return 1;
}

// Where the lineage of `x` leaves the backbone `path` (Taxonomy::LCA, Taxonomy.hpp:775-813): both root
// paths are aligned at their root ends and compared from the top; returns the highest aligned index
// that differs (plen - tlen - 1 ... when the whole aligned window agrees) and, in `node`, x's lineage
// node at that index (~0u when x's path does not reach down to it).  x is not a root-level node.
CFR_HD int tax_divergence(const DevIndex &ix, const u32 *path, int plen, u32 x, u32 &node) {
  const u32 root = (u32)ix.root;
  int tlen = 0;
  u32 t = x;
  do {
    ++tlen;
    t = tax_parent(ix, t);
  } while (t != tax_parent(ix, t));
  ++tlen;  // the pushed root
  int ib, it;
  if (tlen >= plen) {
    it = tlen - plen;
    ib = 0;
  } else {
    it = 0;
    ib = plen - tlen;
  }
  t = x;
  u32 above = ~0u;  // x's lineage node one step below the aligned window (index ib - 1), if it has one
  for (int s = 0; s < it; ++s) {
    above = t;
    t = tax_parent(ix, t);
  }
  int last_mismatch = ib - 1;
  node = above;
  for (int q = ib; q < plen; ++q, ++it) {
    const u32 here = (it == tlen - 1) ? root : t;
    if (here != path[q]) {
      last_mismatch = q;
      node = here;
    }
    if (it < tlen - 1) t = tax_parent(ix, t);
  }
  return last_mismatch;
}

// The child lists of Taxonomy::ReduceTaxIds / LCA (promotedChildTaxIds / lcaChildTaxIds, used by
// --expand-taxid: Taxonomy.hpp:767-773, :811-812, :825-831, :871-878, :938-971), computed after
// tax_reduce() on the same ids.  promoted[0..np) is what tax_reduce returned.  Writes list after
// list to `child` (room for cnt ids; child_cnt[i] = length of list i) and returns the total; lists
// the classifier would not print (Classifier.hpp:823) have length 0.  `scratch` has room for cnt ids.
CFR_HD int tax_expand(const DevIndex &ix, const u64 *tax_ids, int cnt, int k, const u64 *promoted, int np,
                      u64 *scratch, u64 *child, u32 *child_cnt, u64 *err_flags) {
  for (int i = 0; i < np; ++i) child_cnt[i] = 0;
  for (int i = 0; i < cnt; ++i)
    if (tax_ids[i] >= ix.node_cnt) {  // :871-878: every input id, as given
      for (int j = 0; j < cnt; ++j) child[j] = tax_ids[j];
      child_cnt[0] = (u32)cnt;
      return cnt;
    }
  if (k == 1) {
    const u32 root = (u32)ix.root;
    int kk = 0;
    while (kk < cnt && (u32)tax_ids[kk] == root) ++kk;
    if (kk >= cnt) return 0;  // all root: the list stays empty (:756)
    u32 path[CFR_TAX_PATH_CAP];
    int plen = 0;
    u32 t = (u32)tax_ids[kk];
    do {
      if (plen >= CFR_TAX_PATH_CAP - 1) {
        *err_flags |= 1ull;
        return 0;
      }
      path[plen++] = t;
      t = tax_parent(ix, t);
    } while (t != tax_parent(ix, t));
    path[plen++] = root;
    // the answer is the backbone node just above the highest divergence; the list holds the backbone
    // node below it and the nodes where the other lineages leave the backbone right there
    int j = 0;
    for (int i = 0; i < cnt; ++i) {
      const u32 x = (u32)tax_ids[i];
      if (i == kk || x == tax_parent(ix, x)) continue;
      u32 node;
      const int m = tax_divergence(ix, path, plen, x, node);
      if (m + 1 > j) j = m + 1;
    }
    int nc = 0;
    if (j >= 1) scratch[nc++] = path[j - 1];
    for (int i = 0; i < cnt; ++i) {
      const u32 x = (u32)tax_ids[i];
      if (i == kk || x == tax_parent(ix, x)) continue;
      u32 node;
      const int m = tax_divergence(ix, path, plen, x, node);
      if (m + 1 == j && node != ~0u) scratch[nc++] = node;
    }
    sort_u64(scratch, nc);
    nc = unique_u64(scratch, nc);
    for (int i = 0; i < nc; ++i) child[i] = scratch[i];
    child_cnt[0] = (u32)nc;
    return nc;
  }
  // rank promotion: find the level tax_reduce stopped at, then group the level below by promoted node
  const int unknown = ix.rank_num[0];
  int ri = 0, m = 0;
  for (; ri < unknown; ++ri) {
    m = 0;
    for (int i = 0; i < cnt; ++i) {
      const u32 node = tax_node_at_level(ix, (u32)tax_ids[i], ri);
      if (node != ~0u) scratch[m++] = node;
    }
    sort_u64(scratch, m);
    m = unique_u64(scratch, m);
    if (m <= k) break;
  }
  if (ri >= unknown || m == 0 || ri == 0) return 0;  // root by default (:939-940) or nothing promoted (:941)
  m = 0;
  for (int i = 0; i < cnt; ++i) {
    const u32 node = tax_node_at_level(ix, (u32)tax_ids[i], ri - 1);
    if (node != ~0u) scratch[m++] = node;
  }
  sort_u64(scratch, m);
  m = unique_u64(scratch, m);
  int total = 0;
  for (int q = 0; q < m; ++q) {  // :953-970
    const u32 c = (u32)scratch[q];
    u32 t = c;
    int owner = -1;
    while (t != tax_parent(ix, t)) {
      t = tax_parent(ix, t);
      const int rn = tax_level_of(ix, t);
      if (rn > ri) break;
      if (rn == ri) {
        for (int i = 0; i < np; ++i)
          if ((u32)promoted[i] == t) owner = i;
        break;
      }
    }
    if (owner >= 0) child[total++] = ((u64)owner << 32) | c;
  }
  sort_u64(child, total);  // by promoted node, ascending ids inside (the order of the reference's map walk)
  for (int q = 0; q < total; ++q) {
    ++child_cnt[child[q] >> 32];
    child[q] &= 0xffffffffull;
  }
  return total;
}

// ---------------------------------------------------------------------------
// Classifier: scoring (GetClassificationFromHits after the locate walks)
// ---------------------------------------------------------------------------

// sorted-by-seqId table standing in for std::map<size_t,_seqHitRecord>
struct RecTable {
  SeqRec *a;
  int n;
};

CFR_HD int rec_lower_bound(const RecTable &t, u32 seq_id) {
  int lo = 0, hi = t.n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (t.a[mid].seq_id < seq_id) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// operator[] semantics: default-constructs {score 0, hitLength 0} when absent
CFR_HD SeqRec *rec_get(RecTable &t, u32 seq_id, bool &created) {
  const int pos = rec_lower_bound(t, seq_id);
  if (pos < t.n && t.a[pos].seq_id == seq_id) {
    created = false;
    return &t.a[pos];
  }
  for (int i = t.n; i > pos; --i) t.a[i] = t.a[i - 1];
  t.a[pos].seq_id = seq_id;
  t.a[pos].score = 0;
  t.a[pos].hit_length = 0;
  ++t.n;
  created = true;
  return &t.a[pos];
}

// Everything after the locate walks for one read.  `seq_ids` holds, hit after
// hit, the located sequence ids (row_cnt entries per scored hit, same order as
// the row plan).  rec0/rec1/best/tmp are per-read scratch with `arena_rows`
// capacity each (a hit contributes at most row_cnt distinct ids).
CFR_HD void score_read(const DevIndex &ix, const DevParams &p, const FinalHit *hits, int hit_cnt, u32 *seq_ids,
                       SeqRec *rec0, SeqRec *rec1, u64 *best, u64 *tmp, DevResult &res, u64 *out_ids,
                       u64 *err_flags, int *nb_out = nullptr) {
  const int mhl = p.min_hit_len;
  RecTable rec[2] = {{rec0, 0}, {rec1, 0}};
  u32 prev_seq = 0;
  u64 prev_score = 0;
  int prev_len = 0;
  bool mix_strand = false;
  for (int i = 1; i < hit_cnt; ++i)
    if (hits[i].strand != hits[i - 1].strand) {
      mix_strand = true;
      break;
    }
  u32 *ids = seq_ids;
  for (int i = 0; i < hit_cnt; ++i) {
    if (hits[i].l < mhl) continue;
    const u64 score = hit_score(hits[i].l, mhl);
    const int k = (hits[i].strand + 1) / 2;
    int m = (int)hits[i].row_cnt;
    // localSeqIdHit: ascending distinct ids of this hit
    if (m <= 24) {
      for (int a = 1; a < m; ++a) {
        const u32 x = ids[a];
        int b = a - 1;
        while (b >= 0 && ids[b] > x) {
          ids[b + 1] = ids[b];
          --b;
        }
        ids[b + 1] = x;
      }
    } else {  // heapsort on u32
      for (int start = m / 2 - 1; start >= 0; --start) {
        int root = start;
        for (;;) {
          int child = 2 * root + 1;
          if (child >= m) break;
          if (child + 1 < m && ids[child] < ids[child + 1]) ++child;
          if (ids[root] >= ids[child]) break;
          const u32 t2 = ids[root]; ids[root] = ids[child]; ids[child] = t2;
          root = child;
        }
      }
      for (int end = m - 1; end > 0; --end) {
        u32 t2 = ids[0]; ids[0] = ids[end]; ids[end] = t2;
        int root = 0;
        for (;;) {
          int child = 2 * root + 1;
          if (child >= end) break;
          if (child + 1 < end && ids[child] < ids[child + 1]) ++child;
          if (ids[root] >= ids[child]) break;
          t2 = ids[root]; ids[root] = ids[child]; ids[child] = t2;
          root = child;
        }
      }
    }
    const bool uniq = hits[i].ep == hits[i].sp;
    const bool adjacent = !mix_strand && i > 0 && uniq && hits[i - 1].ep == hits[i - 1].sp &&
                          hits[i - 1].strand == hits[i].strand &&
                          hits[i - 1].offset + hits[i - 1].l + 1 == hits[i].offset;
    for (int a = 0; a < m; ++a) {
      const u32 seq_id = ids[a];
      if (a > 0 && seq_id == ids[a - 1]) continue;
      bool created;
      if (adjacent && seq_id == prev_seq) {  // merge adjacent unique hits (Classifier.hpp:673-685)
        SeqRec *r = rec_get(rec[k], seq_id, created);
        r->score -= prev_score;
        prev_len += hits[i].l;
        prev_score = hit_score(prev_len, mhl);
        r->score += prev_score;
        r->hit_length += hits[i].l;
      } else {
        SeqRec *r = rec_get(rec[k], seq_id, created);
        if (created) {
          r->score = score;
          r->hit_length = hits[i].l;
        } else {
          r->score += score;
          r->hit_length += hits[i].l;
        }
        if (uniq) {
          prev_seq = seq_id;
          prev_score = score;
          prev_len = hits[i].l;
        }
      }
    }
    ids += m;
  }

  // best / second best, strand 0 then strand 1, ascending seqId (Classifier.hpp:711-736)
  u64 best_score = 0, second = 0, best_len = 0, second_len = 0;
  for (int k = 0; k <= 1; ++k)
    for (int i = 0; i < rec[k].n; ++i) {
      const SeqRec &r = rec[k].a[i];
      if (r.score > best_score) {
        second = best_score;
        second_len = best_len;
        best_score = r.score;
        best_len = (u64)(long long)r.hit_length;
      } else if (r.score > second) {
        second = r.score;
        second_len = (u64)(long long)r.hit_length;
      }
    }
  res.score = best_score;
  res.secondary_score = second;
  res.hit_length = (int)best_len;

  int nb = 0;
  for (int k = 0; k <= 1; ++k)  // Classifier.hpp:743-757
    for (int i = 0; i < rec[k].n; ++i)
      if (rec[k].a[i].score == best_score) {
        const u64 id = rec[k].a[i].seq_id;
        bool used = false;
        for (int q = 0; q < nb; ++q)
          if (best[q] == id) {
            used = true;
            break;
          }
        if (!used) best[nb++] = id;
      }
  if (nb > 1) res.secondary_score = best_score;
  if (second_len >= p.secondary_len && second < best_score &&
      second >= (u64)(p.secondary_factor * (double)best_score)) {  // Classifier.hpp:763-781
    for (int k = 0; k <= 1; ++k)
      for (int i = 0; i < rec[k].n; ++i)
        if (rec[k].a[i].score == second) {
          const u64 id = rec[k].a[i].seq_id;
          bool used = false;
          for (int q = 0; q < nb; ++q)
            if (best[q] == id) {
              used = true;
              break;
            }
          if (!used) best[nb++] = id;
        }
    res.secondary_score = second;
  }

  if (nb <= p.max_result || p.max_result <= 0) {  // Classifier.hpp:784-797 (-k 0: all of them, never reduced)
    if (nb > p.ids_stride) {
      *err_flags |= 4ull;  // more best-scoring sequences than the unlimited form keeps per read
      nb = p.ids_stride;
    }
    for (int i = 0; i < nb; ++i) out_ids[i] = best[i];
    res.n_assign = nb;
    res.by_rank = 0;
  } else {  // Classifier.hpp:798-841
    for (int i = 0; i < nb; ++i) {
      const u64 s = best[i];
      best[i] = s < ix.seq_cnt ? (u64)ld32(ix.seq_to_tax + s) : ix.node_cnt;  // SeqIdToTaxId
    }
    res.n_assign = tax_reduce(ix, best, nb, p.max_result, tmp, out_ids, err_flags);
    res.by_rank = 1;
  }
  if (nb_out) *nb_out = nb;  // best[0..nb) now holds the compact tax ids that were reduced (by_rank) or the seq ids
}

// ---------------------------------------------------------------------------
// SDUST (Dustmasker.hpp), window 64, threshold 20, 5-symbol alphabet
// ---------------------------------------------------------------------------
//
// The reference keeps the list P of "perfect intervals" as a std::vector that can
// hold >1700 entries on low-complexity reads.  Only two facts about P are ever
// used (Dustmasker.hpp:139-168, :200-224):
//   (1) when the window start passes s, the LONGEST interval starting at s is
//       masked (P.back()), and every interval starting at s is dropped;
//   (2) FindPerfect compares the candidate suffix with the best score/length
//       ratio among intervals whose start is >= the candidate's start.  A new
//       interval is only inserted when its ratio is >= that maximum, so per start
//       the most recent insertion carries both the largest end and the largest
//       ratio of that start.
// Interval ends are always the current window end, so they grow monotonically.
// Hence one slot per start position (64 slots, indexed by start mod 64) is an
// exact replacement for P: {score, span, end-start} of the latest insertion.
// All ratio tests are cross-multiplications of positive integers (span >= 1 and
// score >= 3 for every stored interval), so which of two equal ratios is kept
// does not change any outcome.

// A per-thread byte array stored as a column of 32-bit words with a stride of SW
// words between consecutive words: SW = 1 is a plain array (host / local memory);
// SW = blockDim.x puts every thread's array in its own shared-memory bank, so the
// data-dependent counter updates of a warp never conflict.
template <int SW>
struct ByteCol {
  unsigned char *base;
  CFR_HD unsigned char &operator[](int j) const { return base[(j >> 2) * (SW * 4) + (j & 3)]; }
};

// same column layout for 16-bit entries
template <int SW>
struct HalfCol {
  unsigned char *base;
  CFR_HD unsigned short &operator[](int j) const {
    return *reinterpret_cast<unsigned short *>(base + (j >> 1) * (SW * 4) + (j & 1) * 2);
  }
};

template <int SW>
struct DustStateT {
  // per triplet (base-5 index, 125 used): low byte = count in the window (cw), high byte = count
  // in its suffix v (cv); one load / store updates both
  HalfCol<SW> cc;
  ByteCol<SW> win;             // ring buffer of triplet indices; size <= 62
  unsigned short p_score[64];  // per start slot: score of the latest interval
  unsigned char p_span[64];    // its end - start - 2
  unsigned char p_len[64];     // its end - start
  u64 p_valid;
  int head, size;
  int rv, rw, lv;
};

// plain-array instantiation (host simulation, tests)
struct DustState : DustStateT<1> {
  alignas(4) unsigned char cc_store[256];
  alignas(4) unsigned char win_store[64];
  CFR_HD DustState() {
    cc.base = cc_store;
    win.base = win_store;
  }
};

// DUST input: the mate's bases as uploaded (codes + original N bits)
struct DustIn {
  const u64 *codes;
  const u32 *nmask;  // N bits before masking
  u64 base;
  u64 cw = 0;
  u32 mw = 0;
  u64 widx = ~0ull;
  CFR_HD int operator()(int i) {  // 0..3, or 4 = the catch-all fifth symbol
    const u64 q = base + (u64)i;
    const u64 wi = q >> 5;
    if (wi != widx) {
      cw = ld64(codes + wi);
      mw = ld32(nmask + wi);
      widx = wi;
    }
    const int sh = (int)(q & 31);
    if ((mw >> sh) & 1u) return 4;
    return (int)((cw >> (2 * sh)) & 3ull);
  }
};

// DUST output: N bits of the working mask (+ an optional copy of just the masked
// intervals, used by the parity diagnostics).  Mask words are shared between
// neighbouring reads, hence the atomics.
struct DustOut {
  u32 *mask;
  u32 *dust_bits;  // may be nullptr
  u64 base;
  CFR_HD void set_range(int s, int e) const {  // positions s..e of the mate, inclusive
    u64 q = base + (u64)s;
    const u64 qe = base + (u64)e;
    while (q <= qe) {
      const u64 wi = q >> 5;
      const int lo = (int)(q & 31);
      const u64 wend = (wi << 5) + 31;
      const int hi = (int)((qe < wend ? qe : wend) & 31);
      const u32 bits = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
#if defined(__CUDA_ARCH__)
      atomicOr(mask + wi, bits);
      if (dust_bits) atomicOr(dust_bits + wi, bits);
#else
      mask[wi] |= bits;
      if (dust_bits) dust_bits[wi] |= bits;
#endif
      q = wend + 1;
    }
  }
};

template <int SW>
CFR_HD int dust_win_at(const DustStateT<SW> &d, int i) { return d.win[(d.head + i) & 63]; }

// Dustmasker::SaveMaskedRegions for the single start that can leave the window
template <int SW>
CFR_HD void dust_evict(DustStateT<SW> &d, const DustOut &out, int seg_off, int start) {
  const int slot = start & 63;
  if ((d.p_valid >> slot) & 1ull) {
    out.set_range(seg_off + start, seg_off + start + d.p_len[slot]);
    d.p_valid &= ~(1ull << slot);
  }
}

// ---- SDust (Dustmasker.hpp:312-354) split into resumable pieces so that the
// kernel can run it as a warp-synchronous state machine (cfr_pipeline.cuh) ----

// start of SDust on a segment: clears the counters, primes the first two bases
template <int SW>
CFR_HD void dust_seg_init(DustIn &in, int seg_off, DustStateT<SW> &d, int &c1, int &c2) {
  for (int i = 0; i < 128; i += 2) *reinterpret_cast<u32 *>(&d.cc[i]) = 0;
  d.head = d.size = 0;
  d.rv = d.rw = d.lv = 0;
  d.p_valid = 0;
  c1 = in(seg_off);
  c2 = in(seg_off + 1);
}

CFR_HD int dust_wstart(int wfinish) { return wfinish + 1 > 64 ? wfinish + 1 - 64 : 0; }

// one iteration of SDust's main loop (:330-340) up to the point where the suffix v may
// have to be shrunk (ShiftWindow's inner loop, :125-135); returns true when it must.
// `t` receives the triplet that entered the window.
template <int SW>
CFR_HD bool dust_step(DustIn &in, const DustOut &out, int seg_off, int wfinish, DustStateT<SW> &d, int &c1, int &c2,
                      int &t) {
  const int W = 64, T = 20;
  const int wstart = dust_wstart(wfinish);
  if (wstart > 0) dust_evict(d, out, seg_off, wstart - 1);
  const int c3 = in(seg_off + wfinish);
  t = c1 * 25 + c2 * 5 + c3;
  c1 = c2;
  c2 = c3;
  // ShiftWindow (Dustmasker.hpp:106-136)
  if (d.size >= W - 2) {
    const int old = d.win[d.head];
    unsigned short &eo = d.cc[old];
    int e = (int)eo - 1;  // --cw[old]
    d.rw -= e & 0xff;
    d.head = (d.head + 1) & 63;
    --d.size;
    if (d.lv > d.size) {
      e -= 0x100;  // --cv[old]
      d.rv -= e >> 8;
      --d.lv;
    }
    eo = (unsigned short)e;
  }
  d.win[(d.head + d.size) & 63] = (unsigned char)t;
  ++d.size;
  ++d.lv;
  unsigned short &et = d.cc[t];
  const int e = et;
  d.rw += e & 0xff;
  const int cvt = e >> 8;
  d.rv += cvt;
  et = (unsigned short)(e + 0x101);  // ++cw[t], ++cv[t]
  return (cvt + 1) * 10 > 2 * T;
}

// the suffix no longer satisfies max c(v) <= 2T: drop its head up to the first t (:127-134)
template <int SW>
CFR_HD void dust_shrink(DustStateT<SW> &d, int t) {
  // (the window entry of the next turn is fetched before the counter update of this one: the loop is a
  // chain of dependent shared-memory accesses and this takes one of them off it)
  int s_next = dust_win_at(d, d.size - d.lv);
  for (;;) {
    const int s = s_next;
    if (d.lv > 1) s_next = dust_win_at(d, d.size - d.lv + 1);
    unsigned short &es = d.cc[s];
    const int e = (int)es - 0x100;  // --cv[s]
    es = (unsigned short)e;
    d.rv -= e >> 8;
    --d.lv;
    if (s == t) break;
  }
}

// SDust :340 -- does the window hold a candidate perfect interval?
template <int SW>
CFR_HD bool dust_needs_find_perfect(const DustStateT<SW> &d) { return d.rw * 10 > d.lv * 20; }

// FindPerfect (Dustmasker.hpp:173-242) over the per-start slots
template <int SW>
CFR_HD void dust_find_perfect(int wfinish, DustStateT<SW> &d) {
  const int T = 20;
  const int wstart = dust_wstart(wfinish);
  int rv = d.rv;
  int max_score = 0, max_cnt = 1;
  int folded = wstart + d.size;  // starts >= folded have been folded into (max_score, max_cnt)
  const int first = d.size - d.lv - 1;
  int i = first;
  int tt_next = first >= 0 ? dust_win_at(d, first) : 0;
  for (; i >= 0; --i) {
    const int span = d.size - i - 1;
    // A candidate needs rv * 10 > T * span.  rv counts pairs inside a part of the window, so it
    // never exceeds rw (the pairs of the whole window), and span only grows from here: once
    // T * span >= rw * 10 no candidate is left and the rest of the scan cannot change anything.
    if (d.rw * 10 <= T * span) break;
    const int tt = tt_next;
    if (i > 0) tt_next = dust_win_at(d, i - 1);  // next turn's window entry, off the dependent chain
    unsigned short &ett = d.cc[tt];
    rv += ett >> 8;
    ett = (unsigned short)(ett + 0x100);  // ++cv[tt]
    if (rv * 10 > T * span) {
      const int start = i + wstart;
      while (folded > start) {  // the scan of P from its head (:203-211)
        --folded;
        const int slot = folded & 63;
        if ((d.p_valid >> slot) & 1ull) {
          if ((u64)d.p_score[slot] * (u64)max_cnt > (u64)max_score * (u64)d.p_span[slot]) {
            max_score = d.p_score[slot];
            max_cnt = d.p_span[slot];
          }
        }
      }
      if (rv * max_cnt >= max_score * span) {
        max_score = rv;
        max_cnt = span;
        const int slot = start & 63;
        d.p_score[slot] = (unsigned short)rv;
        d.p_span[slot] = (unsigned char)span;
        d.p_len[slot] = (unsigned char)(wstart + d.size + 1 - start);
        d.p_valid |= 1ull << slot;
      }
    }
  }
  for (int q = first; q > i; --q) d.cc[dust_win_at(d, q)] -= 0x100;  // undo the visited ++cv
}

// the tail loop of Dustmasker.hpp:343-350 (n = segment length): saves every remaining start
template <int SW>
CFR_HD void dust_seg_tail(const DustOut &out, int seg_off, int n, DustStateT<SW> &d) {
  if (!d.p_valid) return;
  int base = dust_wstart(n);
  if (base > 0) --base;  // the last in-loop save ran with wstart(n-1): start base-1 may remain
  // every valid slot exactly once: slot s holds the start in [base, base + 64) that is congruent to s
  u64 pv = d.p_valid;
  while (pv) {
    const int lo32 = (u32)pv ? ctz32((u32)pv) : 32 + ctz32((u32)(pv >> 32));
    pv &= pv - 1ull;
    dust_evict(d, out, seg_off, base + ((lo32 - base) & 63));
  }
}

// true when the mate holds no non-ACGT base at all (then MaskWithBuffer hands the whole
// mate to SDust as one segment: no leading Ns to skip, no N run to split at)
CFR_HD bool dust_all_acgt(const u32 *nmask, u64 base, int n) {
  if (n <= 0) return true;
  const u64 q0 = base, q1 = base + (u64)n - 1;
  const u64 w0 = q0 >> 5, w1 = q1 >> 5;
  u32 any = 0;
  for (u64 w = w0; w <= w1; ++w) {
    u32 m = ld32(nmask + w);
    if (w == w0) m &= ~((1u << (q0 & 31)) - 1u);
    if (w == w1 && (q1 & 31) != 31) m &= (1u << ((q1 & 31) + 1)) - 1u;
    any |= m;
  }
  return any == 0;
}

// ---- register-only screen: can SDUST mask anything in this (all-ACGT) mate? ----
//
// FindPerfect (Dustmasker.hpp:173-242) masks a run of k triplets S only when its
// score exceeds the threshold: sum_t C(c_t,2) * 10 > T * (k-1) with T = 20, i.e.
//     f(S) = sum_t C(c_t,2) - 2(k-1) = 2 + sum_t g(c_t) > 0,   g(c) = c(c-5)/2,
// c_t = multiplicity of triplet class t in S.  g(1..4) = -2,-3,-3,-2, g(5) = 0 and g grows
// from there, so S needs "heavy" classes (c_t >= 5), and every other element of S costs
// at least 1/2: with heavy multiplicities c_1..c_h
//     |S| <= L(S) = 2 + sum_i c_i (c_i - 4)                                   (*)
// (5 identical triplets = 7 identical bases give L = 7; one class with 6 copies L = 14).
// Take the LAST heavy element of S, entering the window at time tau with class t.  The
// window (62 triplets) then holds all of S up to tau, so every heavy class of S has a
// window count >= its multiplicity in S, count[t] >= 5, and by (*) -- L evaluated on the
// window's heavy classes only grows -- the >= 5 copies of t in S lie among the last L
// triplets.  Hence, whenever a triplet enters whose class then has >= 5 copies in the
// window: compute L from the window's heavy classes and count the copies of t among
// the last L triplets; if that count never reaches 5 (and the conservative outs below
// never fire) no perfect interval exists and SDUST masks nothing.
// The window counts are kept bit-sliced in three 64-bit registers (one bit per triplet
// class and plane), so the screen touches no memory beyond the packed read itself.
// Conservative outs: a count of 7 (the planes would wrap), three heavy classes at
// once, or L > 30 (the history examined is one 32-base word).
// Returns true when the full SDUST must run.
// (lo, hi) >> s for 0 <= s < 32, low word
CFR_HD u32 funnel_r(u32 lo, u32 hi, int s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, (unsigned)s);
#else
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}

CFR_HD u32 dust_stream_word(const u64 *codes, u64 q0, int k) {  // 16 bases from q0 + 16k
  const u32 *c32 = reinterpret_cast<const u32 *>(codes);
  const u64 q = q0 + 16ull * (u64)k;
  const u64 wi = q >> 4;
  return funnel_r(ld32(c32 + wi), ld32(c32 + wi + 1), 2 * (int)(q & 15));
}

// bit 2p set iff base p of the 32-base word w equals the 2-bit code b
CFR_HD u64 dust_base_eq(u64 w, u64 b) {
  const u64 REP = 0x5555555555555555ull;
  const u64 x = w ^ (b * REP);
  return ~(x | (x >> 1)) & REP;
}

// the rare part of the screen: class t (mask m) has >= 5 copies in the window after triplet
// i entered.  `hist` = the 32 bases ending with triplet i (triplet i-d starts at base 29-d).
// One copy of this code for all unrolled steps of the screen loop (instruction cache).
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
bool dust_screen_event(u64 b0, u64 b1, u64 b2, u64 m, u32 t, u64 hist, int i) {
  if (b2 & b1 & b0 & m) return true;  // 7 copies
  u64 hh = b2 & (b0 | b1);            // heavy classes
  int L = 2;
  for (int q = 0; q < 2; ++q) {
    if (!hh) break;
    const u64 bit = hh & (~hh + 1ull);
    hh ^= bit;
    const int c = 4 + ((b1 & bit) ? 2 : 0) + ((b0 & bit) ? 1 : 0);
    L += c * (c - 4);
  }
  if (hh || L > 30) return true;  // three heavy classes, or more history than one word
  const int leff = L < i + 1 ? L : i + 1;
  const u64 occ = dust_base_eq(hist, t & 3u) & (dust_base_eq(hist, (t >> 2) & 3u) >> 2) &
                  (dust_base_eq(hist, (t >> 4) & 3u) >> 4);
  return popc64(occ >> (2 * (30 - leff))) >= 5;
}

CFR_HD bool dust_screen(const u64 *codes, u64 q0, int len) {
  const int nt = len - 2;  // triplets; fewer than 5 can never reach the threshold
  if (nt < 5) return false;
  u64 b0 = 0, b1 = 0, b2 = 0;  // bit planes of the per-class window counts
  // sliding view of the mate in 16-base words: w0 holds the bases of the current block,
  // wm4..wm1 the four blocks before it (the triplet leaving the 62-triplet window starts
  // 62 bases back), w1 the next one
  u32 wm4 = 0, wm3 = 0, wm2 = 0, wm1 = 0, w0 = dust_stream_word(codes, q0, 0);
  bool need = false;
  for (int k = 0; 16 * k < nt; ++k) {
    const u32 w1 = 16 * (k + 1) < len ? dust_stream_word(codes, q0, k + 1) : 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 16; ++j) {
      const int i = 16 * k + j;
      if (i < nt) {
        if (i >= 62) {  // triplet i-62 starts at base j+2 of block k-4
          const int jj = j + 2;
          const u32 o = (jj <= 13 ? wm4 >> (2 * jj) : jj <= 15 ? funnel_r(wm4, wm3, 2 * jj) : wm3 >> (2 * (jj - 16))) & 63u;
          const u64 mo = 1ull << o;
          const u64 br0 = ~b0 & mo;  // borrow chain of count[o] -= 1
          b0 ^= mo;
          const u64 br1 = ~b1 & br0;
          b1 ^= br0;
          b2 ^= br1;
        }
        const u32 t = (j <= 13 ? w0 >> (2 * j) : funnel_r(w0, w1, 2 * j)) & 63u;
        const u64 m = 1ull << t;
        const u64 c0 = b0 & m;  // carry chain of count[t] += 1
        b0 ^= m;
        const u64 c1 = b1 & c0;
        b1 ^= c0;
        b2 ^= c1;
        if (b2 & (b0 | b1) & m) {  // class t has >= 5 copies in the window
          // the 32 bases ending with triplet i: from base j+3 of block k-2 on
          u32 lo, hi;
          if (j <= 12) {
            lo = funnel_r(wm2, wm1, 2 * (j + 3));
            hi = funnel_r(wm1, w0, 2 * (j + 3));
          } else {
            lo = funnel_r(wm1, w0, 2 * (j - 13));
            hi = funnel_r(w0, w1, 2 * (j - 13));
          }
          if (dust_screen_event(b0, b1, b2, m, t, (u64)lo | ((u64)hi << 32), i)) need = true;
        }
      }
    }
    wm4 = wm3;
    wm3 = wm2;
    wm2 = wm1;
    wm1 = w0;
    w0 = w1;
  }
  return need;
}

// MaskWithBuffer's segment finder (Dustmasker.hpp:369-401): starting at cursor i,
// the next run [i, last_valid] to hand to SDust, and the cursor after it
CFR_HD void dust_next_segment(DustIn &in, int n, int i, int &last_valid, int &next_i) {
  const int W = 64;
  int n_count = 0, j;
  last_valid = i;
  for (j = i; j < n; ++j) {
    if (in(j) == 4)
      ++n_count;
    else {
      if (n_count > W) break;
      last_valid = j;
      n_count = 0;
    }
  }
  next_i = j;
}

// Dustmasker::MaskWithBuffer + the in-place masking of CentrifugerClass.cpp:281-289,
// sequential form (host simulation / reference for the state machine).
// `in` is the mate as uploaded, `out` the working N mask the searches read.
template <int SW>
CFR_HD void dust_task(DustIn &in, int n, const DustOut &out, DustStateT<SW> &d) {
  if (n < 3) return;
  int i = 0;
  while (i < n && in(i) == 4) ++i;
  while (i < n) {
    int last_valid, next_i;
    dust_next_segment(in, n, i, last_valid, next_i);
    const int seg_n = last_valid - i + 1;
    if (last_valid > i && seg_n >= 3) {
      int c1, c2;
      dust_seg_init(in, i, d, c1, c2);
      for (int wfinish = 2; wfinish < seg_n; ++wfinish) {
        int t;
        if (dust_step(in, out, i, wfinish, d, c1, c2, t)) dust_shrink(d, t);
        if (dust_needs_find_perfect(d)) dust_find_perfect(wfinish, d);
      }
      dust_seg_tail(out, i, seg_n, d);
    }
    i = next_i;
  }
}

}  // namespace cfrb200

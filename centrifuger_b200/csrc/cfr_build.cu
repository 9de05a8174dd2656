// cfr_build.cu -- FM-index construction on the GPU.
//
// Writes <prefix>.1.cfr in the reference's grammar (FMIndex::Save, FMIndex.hpp:571-586) from a
// 2-bit text, replacing on the reference side
//   FMBuilder::Build                      compactds/FMBuilder.hpp:444   (blockwise suffix sorting, BWT, sampled
//                                                                        SA, lookup table, boundary rows)
//   Builder::TransformSampledSAToSeqId    Builder.hpp:27                 (samples become sequence ids)
//   Sequence_RunBlock::Init               compactds/Sequence_RunBlock.hpp:231 (block size, run-block split)
//   Sequence_WaveletTree::Init            compactds/Sequence_WaveletTree.hpp:196
//   DS_Rank9::Init                        compactds/DS_Rank.hpp:204
//   FMIndex::Init / InitAuxData / Save    compactds/FMIndex.hpp:256 / :196 / :571
// The file is byte-identical to what the reference's centrifuger-build writes for the same text
// (tests/test_builder.py compares both, `_space` bookkeeping fields included).
//
// B200-first design, not a translation of the reference's difference-cover sorter:
//   * suffixes are cut into batches by splitter keys drawn from a sample (no histogram pass);
//   * a batch is sorted by 31-base keys with device-wide radix sorts; ties are refined by the next
//     31 bases, restricted to the still-tied suffixes, until every suffix stands alone.  The end of
//     the text compares smaller than any base (SuffixArrayGenerator.hpp:297-306): a key whose window
//     starts past the end carries the suffix length instead of bases and sorts first;
//   * everything derived from the sorted batch (BWT symbols, every sampleRate-th row's sequence id,
//     first row of every W-mer, rows of the sequence boundaries) is produced by passes over the batch
//     while it is in HBM; nothing but the finished arrays leaves the device;
//   * the run-block split, the two wavelet trees and the seven rank9 directories are stream
//     compactions over the BWT: count pass, prefix scan, emit pass.
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <map>
#include <thread>

#include "../../include/centrifuger_b200_build.h"
#include "cfr_build_backend.cuh"

namespace cfrbuild {

// ---------------------------------------------------------------------------
// text access: 2-bit codes, base i in bits [2(i&31), 2(i&31)+2) of word i>>5; at least two zero
// words follow the last base, and the unused bits of the last word are zero
// ---------------------------------------------------------------------------
BK_HD u32 text_base(const u64 *w, u64 i) { return (u32)((w[i >> 5] >> (2 * (i & 31))) & 3ull); }

BK_HD u64 text_bits64(const u64 *w, u64 pos) {  // bases pos .. pos+31, base pos in the low bits
  const u64 wi = pos >> 5;
  const int sh = (int)(pos & 31) * 2;
  const u64 lo = w[wi];
  return sh == 0 ? lo : (lo >> sh) | (w[wi + 1] << (64 - sh));
}

BK_D u64 rev_groups(u64 x) {  // reverse the order of the 32 two-bit groups
  x = brev64(x);
  return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// 31 bases starting at q as a 62-bit number, first base most significant (zero = 'A' past the end)
BK_D u64 key31(const u64 *w, u64 q) { return rev_groups(text_bits64(w, q)) >> 2; }

// sort key of suffix `pos` for the window of 31 bases at depth d.  A window that starts inside the
// text has bit 63 set; one that starts past the end carries the suffix length n - pos (the shorter
// suffix is the smaller one), and sorts before every live window.
BK_D u64 suffix_key(const u64 *w, u64 n, u64 pos, u64 d) {
  const u64 q = pos + 31ull * d;
  return q < n ? (1ull << 63) | key31(w, q) : (n - pos);
}

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Log {
  int verbose;
  double t0;
  void operator()(const char *fmt, ...) const __attribute__((format(printf, 2, 3)));
};
void Log::operator()(const char *fmt, ...) const {
  if (!verbose) return;
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "[cfr-build %7.2fs] ", now_s() - t0);
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}

// ---------------------------------------------------------------------------
// Stage 1: suffix sorting in batches
// ---------------------------------------------------------------------------
struct SortWork {
  u64 cap = 0;
  Buf<u64> pA, pB, kA, kB, sa;
  Buf<u32> iA, iB, gA, gB, sA, sB, h, hg, off;
  void alloc(u64 c) {
    cap = c;
    pA.alloc(c); pB.alloc(c); kA.alloc(c); kB.alloc(c); sa.alloc(c);
    iA.alloc(c); iB.alloc(c); gA.alloc(c); gB.alloc(c); sA.alloc(c); sB.alloc(c);
    h.alloc(c + 1); hg.alloc(c + 1); off.alloc(c + 1);
  }
  void release() {
    pA.release(); pB.release(); kA.release(); kB.release(); sa.release();
    iA.release(); iB.release(); gA.release(); gB.release(); sA.release(); sB.release();
    h.release(); hg.release(); off.release();
    cap = 0;
  }
  static u64 bytes_per_elem() { return 5 * 8 + 9 * 4; }
};

// positions whose 31-base key lies in [lo, hi) -> out (any order); returns how many.
// Two passes over the text in chunks: count, scan, write.
u64 select_positions(const u64 *text, u64 n, u64 lo, u64 hi, bool all, u64 *out, u64 cap, Buf<u64> &chunk_cnt) {
  const u64 CH = 4096;
  const u64 nchunk = (n + CH - 1) / CH;
  if (chunk_cnt.n < nchunk + 1) chunk_cnt.alloc(nchunk + 1);
  u64 *cnt = chunk_cnt.p;
  const u64 mask62 = (1ull << 62) - 1;
  if (all) {
    if (n > cap) throw BuildError(-1, "builder: batch capacity too small");
    par_for(n, [=] BK_LAMBDA(u64 i) { out[i] = i; });
    return n;
  }
  par_for(nchunk, [=] BK_LAMBDA(u64 c) {
    const u64 i0 = c * CH, i1 = i0 + CH < n ? i0 + CH : n;
    u64 k = key31(text, i0), m = 0;
    for (u64 i = i0; i < i1; ++i) {
      m += (k >= lo && k < hi) ? 1 : 0;
      k = ((k << 2) | (u64)text_base(text, i + 31)) & mask62;
    }
    cnt[c] = m;
  });
  const u64 total = exclusive_sum(cnt, cnt, nchunk);
  if (total > cap) return total;  // caller re-plans with more batches
  par_for(nchunk, [=] BK_LAMBDA(u64 c) {
    const u64 i0 = c * CH, i1 = i0 + CH < n ? i0 + CH : n;
    u64 k = key31(text, i0), o = cnt[c];
    for (u64 i = i0; i < i1; ++i) {
      if (k >= lo && k < hi) out[o++] = i;
      k = ((k << 2) | (u64)text_base(text, i + 31)) & mask62;
    }
  });
  return total;
}

// Sort the m suffixes in W.pA; the sorted positions end up in W.sa[0..m).
void sort_batch(const u64 *text, u64 n, SortWork &W, u64 m, const Log &log, u64 &max_depth) {
  u64 *P = W.pA.p, *Palt = W.pB.p, *K = W.kA.p, *Kalt = W.kB.p;
  u32 *I = W.iA.p, *Ialt = W.iB.p, *G = W.gA.p, *Galt = W.gB.p, *S = W.sA.p, *Salt = W.sB.p;
  u32 *H = W.h.p, *HG = W.hg.p, *OFF = W.off.p;
  u64 *SA = W.sa.p;
  {
    u64 *k = K;
    const u64 *p = P;
    par_for(m, [=] BK_LAMBDA(u64 j) { k[j] = suffix_key(text, n, p[j], 0); });
    sort_pairs(K, Kalt, P, Palt, m, 0, 62);
    u32 *s = S, *g = G;
    par_for(m, [=] BK_LAMBDA(u64 j) {
      s[j] = (u32)j;
      g[j] = 0;
    });
  }
  u64 mc = m, d = 0;
  for (;;) {
    {
      const u64 *k = K, *p = P;
      const u32 *g = G, *s = S;
      u32 *h = H, *hg = HG, *off = OFF;
      par_for(mc, [=] BK_LAMBDA(u64 j) {
        const u32 head = (j == 0 || g[j] != g[j - 1] || k[j] != k[j - 1]) ? 1u : 0u;
        h[j] = head;
        hg[j] = head ? (u32)j : 0u;
      });
      par_for(mc, [=] BK_LAMBDA(u64 j) {
        const bool single = h[j] && (j + 1 == mc || h[j + 1]);
        if (single) SA[s[j]] = p[j];
        off[j] = single ? 0u : 1u;
      });
    }
    inclusive_max_u32(HG, HG, mc);
    const u64 mu = exclusive_sum(OFF, OFF, mc);
    if (mu == 0) break;
    ++d;
    {
      const u64 *p = P;
      const u32 *s = S, *h = H, *hg = HG, *off = OFF;
      u64 *p2 = Palt, *k2 = Kalt;
      u32 *s2 = Salt, *g2 = Galt, *i2 = I;
      par_for(mc, [=] BK_LAMBDA(u64 j) {
        const bool single = h[j] && (j + 1 == mc || h[j + 1]);
        if (single) return;
        const u32 o = off[j];
        p2[o] = p[j];
        s2[o] = s[j];
        g2[o] = hg[j];
        k2[o] = suffix_key(text, n, p[j], d);
        i2[o] = o;
      });
    }
    // now: Palt / Salt / Galt / Kalt hold the mu unresolved suffixes in slot order, I = identity
    std::swap(P, Palt);
    std::swap(S, Salt);
    std::swap(G, Galt);
    std::swap(K, Kalt);
    sort_pairs(K, Kalt, I, Ialt, mu, 0, 64);  // by the next 31 bases ...
    {
      const u32 *g = G, *i = I;
      u32 *gs = Galt;
      par_for(mu, [=] BK_LAMBDA(u64 j) { gs[j] = g[i[j]]; });
    }
    int gbits = 1;
    while (gbits < 32 && (1ull << gbits) < mc) ++gbits;
    {
      u32 *gk = Galt, *gk_alt = G;  // G (slot order) is not needed any more
      sort_pairs(gk, gk_alt, I, Ialt, mu, 0, gbits);  // ... then, stably, by the tie group they came from
      G = gk;
      Galt = gk_alt;
    }
    {
      const u64 *p = P;
      const u32 *i = I;
      u64 *p2 = Palt, *k2 = K;  // K's content (sorted keys, wrong order now) is rebuilt from the text
      par_for(mu, [=] BK_LAMBDA(u64 j) {
        const u64 pos = p[i[j]];
        p2[j] = pos;
        k2[j] = suffix_key(text, n, pos, d);
      });
    }
    std::swap(P, Palt);
    mc = mu;
    if (d > max_depth) max_depth = d;
    if (d > 100000000ull) throw BuildError(-1, "builder: tie refinement did not converge");
  }
  (void)log;
}

// ---------------------------------------------------------------------------
// Stage 2: Sequence_RunBlock::Init + the wavelet trees + rank9, from the finished BWT
// ---------------------------------------------------------------------------

// Sequence_RunBlock::GetRunBlockLength (Sequence_RunBlock.hpp:26-47) on a window of symbols that starts
// at text row `base`: S(i) = win[i - base]
struct SymWindow {
  const unsigned char *win;
  u64 base, len;  // symbols available: rows [base, base + len)
  unsigned char at(u64 i) const { return win[i - base]; }
};
static u64 run_block_length(const SymWindow &S, u64 n, u64 s, u64 e, u64 b) {
  u64 total = 0;
  for (u64 i = s; i <= e && i < n; i += b) {
    const unsigned char c = S.at(i);
    bool run = true;
    u64 j;
    for (j = i + 1; j < i + b && j < n; ++j)
      if (S.at(j) != c) {
        run = false;
        break;
      }
    if (run) total += (j - i);
  }
  return total;
}

// Sequence_RunBlock::ComputeBlockSize (:124-168) with EstimateSpace (:49-80) and
// EstimateAverageRunLength (:83-121).  `chunks[k]` holds the symbols of the k-th test window.
struct BlockSizeEstimator {
  u64 n;
  u64 len = 1024, test_cases = 1024;
  bool whole;                    // len * test_cases >= n: one window covering everything
  std::vector<SymWindow> chunks; // sampled windows (each long enough for a block that starts at its last row)
  std::vector<std::vector<unsigned char>> store;

  u64 stride() const { return (n + test_cases - 1) / test_cases; }
  u64 estimate_space(u64 b, int abits) const {
    u64 rbl = 0, m = 0;
    if (whole) {
      rbl = run_block_length(chunks[0], n, 0, n - 1, b);
      m = n;
    } else {
      u64 k = 0;
      for (u64 i = 0; i < n; i += stride(), ++k) {
        u64 e = i + len - 1;
        if (e >= n) e = n - 1;
        rbl += run_block_length(chunks[k], n, i, i + len - 1, b);
        m += e - i + 1;
      }
    }
    const u64 rbc = (rbl + b - 1) / b;
    if (b > 1) return (m + b - 1) / b + (u64)abits * (rbc + m - rbl);
    return (u64)abits * m;
  }
  double average_run_length() const {
    u64 r = 0, m = 0;
    if (whole) {
      unsigned char c = chunks[0].at(0);
      for (u64 i = 1; i < n; ++i) {
        const unsigned char t = chunks[0].at(i);
        if (t != c) {
          ++r;
          c = t;
        }
      }
      ++r;
      m = n;
    } else {
      u64 k = 0;
      for (u64 i = 0; i < n; i += stride(), ++k) {
        u64 e = i + len - 1;
        if (e >= n) e = n - 1;
        unsigned char c = chunks[k].at(i);
        for (u64 j = i + 1; j <= e; ++j) {
          const unsigned char t = chunks[k].at(j);
          if (t != c) {
            ++r;
            c = t;
          }
        }
        ++r;
        m += e - i + 1;
      }
    }
    return (double)m / (double)r;
  }
  u64 compute(int abits) const {
    u64 best_space = 0, best = 0;
    const u64 m = len;
    for (u64 i = 1; i <= m; i *= 2) {
      const u64 sp = estimate_space(i, abits);
      if (best_space == 0 || sp < best_space) {
        best_space = sp;
        best = i;
      }
    }
    if (best <= m) {
      if (best >= 2) {
        const u64 sp = estimate_space(best / 2 * 3, abits);
        if (sp < best_space) {
          best_space = sp;
          best = best / 2 * 3;
        }
      }
      const u64 test = (u64)std::ceil(std::sqrt(average_run_length()));
      if (test > 2) {
        const u64 sp = estimate_space(test, abits);
        if (sp < best_space) {
          best_space = sp;
          best = test;
        }
      }
    }
    return best;
  }
};

// one rank9 bitvector under construction on the device
struct DevBits {
  Buf<u64> B, R;
  u64 nbits = 0, words = 0, rwords = 0;
  void alloc(u64 bits) {
    nbits = bits;
    words = (bits + 63) / 64;
    rwords = ((words + 7) / 8) * 2;
    B.alloc(words + 2);
    B.zero();
  }
};

// DS_Rank9::Init (DS_Rank.hpp:204-243): per 8 words an absolute count and seven 9-bit relative counts;
// the unused slots of the last block repeat the final count unless that block holds a single word
void build_rank9(DevBits &v, Buf<u64> &scratch) {
  if (v.nbits == 0) return;
  const u64 words = v.words, blocks = (words + 7) / 8;
  v.R.alloc(v.rwords + 2);
  v.R.zero();
  if (scratch.n < blocks + 1) scratch.alloc(blocks + 1);
  const u64 *B = v.B.p;
  u64 *bc = scratch.p, *R = v.R.p;
  par_for(blocks, [=] BK_LAMBDA(u64 b) {
    u64 c = 0;
    for (u64 i = b * 8; i < b * 8 + 8 && i < words; ++i) c += (u64)popc64(B[i]);
    bc[b] = c;
  });
  // the scans of the builder take at most 2^31 - 1 entries: directories of more than 2^34 words
  // (2^40 bits) do not occur below the occ-sector limit of n < 2^40 rows
  exclusive_sum(bc, bc, blocks);
  par_for(blocks, [=] BK_LAMBDA(u64 b) {
    R[2 * b] = bc[b];
    u64 sub = 0, local = 0;
    u64 i = b * 8;
    for (; i < b * 8 + 8 && i < words; ++i) {
      const int br = (int)(i & 7);
      if (br > 0) sub |= local << ((br - 1) * 9);
      local += (u64)popc64(B[i]);
    }
    if (i == words && ((i - 1) & 7) > 0)
      for (; i & 7; ++i) sub |= local << (((int)(i & 7) - 1) * 9);
    R[2 * b + 1] = sub;
  });
}

// append `cnt` (<= 64) bits to a bit stream at bit offset `at`; the stream's words may be shared with a
// neighbouring writer, hence the atomics
BK_D void put_bits(u64 *dst, u64 at, u64 bits, int cnt) {
  if (cnt <= 0) return;
  const int sh = (int)(at & 63);
  a_or64(dst + (at >> 6), bits << sh);
  if (sh + cnt > 64) a_or64(dst + (at >> 6) + 1, bits >> (64 - sh));
}

// a small register bit buffer that flushes whole 64-bit chunks through put_bits
struct BitWriter {
  u64 *dst;
  u64 at;   // bit offset of the first buffered bit
  u64 acc;
  int fill;
  BK_D void init(u64 *d, u64 start) {
    dst = d;
    at = start;
    acc = 0;
    fill = 0;
  }
  BK_D void push(u32 bit) {
    acc |= (u64)bit << fill;
    if (++fill == 64) {
      put_bits(dst, at, acc, 64);
      at += 64;
      acc = 0;
      fill = 0;
    }
  }
  BK_D void flush() {
    put_bits(dst, at, acc, fill);
    at += (u64)fill;
    acc = 0;
    fill = 0;
  }
};

struct RunBlockParts {
  u64 b = 0, block_cnt = 0;
  DevBits type;     // _useRunBlock
  DevBits wt[2][3]; // [0] = _waveletSeq (plain blocks), [1] = _runBlockSeq; node 0 root, 1 = A/C, 2 = G/T
  u64 seq_n[2] = {0, 0};
};

// rows are processed in units of UNIT rows by one thread each
static const u64 UNIT = 4096;

void build_run_blocks(const u64 *bwt, u64 n, u64 b, RunBlockParts &out, const Log &log) {
  out.b = b;
  const u64 block_cnt = (n + b - 1) / b;
  out.block_cnt = block_cnt;
  out.type.alloc(block_cnt);
  u64 *TB = out.type.B.p;
  // ---- block types (Sequence_RunBlock.hpp:250-268): a block is a run block when all its symbols are equal
  if (b <= 1024) {
    const u64 twords = out.type.words;
    par_for(twords, [=] BK_LAMBDA(u64 w) {
      u64 bits = 0;
      for (u64 k = w * 64; k < w * 64 + 64 && k < block_cnt; ++k) {
        const u64 r0 = k * b, r1 = r0 + b < n ? r0 + b : n;
        const u32 c = text_base(bwt, r0);
        bool run = true;
        for (u64 r = r0 + 1; r < r1; ++r)
          if (text_base(bwt, r) != c) {
            run = false;
            break;
          }
        if (run) bits |= 1ull << (k & 63);
      }
      TB[w] = bits;
    });
  } else {
    par_for(out.type.words, [=] BK_LAMBDA(u64 w) {
      const u64 left = block_cnt - w * 64;
      TB[w] = left >= 64 ? ~0ull : ((1ull << left) - 1ull);
    });
    const u64 nunit = (n + UNIT - 1) / UNIT;
    par_for(nunit, [=] BK_LAMBDA(u64 u) {
      const u64 r0 = u * UNIT, r1 = r0 + UNIT < n ? r0 + UNIT : n;
      u64 k = r0 / b;
      u32 c = text_base(bwt, k * b);
      u64 kend = (k + 1) * b;
      bool cleared = false;
      for (u64 r = r0; r < r1; ++r) {
        if (r == kend) {
          ++k;
          kend += b;
          c = text_base(bwt, r);
          cleared = false;
        }
        if (!cleared && text_base(bwt, r) != c) {
          a_and64(TB + (k >> 6), ~(1ull << (k & 63)));
          cleared = true;
        }
      }
    });
  }
  // ---- per-unit counts: elements each unit appends to the two sequences and to their A/C nodes
  const u64 nunit = (n + UNIT - 1) / UNIT;
  Buf<u64> cnt(4 * (nunit + 1));  // [plain n, run n, plain high-bit-0, run high-bit-0] per unit, as four arrays
  u64 *c_pn = cnt.p, *c_rn = cnt.p + (nunit + 1), *c_p0 = cnt.p + 2 * (nunit + 1), *c_r0 = cnt.p + 3 * (nunit + 1);
  par_for(nunit, [=] BK_LAMBDA(u64 u) {
    const u64 r0 = u * UNIT, r1 = r0 + UNIT < n ? r0 + UNIT : n;
    u64 k = r0 / b, kend = (k + 1) * b;
    bool run = (TB[k >> 6] >> (k & 63)) & 1ull;
    u64 pn = 0, rn = 0, p0 = 0, rz = 0;
    for (u64 r = r0; r < r1; ++r) {
      if (r == kend) {
        ++k;
        kend += b;
        run = (TB[k >> 6] >> (k & 63)) & 1ull;
      }
      const u32 c = text_base(bwt, r);
      if (!run) {
        ++pn;
        p0 += (c >> 1) ? 0 : 1;
      } else if (r == k * b) {
        ++rn;
        rz += (c >> 1) ? 0 : 1;
      }
    }
    c_pn[u] = pn;
    c_rn[u] = rn;
    c_p0[u] = p0;
    c_r0[u] = rz;
  });
  const u64 plain_n = exclusive_sum(c_pn, c_pn, nunit);
  const u64 run_n = exclusive_sum(c_rn, c_rn, nunit);
  const u64 plain_0 = exclusive_sum(c_p0, c_p0, nunit);
  const u64 run_0 = exclusive_sum(c_r0, c_r0, nunit);
  out.seq_n[0] = plain_n;
  out.seq_n[1] = run_n;
  log("run-block split: b = %llu, %llu blocks, %llu plain symbols, %llu run blocks", b, block_cnt, plain_n, run_n);
  const u64 sizes[2][3] = {{plain_n, plain_0, plain_n - plain_0}, {run_n, run_0, run_n - run_0}};
  for (int t = 0; t < 2; ++t)
    for (int k = 0; k < 3; ++k)
      if (sizes[t][0] > 0) out.wt[t][k].alloc(sizes[t][k]);
  u64 *P0 = out.wt[0][0].B.p, *P1 = out.wt[0][1].B.p, *P2 = out.wt[0][2].B.p;
  u64 *R0 = out.wt[1][0].B.p, *R1 = out.wt[1][1].B.p, *R2 = out.wt[1][2].B.p;
  // ---- emit: root bit = high code bit; the low bit goes to the A/C node (high 0) or the G/T node (high 1)
  par_for(nunit, [=] BK_LAMBDA(u64 u) {
    const u64 r0 = u * UNIT, r1 = r0 + UNIT < n ? r0 + UNIT : n;
    u64 k = r0 / b, kend = (k + 1) * b;
    bool run = (TB[k >> 6] >> (k & 63)) & 1ull;
    BitWriter pr, pa, pg, rr, ra, rg;
    pr.init(P0, c_pn[u]);
    pa.init(P1, c_p0[u]);
    pg.init(P2, c_pn[u] - c_p0[u]);
    rr.init(R0, c_rn[u]);
    ra.init(R1, c_r0[u]);
    rg.init(R2, c_rn[u] - c_r0[u]);
    for (u64 r = r0; r < r1; ++r) {
      if (r == kend) {
        ++k;
        kend += b;
        run = (TB[k >> 6] >> (k & 63)) & 1ull;
      }
      const u32 c = text_base(bwt, r);
      if (!run) {
        pr.push(c >> 1);
        if (c >> 1) pg.push(c & 1); else pa.push(c & 1);
      } else if (r == k * b) {
        rr.push(c >> 1);
        if (c >> 1) rg.push(c & 1); else ra.push(c & 1);
      }
    }
    pr.flush(); pa.flush(); pg.flush(); rr.flush(); ra.flush(); rg.flush();
  });
  Buf<u64> scratch;
  build_rank9(out.type, scratch);
  for (int t = 0; t < 2; ++t)
    for (int k = 0; k < 3; ++k) build_rank9(out.wt[t][k], scratch);
}

// ---------------------------------------------------------------------------
// file writer (host): the grammar of FMIndex::Save and its members' Save()
// ---------------------------------------------------------------------------
struct FileOut {
  FILE *fp;
  u64 written = 0;
  explicit FileOut(const std::string &path) {
    fp = fopen(path.c_str(), "wb");
    if (!fp) throw BuildError(-2, "cannot create " + path);
    setvbuf(fp, nullptr, _IOFBF, 8 << 20);
  }
  ~FileOut() {
    if (fp) fclose(fp);
  }
  void raw(const void *p, size_t bytes) {
    if (bytes && fwrite(p, 1, bytes, fp) != bytes) throw BuildError(-2, "short write (disk full?)");
    written += bytes;
  }
  void u64v(u64 v) { raw(&v, 8); }
  void i32v(int32_t v) { raw(&v, 4); }
  void u8v(unsigned char v) { raw(&v, 1); }
  // device words -> file, through a bounded host staging buffer
  void dev_words(const u64 *dev, u64 words) {
    const u64 CH = 8ull << 20;  // 64 MiB of words
    std::vector<u64> stage((size_t)std::min<u64>(CH, words ? words : 1));
    for (u64 o = 0; o < words; o += CH) {
      const u64 c = std::min<u64>(CH, words - o);
      bk_to_host(stage.data(), dev + o, c * 8);
      raw(stage.data(), c * 8);
    }
  }
  void close() {
    if (fp && fclose(fp) != 0) {
      fp = nullptr;
      throw BuildError(-2, "close failed (disk full?)");
    }
    fp = nullptr;
  }
};

// Alphabet::Save (Alphabet.hpp:194-205) of the plain "ACGT" coder (InitFromList, :51-69)
static void write_alphabet_acgt(FileOut &f) {
  f.u64v(4);  // _space = sizeof(ALPHABET) * n
  f.i32v(1);  // ALPHABET_CODE_PLAIN
  f.u64v(4);
  f.raw("ACGT", 4);
  int32_t code[256];
  int16_t code_len[256];
  memset(code, 0, sizeof(code));
  memset(code_len, 0, sizeof(code_len));
  const char *s = "ACGT";
  for (int i = 0; i < 4; ++i) {
    code[(int)s[i]] = i;
    code_len[(int)s[i]] = 2;
  }
  f.raw(code, sizeof(code));
  f.raw(code_len, sizeof(code_len));
}

// `_space` of a Bitvector_Plain after Init() (Bitvector_Plain.hpp:104-114): the words, plus rank9's
// directory; the select structure is empty (DS_SELECT_SPEED_NO)
static u64 bv_space(const DevBits &v) { return v.words * 8 + v.rwords * 8; }

// Bitvector_Plain::Save (Bitvector_Plain.hpp:182-196)
static void write_bitvector(FileOut &f, const DevBits &v) {
  f.u64v(bv_space(v));
  f.u64v(v.nbits);
  f.i32v(0);  // _rb
  f.i32v(0);  // _sb
  f.i32v(0);  // _selectSpeed = DS_SELECT_SPEED_NO
  f.i32v(3);  // _selectTypeSupport
  if (v.nbits > 0) {
    f.dev_words(v.B.p, v.words);
    f.u64v(v.rwords * 8);  // DS_Rank9::_space
    f.u64v(v.words);       // _wordCnt
    f.dev_words(v.R.p, v.rwords);
    f.u64v(0);        // DS_Select::_space
    f.u64v(v.nbits);  // DS_Select::_n
    f.i32v(0);        // speed
  }
}

// bookkeeping constants of the reference's classes as they appear in every file it writes
// (tools/cfr_dump.py on reference-built indexes): 3 wavelet nodes of 592 bytes; sizeof(the tree) = 1648
static const u64 WT_NODES_BYTES = 1776, WT_SIZEOF = 1648;

static u64 wt_space(const DevBits node[3], u64 n) {
  if (n == 0) return 0;
  return WT_NODES_BYTES + bv_space(node[0]) + bv_space(node[1]) + bv_space(node[2]);
}

// Sequence_WaveletTree::Save (Sequence_WaveletTree.hpp:303-318); an empty tree is the bare header
static void write_wavelet(FileOut &f, const DevBits node[3], u64 n) {
  f.u64v(wt_space(node, n));
  f.u64v(n);
  if (n == 0) {
    f.u64v(0);  // Alphabet: _space, _method, _n of a default-constructed coder
    f.i32v(0);
    f.u64v(0);
    f.i32v(0);  // _tNodeCnt
    f.i32v(3);  // _selectSpeed as constructed
    return;
  }
  write_alphabet_acgt(f);
  f.i32v(3);
  f.i32v(0);
  for (int i = 0; i < 3; ++i) {
    f.u64v(i == 2 ? 1 : 0);   // prefix
    f.i32v(i == 0 ? 0 : 1);   // prefixLen
    f.i32v(i == 0 ? 1 : -1);  // children
    f.i32v(i == 0 ? 2 : -1);
    write_bitvector(f, node[i]);
  }
}

struct BuildInputs {
  const u64 *text;  // device (or host-twin) words, padded
  u64 n;
  std::vector<u64> genome_len, genome_seq_id;
  cfr_build_params prm;
};

struct Stats {
  u64 batches = 0, max_depth = 0;
  double t_sort = 0, t_post = 0, t_rb = 0, t_write = 0;
};

void build_index_file(const BuildInputs &in, const std::string &out_path, Stats &st, const Log &log) {
  const u64 n = in.n;
  const u64 *text = in.text;
  const int W = in.prm.precompute_width;
  const u64 rate = (u64)in.prm.sample_rate;
  if (W < 1 || W > 14) throw BuildError(-1, "builder: precompute_width must be in 1..14");
  if (rate < 1) throw BuildError(-1, "builder: sample_rate must be >= 1");
  if (n < (u64)W + 2) throw BuildError(-1, "builder: text too short");
  if (n >= (1ull << 40)) throw BuildError(-4, "builder: n must be below 2^40");
  const u64 ngen = in.genome_len.size();
  std::vector<u64> gstart(ngen + 1, 0);
  for (u64 g = 0; g < ngen; ++g) gstart[g + 1] = gstart[g] + in.genome_len[g];
  if (gstart[ngen] != n) throw BuildError(-1, "builder: genome lengths do not add up to n");

  // ---- outputs of stage 1
  const u64 bwt_words = (n + 31) / 32;
  Buf<u64> bwt(bwt_words + 2);
  bwt.zero();
  const u64 sample_n = (n + rate - 1) / rate;
  Buf<u32> samp(sample_n + 1);
  Buf<u32> samp_max(1);
  samp_max.zero();
  const u64 lut_n = 1ull << (2 * W);
  Buf<u64> lut_first(lut_n);
  bk_fill_ff(lut_first.p, lut_n * 8);
  Buf<u64> scal(64);  // [0] firstISA, [1] number of short-suffix rows, [2..] those rows
  scal.zero();
  // sequence boundaries (Builder.hpp:229-238): text positions psum - W - 1
  std::vector<u64> sel_pos;
  for (u64 g = 0; g + 1 < ngen; ++g)
    if (gstart[g + 1] >= (u64)W + 1) sel_pos.push_back(gstart[g + 1] - (u64)W - 1);
  std::sort(sel_pos.begin(), sel_pos.end());
  sel_pos.erase(std::unique(sel_pos.begin(), sel_pos.end()), sel_pos.end());
  const u64 nsel = sel_pos.size();
  Buf<u64> d_sel_pos(nsel + 1), d_sel_row(nsel + 1);
  if (nsel) bk_to_dev(d_sel_pos.p, sel_pos.data(), nsel * 8);
  bk_fill_ff(d_sel_row.p, (nsel + 1) * 8);
  const int COARSE = 16;  // one bit per 2^16 text positions: does the window hold a boundary?
  const u64 coarse_words = ((n >> COARSE) + 64) / 64;
  Buf<u64> d_coarse(coarse_words + 1);
  {
    std::vector<u64> cw(coarse_words + 1, 0);
    for (u64 p : sel_pos) cw[(p >> COARSE) >> 6] |= 1ull << ((p >> COARSE) & 63);
    bk_to_dev(d_coarse.p, cw.data(), cw.size() * 8);
  }
  Buf<u64> d_gstart(ngen + 1);
  bk_to_dev(d_gstart.p, gstart.data(), (ngen + 1) * 8);
  Buf<u32> d_gseq(ngen + 1);
  {
    std::vector<u32> gs(ngen + 1, 0);
    for (u64 g = 0; g < ngen; ++g) {
      if (in.genome_seq_id[g] > 0xffffffffull) throw BuildError(-4, "builder: sequence ids above 2^32");
      gs[g] = (u32)in.genome_seq_id[g];
    }
    bk_to_dev(d_gseq.p, gs.data(), gs.size() * 4);
  }

  // ---- batch plan: splitter keys from a sample of the suffixes
  u64 cap = in.prm.max_batch_rows;
  if (cap == 0) {
    const u64 fr = (u64)bk_free_bytes();
    const u64 reserve = (2ull << 30) + n / 8;  // stage-2 scratch is allocated after the sorter is gone
    cap = fr > reserve ? (fr - reserve) / (SortWork::bytes_per_elem() + 4) : (1ull << 20);
  }
  cap = std::min<u64>(cap, (1ull << 31) - 16);
  cap = std::max<u64>(cap, 1024);
  SortWork Wk;
  Wk.alloc(std::min<u64>(cap, n));
  cap = Wk.cap;
  Buf<u64> chunk_cnt;
  double t0 = now_s();
  u64 nb = n <= cap ? 1 : (u64)((double)n * 1.10 / (double)cap) + 1;
  std::vector<u64> split;
  for (int attempt = 0;; ++attempt) {
    split.assign(1, 0);
    if (nb > 1) {
      const u64 S = std::min<u64>(n, std::max<u64>(1ull << 16, nb * 4096));
      Buf<u64> sk(S), sv(S), sk2(S), sv2(S);
      u64 *skp = sk.p;
      const u64 step = n / S;
      par_for(S, [=] BK_LAMBDA(u64 i) { skp[i] = key31(text, i * step + ((i * 0x9E3779B97F4A7C15ull) >> 40) % step); });
      u64 *a = sk.p, *a2 = sk2.p, *v = sv.p, *v2 = sv2.p;
      sort_pairs(a, a2, v, v2, S, 0, 62);
      std::vector<u64> hs(S);
      bk_to_host(hs.data(), a, S * 8);
      for (u64 j = 1; j < nb; ++j) {
        const u64 s = hs[(size_t)(j * S / nb)];
        if (s > split.back()) split.push_back(s);
      }
    }
    split.push_back(1ull << 62);
    // dry run of the plan is folded into the real run below: a batch that overflows re-plans
    bool ok = true;
    u64 row0 = 0;
    st.batches = split.size() - 1;
    log("suffix sort: n = %llu, %llu batch(es), capacity %llu suffixes", n, st.batches, cap);
    for (size_t bi = 0; bi + 1 < split.size(); ++bi) {
      const double tb0 = now_s();
      const u64 m = select_positions(text, n, split[bi], split[bi + 1], split.size() == 2, Wk.pA.p, cap, chunk_cnt);
      if (m > cap) {
        ok = false;
        break;
      }
      if (m == 0) continue;
      sort_batch(text, n, Wk, m, log, st.max_depth);
      const double tb1 = now_s();
      // ---- everything the index needs from this stretch of the suffix array
      const u64 *SA = Wk.sa.p;
      u64 *BW = bwt.p, *SC = scal.p, *LF = lut_first.p, *SELR = d_sel_row.p;
      const u64 *SELP = d_sel_pos.p, *CO = d_coarse.p, *GS = d_gstart.p;
      const u32 *GQ = d_gseq.p;
      u32 *SM = samp.p, *SMX = samp_max.p;
      const u64 w_first = row0 >> 5, w_last = (row0 + m - 1) >> 5;
      // BWT[row] = T[SA[row] - 1]; the row of the whole text takes T[n - 1] and is firstISA (FMBuilder.hpp:249-256)
      par_for(w_last - w_first + 1, [=] BK_LAMBDA(u64 wi) {
        const u64 w = w_first + wi;
        const u64 r0 = w * 32 > row0 ? w * 32 : row0, r1 = (w + 1) * 32 < row0 + m ? (w + 1) * 32 : row0 + m;
        u64 val = 0;
        for (u64 r = r0; r < r1; ++r) {
          const u64 p = SA[r - row0];
          if (p == 0) SC[0] = r;
          val |= (u64)text_base(text, p == 0 ? n - 1 : p - 1) << (2 * (r & 31));
        }
        if (r1 - r0 == 32) BW[w] = val; else a_or64(BW + w, val);
      });
      // every rate-th row keeps the id of the sequence its suffix lies in, W + 1 bases further on when that
      // is still inside the text (Builder.hpp:37-44: "fuzzy boundary")
      const u64 s_first = (row0 + rate - 1) / rate, s_end = (row0 + m + rate - 1) / rate;
      par_for(s_end - s_first, [=] BK_LAMBDA(u64 si) {
        const u64 r = (s_first + si) * rate;
        const u64 p = SA[r - row0];
        const u64 q = p + (u64)W + 1 < n ? p + (u64)W + 1 : p;
        u64 lo = 0, hi = ngen;  // last genome whose start is <= q
        while (hi - lo > 1) {
          const u64 mid = (lo + hi) >> 1;
          if (GS[mid] <= q) lo = mid; else hi = mid;
        }
        const u32 id = GQ[lo];
        SM[s_first + si] = id;
        a_max32(SMX, id);
      });
      // first row of every W-mer (FMBuilder.hpp:262-285); suffixes shorter than W are noted
      par_for(m, [=] BK_LAMBDA(u64 j) {
        const u64 p = SA[j];
        const u64 mask = (1ull << (2 * W)) - 1ull;
        if (p + (u64)W > n) {
          const u64 slot = a_add64(SC + 1, 1);
          if (slot < 60) SC[2 + slot] = row0 + j;
          return;
        }
        const u64 w = text_bits64(text, p) & mask;
        bool cand = j == 0;
        if (!cand) {
          const u64 pp = SA[j - 1];
          cand = pp + (u64)W > n || (text_bits64(text, pp) & mask) != w;
        }
        if (cand) a_min64(LF + w, row0 + j);
      });
      // rows of the sequence boundaries (FMBuilder.hpp:293-297)
      if (nsel)
        par_for(m, [=] BK_LAMBDA(u64 j) {
          const u64 p = SA[j];
          const u64 cb = p >> COARSE;
          if (!((CO[cb >> 6] >> (cb & 63)) & 1ull)) return;
          u64 lo = 0, hi = nsel;
          while (lo < hi) {
            const u64 mid = (lo + hi) >> 1;
            if (SELP[mid] < p) lo = mid + 1; else hi = mid;
          }
          if (lo < nsel && SELP[lo] == p) SELR[lo] = row0 + j;
        });
      bk_sync();
      st.t_sort += tb1 - tb0;
      st.t_post += now_s() - tb1;
      row0 += m;
    }
    if (ok) {
      if (row0 != n) throw BuildError(-1, "builder: batches do not cover the text");
      break;
    }
    if (attempt > 6) throw BuildError(-6, "builder: cannot fit a batch into the work area");
    nb = nb * 3 / 2 + 1;
    bwt.zero();
    bk_fill_ff(lut_first.p, lut_n * 8);
    scal.zero();
    samp_max.zero();
    log("a batch exceeded the work area: re-planning with %llu batches", nb);
  }
  log("suffix sort done in %.2f s (sort %.2f s, derive %.2f s, deepest tie %llu bases)", now_s() - t0, st.t_sort, st.t_post,
      31 * (st.max_depth + 1));
  // the sorter's work area goes away before the run-block stage allocates
  Wk.release();
  chunk_cnt.release();
  bk_release_temp();

  // ---- host-side pieces of the auxiliary data
  std::vector<u64> hs = scal.host();
  const u64 first_isa = hs[0];
  const u64 n_short = hs[1];
  if (n_short > 60) throw BuildError(-1, "builder: more short suffixes than W - 1");
  std::vector<u64> short_rows(hs.begin() + 2, hs.begin() + 2 + n_short);
  std::sort(short_rows.begin(), short_rows.end());
  const unsigned char last_code = 0;
  (void)last_code;
  // precomputedRange[w] = {first row, rows} (FMBuilder.hpp:262-285): rows of one W-mer are contiguous;
  // the few suffixes shorter than W sit between ranges and are not counted
  std::vector<u64> lf = lut_first.host();
  std::vector<u64> lut(2 * lut_n, 0);
  {
    std::vector<std::pair<u64, u64>> present;  // (first row, w)
    present.reserve(lut_n);
    for (u64 w = 0; w < lut_n; ++w)
      if (lf[w] != ~0ull) present.push_back(std::make_pair(lf[w], w));
    std::sort(present.begin(), present.end());
    for (size_t i = 0; i < present.size(); ++i) {
      const u64 a = present[i].first, e = i + 1 < present.size() ? present[i + 1].first : n;
      u64 shorts = 0;
      for (u64 r : short_rows) shorts += (r >= a && r < e) ? 1 : 0;
      lut[2 * present[i].second] = a;
      lut[2 * present[i].second + 1] = e - a - shorts;
    }
  }
  // selectedSA: row -> id of the sequence that starts W + 1 bases after the boundary position
  // (Builder.hpp:48-52), ascending row (std::map order, FMIndex.hpp:122-127)
  std::vector<std::pair<u64, u64>> selected;
  {
    std::vector<u64> rows = d_sel_row.host();
    for (u64 k = 0; k < nsel; ++k) {
      if (rows[k] == ~0ull) throw BuildError(-1, "builder: a sequence boundary row was not found");
      const u64 q = sel_pos[k] + (u64)W + 1;
      const u64 g = (u64)(std::upper_bound(gstart.begin(), gstart.begin() + ngen, q) - gstart.begin()) - 1;
      selected.push_back(std::make_pair(rows[k], in.genome_seq_id[g]));
    }
    std::sort(selected.begin(), selected.end());
  }
  // sampled SA, bit-packed (FixedSizeElemArray::InitFromArray, FixedSizeElemArray.hpp:72-90)
  u32 smax = 0;
  bk_to_host(&smax, samp_max.p, 4);
  int sa_bits = 1;
  while (sa_bits < 32 && (smax >> sa_bits) != 0) ++sa_bits;
  const u64 sa_words = (sample_n * (u64)sa_bits + 63) / 64;
  Buf<u64> sa_packed(sa_words + 1);
  {
    const u32 *SM = samp.p;
    u64 *OUT = sa_packed.p;
    const int l = sa_bits;
    par_for(sa_words, [=] BK_LAMBDA(u64 w) {
      const u64 bit0 = w * 64;
      u64 e = bit0 / (u64)l, val = 0;
      for (; e < sample_n && e * (u64)l < bit0 + 64; ++e) {
        const u64 at = e * (u64)l;
        const u64 x = (u64)SM[e];
        if (at >= bit0) val |= x << (at - bit0);
        else val |= x >> (bit0 - at);
      }
      OUT[w] = val;
    });
  }
  samp.release();

  // ---- C array (FMIndex.hpp:283-292) and the symbol of the last text position
  u64 C[5] = {0, 0, 0, 0, 0};
  {
    Buf<u64> cnt4(4);
    cnt4.zero();
    u64 *c4 = cnt4.p;
    const u64 *BW = bwt.p;
    const u64 nunit = (n + UNIT - 1) / UNIT;
    par_for(nunit, [=] BK_LAMBDA(u64 u) {
      const u64 r0 = u * UNIT, r1 = r0 + UNIT < n ? r0 + UNIT : n;
      u64 c[4] = {0, 0, 0, 0};
      for (u64 r = r0; r < r1; ++r) ++c[text_base(BW, r)];
      for (int k = 0; k < 4; ++k) a_add64(c4 + k, c[k]);
    });
    std::vector<u64> h = cnt4.host();
    for (int k = 0; k < 4; ++k) C[k + 1] = C[k] + h[k];
  }
  u64 last_word = 0;
  bk_to_host(&last_word, text + ((n - 1) >> 5), 8);
  const char last_chr = "ACGT"[(last_word >> (2 * ((n - 1) & 31))) & 3ull];

  // ---- block size (Sequence_RunBlock::ComputeBlockSize) on sampled windows of the BWT
  double t1 = now_s();
  u64 b = in.prm.rbbwt_b;
  if (b == 0) {
    BlockSizeEstimator est;
    est.n = n;
    est.whole = est.len * est.test_cases >= n;
    const u64 *BW = bwt.p;
    if (est.whole) {
      Buf<unsigned char> d(n);
      unsigned char *dp = d.p;
      par_for(n, [=] BK_LAMBDA(u64 i) { dp[i] = (unsigned char)text_base(BW, i); });
      est.store.push_back(d.host());
      est.chunks.push_back(SymWindow{est.store.back().data(), 0, n});
    } else {
      // a block that starts at the last row of a window can reach 1.5 * 1024 rows further
      const u64 WIN = est.len + 2048, stride = est.stride();
      const u64 nwin = (n + stride - 1) / stride;
      Buf<unsigned char> d(nwin * WIN);
      unsigned char *dp = d.p;
      par_for(nwin * WIN, [=] BK_LAMBDA(u64 i) {
        const u64 r = (i / WIN) * stride + (i % WIN);
        dp[i] = r < n ? (unsigned char)text_base(BW, r) : 0;
      });
      est.store.push_back(d.host());
      for (u64 k = 0; k < nwin; ++k) est.chunks.push_back(SymWindow{est.store.back().data() + k * WIN, k * stride, WIN});
    }
    b = est.compute(2);
  }
  if (b == 1) b = n;  // Sequence_RunBlock.hpp:245-246
  RunBlockParts rb;
  build_run_blocks(bwt.p, n, b, rb, log);
  bk_sync();
  bwt.release();
  st.t_rb = now_s() - t1;

  // ---- write the file
  t1 = now_s();
  FileOut f(out_path);
  f.u64v(n);
  f.u64v(2);  // _plainAlphabetBits
  f.u64v(first_isa);
  f.u8v((unsigned char)last_chr);
  // Sequence_RunBlock::Save (Sequence_RunBlock.hpp:468-476)
  const u64 sp_type = bv_space(rb.type);
  u64 rb_space = sp_type;
  for (int t = 0; t < 2; ++t)  // += tree.GetSpace() - sizeof(tree) = _space + alphabet bytes + sizeof(pointer) - sizeof(tree)
    rb_space += wt_space(rb.wt[t], rb.seq_n[t]) + (rb.seq_n[t] ? 4 : 0) + 8 - WT_SIZEOF;
  f.u64v(rb_space);
  f.u64v(n);
  write_alphabet_acgt(f);
  f.u64v(b);
  f.u64v(rb.block_cnt);
  write_bitvector(f, rb.type);
  write_wavelet(f, rb.wt[0], rb.seq_n[0]);
  write_wavelet(f, rb.wt[1], rb.seq_n[1]);
  write_alphabet_acgt(f);  // FMIndex::_alphabets
  write_alphabet_acgt(f);  // FMIndex::_plainAlphabetCoder
  for (int k = 0; k < 5; ++k) f.u64v(C[k]);
  // _FMIndexAuxData::Save (FMIndex.hpp:100-134)
  f.u64v(n);
  f.i32v(0);  // sampleStrategy
  f.i32v((int32_t)rate);
  f.u64v(sample_n);
  f.u64v((u64)W);
  f.u64v(lut_n);
  f.u64v(in.genome_seq_id[0]);  // adjustedSA0 (Builder.hpp:45)
  f.u64v(sa_words);             // FixedSizeElemArray::_size
  f.i32v(sa_bits);
  f.u64v(sample_n);
  f.dev_words(sa_packed.p, sa_words);
  f.raw(lut.data(), lut.size() * 8);
  f.u64v(0);  // maxLcp
  f.u64v((u64)selected.size());
  f.i32v(1024);  // selectedSAFilterSampleRate
  for (auto &pr : selected) {
    f.u64v(pr.first);
    f.u64v(pr.second);
  }
  f.u8v(0);  // hasEndMarker
  f.close();
  st.t_write = now_s() - t1;
  log("wrote %s: %llu bytes (run-block stage %.2f s, file %.2f s)", out_path.c_str(), f.written, st.t_rb, st.t_write);
}

// ---------------------------------------------------------------------------
// synthetic collections (bench workloads): position-addressable, so reads can be drawn from the
// same definition without holding the text on the host
// ---------------------------------------------------------------------------
BK_HD u64 mix64(u64 x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// base i of strain t of species s: the species sequence is uniform random; a strain replaces a base
// with probability div_ppm / 10^6 by one of the three others
BK_HD u32 synth_base(u64 seed, u64 s, u64 t, u64 i, u64 div_ppm) {
  const u64 a = mix64(seed ^ mix64(s * 0x100000001B3ull + 1) ^ (i * 0xD6E8FEB86659FD93ull));
  u32 base = (u32)(a & 3);
  const u64 m = mix64((seed + 0x5851F42D4C957F2Dull) ^ mix64((s << 20) + t + 7) ^ (i * 0xA0761D6478BD642Full));
  if ((m & 0xffffffffull) % 1000000ull < div_ppm) base = (base + 1 + (u32)((m >> 40) % 3)) & 3;
  return base;
}

}  // namespace cfrbuild

using namespace cfrbuild;

// ---------------------------------------------------------------------------
// C ABI (include/centrifuger_b200_build.h)
// ---------------------------------------------------------------------------
namespace {
thread_local std::string g_build_err;
int build_fail(int code, const std::string &m) {
  g_build_err = m;
  return code;
}
}  // namespace

extern "C" {

const char *cfr_build_last_error(void) { return g_build_err.c_str(); }

void cfr_build_default_params(cfr_build_params *p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->sample_rate = 16;       // --offrate 4 (CentrifugerBuild.cpp:81)
  p->precompute_width = 10;  // --ftabchars 10
  p->rbbwt_b = 0;            // --rbbwt-b 0: automatic
  p->max_batch_rows = 0;
  p->device = 0;
  p->verbose = 0;
}

static int run_build(BuildInputs &in, const char *out_path, cfr_build_stats *stats) {
  Log log{in.prm.verbose, now_s()};
  Stats st;
  try {
    build_index_file(in, out_path, st, log);
  } catch (const BuildError &e) {
    return build_fail(e.code, e.what());
  } catch (const std::exception &e) {
    return build_fail(-6, e.what());
  }
  if (stats) {
    stats->batches = st.batches;
    stats->max_tie_depth_bases = 31 * (st.max_depth + 1);
    stats->sort_seconds = st.t_sort;
    stats->derive_seconds = st.t_post;
    stats->runblock_seconds = st.t_rb;
    stats->write_seconds = st.t_write;
  }
  return 0;
}

int cfr_build_fm_index(const uint8_t *codes, uint64_t n, const uint64_t *genome_lens, const uint64_t *genome_seq_ids,
                       uint64_t n_genomes, const cfr_build_params *p, const char *out_path, cfr_build_stats *stats) {
  if (!codes || !genome_lens || !genome_seq_ids || !out_path || n_genomes == 0) return build_fail(-1, "null argument");
  cfr_build_params prm;
  if (p) prm = *p; else cfr_build_default_params(&prm);
  try {
#if !defined(CFR_HOSTSIM)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return build_fail(-5, "no CUDA device available (the builder has no CPU path)");
    BK_CUDA(cudaSetDevice(prm.device));
#endif
    const u64 words = (n + 31) / 32 + 2;
    std::vector<u64> packed(words, 0);
    for (u64 i = 0; i < n; ++i) {
      if (codes[i] > 3) return build_fail(-1, "codes must be 0..3");
      packed[i >> 5] |= (u64)codes[i] << (2 * (i & 31));
    }
    Buf<u64> text(words);
    bk_to_dev(text.p, packed.data(), words * 8);
    std::vector<u64>().swap(packed);
    BuildInputs in;
    in.text = text.p;
    in.n = n;
    in.genome_len.assign(genome_lens, genome_lens + n_genomes);
    in.genome_seq_id.assign(genome_seq_ids, genome_seq_ids + n_genomes);
    in.prm = prm;
    return run_build(in, out_path, stats);
  } catch (const BuildError &e) {
    return build_fail(e.code, e.what());
  } catch (const std::exception &e) {
    return build_fail(-6, e.what());
  }
}

int cfr_build_synthetic_fm_index(uint64_t species, uint64_t strains, uint64_t genome_len, uint64_t div_ppm, uint64_t seed,
                                 const cfr_build_params *p, const char *out_path, cfr_build_stats *stats) {
  if (!out_path || species == 0 || strains == 0 || genome_len == 0) return build_fail(-1, "bad argument");
  cfr_build_params prm;
  if (p) prm = *p; else cfr_build_default_params(&prm);
  try {
#if !defined(CFR_HOSTSIM)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return build_fail(-5, "no CUDA device available (the builder has no CPU path)");
    BK_CUDA(cudaSetDevice(prm.device));
#endif
    const u64 ngen = species * strains, n = ngen * genome_len;
    const u64 words = (n + 31) / 32 + 2;
    Buf<u64> text(words);
    u64 *tw = text.p;
    par_for(words, [=] BK_LAMBDA(u64 w) {
      u64 v = 0;
      for (u64 i = w * 32; i < w * 32 + 32 && i < n; ++i) {
        const u64 g = i / genome_len, off = i - g * genome_len;
        v |= (u64)synth_base(seed, g / strains, g % strains, off, div_ppm) << (2 * (i & 31));
      }
      tw[w] = v;
    });
    BuildInputs in;
    in.text = text.p;
    in.n = n;
    in.genome_len.assign(ngen, genome_len);
    in.genome_seq_id.resize(ngen);
    for (u64 g = 0; g < ngen; ++g) in.genome_seq_id[g] = g;
    in.prm = prm;
    return run_build(in, out_path, stats);
  } catch (const BuildError &e) {
    return build_fail(e.code, e.what());
  } catch (const std::exception &e) {
    return build_fail(-6, e.what());
  }
}

void cfr_synth_bases(uint64_t species_index, uint64_t strain_index, uint64_t offset, uint64_t count, uint64_t div_ppm,
                     uint64_t seed, uint8_t *out_codes) {
  for (uint64_t i = 0; i < count; ++i) out_codes[i] = (uint8_t)synth_base(seed, species_index, strain_index, offset + i, div_ppm);
}

// `n` fragments of `length` bases: fragment i = bases [offset[i], offset[i] + length) of genome genome[i]
// (= species * strains + strain), one row of `out` each; spread over the host threads
void cfr_synth_fragments(const uint64_t *genome, const uint64_t *offset, uint64_t n, uint64_t length, uint64_t strains,
                         uint64_t div_ppm, uint64_t seed, uint8_t *out_codes) {
  const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([=]() {
      for (uint64_t i = t; i < n; i += nt) {
        uint8_t *o = out_codes + i * length;
        const uint64_t s = genome[i] / strains, k = genome[i] % strains;
        for (uint64_t j = 0; j < length; ++j) o[j] = (uint8_t)synth_base(seed, s, k, offset[i] + j, div_ppm);
      }
    });
  for (auto &x : th) x.join();
}

}  // extern "C"

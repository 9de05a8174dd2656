// cfr_pipeline.cuh -- the stages of one classification pass over a read chunk.
//
// Stage              unit of work               reference code it replaces
//   encode_stage       32 bases                   (layout: bytes -> 2-bit codes + N bits)
//   dust_screen_stage  one mate                   proves "Dustmasker masks nothing" for most mates (cfr_core.cuh)
//   dust_tasks         one mate the screen kept   CentrifugerClass.cpp:276-316 (Dustmasker::MaskWithBuffer)
//   search_tasks       one (read, mate, strand)   Classifier::GetHitsFromRead              Classifier.hpp:274
//   select_plan /      one read                   AdjustHitBoundaryFromStrandHits + strand pick :303,:509
//   select_write_rows                             + the row plan of GetClassificationFromHits :620-666
//                                                 (+ the rows themselves when the dense locate table covers every row)
//   locate_rows        one BWT row                FMIndex::BackwardToSampledSA             FMIndex.hpp:514
//   score_stage        one read                   GetClassificationFromHits :668-843 + Taxonomy::ReduceTaxIds
//
// Each stage is a plain function (compiled for the device by nvcc and for the host by the tests'
// simulation harness, where a "warp" is one lane); cfr_kernels.cuh wraps them in __global__ kernels
// that claim their work dynamically.
#pragma once
#include <type_traits>

#include "cfr_core.cuh"

namespace cfrb200 {

#define CFR_ROW_SENTINEL (~0ull)

struct ChunkDev {
  u64 n_reads;
  int mates;  // 1 or 2
  int cap_h;  // hit slots per strand task
  const unsigned char *seq_raw;  // bases as uploaded (input of encode_stage only)
  u64 n_words;                   // 32-base words covering the batch buffer
  u64 *codes;                    // 2-bit codes, 32 bases per word
  u32 *mask_raw;                 // N bits as uploaded
  u32 *mask;                     // N bits the searches see (= mask_raw | DUST intervals)
  u32 *dust_bits;                // optional: just the DUST intervals (diagnostics)
  const u64 *off[2];             // per mate: n_reads + 1 offsets as given by the caller
  u64 off_bias[2];               // position in seq = off[m][i] - off_bias[m]
  u64 uni_len[2], uni_pos0[2];   // every read of mate m has uni_len[m] bases (0 = lengths differ): read i stands at
                                 // uni_pos0[m] + i * uni_len[m] -- the search kernel then needs no offset loads
  Hit *strand_hits;              // [n_reads * 2*mates * cap_h]
  int *strand_nhits;             // [n_reads * 2*mates]
  FinalHit *fhits;               // [n_reads * 2*mates * cap_h]
  ReadWork *work;                // [n_reads]
  // locate / scoring arena, one slot per planned BWT row
  u64 arena_cap;
  u64 *arena_used;
  u64 *arena_valid;   // lowest arena row of a read that did NOT fit: rows below it are all written
  u64 *task_counter;  // next strand task (dynamic fetch by the search warps)
  u64 *row_counter;   // next arena row (dynamic fetch by the locate warps)
  u64 *dust_counter;  // next mate (dynamic fetch by the DUST warps)
  u32 *dust_list;     // mates the register-only screen could not clear (nullptr = every mate runs SDUST)
  u32 *dust_list_n;   // number of entries in dust_list
  u64 *rows;
  u32 *seq_ids;
  SeqRec *rec0, *rec1;
  u64 *best, *tmp;
  // outputs
  DevResult *results;  // [n_reads]
  u64 *out_ids;        // [n_reads * max_result]
  u64 *taxon_counts;   // [node_cnt + 3]
  DevCounters *counters;
  // optional (--expand-taxid): the ids promoted into each reported id, Classifier.hpp:807-838
  u32 *exp_cnt;   // [n_reads * max_result] list lengths (nullptr = not requested)
  u64 *exp_off;   // [n_reads] start of the read's lists in exp_ids
  u64 *exp_ids;   // compact tax ids, list after list
  u64 exp_cap;
  u64 *exp_used;  // next free entry of exp_ids
  // reads that did not fit the arena in this pass
  u32 *deferred;
  u32 *n_deferred;
  const u32 *read_list;  // nullptr = identity
  u64 n_list;
};

CFR_HD u64 chunk_read_id(const ChunkDev &B, u64 t) { return B.read_list ? (u64)B.read_list[t] : t; }

// ---- warp-level helpers shared by the state-machine stages ----
#if defined(__CUDA_ARCH__)
#define CFR_BALLOT(pred) __ballot_sync(0xffffffffu, (pred))
#define CFR_SYNCWARP() __syncwarp()
#else
#define CFR_BALLOT(pred) ((pred) ? 1u : 0u)
#define CFR_SYNCWARP() ((void)0)
#endif

CFR_HD int popc32(u32 x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// Warp-aggregated claim of work items from a global counter: the lanes that `want`
// one (LANES adjacent lanes share an item) get consecutive indices with a single
// atomic per warp.  Must be called by all 32 lanes.
template <int LANES>
CFR_HD u64 warp_claim(u64 *counter, bool want) {
#if defined(__CUDA_ARCH__)
  const u32 m = __ballot_sync(0xffffffffu, want);
  if (m == 0) return 0;
  const int lane = threadIdx.x & 31, leader = __ffs((int)m) - 1;
  u64 base = 0;
  if (lane == leader) base = atomicAdd(counter, (u64)(__popc(m) / LANES));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + (u64)(__popc(m & ((1u << lane) - 1u)) / LANES);
#else
  return want ? (*counter)++ : 0;
#endif
}


// n consecutive entries of a shared output area
CFR_HD u64 claim_entries(u64 *counter, u64 n) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(counter, n);
#else
  const u64 at = *counter;
  *counter += n;
  return at;
#endif
}

// ------------------------------------------------------------------ encode
// word w of the batch buffer: 32 uploaded bytes -> 2-bit codes + N bits
CFR_HD void encode_stage(const ChunkDev &B, u64 w, u64 total_bytes) {
  const u64 b0 = w * 32;
  const int n = b0 + 32 <= total_bytes ? 32 : (b0 < total_bytes ? (int)(total_bytes - b0) : 0);
  // two aligned 16-byte loads (the buffer is padded past total_bytes), then bytes from registers
  const u64x2 v0 = ld128(reinterpret_cast<const u64x2 *>(B.seq_raw + b0));
  const u64x2 v1 = ld128(reinterpret_cast<const u64x2 *>(B.seq_raw + b0) + 1);
  const u64 q[4] = {v0.x, v0.y, v1.x, v1.y};
  u64 codes = 0;
  u32 nm = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 32; ++i) {  // branch-free: A,C,G,T = 0x41,0x43,0x47,0x54 -> ((b>>1)&3) = 0,1,3,2
    const u32 b = (u32)(q[i >> 3] >> ((i & 7) * 8)) & 0xffu;
    const u32 x = (b >> 1) & 3u;
    const u32 c = x ^ (x >> 1);
    const u32 valid = (b == 'A') | (b == 'C') | (b == 'G') | (b == 'T');
    nm |= (valid ^ 1u) << i;
    codes |= (u64)(c & (0u - valid)) << (2 * i);
  }
  if (n < 32) {  // padding reads as N
    const u32 keep = n == 0 ? 0u : ((1u << n) - 1u);
    nm |= ~keep;
    codes &= n == 0 ? 0ull : ((1ull << (2 * n)) - 1ull);
  }
  B.codes[w] = codes;
  B.mask_raw[w] = nm;
  if (B.mask != B.mask_raw) B.mask[w] = nm;
  if (B.dust_bits) B.dust_bits[w] = 0;
}

// ------------------------------------------------------------------ dust
// SDUST over all mates of the chunk as a warp-synchronous state machine: lanes in
// STEP advance their window by one base together; FindPerfect (long, data
// dependent, needed by a minority of positions) and the per-mate set-up run when a
// quorum of lanes waits for them.  Mates are claimed dynamically.
enum { CFR_DS_FETCH = 0, CFR_DS_SEG = 1, CFR_DS_STEP = 2, CFR_DS_SLOW = 3, CFR_DS_DONE = 4, CFR_DS_SHRINK = 5 };

template <int SW>
CFR_HD void dust_tasks(const ChunkDev &B, const u64 ntask, DustStateT<SW> &d, const int quorum, const bool active = true,
                       const bool defer_shrink = false) {
  DustIn in{B.codes, B.mask_raw, 0};
  DustOut out{B.mask, B.dust_bits, 0};
  // `active` = false parks the lane: the deferred loops (FindPerfect, shrink) of one mate stall the
  // other mates of its warp, so a launch over few, all slow mates spreads them over more warps
  int st = active ? CFR_DS_FETCH : CFR_DS_DONE;
  int len = 0, cursor = 0, seg_off = 0, seg_n = 0, wfinish = 0, c1 = 0, c2 = 0, t_last = 0;
  for (;;) {
    const u32 m_step = CFR_BALLOT(st == CFR_DS_STEP);
    const u32 m_slow = CFR_BALLOT(st == CFR_DS_SLOW || st == CFR_DS_SHRINK);
    const u32 m_trn = CFR_BALLOT(st == CFR_DS_FETCH || st == CFR_DS_SEG);
    if ((m_step | m_slow | m_trn) == 0) break;
    const int alive = popc32(m_step | m_slow | m_trn);
    const int q_now = (alive + 3) / 4 < quorum ? ((alive + 3) / 4 < 1 ? 1 : (alive + 3) / 4) : quorum;
    if (m_trn != 0 && ((m_step | m_slow) == 0 || popc32(m_trn) >= q_now)) {
      for (int tries = 0; tries < 3; ++tries) {
        if (CFR_BALLOT(st == CFR_DS_FETCH || st == CFR_DS_SEG) == 0) break;
        const u64 claimed = warp_claim<1>(B.dust_counter, st == CFR_DS_FETCH);
        if (st == CFR_DS_FETCH) {
          if (claimed >= ntask) {
            st = CFR_DS_DONE;
          } else {
            const u64 task = B.dust_list ? (u64)B.dust_list[claimed] : claimed;
            const u64 read = task / (u64)B.mates;
            const int mate = (int)(task % (u64)B.mates);
            const u64 base = B.off[mate][read] - B.off_bias[mate];
            len = (int)(B.off[mate][read + 1] - B.off[mate][read]);
            in.base = base;
            in.widx = ~0ull;
            out.base = base;
            if (len >= 3) {
              if (dust_all_acgt(B.mask_raw, base, len)) {  // common case: one segment, the whole mate
                seg_off = 0;
                seg_n = len;
                cursor = len;
                dust_seg_init(in, seg_off, d, c1, c2);
                wfinish = 2;
                st = CFR_DS_STEP;
              } else {  // MaskWithBuffer: skip the leading Ns (Dustmasker.hpp:365-367)
                cursor = 0;
                while (cursor < len && in(cursor) == 4) ++cursor;
                st = CFR_DS_SEG;
              }
            }
          }
        }
        if (st == CFR_DS_SEG) {
          if (cursor >= len) {
            st = CFR_DS_FETCH;
          } else {
            int last_valid, next_i;
            dust_next_segment(in, len, cursor, last_valid, next_i);
            seg_off = cursor;
            seg_n = last_valid - cursor + 1;
            cursor = next_i;
            if (last_valid > seg_off && seg_n >= 3) {
              dust_seg_init(in, seg_off, d, c1, c2);
              wfinish = 2;
              st = CFR_DS_STEP;
            }
          }
        }
      }
    }
    bool advance = false;
    // FindPerfect (long, data dependent) for the lanes that need it
    // the two data-dependent loops of a step -- shrinking the suffix and FindPerfect -- for the lanes
    // that wait for them
    if (m_slow != 0 && (m_step == 0 || popc32(m_slow) >= q_now)) {
      if (st == CFR_DS_SHRINK) {
        dust_shrink(d, t_last);
        st = CFR_DS_SLOW;
        if (!dust_needs_find_perfect(d)) advance = true;
      }
      if (st == CFR_DS_SLOW && !advance) {
        dust_find_perfect(wfinish, d);
        advance = true;
      }
    }
    if (st == CFR_DS_STEP) {
      const bool shrink = dust_step(in, out, seg_off, wfinish, d, c1, c2, t_last);
      if (shrink && defer_shrink) {
        st = CFR_DS_SHRINK;  // mates that reach the full masker after the screen shrink often: batch the loop
      } else {
        if (shrink) dust_shrink(d, t_last);  // rare on unscreened input and short: inline
        if (dust_needs_find_perfect(d))
          st = CFR_DS_SLOW;
        else
          advance = true;
      }
    }
    if (advance) {
      ++wfinish;
      st = CFR_DS_STEP;
      if (wfinish >= seg_n) {
        dust_seg_tail(out, seg_off, seg_n, d);
        st = CFR_DS_SEG;
      }
    }
  }
}

// The register-only screen (dust_screen, cfr_core.cuh) for mate `task`: true when the mate
// must go through the full SDUST (it holds a non-ACGT byte, or the screen cannot rule
// out a masked interval).
CFR_HD bool dust_screen_stage(const ChunkDev &B, u64 task) {
  const u64 read = task / (u64)B.mates;
  const int mate = (int)(task % (u64)B.mates);
  const u64 base = B.off[mate][read] - B.off_bias[mate];
  const int len = (int)(B.off[mate][read + 1] - B.off[mate][read]);
  if (len < 3) return false;  // MaskWithBuffer returns at once
  if (!dust_all_acgt(B.mask_raw, base, len)) return true;
  return dust_screen(B.codes, base, len);
}

// sequential form: one mate (host-side diagnostics)
template <int SW>
CFR_HD void dust_stage(const ChunkDev &B, u64 task, DustStateT<SW> &d) {
  const u64 read = task / (u64)B.mates;
  const int mate = (int)(task % (u64)B.mates);
  const u64 base = B.off[mate][read] - B.off_bias[mate];
  const int len = (int)(B.off[mate][read + 1] - B.off[mate][read]);
  DustIn in{B.codes, B.mask_raw, base};
  const DustOut out{B.mask, B.dust_bits, base};
  dust_task(in, len, out, d);
}

// ------------------------------------------------------------------ search
// task = read * (2*mates) + mate*2 + s, s = 1: the mate as read (strandHits[1]),
// s = 0: its reverse complement (strandHits[0]).
//
// GetHitsFromRead + BackwardSearch (Classifier.hpp:274-293, FMIndex.hpp:388-422,
// 487-510) as a warp-synchronous state machine.  Every iteration the lanes that
// are inside a search do ONE BackwardExtend together (straight-line code).  The
// rare events -- closing a search (record the hit, skip the mismatching base),
// starting the next one (lookup-table probe of the last W bases) and fetching the
// next task -- are deferred until a quorum of lanes is waiting for them (or nobody
// can extend), so that block runs with many lanes instead of one or two.
// How many waiting tasks trigger the deferred block: the configured quorum, but never
// more than a quarter of the tasks the warp still has alive (so a draining warp does
// not stall its last lanes behind a quorum it can no longer reach).
CFR_HD int adaptive_quorum(int quorum, u32 alive_mask, int lanes_per_task) {
  const int alive = popc32(alive_mask) / lanes_per_task;
  const int q = (alive + 3) / 4;
  return (q < quorum ? (q < 1 ? 1 : q) : quorum) * lanes_per_task;
}

// CFR_ST_PROBE (single-wait pair policy only): the search has its wide-table key (kept in sp) and waits for the entry
enum { CFR_ST_EXTEND = 0, CFR_ST_CLOSE = 1, CFR_ST_FETCH = 2, CFR_ST_DONE = 3, CFR_ST_PROBE = 4 };

// the pair policies implement extend2; the others never reach the call (Bwt::PAIR == 0)
template <class Bwt>
CFR_HD typename std::enable_if<(Bwt::PAIR != 0), int>::type pair_extend2(const DevIndex &ix, bool go, int c1, int c2, u64 &sp,
                                                                         u64 &ep, OpCount &oc) {
  return Bwt::extend2(ix, go, c1, c2, sp, ep, oc);
}
template <class Bwt>
CFR_HD typename std::enable_if<(Bwt::PAIR == 0), int>::type pair_extend2(const DevIndex &, bool, int, int, u64 &, u64 &, OpCount &) {
  return 0;
}
// the single-wait pair policy (Bwt::PAIR == 2, device only) implements round(); the others never reach the call
template <class Bwt>
CFR_HD typename std::enable_if<(Bwt::PAIR == 2), int>::type pair_round(const DevIndex &ix, bool go, int c1, int c2, u64 &sp, u64 &ep,
                                                                       bool probe, u64 key, u64x2 &entry, OpCount &oc) {
#if defined(__CUDA_ARCH__)
  return Bwt::round(ix, go, c1, c2, sp, ep, probe, key, entry, oc);
#else
  return 0;
#endif
}
template <class Bwt>
CFR_HD typename std::enable_if<(Bwt::PAIR != 2), int>::type pair_round(const DevIndex &, bool, int, int, u64 &, u64 &, bool, u64, u64x2 &,
                                                                       OpCount &) {
  return 0;
}

template <class Bwt>
CFR_HD void search_tasks(const DevIndex &ix, const DevParams &P, const ChunkDev &B, const u64 ntask, OpCount &oc) {
  const int W = ix.pre_width, WW = ix.wide_width, mhl = P.min_hit_len;
  const int sshift = B.mates == 2 ? 2 : 1;  // strand tasks per read = 2 * mates = 1 << sshift
  typedef typename Bwt::pos_t pos_t;
  const pos_t n_rows = (pos_t)ix.n;
  StrandSeq s{B.codes, B.mask, 0, 0, 0};
  u64 cur = 0;
  pos_t sp = 0, ep = 0;
  int nh = 0, remaining = 0, l = 0;
  int l0 = 0;  // bases of the current search that came out of a lookup table (no BackwardExtend ran for them)
  int st = CFR_ST_FETCH;
  for (;;) {
    const u32 ext = CFR_BALLOT(st == CFR_ST_EXTEND || st == CFR_ST_PROBE);
    const u32 trn = CFR_BALLOT(st == CFR_ST_CLOSE || st == CFR_ST_FETCH);
    if ((ext | trn) == 0) break;
    // warp-uniform: does the deferred transition block run in this iteration?
    const bool transit = trn != 0 && (ext == 0 || popc32(trn) >= adaptive_quorum(P.quorum, ext | trn, (int)Bwt::LANES));
    if (transit) {
      // ---- transition block (warp-uniform entry): CLOSE -> (FETCH ->) start of the next search
      bool start = false;
      if (st == CFR_ST_CLOSE) {  // back in GetHitsFromRead
        if (Bwt::STEPS_COUNTED_AT_CLOSE && l >= l0 && l0 >= W) {
          // BackwardExtend calls of this search: l - l0 that succeeded, plus the one that failed when
          // the search stopped on an ACGT base before the start of the strand (the cursor is still on it)
          oc.xext += (u32)(l - l0) + ((l < remaining && s.peek() <= 3) ? 1u : 0u);
        }
        if (l >= mhl && sp <= ep && nh < B.cap_h) {
          if (Bwt::leader()) {
            Hit &o = B.strand_hits[cur * (u64)B.cap_h + (u64)nh];
            o.sp = sp;
            o.ep = ep;
            o.l = l;
            o.offset = s.len - remaining;
          }
          ++nh;
        }
        remaining -= (l + 1);
        if (remaining >= mhl) {
          start = true;
        } else {
          if (Bwt::leader()) B.strand_nhits[cur] = nh;
          st = CFR_ST_FETCH;
        }
      }
      const u64 claimed = warp_claim<Bwt::LANES>(B.task_counter, st == CFR_ST_FETCH);
      if (st == CFR_ST_FETCH) {
        if (claimed >= ntask) {
          st = CFR_ST_DONE;
        } else {
          cur = claimed;
          const u64 read = cur >> sshift;
          const int w = (int)(cur & ((1ull << sshift) - 1ull));
          const int mate = w >> 1;
          const u64 ulen = mate ? B.uni_len[1] : B.uni_len[0];
          if (ulen) {  // one read length: no loads on the way to the first search of the task
            s.base = (mate ? B.uni_pos0[1] : B.uni_pos0[0]) + read * ulen;
            s.len = (int)ulen;
          } else {
            const u64 *off = mate ? B.off[1] : B.off[0];
            const u64 o0 = off[read], o1 = off[read + 1];
            s.base = o0 - (mate ? B.off_bias[1] : B.off_bias[0]);
            s.len = (int)(o1 - o0);
          }
          s.rc = (w & 1) ? 0 : 1;
          s.widx = ~0ull;
          nh = 0;
          remaining = s.len;
          sp = ep = 0;
          st = CFR_ST_CLOSE;  // a strand shorter than minHitLen closes at once with no hits
          l = -1;             // (remaining -= l + 1 leaves it unchanged; xext sees l < W)
          if (remaining >= mhl) start = true;
        }
      }
      // the lanes that continue a strand and the lanes that just fetched one start their searches
      // together (without the barrier the compiler runs the block below once per group)
      CFR_SYNCWARP();
      if (start) {  // FMIndex::BackwardSearch up to the initial range
        st = CFR_ST_CLOSE;
        l0 = W;
        bool wide_done = false;
        if (WW > 0 && remaining >= WW) {  // the wide table answers for the last WW bases unless one is not ACGT
          u64 key;
          int nvalid;
          if (Bwt::PAIR == 2 && s.init_key(remaining, WW, key, nvalid)) {
            // the entry is fetched with this iteration's lines (pair_round below); the cursor is placed where the
            // search continues if it does (the table covered WW bases), so its words are loading meanwhile
            ++oc.search;
            st = CFR_ST_PROBE;
            sp = (pos_t)key;
            wide_done = true;
            l0 = WW;
            if (WW < remaining) s.seek(remaining - 1 - WW);
          } else if (Bwt::PAIR != 2 && s.init_key(remaining, WW, key, nvalid)) {
            ++oc.search;
            const u64x2 e = ld128(ix.wide + key);
            l = (int)(e.y >> 56);
            sp = (pos_t)e.x;
            ep = (pos_t)(e.y & 0xffffffffffffffull);
            wide_done = true;
            l0 = WW;  // l < WW: the search ended inside the table (xext sees l < l0)
            if (l == WW && l < remaining) {
              st = CFR_ST_EXTEND;
              s.seek(remaining - 1 - l);
            }
          }
        }
        if (wide_done) {
        } else if (remaining < W) {
          l = 0;
        } else {
          ++oc.search;
          if (W > 0) {
            u64 key;
            int nvalid;
            if (!s.init_key(remaining, W, key, nvalid)) {
              sp = 1;
              ep = 0;
              l = nvalid;
            } else {
              const u64x2 e = ld128(ix.lookup + key);
              if (e.y == 0) {
                sp = 1;
                ep = 0;
                l = W - 1;
              } else {
                sp = (pos_t)e.x;
                ep = (pos_t)(e.x + e.y - 1);
                l = W;
                if (l < remaining) {
                  st = CFR_ST_EXTEND;
                  s.seek(remaining - 1 - l);
                }
              }
            }
          } else {
            sp = 0;
            ep = n_rows - 1;
            l = 0;
            if (l < remaining) {
              st = CFR_ST_EXTEND;
              s.seek(remaining - 1 - l);
            }
          }
        }
      }
    }
    if (Bwt::PAIR == 2) {
      // one memory round for the whole warp: the lines of the lanes that extend and the table entries of the lanes
      // that start a search arrive together
      const bool act = st == CFR_ST_EXTEND, probe = st == CFR_ST_PROBE;
      int c1 = 0, c2 = -1;
      if (act) {
        c1 = s.peek();  // the cursor stands on strand position remaining - 1 - l
        st = CFR_ST_CLOSE;
        if (c1 <= 3 && l + 1 < remaining) {
          s.advance();
          c2 = s.peek();
          if (c2 > 3) c2 = -1;  // the search ends on this base (FMIndex.hpp:500)
        }
      }
      const bool go = act && c1 <= 3;
      u64 psp = (u64)sp, pep = (u64)ep;
      u64x2 e;
      e.x = e.y = 0;
      const int done = pair_round<Bwt>(ix, go, c1, c2, psp, pep, probe, (u64)sp, e, oc);
      if (go) {
        sp = (pos_t)psp;
        ep = (pos_t)pep;
        l += done;
        if (done == 2 && l < remaining) {
          st = CFR_ST_EXTEND;
          s.advance();
        }
      }
      if (probe) {  // where FMIndex::BackwardSearch stands after the last WW bases
        l = (int)(e.y >> 56);
        sp = (pos_t)e.x;
        ep = (pos_t)(e.y & 0xffffffffffffffull);
        st = (l == WW && l < remaining) ? CFR_ST_EXTEND : CFR_ST_CLOSE;
      }
    } else if (Bwt::PAIR) {
      // up to two FMIndex::BackwardExtend steps from one line per boundary; the call is warp-uniform (the lines
      // are fetched cooperatively), lanes outside a search pass go = false
      const bool act = st == CFR_ST_EXTEND;
      int c1 = 0, c2 = -1;
      if (act) {
        c1 = s.peek();  // the cursor stands on strand position remaining - 1 - l
        st = CFR_ST_CLOSE;
        if (c1 <= 3 && l + 1 < remaining) {
          s.advance();
          c2 = s.peek();
          if (c2 > 3) c2 = -1;  // the search ends on this base (FMIndex.hpp:500)
        }
      }
      const bool go = act && c1 <= 3;
      u64 psp = (u64)sp, pep = (u64)ep;
      const int done = pair_extend2<Bwt>(ix, go, c1, c2, psp, pep, oc);
      if (go) {
        sp = (pos_t)psp;
        ep = (pos_t)pep;
        l += done;
        if (done == 2 && l < remaining) {
          st = CFR_ST_EXTEND;
          s.advance();
        }
      }
    } else if (st == CFR_ST_EXTEND) {  // one FMIndex::BackwardExtend; l < remaining holds here
      const int c = s.peek();   // the cursor stands on strand position remaining - 1 - l
      st = CFR_ST_CLOSE;
      if (c <= 3) {
        pos_t nsp, nep;
        Bwt::extend_step(ix, c, sp, ep, nsp, nep, oc);
        if (!(nsp > nep || nep > n_rows)) {
          sp = nsp;
          ep = nep;
          ++l;
          if (l < remaining) {
            st = CFR_ST_EXTEND;
            s.advance();
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ locate
// FMIndex::BackwardToSampledSA for the arena rows.  Same scheme: lanes that are
// walking do one LF step together; resolving a sampled row (sampled-SA read /
// selected-SA look-up), storing the sequence id and claiming the next row wait
// for a quorum.
enum { CFR_LS_WALK = 0, CFR_LS_CHECK = 1, CFR_LS_NEED = 2, CFR_LS_DONE = 3 };

template <class Bwt>
CFR_HD void locate_rows(const DevIndex &ix, const DevParams &P, const ChunkDev &B, const u64 used, OpCount &oc) {
  typedef typename Bwt::pos_t pos_t;
  u64 cur = 0;
  pos_t i = 0;
  int st = CFR_LS_NEED;
  for (;;) {
    const u32 walk = CFR_BALLOT(st == CFR_LS_WALK);
    const u32 trn = CFR_BALLOT(st == CFR_LS_CHECK || st == CFR_LS_NEED);
    if ((walk | trn) == 0) break;
    if (trn != 0 && (walk == 0 || popc32(trn) >= adaptive_quorum(P.quorum, walk | trn, (int)Bwt::LANES))) {
      if (st == CFR_LS_CHECK) {  // FMIndex::GetSampledSA, literally (behind the dense table, when built)
        u64 sa;
        if (get_located(ix, i, sa)) {
          if (Bwt::leader()) B.seq_ids[cur] = (u32)sa;
          ++oc.locate;
          st = CFR_LS_NEED;
        } else {  // filter bit set but the row is not a selected one: keep walking
          i = Bwt::lf(ix, i, oc);
          ++oc.lf;
          st = CFR_LS_WALK;
        }
      }
      CFR_SYNCWARP();
      const u64 claimed = warp_claim<Bwt::LANES>(B.row_counter, st == CFR_LS_NEED);
      if (st == CFR_LS_NEED) {
        if (claimed >= used) {
          st = CFR_LS_DONE;
        } else {
          cur = claimed;
          const u64 row = B.rows[cur];
          i = (pos_t)row;
          if (row != CFR_ROW_SENTINEL) st = CFR_LS_WALK;
        }
      }
    }
    if (st == CFR_LS_WALK) {
      // cheap pre-test of GetSampledSA's three conditions; the loads happen in the transition block
      bool maybe = i == (pos_t)ix.first_isa || is_sampled_row(ix, i) || is_dense_row(ix, i);
      if (!maybe && ix.sel_filter) {
        const pos_t fb = filter_bit_index(ix, i);
        maybe = (ld64(ix.sel_filter + (fb >> 6)) >> (fb & 63)) & 1ull;
      }
      if (maybe) {
        st = CFR_LS_CHECK;
      } else {
        i = Bwt::lf(ix, i, oc);
        ++oc.lf;
      }
    }
  }
}

// ------------------------------------------------------------------ select
// Phase A: boundary adjustment, strand choice, final hit list, row plan.
// Returns the number of arena rows the read needs.
template <class Bwt>
CFR_HD u32 select_plan(const DevIndex &ix, const DevParams &P, const ChunkDev &B, u64 read, OpCount &oc) {
  const int S = 2 * B.mates;
  const int mhl = P.min_hit_len;
  Hit *h[2][2];
  int n[2][2];
  int qlen = 0;
  for (int m = 0; m < B.mates; ++m) {
    const u64 base = B.off[m][read] - B.off_bias[m];
    const int len = (int)(B.off[m][read + 1] - B.off[m][read]);
    qlen += len;
    for (int s = 0; s < 2; ++s) {
      const u64 task = read * (u64)S + (u64)(m * 2 + s);
      h[m][s] = B.strand_hits + task * (u64)B.cap_h;
      n[m][s] = B.strand_nhits[task];
    }
    adjust_hit_boundary<Bwt>(ix, StrandSeq{B.codes, B.mask, base, len, 0}, h[m][0], n[m][0], h[m][1], n[m][1], oc);
  }
  B.results[read].query_length = qlen;
  // template strand k: mate-1 hits of strand k, then mate-2 hits of strand 1-k (Classifier.hpp:551-552)
  u64 score[2] = {0, 0};
  for (int k = 0; k < 2; ++k) {
    for (int i = 0; i < n[0][k]; ++i) score[k] += hit_score(h[0][k][i].l, mhl);
    if (B.mates == 2)
      for (int i = 0; i < n[1][1 - k]; ++i) score[k] += hit_score(h[1][1 - k][i].l, mhl);
  }
  int order[2], norder;
  if (score[1] > score[0] + score[0] / 100) {
    order[0] = 1;
    norder = 1;
  } else if (score[0] > score[1] + score[1] / 100) {
    order[0] = 0;
    norder = 1;
  } else {
    order[0] = 1;
    order[1] = 0;
    norder = 2;
  }
  FinalHit *fh = B.fhits + read * (u64)S * (u64)B.cap_h;
  u32 nh = 0;
  u64 rows = 0;
  for (int q = 0; q < norder; ++q) {
    const int k = order[q];
    for (int part = 0; part < B.mates; ++part) {
      const Hit *src = part == 0 ? h[0][k] : h[1][1 - k];
      const int cnt = part == 0 ? n[0][k] : n[1][1 - k];
      for (int i = 0; i < cnt; ++i) {
        FinalHit f;
        f.sp = src[i].sp;
        f.ep = src[i].ep;
        f.l = src[i].l;
        f.offset = src[i].offset;
        f.strand = 2 * k - 1;
        f.row_cnt = 0;
        if (f.l >= mhl) {
          const RowPlan rp = plan_rows(f.sp, f.ep, P);
          f.row_cnt = rp.total > 0xffffffffull ? 0xffffffffu : (u32)rp.total;
          rows += rp.total;
        }
        fh[nh++] = f;
      }
    }
  }
  B.work[read].n_hits = nh;
  B.work[read].arena_rows = rows > 0xffffffffull ? 0xffffffffu : (u32)rows;
  return B.work[read].arena_rows;
}

// Phase B: with the arena slice known, expand the planned rows.
// When the dense locate table answers every row (dense_shift == 0) the sequence ids are written here
// and the locate launch is skipped (locate_in_select); returns the number of rows resolved that way.
CFR_HD bool locate_in_select(const DevIndex &ix) { return ix.dense_shift == 0; }

CFR_HD u32 select_write_rows(const DevIndex &ix, const DevParams &P, const ChunkDev &B, u64 read, u64 arena_base,
                             bool fits) {
  const int S = 2 * B.mates;
  ReadWork &w = B.work[read];
  w.arena_base = arena_base;
  w.status = fits ? 0 : 1;
  if (!fits) return 0;
  const FinalHit *fh = B.fhits + read * (u64)S * (u64)B.cap_h;
  const bool direct = locate_in_select(ix);
  u64 o = arena_base;
  for (u32 i = 0; i < w.n_hits; ++i) {
    if (fh[i].row_cnt == 0) continue;
    const RowPlan rp = plan_rows(fh[i].sp, fh[i].ep, P);
    for (u64 t = 0; t < rp.total; ++t) {
      const u64 row = plan_row_at(fh[i].sp, fh[i].ep, rp, t);
      if (direct)
        B.seq_ids[o++] = dense_read(ix, row);
      else
        B.rows[o++] = row;
    }
  }
  return direct ? (u32)(o - arena_base) : 0u;
}

// ------------------------------------------------------------------ score
// returns the number of assignments (for the classified counter)
CFR_HD int score_stage(const DevIndex &ix, const DevParams &P, const ChunkDev &B, u64 read, u64 *err_flags) {
  const int S = 2 * B.mates;
  const ReadWork &w = B.work[read];
  const FinalHit *fh = B.fhits + read * (u64)S * (u64)B.cap_h;
  DevResult res = B.results[read];  // query_length was filled by select_plan
  res.score = res.secondary_score = 0;
  res.hit_length = 0;
  res.n_assign = 0;
  res.by_rank = 0;
  u64 *out = B.out_ids + read * (u64)P.ids_stride;
  const u64 a = w.arena_base;
  int nb = 0;
  score_read(ix, P, fh, (int)w.n_hits, B.seq_ids + a, B.rec0 + a, B.rec1 + a, B.best + a, B.tmp + a, res, out,
             err_flags, &nb);
  for (int i = res.n_assign; i < P.ids_stride; ++i) out[i] = 0;  // unused id slots read as 0
  if (B.exp_cnt) {
    u32 *cc = B.exp_cnt + read * (u64)P.ids_stride;
    for (int i = 0; i < P.ids_stride; ++i) cc[i] = 0;
    B.exp_off[read] = 0;
    if (res.by_rank) {  // the scoring records are done with: their space holds the lists until they are copied out
      u64 *lists = reinterpret_cast<u64 *>(B.rec0 + a);
      const int total = tax_expand(ix, B.best + a, nb, P.max_result, out, res.n_assign, B.tmp + a, lists, cc, err_flags);
      if (total > 0) {
        const u64 at = claim_entries(B.exp_used, (u64)total);
        if (at + (u64)total <= B.exp_cap) {
          for (int i = 0; i < total; ++i) B.exp_ids[at + i] = lists[i];
          B.exp_off[read] = at;
        } else {
          *err_flags |= 2ull;
        }
      }
    }
  }
  B.results[read] = res;
  return res.n_assign;
}

}  // namespace cfrb200

// cfr_kernels.cuh -- __global__ wrappers around the stage functions of
// cfr_pipeline.cuh.  sm_100a only.  All kernels are grid-stride so the launch
// grid can be sized as a multiple of the SM count independent of the batch.
#pragma once
#include <cuda_runtime.h>

#include "cfr_pipeline.cuh"

namespace cfrb200 {

__device__ __forceinline__ void flush_counts(OpCount oc, DevCounters *c) {
  const unsigned full = 0xffffffffu;
  oc_fold(oc);
  const u32 r = __reduce_add_sync(full, oc.rank), a = __reduce_add_sync(full, oc.access),
            s = __reduce_add_sync(full, oc.search), l = __reduce_add_sync(full, oc.locate),
            f = __reduce_add_sync(full, oc.lf), e = __reduce_add_sync(full, oc.extend);
  if ((threadIdx.x & 31) == 0) {
    if (r) atomicAdd(&c->n_rank, (u64)r);
    if (a) atomicAdd(&c->n_access, (u64)a);
    if (s) atomicAdd(&c->n_search, (u64)s);
    if (l) atomicAdd(&c->n_locate, (u64)l);
    if (f) atomicAdd(&c->n_lf, (u64)f);
    if (e) atomicAdd(&c->n_extend, (u64)e);
  }
}

// bytes -> 2-bit codes + N bits, one 32-base word per thread (coalesced 32-byte reads)
__global__ void __launch_bounds__(256) k_encode(const __grid_constant__ ChunkDev B, const u64 total_bytes) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < B.n_words; w += stride) encode_stage(B, w, total_bytes);
}

// read offsets of a batch whose reads all have one length: off[i] = first + i * len
__global__ void k_fill_offsets(u64 *off, u64 n, u64 first, u64 len) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) off[i] = first + i * len;
}

// diagnostics: the uploaded bytes with the DUST intervals replaced by 'N'
__global__ void k_apply_dust(const __grid_constant__ ChunkDev B, unsigned char *out, const u64 total_bytes) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < total_bytes; q += stride)
    out[q] = ((B.dust_bits[q >> 5] >> (q & 31)) & 1u) ? (unsigned char)'N' : B.seq_raw[q];
}

// SDUST screen: one mate per thread, registers only.  Mates that may hold a masked interval are
// appended to B.dust_list (one atomic per warp); k_dust then runs the full SDUST on those only.
// MINB (here and in k_dust / k_score): resident blocks per SM the register allocation must allow.  32-register variants
// (MINB 16) fit twice as often beside the search kernel of a neighbouring batch but spill: SDUST alone 2.0 -> 2.7 ms per
// batch and the overlapped step 7.5 -> 8.6 ms (profiles/r02_sweeps.md) -- not used.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_dust_screen(const __grid_constant__ ChunkDev B) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const u64 ntask = B.n_reads * (u64)B.mates;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t - lane < ntask; t += stride) {
    const bool need = t < ntask && dust_screen_stage(B, t);
    const u32 m = __ballot_sync(full, need);
    if (m == 0) continue;
    u32 base = 0;
    if (lane == 0) base = atomicAdd(B.dust_list_n, (u32)__popc(m));
    base = __shfl_sync(full, base, 0);
    if (need) B.dust_list[base + (u32)__popc(m & ((1u << lane) - 1u))] = (u32)t;
  }
}

// SDUST: one mate per thread.  The data-dependent triplet counters and the window
// ring live in shared memory, one bank column per ACTIVE thread (80 words each), so their
// updates are conflict-free single wavefronts.  AL = lanes per warp that take mates:
//   32  every lane (unscreened input: most mates never leave the common step): 40 KiB per block
//   16  the launch behind the screen: every mate runs the serial loops (shrink, FindPerfect),
//       which stall the other mates of its warp, so the mates are spread over twice as many warps;
//       columns only for the active lanes (20 KiB per block) let twice as many blocks be resident
enum { CFR_DUST_THREADS = 128 };
template <int AL, int MINB>
__global__ void __launch_bounds__(CFR_DUST_THREADS, MINB) k_dust(const __grid_constant__ ChunkDev B, const int quorum) {
  extern __shared__ u32 dust_sm[];
  constexpr int COLS = (CFR_DUST_THREADS / 32) * AL;  // active threads per block
  const int lane = threadIdx.x & 31;
  const int col = (threadIdx.x >> 5) * AL + (lane < AL ? lane : 0);
  DustStateT<COLS> d;
  d.cc.base = reinterpret_cast<unsigned char *>(&dust_sm[col]);               // 64 words
  d.win.base = reinterpret_cast<unsigned char *>(&dust_sm[64 * COLS + col]);  // 16 words
  // mates are claimed dynamically from B.dust_counter (through B.dust_list when the screen ran)
  dust_tasks(B, B.dust_list ? (u64)*B.dust_list_n : B.n_reads * (u64)B.mates, d, quorum, lane < AL,
             B.dust_list != nullptr);
}
template <int AL>
constexpr int dust_smem_bytes() { return 80 * (CFR_DUST_THREADS / 32) * AL * 4; }

// MINB = resident blocks per SM the register allocation must allow (occupancy knob)
template <class Bwt, int MINB>
__global__ void __launch_bounds__(128, MINB) k_search(const __grid_constant__ DevIndex ix,
                                                      const __grid_constant__ DevParams P,
                                                      const __grid_constant__ ChunkDev B) {
  OpCount oc{};
  const u64 ntask = B.n_reads * (u64)(2 * B.mates);
  if constexpr (Bwt::PAIR == 2) Bwt::block_init();
  search_tasks<Bwt>(ix, P, B, ntask, oc);  // tasks are claimed dynamically from B.task_counter
  if (!Bwt::leader()) oc = OpCount{};
  flush_counts(oc, B.counters + CFR_STAGE_SEARCH);
}

// One read per thread.  The arena slice of each read is reserved with ONE atomic
// per warp (warp-wide inclusive scan of the row counts).
template <class Bwt>
__global__ void __launch_bounds__(128) k_select(const __grid_constant__ DevIndex ix, const __grid_constant__ DevParams P,
                                                const __grid_constant__ ChunkDev B, const int first_pass) {
  OpCount oc{};
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t - lane < B.n_list; t += stride) {
    const bool active = t < B.n_list;
    const u64 read = active ? chunk_read_id(B, t) : 0;
    u32 r = 0;
    if (active) r = first_pass ? select_plan<Bwt>(ix, P, B, read, oc) : B.work[read].arena_rows;
    u64 incl = r;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u64 v = __shfl_up_sync(full, incl, d);
      if (lane >= d) incl += v;
    }
    const u64 total = __shfl_sync(full, incl, 31);
    u64 base0 = 0;
    if (lane == 0 && total) base0 = atomicAdd(B.arena_used, total);
    base0 = __shfl_sync(full, base0, 0);
    if (active) {
      const u64 my_base = base0 + incl - r;
      const bool fits = my_base + r <= B.arena_cap;
      oc.locate += select_write_rows(ix, P, B, read, my_base, fits);
      if (!fits) {
        B.deferred[atomicAdd(B.n_deferred, 1u)] = (u32)read;
        atomicMin(B.arena_valid, my_base);
      }
    }
  }
  flush_counts(oc, B.counters + CFR_STAGE_SELECT);
}

template <class Bwt>
__global__ void __launch_bounds__(128) k_locate(const __grid_constant__ DevIndex ix, const __grid_constant__ DevParams P,
                                                const __grid_constant__ ChunkDev B) {
  OpCount oc{};
  // reservations are monotone, so the rows below the first read that did not fit are exactly the written ones
  u64 used = *B.arena_used;
  if (used > *B.arena_valid) used = *B.arena_valid;
  if (used > B.arena_cap) used = B.arena_cap;
  locate_rows<Bwt>(ix, P, B, used, oc);  // rows are claimed dynamically from B.row_counter
  if (!Bwt::leader()) oc = OpCount{};
  flush_counts(oc, B.counters + CFR_STAGE_LOCATE);
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_score(const __grid_constant__ DevIndex ix, const __grid_constant__ DevParams P,
                                                     const __grid_constant__ ChunkDev B) {
  const unsigned full = 0xffffffffu;
  u64 err = 0;
  u32 n_done = 0, n_cls = 0;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < B.n_list; t += stride) {
    const u64 read = chunk_read_id(B, t);
    if (B.work[read].status != 0) continue;
    const int na = score_stage(ix, P, B, read, &err);
    ++n_done;
    if (na > 0) {
      ++n_cls;
      const DevResult &res = B.results[read];
      const u64 *ids = B.out_ids + read * (u64)P.ids_stride;
      const int m = na < P.ids_stride ? na : P.ids_stride;
      for (int i = 0; i < m; ++i) {
        u64 ct = ids[i];
        if (!res.by_rank) ct = ct < ix.seq_cnt ? (u64)ld32(ix.seq_to_tax + ct) : ix.node_cnt;
        if (ct > ix.node_cnt) ct = ix.node_cnt;
        atomicAdd(&B.taxon_counts[ct], 1ull);
      }
    }
  }
  n_done = __reduce_add_sync(full, n_done);
  n_cls = __reduce_add_sync(full, n_cls);
  if ((threadIdx.x & 31) == 0) {
    if (n_done) {
      atomicAdd(&B.counters[CFR_STAGE_SCORE].n_reads, (u64)n_done);
      atomicAdd(&B.taxon_counts[ix.node_cnt + 1], (u64)n_done);
    }
    if (n_cls) atomicAdd(&B.taxon_counts[ix.node_cnt + 2], (u64)n_cls);
  }
  if (err) atomicOr(&B.counters[CFR_STAGE_SCORE].error_flags, err);
}

// the largest arena slice any read of the list needs (a read that alone exceeds the arena: the arena is regrown to it)
__global__ void __launch_bounds__(256) k_max_arena_rows(const __grid_constant__ ChunkDev B, unsigned long long *out) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  u32 m = 0;
  for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < B.n_list; t += stride) {
    const u32 r = B.work[chunk_read_id(B, t)].arena_rows;
    m = r > m ? r : m;
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, (unsigned long long)m);
}

// Index transcoding at load time: run-block BWT (as stored) -> 32-byte occ sectors.
// One thread per sector: 3 exclusive Sequence_RunBlock::Rank queries give the
// counters, 64 Sequence_RunBlock::Access queries give the symbols -- the literal
// device port of the reference's rank/access does the decoding.
__global__ void __launch_bounds__(128) k_transcode(const __grid_constant__ DevIndex ix, OccLine *out, u64 n_lines) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 L = (u64)blockIdx.x * blockDim.x + threadIdx.x; L < n_lines; L += stride) {
    const u64 p0 = L * 64;
    u64 cnt[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) cnt[c] = p0 <= ix.n ? rb_rank(ix, c, p0, 0) : 0;
    u64 lo = 0, hi = 0;
    for (int w = 0; w < 64; ++w) {
      const u64 pos = p0 + (u64)w;
      if (pos >= ix.n) break;
      const int s = rb_access(ix, pos);
      lo |= (u64)(s & 1) << w;
      hi |= (u64)(s >> 1) << w;
    }
    out[L] = occ_pack(lo, hi, cnt[0], cnt[1], cnt[2]);
  }
}

// ---- quantification: a read's assignment as the quantifier keys it
// (Quantifier.hpp:515-622): record of W = K + 1 words per read --
//   word 0      number of targets | weight class << 8 | unique flag << 16; 0xffffffff for a read that does not count
//               (unclassified, or below --min-score / --min-length)
//   word 1..K   compact taxonomy ids of the targets in row order (unknown -> the root), 0xffffffff beyond
// weight class d: weight 4^-d, CalculateAssignmentWeight (:283-293)
__global__ void __launch_bounds__(256) k_quant_keys(const __grid_constant__ DevIndex ix, const DevResult *res, const u64 *ids, u64 n,
                                                    int K, u64 min_score, u64 min_hit, u32 *words) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const int W = K + 1;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const DevResult r = res[i];
    u32 *w = words + i * (u64)W;
    const int na = r.n_assign < K ? r.n_assign : K;
    if (na <= 0 || (u64)(long long)r.hit_length < min_hit || r.score < min_score) {
      for (int j = 0; j < W; ++j) w[j] = 0xffffffffu;
      continue;
    }
    int diff = r.query_length - r.hit_length;
    const int slack = (int)((double)(u64)(long long)r.query_length * 0.01);
    int wc = 0;
    if (diff >= slack) {
      diff -= slack;
      wc = diff > 10 ? 11 : diff;
    }
    w[0] = (u32)na | ((u32)wc << 8) | (r.score > r.secondary_score ? 1u << 16 : 0u);
    for (int j = 0; j < K; ++j) {
      u32 t = 0xffffffffu;
      if (j < na) {
        const u64 id = ids[i * (u64)K + j];
        u64 ct = r.by_rank ? id : (id < ix.seq_cnt ? (u64)ix.seq_to_tax[id] : ix.node_cnt);
        if (ct >= ix.node_cnt) ct = ix.root;  // printed as the root's id, read back as the root (Taxonomy.hpp:633-652)
        t = (u32)ct;
      }
      w[1 + j] = t;
    }
  }
}

// one pass of the record sort: the 64-bit key made of words (2p, 2p+1) of the records in their current order
__global__ void __launch_bounds__(256) k_quant_gather(const u32 *words, const u32 *idx, u64 n, int W, int p, u64 *keys, u32 *idx_init) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const u32 i = idx ? idx[j] : (u32)j;
    if (idx_init) idx_init[j] = (u32)j;
    const u32 *w = words + (u64)i * W;
    const u64 lo = w[2 * p], hi = 2 * p + 1 < W ? w[2 * p + 1] : 0;
    keys[j] = lo | (hi << 32);
  }
}

__global__ void __launch_bounds__(256) k_quant_heads(const u32 *words, const u32 *idx, u64 n, int W, unsigned char *flags) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    bool head = j == 0;
    if (!head) {
      const u32 *a = words + (u64)idx[j] * W, *b = words + (u64)idx[j - 1] * W;
      for (int k = 0; k < W; ++k) head |= a[k] != b[k];
    }
    flags[j] = head ? 1 : 0;
  }
}

// entry e: the record of run e followed by its length
__global__ void __launch_bounds__(256) k_quant_emit(const u32 *words, const u32 *idx, const u32 *heads, const u32 *num, u64 n, int W,
                                                    u32 *out) {
  const u64 m = *num, stride = (u64)gridDim.x * blockDim.x;
  for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += stride) {
    const u32 h0 = heads[e], h1 = e + 1 < m ? heads[e + 1] : (u32)n;
    const u32 *w = words + (u64)idx[h0] * W;
    u32 *o = out + e * (u64)(W + 1);
    for (int k = 0; k < W; ++k) o[k] = w[k];
    o[W] = h1 - h0;
  }
}

// Pair lines at load time (cfr_core.cuh, layout 3): planes of every line, totals per chunk of 64 lines,
// [exclusive scan of the totals on the host side of the launch sequence], superblock table, counters
__global__ void __launch_bounds__(128) k_pair_planes(const __grid_constant__ DevIndex ix, PairLine *lines, u64 n_lines) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 L = (u64)blockIdx.x * blockDim.x + threadIdx.x; L < n_lines; L += stride) {
    u64 a, b, c, d;
    pair_line_planes(ix, L, a, b, c, d);
    u64 *w = lines[L].w;
    w[0] = a;
    w[1] = b;
    w[2] = c;
    w[3] = d;
  }
}

__global__ void __launch_bounds__(128) k_pair_totals(const __grid_constant__ DevIndex ix, const PairLine *lines, u64 n_lines,
                                                     u64 *tot, u64 n_chunk) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 ch = (u64)blockIdx.x * blockDim.x + threadIdx.x; ch < n_chunk; ch += stride) {
    u64 sum[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) sum[k] = 0;
    for (u64 L = ch * CFR_PAIR_CHUNK; L < (ch + 1) * CFR_PAIR_CHUNK && L < n_lines; ++L) {
      const u64 *w = lines[L].w;
      u32 cnt[20];
      pair_line_counts(w[0], w[1], w[2], w[3], pair_valid_mask(ix, L), cnt);
#pragma unroll
      for (int k = 0; k < 20; ++k) sum[k] += cnt[k];
    }
#pragma unroll
    for (int k = 0; k < 20; ++k) tot[(u64)k * n_chunk + ch] = sum[k];
  }
}

__global__ void k_pair_sb(const u64 *tot, u64 n_chunk, u64 *sb_table, u64 n_sb) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_sb * 20) return;
  const u64 sb = t / 20, k = t % 20;
  sb_table[t] = tot[k * n_chunk + ((sb << CFR_PAIR_SB_SHIFT) / CFR_PAIR_CHUNK)];
}

__global__ void __launch_bounds__(128) k_pair_counters(const __grid_constant__ DevIndex ix, PairLine *lines, u64 n_lines,
                                                       const u64 *tot, u64 n_chunk, const u64 *sb_table) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 ch = (u64)blockIdx.x * blockDim.x + threadIdx.x; ch < n_chunk; ch += stride)
    pair_chunk_counters(ix, lines, n_lines, ch, tot, n_chunk, sb_table);
}

__global__ void k_pair_consts(const __grid_constant__ DevIndex ix, u64 *out18) {
  if (blockIdx.x == 0 && threadIdx.x == 0) pair_constants(ix, out18);
}

// Wide lookup table at load time: one entry per WW-mer (see wide_lookup_entry)
template <class Bwt>
__global__ void __launch_bounds__(128) k_build_wide(const __grid_constant__ DevIndex ix, u64x2 *out, int WW) {
  const u64 n_keys = 1ull << (2 * WW);
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 key = (u64)blockIdx.x * blockDim.x + threadIdx.x; key < n_keys; key += stride)
    out[key] = wide_lookup_entry<Bwt>(ix, key, WW);
}

// Dense locate table at load time: the literal walk (FMIndex::BackwardToSampledSA over the stored
// samples only) from every row that is a multiple of 2^shift
// One level of the table: the rows that are multiples of 2^level (`skip_coarser`: except the multiples of 2^(level+1),
// which the level before wrote).  `ix` carries the table as built so far -- dense_shift = level + 1, or -1 at the first
// level -- so a walk ends at the first row of a coarser level with the answer the literal walk from there found: the
// levels together cost ~4 LF steps per row instead of the 15 of a walk to the stored samples.
template <class Bwt>
__global__ void __launch_bounds__(128) k_build_dense(const __grid_constant__ DevIndex ix, u32 *out, int level, int skip_coarser,
                                                     u64 n_rows, int e16) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  OpCount oc{};
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n_rows; j += stride) {
    if (skip_coarser && (j & 1ull) == 0) continue;
    u64 i = j << level, sa = 0;
    while (!get_located(ix, i, sa)) i = Bwt::lf(ix, i, oc);
    const u64 at = (j << level) >> ix.dense_idx_shift;
    if (e16) reinterpret_cast<unsigned short *>(out)[at] = (unsigned short)sa;
    else out[at] = (u32)sa;
  }
}

// ---- diagnostics for the parity tests ----
template <class Bwt>
__global__ void k_debug_rank(const __grid_constant__ DevIndex ix, const unsigned char *codes, const u64 *pos,
                             const int *incl, u64 n, u64 *out) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = Bwt::rank(ix, codes[i], pos[i], incl[i]);
}

template <class Bwt>
__global__ void k_debug_access(const __grid_constant__ DevIndex ix, const u64 *pos, u64 n, unsigned char *out) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (unsigned char)Bwt::access(ix, pos[i]);
}

template <class Bwt>
__global__ void k_debug_locate(const __grid_constant__ DevIndex ix, const u64 *rows, u64 n, u64 *out) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  OpCount oc{};
  if (i < n) out[i] = locate_row<Bwt>(ix, rows[i], oc);
}

}  // namespace cfrb200

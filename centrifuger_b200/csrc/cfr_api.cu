// cfr_api.cu -- the C ABI of include/centrifuger_b200.h: index load into HBM,
// batch pipeline orchestration, result hand-back.  Host C++ + CUDA runtime only.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <chrono>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <strings.h>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/centrifuger_b200.h"
#include "cfr_format.hpp"
#include "cfr_kernels.cuh"
#include "cfr_quant.hpp"

using namespace cfrb200;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return fail(e_ == cudaErrorMemoryAllocation ? CFR_ERR_NOMEM : CFR_ERR_CUDA,               \
                  std::string(#expr) + ": " + cudaGetErrorString(e_));                          \
  } while (0)

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {  // grow-only
    if (need <= bytes) return CFR_OK;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    size_t want = need + need / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return fail(CFR_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    bytes = want;
    return CFR_OK;
  }
  // grow keeping the first `keep` bytes (ordered after the work already enqueued on `s`)
  int grow_keeping(size_t need, size_t keep, cudaStream_t s) {
    if (need <= bytes) return CFR_OK;
    void *q = nullptr;
    const size_t want = need + need / 8 + 256;
    cudaError_t e = cudaMalloc(&q, want);
    if (e != cudaSuccess) return fail(CFR_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    if (p && keep) cudaMemcpyAsync(q, p, std::min(keep, bytes), cudaMemcpyDeviceToDevice, s);
    cudaStreamSynchronize(s);
    if (p) cudaFree(p);
    p = q;
    bytes = want;
    return CFR_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

}  // namespace

// One device-resident chunk of reads with all the work areas its pipeline needs.
struct cfr_device_batch {
  bool quant_counted = false;  // quantification has seen this classification of the batch
  u64 n_reads = 0;
  int mates = 1;
  int cap_h = 1;
  u64 seq_bytes = 0, total_bases = 0;
  u64 off_bias[2] = {0, 0};
  u64 uni_len[2] = {0, 0}, uni_pos0[2] = {0, 0};
  u64 arena_cap = 0;
  DevBuf seq_raw, codes, mask_raw, mask, off, strand_hits, strand_nhits, fhits, work, rows, seq_ids, rec0, rec1, best, tmp,
      results, out_ids, deferred, dust_list, scalars;  // scalars: {u64 arena_used, u32 n_deferred, pad}
  DevBuf dust_bits, masked;  // only when the caller wants the masked reads back (cfr_submit_batch_masked)
  DevBuf exp_cnt, exp_off, exp_ids;  // only with cfr_params.expand_taxid
  int ticket = -1;                   // streaming ticket this slot last served
  bool want_masked = false;
  bool classified = false;
  bool packed = false;  // uploaded as 2-bit codes + N bits (cfr_submit_packed): no k_encode pass
  cudaEvent_t ev_lane[4] = {nullptr, nullptr, nullptr, nullptr};  // in, masked, searched, scored (stage lanes, run_first)
  void release() {
    for (cudaEvent_t &e : ev_lane) {
      if (e) cudaEventDestroy(e);
      e = nullptr;
    }
    DevBuf *all[] = {&seq_raw, &codes, &mask_raw, &mask, &off, &strand_hits, &strand_nhits, &fhits, &work, &rows, &seq_ids,
                     &rec0, &rec1, &best, &tmp, &results, &out_ids, &deferred, &dust_list, &scalars, &dust_bits, &masked,
                     &exp_cnt, &exp_off, &exp_ids};
    for (DevBuf *b : all) b->release();
  }
};

struct cfr_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cfr_params params;
  CfrIndexFile file;
  DevIndex ix;
  DevParams P;
  int layout = CFR_LAYOUT_RUNBLOCK;
  std::vector<void *> index_allocs;
  size_t hbm_bytes = 0;
  std::vector<size_t> index_alloc_bytes;  // parallel to index_allocs
  size_t occ_bytes = 0, wide_bytes = 0, dense_bytes = 0, runblock_bytes = 0, pair_bytes = 0;
  double open_seconds = 0.0;
  int sm_count = 148;
  int search_blocks = 10;  // resident 128-thread blocks per SM targeted by k_search (CFR_B200_SEARCH_BLOCKS)
  // how the pair-line search kernel waits for memory (CFR_B200_PAIR_FETCH): 3 = one wait per iteration (lines, far lines and
  // wide-table entries in one group of cp.async copies), 2 = cp.async staging rounds per request kind, 1 = LDG + STS rounds
  int pair_fetch = 3;
  // the same for the pair-line search kernel (CFR_B200_PAIR_SEARCH_BLOCKS): six blocks keep the memory system as busy as
  // eight do (5.4 vs 5.5 ms alone) and leave registers for the latency-bound kernels of the neighbouring batches (SDUST,
  // scoring) on the same SMs; five make the kernel itself slower (5.9 ms) for no gain of the whole step (profiles/r02_sweeps.md)
  int pair_search_blocks = 6;
  int occ_load = 4;    // how k_search / k_locate fetch a sector: 4 = one 256-bit load, 0 = two 128-bit loads (CFR_B200_OCC_LOAD)
  bool pos32 = false;  // 32-bit BWT positions in k_search / k_locate (collections below 2^32 rows; CFR_B200_POS64=1 disables)
  int dust_quorum = 0;  // quorum of the SDUST state machine (0 = the search quorum; CFR_B200_DUST_QUORUM)
  int dust_lanes = 16;  // lanes per warp that take mates in the post-screen SDUST launch (CFR_B200_DUST_LANES)
  bool dust_screen = true;  // register-only screen in front of the full SDUST (CFR_B200_DUST_SCREEN=0 disables)
  u64 *d_taxon = nullptr;
  u64 *d_taxon_reduced = nullptr;  // snapshot the NCCL all-reduce works on
  // quantification (cfr_quant_enable): every finished batch is coalesced on the device, the distinct
  // assignment records and their multiplicities accumulate here
  struct Quant {
    bool on = false;
    u64 min_score = 0, min_hit = 0;
    DevBuf words, keys_a, keys_b, idx_a, idx_b, flags, heads, num, out, tmp;
    std::map<std::vector<u32>, u64> table;
    u64 batches = 0, entries_moved = 0;
  } quant;
  DevCounters *d_counters = nullptr;
  u64 launches = 0;
  u64 host_bases = 0;
  u64 h2d_bytes = 0, d2h_bytes = 0;  // bytes the batch paths really moved over the host link
  // cfr_classify_batch pipeline: two chunk slots, H2D / compute / D2H on three streams
  // NSLOT batches can be in flight in the streaming form (upload of batch i+2 next to the kernels of
  // i+1 and the download of i); cfr_classify_batch's chunk pipeline uses the first two
  enum { NSLOT = 3 };
  cfr_device_batch slots[NSLOT];
  cudaStream_t s_in = nullptr, s_out = nullptr, s_comp[NSLOT] = {nullptr, nullptr, nullptr};
  // Stage lanes (CFR_B200_LANES=0 disables): the stages of a batch run on three streams BY KIND -- encode + SDUST, search,
  // select + locate + score -- chained by events, so that the memory-bound search of batch i always has the latency-bound
  // stages of batches i+1 / i-1 next to it on the SMs, whatever streams the caller's batches arrive on
  cudaStream_t lane[3] = {nullptr, nullptr, nullptr};
  bool lanes = true;
  cudaEvent_t ev_start = nullptr, ev_h2d[NSLOT] = {nullptr, nullptr, nullptr}, ev_comp[NSLOT] = {nullptr, nullptr, nullptr},
              ev_d2h[NSLOT] = {nullptr, nullptr, nullptr};
  struct PinnedScalars {
    u64 used;
    u32 n_def;
    u32 pad;
  } *pinned_scalars = nullptr;  // [NSLOT], cudaHostAlloc
  // cfr_submit_batch / cfr_wait_batch: one job per slot
  struct Job {
    int ticket = -1;  // -1 = slot idle
    cfr_result *results = nullptr;
    uint64_t *ids = nullptr;
  } jobs[NSLOT];
  int next_ticket = 0;
  // CFR_B200_TRACE=1: device timeline of the streaming path (printed by cfr_wait_batch)
  bool trace = false;
  cudaEvent_t tr_base = nullptr, tr_ev[NSLOT][4] = {};
  double tr_host_submit[NSLOT] = {};
  // stage profiling (CUDA events on the launch stream)
  bool profile = false;
  struct EvPair {
    cudaEvent_t a, b;
    int stage;
  };
  std::vector<EvPair> ev_pending;
  std::vector<cudaEvent_t> ev_pool;
  double stage_ms[CFR_N_STAGES] = {0, 0, 0, 0, 0, 0};
  u64 stage_launches[CFR_N_STAGES] = {0, 0, 0, 0, 0, 0};
};

namespace {

int dev_alloc(cfr_handle *h, void **out, size_t bytes) {
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
  if (e != cudaSuccess) return fail(CFR_ERR_NOMEM, std::string("cudaMalloc(index): ") + cudaGetErrorString(e));
  h->index_allocs.push_back(p);
  h->index_alloc_bytes.push_back(bytes);
  h->hbm_bytes += bytes;
  *out = p;
  return CFR_OK;
}

// give one index array back (the run-block arrays once the occ sectors exist)
void dev_release(cfr_handle *h, const void *p) {
  if (!p) return;
  for (size_t i = 0; i < h->index_allocs.size(); ++i)
    if (h->index_allocs[i] == p) {
      cudaFree(h->index_allocs[i]);
      h->hbm_bytes -= h->index_alloc_bytes[i];
      h->index_allocs.erase(h->index_allocs.begin() + i);
      h->index_alloc_bytes.erase(h->index_alloc_bytes.begin() + i);
      return;
    }
}

// upload `bytes` from an (unaligned) host view, padded with `pad` zero bytes
int dev_upload(cfr_handle *h, const void *src, size_t bytes, size_t pad, void **out) {
  int st = dev_alloc(h, out, bytes + pad);
  if (st) return st;
  if (pad) CUDA_TRY(cudaMemset((char *)*out + bytes, 0, pad));
  if (bytes) CUDA_TRY(cudaMemcpy(*out, src, bytes, cudaMemcpyHostToDevice));
  return CFR_OK;
}

int upload_bv(cfr_handle *h, const BvView &v, DevBV &d) {
  d.n = v.nbits;
  d.B = nullptr;
  d.R = nullptr;
  if (v.nbits == 0) return CFR_OK;
  void *p;
  int st = dev_upload(h, v.B, v.words * 8, 16, &p);
  if (st) return st;
  d.B = (const u64 *)p;
  st = dev_upload(h, v.R, v.rwords * 8, 16, &p);
  if (st) return st;
  d.R = (const u64 *)p;
  return CFR_OK;
}

int upload_wt(cfr_handle *h, const WtView &t, DevWT &d) {
  memset(&d, 0, sizeof(d));
  d.n = t.n;
  for (int i = 0; i < 3; ++i) {
    d.child[i][0] = t.child[i][0];
    d.child[i][1] = t.child[i][1];
    if (i < t.node_cnt) {
      int st = upload_bv(h, t.node[i], d.node[i]);
      if (st) return st;
    }
  }
  return CFR_OK;
}

int grid_for(const cfr_handle *h, u64 tasks, int threads, int blocks_per_sm) {
  u64 need = (tasks + threads - 1) / threads;
  u64 cap = (u64)h->sm_count * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)std::min(need, cap);
}

// DUST over the chunk: the register-only screen clears most mates, the full SDUST state
// machine runs on the rest (scalars must be zero: dust_counter, dust_list_n)
void launch_dust(cfr_handle *h, const ChunkDev &B, cudaStream_t s) {
  const u64 ntask = B.n_reads * (u64)B.mates;
  if (B.dust_list) {
    k_dust_screen<8><<<grid_for(h, ntask, 128, 16), 128, 0, s>>>(B);
    ++h->launches;
  }
  // after the screen only the few mates that need the whole algorithm are left: 16 lanes per warp
  const int q = h->dust_quorum ? h->dust_quorum : h->P.quorum;
  if (B.dust_list && h->dust_lanes <= 16) {
    k_dust<16, 10><<<grid_for(h, ntask, CFR_DUST_THREADS, 10), CFR_DUST_THREADS, dust_smem_bytes<16>(), s>>>(B, q);
  } else {
    k_dust<32, 10><<<grid_for(h, ntask, CFR_DUST_THREADS, 5), CFR_DUST_THREADS, dust_smem_bytes<32>(), s>>>(B, q);
  }
  ++h->launches;
}

int upload_index(cfr_handle *h) {
  CfrIndexFile &f = h->file;
  DevIndex &ix = h->ix;
  memset(&ix, 0, sizeof(ix));
  ix.n = f.n;
  ix.first_isa = f.first_isa;
  ix.last_code = base_code((unsigned char)f.last_chr);
  for (int i = 0; i < 5; ++i) ix.C[i] = f.C[i];
  ix.b = f.b;
  ix.block_cnt = f.block_cnt;
  int st;
  if ((st = upload_bv(h, f.block_type, ix.block_type))) return st;
  if ((st = upload_wt(h, f.plain, ix.plain))) return st;
  if ((st = upload_wt(h, f.run, ix.run))) return st;
  ix.sample_rate = f.sample_rate;
  ix.sample_shift = (f.sample_rate & (f.sample_rate - 1)) == 0 ? __builtin_ctz((unsigned)f.sample_rate) : -1;
  ix.sa_bits = f.sa_bits;
  void *p;
  if ((st = dev_upload(h, f.sa_w, f.sa_words * 8, 16, &p))) return st;
  ix.sampled_sa = (const u64 *)p;
  ix.adjusted_sa0 = f.adjusted_sa0;
  if ((st = dev_upload(h, f.sel, f.sel_cnt * 16, 16, &p))) return st;
  ix.sel = (const u64x2 *)p;
  ix.sel_cnt = f.sel_cnt;
  ix.sel_filter_rate = f.sel_filter_rate;
  ix.filter_shift = (f.sel_filter_rate & (f.sel_filter_rate - 1)) == 0 ? __builtin_ctz((unsigned)f.sel_filter_rate) : -1;
  ix.sel_filter = nullptr;
  if (f.sel_cnt > 0) {  // rebuilt at load exactly as FMIndex.hpp:163-176 does
    const u64 fbits = (f.n + (u64)f.sel_filter_rate - 1) / (u64)f.sel_filter_rate;
    std::vector<u64> filt(fbits / 64 + 2, 0);
    for (u64 i = 0; i < f.sel_cnt; ++i) {
      const u64 fb = load_u64(f.sel + i * 16) / (u64)f.sel_filter_rate;
      filt[fb >> 6] |= 1ull << (fb & 63);
    }
    if ((st = dev_upload(h, filt.data(), filt.size() * 8, 0, &p))) return st;
    ix.sel_filter = (const u64 *)p;
  }
  ix.dense_shift = -1;
  ix.dense_idx_shift = 0;
  ix.pre_width = (int)f.precompute_width;
  if ((st = dev_upload(h, f.lookup, f.precompute_size * 16, 16, &p))) return st;
  ix.lookup = (const u64x2 *)p;
  // taxonomy
  const TaxonomyHost &t = f.tax;
  if (t.node_cnt >= 0xfffffff0ull || t.seq_cnt + t.extra_seq_cnt >= 0xfffffff0ull)
    return fail(CFR_ERR_UNSUPPORTED, "taxonomy too large for 32-bit device ids");
  ix.node_cnt = t.node_cnt;
  ix.seq_cnt = t.seq_cnt;
  ix.root = t.root;
  std::vector<u32> parent(t.node_cnt + 1, 0), s2t(t.seq_cnt + 1, 0);
  std::vector<unsigned char> rank(t.node_cnt + 1, 0);
  for (u64 i = 0; i < t.node_cnt; ++i) {
    parent[i] = (u32)t.parent[i];
    rank[i] = t.rank[i];
  }
  for (u64 i = 0; i < t.seq_cnt; ++i) s2t[i] = t.seq_to_tax[i] >= t.node_cnt ? (u32)t.node_cnt : (u32)t.seq_to_tax[i];
  if ((st = dev_upload(h, parent.data(), parent.size() * 4, 0, &p))) return st;
  ix.parent = (const u32 *)p;
  if ((st = dev_upload(h, rank.data(), rank.size(), 16, &p))) return st;
  ix.rank = (const unsigned char *)p;
  if ((st = dev_upload(h, s2t.data(), s2t.size() * 4, 0, &p))) return st;
  ix.seq_to_tax = (const u32 *)p;
  init_tax_rank_num(ix.rank_num);
  return CFR_OK;
}

int build_occ_lines(cfr_handle *h) {
  const u64 n_lines = h->ix.n / 64 + 1;
  void *p;
  int st = dev_alloc(h, &p, n_lines * sizeof(OccLine));
  if (st) return st;
  k_transcode<<<grid_for(h, n_lines, 128, 16), 128, 0, h->stream>>>(h->ix, (OccLine *)p, n_lines);
  ++h->launches;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->ix.occ = (const OccLine *)p;
  return CFR_OK;
}

// Pair lines (DevIndex::pairs, cfr_core.cuh "Layout 3"): built from the occ sectors; k_search then does two
// BackwardExtend steps per DRAM line.  2 bytes per BWT row.
int build_pair_lines(cfr_handle *h) {
  const u64 n_lines = h->ix.n / 64 + 1, n_chunk = (n_lines + CFR_PAIR_CHUNK - 1) / CFR_PAIR_CHUNK;
  const u64 n_sb = ((n_lines - 1) >> CFR_PAIR_SB_SHIFT) + 1;
  void *p_lines, *p_sb;
  int st = dev_alloc(h, &p_lines, n_lines * sizeof(PairLine));
  if (st) return st;
  if ((st = dev_alloc(h, &p_sb, n_sb * 20 * 8))) return st;
  DevBuf tot, k18, tmp;
  if ((st = tot.ensure(20 * n_chunk * 8))) return st;
  if ((st = k18.ensure(18 * 8))) return st;
  PairLine *lines = (PairLine *)p_lines;
  u64 *d_tot = (u64 *)tot.p;
  k_pair_planes<<<grid_for(h, n_lines, 128, 16), 128, 0, h->stream>>>(h->ix, lines, n_lines);
  k_pair_totals<<<grid_for(h, n_chunk, 128, 16), 128, 0, h->stream>>>(h->ix, lines, n_lines, d_tot, n_chunk);
  CUDA_TRY(cudaGetLastError());
  if (n_chunk >= (1ull << 31)) return fail(CFR_ERR_UNSUPPORTED, "pair layout: too many chunks");
  size_t tb = 0;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_tot, d_tot, (int)n_chunk, h->stream));
  if ((st = tmp.ensure(tb))) return st;
  for (int k = 0; k < 20; ++k)
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tb, d_tot + (u64)k * n_chunk, d_tot + (u64)k * n_chunk, (int)n_chunk, h->stream));
  k_pair_sb<<<(unsigned)((n_sb * 20 + 127) / 128), 128, 0, h->stream>>>(d_tot, n_chunk, (u64 *)p_sb, n_sb);
  k_pair_counters<<<grid_for(h, n_chunk, 128, 16), 128, 0, h->stream>>>(h->ix, lines, n_lines, d_tot, n_chunk, (const u64 *)p_sb);
  k_pair_consts<<<1, 32, 0, h->stream>>>(h->ix, (u64 *)k18.p);
  h->launches += 25;
  CUDA_TRY(cudaGetLastError());
  u64 host18[18];
  CUDA_TRY(cudaMemcpyAsync(host18, k18.p, sizeof(host18), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 16; ++i) h->ix.pair_D[i] = host18[i];
  h->ix.pair_E = (int)host18[16];
  h->ix.pair_F = (int)host18[17];
  h->ix.pairs = lines;
  h->ix.pair_sb = (const u64 *)p_sb;
  h->pair_bytes = n_lines * sizeof(PairLine);
  tot.release();
  k18.release();
  tmp.release();
  return CFR_OK;
}

// Wide lookup table (DevIndex::wide): for an index that lives in HBM every rank is a DRAM line, and the
// first steps of a search -- wide ranges, two sectors each -- are the same for every read that ends
// in the same WW-mer.  One probe of a 4^WW-entry table replaces the W-mer probe and WW - W extends.
int build_wide_lookup(cfr_handle *h, int WW) {
  if (WW <= h->ix.pre_width || WW > 15 || h->ix.pre_width <= 0) return CFR_OK;
  const u64 n_keys = 1ull << (2 * WW);
  void *p;
  int st = dev_alloc(h, &p, n_keys * sizeof(u64x2));
  if (st) return st;
  const int grid = grid_for(h, n_keys, 128, 16);
  if (h->layout == CFR_LAYOUT_OCCLINE) k_build_wide<BwtOccLine><<<grid, 128, 0, h->stream>>>(h->ix, (u64x2 *)p, WW);
  else k_build_wide<BwtRunBlock><<<grid, 128, 0, h->stream>>>(h->ix, (u64x2 *)p, WW);
  ++h->launches;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->ix.wide = (const u64x2 *)p;
  h->ix.wide_width = WW;
  h->wide_bytes = n_keys * sizeof(u64x2);
  return CFR_OK;
}

// Dense locate table (DevIndex::dense): the stored samples are 2^offrate rows apart, so a locate walks
// 2^offrate - 1 LF steps on average, each one a sector from wherever the BWT lives.  HBM has room
// for a denser table; its entries are computed by the reference's own walk, so results cannot change.
int build_dense_locate(cfr_handle *h, int shift, bool e16) {
  if (shift < 0 || h->ix.sample_shift < 0 || shift >= h->ix.sample_shift) return CFR_OK;
  const u64 n_rows = ((h->ix.n - 1) >> shift) + 1;  // rows 0 .. n-1 only: row n does not exist
  const u64 esz = e16 ? 2 : 4;
  void *p;
  int st = dev_alloc(h, &p, n_rows * esz + 16);
  if (st) return st;
  // level by level, coarsest first: every 2^(sample_shift-1)-th row walks to the stored samples, the rows of each finer
  // level walk to the level before (k_build_dense)
  h->ix.dense = (const u32 *)p;
  h->ix.dense16 = e16 ? 1 : 0;
  h->ix.dense_idx_shift = shift;
  h->ix.dense_shift = -1;
  for (int level = h->ix.sample_shift - 1; level >= shift; --level) {
    const u64 rows_l = ((h->ix.n - 1) >> level) + 1;
    const int grid = grid_for(h, rows_l, 128, 16);
    const int skip = h->ix.dense_shift >= 0 ? 1 : 0;
    if (h->layout == CFR_LAYOUT_OCCLINE) k_build_dense<BwtOccLine><<<grid, 128, 0, h->stream>>>(h->ix, (u32 *)p, level, skip, rows_l, e16 ? 1 : 0);
    else k_build_dense<BwtRunBlock><<<grid, 128, 0, h->stream>>>(h->ix, (u32 *)p, level, skip, rows_l, e16 ? 1 : 0);
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->ix.dense_shift = level;
  }
  h->ix.dense_shift = shift;
  h->ix.dense16 = e16 ? 1 : 0;
  h->dense_bytes = n_rows * esz;
  return CFR_OK;
}

cudaStream_t pick_stream(cfr_handle *h, void *stream) { return stream ? (cudaStream_t)stream : h->stream; }

cudaEvent_t ev_get(cfr_handle *h) {
  if (!h->ev_pool.empty()) {
    cudaEvent_t e = h->ev_pool.back();
    h->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

// brackets the work enqueued between construction and destruction
struct StageScope {
  cfr_handle *h;
  cudaStream_t s;
  int stage;
  cudaEvent_t a = nullptr;
  StageScope(cfr_handle *h_, cudaStream_t s_, int stage_) : h(h_), s(s_), stage(stage_) {
    ++h->stage_launches[stage];
    if (h->profile) {
      a = ev_get(h);
      cudaEventRecord(a, s);
    }
  }
  ~StageScope() {
    if (a) {
      cudaEvent_t b = ev_get(h);
      cudaEventRecord(b, s);
      h->ev_pending.push_back({a, b, stage});
    }
  }
};

// -------------------------------------------------------------------------------------
// batch upload: copies reads [r0, r1) of `in` into device batch `b` and sizes the work areas
// -------------------------------------------------------------------------------------
void fill_chunk(cfr_handle *h, cfr_device_batch *b, ChunkDev &B);

// `pk` != nullptr: the reads arrive packed (cfr_packed_batch; `in` then carries only its offsets, r0 = 0)
int upload_chunk(cfr_handle *h, const cfr_read_batch *in, u64 r0, u64 r1, cfr_device_batch *b, cudaStream_t s,
                 const cfr_packed_batch *pk = nullptr) {
  const u64 n = r1 - r0;
  const int mates = (pk ? pk->off2 != nullptr : in->seq2 != nullptr) ? 2 : 1;
  if (mates == 2 && !in->off2) return fail(CFR_ERR_ARG, "seq2 given without off2");
  b->n_reads = n;
  b->mates = mates;
  b->classified = false;
  b->packed = pk != nullptr;
  const u64 s1 = in->off1[r0], e1 = in->off1[r1];
  const u64 s2 = mates == 2 ? in->off2[r0] : 0, e2 = mates == 2 ? in->off2[r1] : 0;
  if (e1 < s1 || e2 < s2) return fail(CFR_ERR_ARG, "read offsets are not ascending");
  const u64 len1 = e1 - s1, len2 = e2 - s2;
  const u64 pos2 = (len1 + 31) & ~31ull;
  b->seq_bytes = pos2 + len2;
  b->total_bases = len1 + len2;
  b->off_bias[0] = s1;
  b->off_bias[1] = s2 - pos2;  // wraps modulo 2^64 by design; positions are computed as off - bias
  int max_len = 0;
  // uniform[m]: every read of mate m has the same length (the usual case for short-read data): the
  // offsets are then an arithmetic progression written on the device instead of crossing PCIe
  bool uniform[2] = {n > 0, n > 0 && mates == 2};
  const u64 ulen[2] = {n ? in->off1[r0 + 1] - in->off1[r0] : 0, (n && mates == 2) ? in->off2[r0 + 1] - in->off2[r0] : 0};
  for (u64 i = r0; i < r1; ++i) {
    const u64 l = in->off1[i + 1] - in->off1[i];
    if (l > 0x3fffffffull) return fail(CFR_ERR_ARG, "read longer than 2^30");
    if ((int)l > max_len) max_len = (int)l;
    uniform[0] = uniform[0] && l == ulen[0];
  }
  if (mates == 2)
    for (u64 i = r0; i < r1; ++i) {
      const u64 l = in->off2[i + 1] - in->off2[i];
      if (l > 0x3fffffffull) return fail(CFR_ERR_ARG, "read longer than 2^30");
      if ((int)l > max_len) max_len = (int)l;
      uniform[1] = uniform[1] && l == ulen[1];
    }
  b->cap_h = std::max(1, max_hits_for_len(max_len, h->P.min_hit_len));
  b->uni_len[0] = uniform[0] ? ulen[0] : 0;
  b->uni_len[1] = uniform[1] ? ulen[1] : 0;
  b->uni_pos0[0] = 0;
  b->uni_pos0[1] = pos2;
  const u64 S = 2 * (u64)mates;
  int st;
  const u64 n_words = b->seq_bytes / 32 + 2;
  if (pk) {
    if (n && (s1 != 0 || (mates == 2 && s2 != pos2))) return fail(CFR_ERR_ARG, "packed batch: off1[0] must be 0 and off2[0] the next multiple of 32 after mate 1");
    if (pk->n_words < n_words) return fail(CFR_ERR_ARG, "packed batch: codes / nmask hold fewer words than the offsets need");
  }
  if (!pk && (st = b->seq_raw.ensure(b->seq_bytes + 64))) return st;
  if ((st = b->codes.ensure(n_words * 8))) return st;
  if ((st = b->mask_raw.ensure(n_words * 4))) return st;
  if ((st = b->mask.ensure(n_words * 4))) return st;
  if ((st = b->off.ensure((n + 1) * 8 * 2))) return st;
  if ((st = b->strand_hits.ensure(n * S * b->cap_h * sizeof(Hit)))) return st;
  if ((st = b->strand_nhits.ensure(n * S * sizeof(int)))) return st;
  if ((st = b->fhits.ensure(n * S * b->cap_h * sizeof(FinalHit)))) return st;
  if ((st = b->work.ensure(n * sizeof(ReadWork)))) return st;
  u64 arena = h->params.arena_rows ? h->params.arena_rows : std::max<u64>(n * 32, 1u << 16);
  b->arena_cap = arena;
  if ((st = b->rows.ensure(arena * 8))) return st;
  if ((st = b->seq_ids.ensure(arena * 4))) return st;
  if ((st = b->rec0.ensure(arena * sizeof(SeqRec)))) return st;
  if ((st = b->rec1.ensure(arena * sizeof(SeqRec)))) return st;
  if ((st = b->best.ensure(arena * 8))) return st;
  if ((st = b->tmp.ensure(arena * 8))) return st;
  if ((st = b->results.ensure(n * sizeof(DevResult)))) return st;
  if ((st = b->out_ids.ensure(n * (u64)h->P.ids_stride * 8))) return st;
  if ((st = b->deferred.ensure(n * 4 * 2))) return st;
  if ((st = b->dust_list.ensure(n * 4 * (u64)mates))) return st;
  if ((st = b->scalars.ensure(64))) return st;
  if (b->want_masked) {
    if ((st = b->dust_bits.ensure(n_words * 4))) return st;
    if ((st = b->masked.ensure(b->seq_bytes + 64))) return st;
  }
  if (h->params.expand_taxid) {  // a read's lists hold at most as many ids as it has arena rows
    if ((st = b->exp_cnt.ensure(n * (u64)h->P.ids_stride * 4))) return st;
    if ((st = b->exp_off.ensure(n * 8))) return st;
    if ((st = b->exp_ids.ensure(arena * 8))) return st;
  }
  // H2D
  if (pk) {
    CUDA_TRY(cudaMemcpyAsync(b->codes.p, pk->codes, n_words * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->mask_raw.p, pk->nmask, n_words * 4, cudaMemcpyHostToDevice, s));
    h->h2d_bytes += n_words * 12;
  } else {
    if (len1) CUDA_TRY(cudaMemcpyAsync(b->seq_raw.p, in->seq1 + s1, len1, cudaMemcpyHostToDevice, s));
    if (len2) CUDA_TRY(cudaMemcpyAsync((char *)b->seq_raw.p + pos2, in->seq2 + s2, len2, cudaMemcpyHostToDevice, s));
    h->h2d_bytes += len1 + len2;
  }
  for (int m = 0; m < mates; ++m) {
    u64 *dst = (u64 *)b->off.p + (m ? n + 1 : 0);
    const uint64_t *src = (m ? in->off2 : in->off1) + r0;
    if (uniform[m]) {
      k_fill_offsets<<<grid_for(h, n + 1, 256, 8), 256, 0, s>>>(dst, n + 1, src[0], ulen[m]);
      ++h->launches;
    } else {
      CUDA_TRY(cudaMemcpyAsync(dst, src, (n + 1) * 8, cudaMemcpyHostToDevice, s));
      h->h2d_bytes += (n + 1) * 8;
    }
  }
  if (!pk && n) {
    // layout is part of the upload: bytes -> 2-bit codes + N bits right behind the copy, on the copy's stream; a batch
    // that stays resident is encoded once however often it is classified
    ChunkDev B;
    fill_chunk(h, b, B);
    k_encode<<<grid_for(h, B.n_words, 256, 8), 256, 0, s>>>(B, b->seq_bytes);
    ++h->launches;
  }
  CUDA_TRY(cudaGetLastError());
  return CFR_OK;
}

void fill_chunk(cfr_handle *h, cfr_device_batch *b, ChunkDev &B) {
  memset(&B, 0, sizeof(B));
  B.n_reads = b->n_reads;
  B.mates = b->mates;
  B.cap_h = b->cap_h;
  B.seq_raw = (const unsigned char *)b->seq_raw.p;
  B.n_words = b->seq_bytes / 32 + 2;
  B.codes = (u64 *)b->codes.p;
  B.mask_raw = (u32 *)b->mask_raw.p;
  B.mask = h->params.dust ? (u32 *)b->mask.p : (u32 *)b->mask_raw.p;
  B.dust_bits = b->want_masked ? (u32 *)b->dust_bits.p : nullptr;
  B.off[0] = (const u64 *)b->off.p;
  B.off[1] = (const u64 *)b->off.p + (b->n_reads + 1);
  B.off_bias[0] = b->off_bias[0];
  B.off_bias[1] = b->off_bias[1];
  for (int m = 0; m < 2; ++m) {
    B.uni_len[m] = b->uni_len[m];
    B.uni_pos0[m] = b->uni_pos0[m];
  }
  B.strand_hits = (Hit *)b->strand_hits.p;
  B.strand_nhits = (int *)b->strand_nhits.p;
  B.fhits = (FinalHit *)b->fhits.p;
  B.work = (ReadWork *)b->work.p;
  B.arena_cap = b->arena_cap;
  B.arena_used = (u64 *)b->scalars.p;
  B.n_deferred = (u32 *)((char *)b->scalars.p + 8);
  B.task_counter = (u64 *)((char *)b->scalars.p + 16);
  B.row_counter = (u64 *)((char *)b->scalars.p + 24);
  B.dust_counter = (u64 *)((char *)b->scalars.p + 32);
  B.arena_valid = (u64 *)((char *)b->scalars.p + 40);
  B.dust_list_n = (u32 *)((char *)b->scalars.p + 48);
  B.dust_list = h->dust_screen ? (u32 *)b->dust_list.p : nullptr;
  B.rows = (u64 *)b->rows.p;
  B.seq_ids = (u32 *)b->seq_ids.p;
  B.rec0 = (SeqRec *)b->rec0.p;
  B.rec1 = (SeqRec *)b->rec1.p;
  B.best = (u64 *)b->best.p;
  B.tmp = (u64 *)b->tmp.p;
  B.results = (DevResult *)b->results.p;
  B.out_ids = (u64 *)b->out_ids.p;
  B.taxon_counts = h->d_taxon;
  B.counters = h->d_counters;
  B.deferred = (u32 *)b->deferred.p;
  if (h->params.expand_taxid) {
    B.exp_cnt = (u32 *)b->exp_cnt.p;
    B.exp_off = (u64 *)b->exp_off.p;
    B.exp_ids = (u64 *)b->exp_ids.p;
    B.exp_cap = b->exp_ids.bytes / 8;
    B.exp_used = (u64 *)((char *)b->scalars.p + 56);
  }
  B.read_list = nullptr;
  B.n_list = b->n_reads;
}

// one select -> locate -> score pass over B.read_list
// Bwt = scalar layout policy (select stage), BwtWide = policy of the search / locate kernels
template <class Bwt, class BwtWide>
int run_pass(cfr_handle *h, const ChunkDev &B, int first_pass, cudaStream_t s) {
  {
    StageScope sc(h, s, CFR_STAGE_OTHER);
    CUDA_TRY(cudaMemsetAsync(B.arena_used, 0, 16, s));
    CUDA_TRY(cudaMemsetAsync(B.row_counter, 0, 8, s));
    CUDA_TRY(cudaMemsetAsync(B.arena_valid, 0xff, 8, s));
  }
  {
    StageScope sc(h, s, CFR_STAGE_SELECT);
    k_select<Bwt><<<grid_for(h, B.n_list, 128, 16), 128, 0, s>>>(h->ix, h->P, B, first_pass);
  }
  const bool need_locate = !locate_in_select(h->ix);  // the dense table answered every row in k_select
  if (need_locate) {
    StageScope sc(h, s, CFR_STAGE_LOCATE);
    k_locate<BwtWide><<<grid_for(h, B.arena_cap * BwtWide::LANES, 128, 16), 128, 0, s>>>(h->ix, h->P, B);
  }
  {
    StageScope sc(h, s, CFR_STAGE_SCORE);
    k_score<10><<<grid_for(h, B.n_list, 128, 16), 128, 0, s>>>(h->ix, h->P, B);
  }
  h->launches += need_locate ? 3 : 2;
  CUDA_TRY(cudaGetLastError());
  return CFR_OK;
}

// BwtSearch = policy of the search kernel (the pair lines when they were built)
template <class Bwt, class BwtWide, class BwtSearch = BwtWide>
int run_first(cfr_handle *h, cfr_device_batch *b, cudaStream_t s) {
  ChunkDev B;
  fill_chunk(h, b, B);
  if (b->n_reads == 0) return CFR_OK;
  // the three stage lanes (or the caller's stream for everything)
  cudaStream_t s_pre = s, s_search = s, s_post = s;
  if (h->lanes) {
    for (cudaStream_t &l : h->lane)
      if (!l) CUDA_TRY(cudaStreamCreateWithFlags(&l, cudaStreamNonBlocking));
    for (cudaEvent_t &e : b->ev_lane)
      if (!e) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    s_pre = h->lane[0];
    s_search = h->lane[1];
    s_post = h->lane[2];
    CUDA_TRY(cudaEventRecord(b->ev_lane[0], s));  // ordered after what the caller's stream holds so far
    CUDA_TRY(cudaStreamWaitEvent(s_pre, b->ev_lane[0], 0));
  }
  CUDA_TRY(cudaMemsetAsync(b->scalars.p, 0, 64, s_pre));
  {  // the bases were encoded at upload (or by the producer): the masks the searches see start as the uploaded ones
    StageScope sc(h, s_pre, CFR_STAGE_OTHER);
    if (B.mask != B.mask_raw) CUDA_TRY(cudaMemcpyAsync(B.mask, B.mask_raw, B.n_words * 4, cudaMemcpyDeviceToDevice, s_pre));
    if (B.dust_bits) CUDA_TRY(cudaMemsetAsync(B.dust_bits, 0, B.n_words * 4, s_pre));
  }
  if (h->params.dust) {
    StageScope sc(h, s_pre, CFR_STAGE_DUST);
    launch_dust(h, B, s_pre);
  }
  if (b->want_masked) {  // the reads as the searches see them, for the caller's --un / --cl files
    k_apply_dust<<<grid_for(h, b->seq_bytes, 256, 8), 256, 0, s_pre>>>(B, (unsigned char *)b->masked.p, b->seq_bytes);
    ++h->launches;
  }
  CUDA_TRY(cudaMemsetAsync(B.task_counter, 0, 8, s_pre));
  if (h->lanes) {
    CUDA_TRY(cudaEventRecord(b->ev_lane[1], s_pre));
    CUDA_TRY(cudaStreamWaitEvent(s_search, b->ev_lane[1], 0));
  }
  {
    StageScope sc(h, s_search, CFR_STAGE_SEARCH);
    // the pair policy keeps four lanes per task and more live registers: 8 resident blocks per SM do not spill
    const int sblocks = BwtSearch::PAIR ? h->pair_search_blocks : h->search_blocks;
    const int g = grid_for(h, B.n_reads * 2 * B.mates * BwtSearch::LANES, 128, sblocks);
    if (sblocks >= 12) k_search<BwtSearch, 12><<<g, 128, 0, s_search>>>(h->ix, h->P, B);
    else if (sblocks >= 10) k_search<BwtSearch, 10><<<g, 128, 0, s_search>>>(h->ix, h->P, B);
    else k_search<BwtSearch, 8><<<g, 128, 0, s_search>>>(h->ix, h->P, B);
    ++h->launches;
  }
  CUDA_TRY(cudaGetLastError());
  if (h->lanes) {
    CUDA_TRY(cudaEventRecord(b->ev_lane[2], s_search));
    CUDA_TRY(cudaStreamWaitEvent(s_post, b->ev_lane[2], 0));
  }
  const int st = run_pass<Bwt, BwtWide>(h, B, 1, s_post);
  if (st) return st;
  if (h->lanes) {  // the caller's stream continues when the batch is scored
    CUDA_TRY(cudaEventRecord(b->ev_lane[3], s_post));
    CUDA_TRY(cudaStreamWaitEvent(s, b->ev_lane[3], 0));
  }
  return CFR_OK;
}

// after the first pass: re-run select/locate/score for reads that did not fit the arena
template <class Bwt, class BwtWide>
int finish_deferred(cfr_handle *h, cfr_device_batch *b, cudaStream_t s) {
  if (b->n_reads == 0) return CFR_OK;
  ChunkDev B;
  fill_chunk(h, b, B);
  u32 *lists[2] = {(u32 *)b->deferred.p, (u32 *)b->deferred.p + b->n_reads};
  int cur = 0;
  for (int iter = 0; iter < 1 << 20; ++iter) {
    struct {
      u64 used;
      u32 n_def;
      u32 pad;
      u64 other[5];
      u64 exp_used;
    } sc;
    static_assert(sizeof(sc) == 64, "scalars block");
    CUDA_TRY(cudaMemcpyAsync(&sc, b->scalars.p, 64, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (sc.n_def == 0) return CFR_OK;
    B.read_list = lists[cur];
    const bool none_fitted = (u64)sc.n_def == B.n_list;
    B.n_list = sc.n_def;
    cur ^= 1;
    B.deferred = lists[cur];
    if (none_fitted) {
      // nothing fitted: the first read of the list alone exceeds the arena (a hit with a wide range under -k 0 or
      // --hitk-factor 0, a small last batch).  The reference classifies such reads, so the arena grows to the largest
      // slice a deferred read needs; the rows of the reads scored so far are done with.
      unsigned long long *d_max = reinterpret_cast<unsigned long long *>((char *)b->scalars.p + 16);  // the search kernel's task counter: done with
      unsigned long long need = 0;
      CUDA_TRY(cudaMemsetAsync(d_max, 0, 8, s));
      k_max_arena_rows<<<grid_for(h, B.n_list, 256, 4), 256, 0, s>>>(B, d_max);
      ++h->launches;
      CUDA_TRY(cudaMemcpyAsync(&need, d_max, 8, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      if (need <= b->arena_cap) return fail(CFR_ERR_OVERFLOW, "deferral made no progress although every read fits the arena");
      const u64 arena = need + need / 4;
      int st;
      if ((st = b->rows.ensure(arena * 8)) || (st = b->seq_ids.ensure(arena * 4)) || (st = b->rec0.ensure(arena * sizeof(SeqRec))) ||
          (st = b->rec1.ensure(arena * sizeof(SeqRec))) || (st = b->best.ensure(arena * 8)) || (st = b->tmp.ensure(arena * 8)))
        return fail(CFR_ERR_OVERFLOW, "a single read needs more locate rows than the device can hold");
      b->arena_cap = arena;
      const u32 *keep_list = B.read_list;
      u32 *keep_def = B.deferred;
      const u64 keep_n = B.n_list;
      fill_chunk(h, b, B);
      B.read_list = keep_list;
      B.deferred = keep_def;
      B.n_list = keep_n;
    }
    if (h->params.expand_taxid) {
      // the lists of the reads scored so far stay in exp_ids until they are fetched: room for one more arena of them
      if (b->exp_ids.grow_keeping((sc.exp_used + b->arena_cap) * 8, sc.exp_used * 8, s)) return CFR_ERR_NOMEM;
      B.exp_ids = (u64 *)b->exp_ids.p;
      B.exp_cap = b->exp_ids.bytes / 8;
    }
    int st = run_pass<Bwt, BwtWide>(h, B, 0, s);
    if (st) return st;
  }
  return fail(CFR_ERR_OVERFLOW, "deferral loop did not converge");
}

int check_device_errors(cfr_handle *h, cudaStream_t s) {
  u64 flags = 0;
  CUDA_TRY(cudaMemcpyAsync(&flags, &h->d_counters[CFR_STAGE_SCORE].error_flags, 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (flags & 1ull) return fail(CFR_ERR_OVERFLOW, "taxonomy lineage deeper than the device path capacity");
  if (flags & 2ull) return fail(CFR_ERR_OVERFLOW, "expanded tax id lists exceed the batch's list area; raise cfr_params.arena_rows");
  if (flags & 4ull) return fail(CFR_ERR_OVERFLOW, "a read has more best-scoring sequences than cfr_params.unlimited_cap keeps (-k 0)");
  return CFR_OK;
}

// D2H of the --expand-taxid lists of a finished batch
int fetch_expanded(cfr_handle *h, cfr_device_batch *b, uint32_t *exp_cnt, uint64_t *exp_off, uint64_t *exp_ids,
                   uint64_t exp_cap, uint64_t *exp_n, cudaStream_t s) {
  if (!exp_cnt || !exp_off || !exp_n || (exp_cap && !exp_ids)) return fail(CFR_ERR_ARG, "null argument");
  if (!h->params.expand_taxid) return fail(CFR_ERR_ARG, "the handle was opened without cfr_params.expand_taxid");
  *exp_n = 0;
  if (b->n_reads == 0) return CFR_OK;
  u64 used = 0;
  CUDA_TRY(cudaMemcpyAsync(&used, (char *)b->scalars.p + 56, 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  *exp_n = used;
  if (used > exp_cap) return fail(CFR_ERR_OVERFLOW, "exp_ids is too small for this batch (see *exp_n)");
  const u64 k = (u64)h->P.ids_stride;
  CUDA_TRY(cudaMemcpyAsync(exp_cnt, b->exp_cnt.p, b->n_reads * k * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(exp_off, b->exp_off.p, b->n_reads * 8, cudaMemcpyDeviceToHost, s));
  if (used) CUDA_TRY(cudaMemcpyAsync(exp_ids, b->exp_ids.p, used * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  h->d2h_bytes += b->n_reads * (k * 4 + 8) + used * 8 + 8;
  return CFR_OK;
}

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

void cfr_default_params(cfr_params *p) {
  memset(p, 0, sizeof(*p));
  p->max_result = 1;
  p->min_hit_len = 0;
  p->max_result_per_hit_factor = 40;
  p->dust = 1;
  p->consider_secondary_hit_len = 2000;
  p->consider_secondary_score_factor = 0.995;
  p->layout = CFR_LAYOUT_AUTO;
  p->max_batch_reads = 0;
  p->arena_rows = 0;
  p->expand_taxid = 0;
}

const char *cfr_last_error(void) { return g_err.c_str(); }

int cfr_open(const char *idx_prefix, const cfr_params *p, int device, cfr_handle **out) {
  if (!idx_prefix || !out) return fail(CFR_ERR_ARG, "null argument");
  *out = nullptr;
  const auto t_open0 = std::chrono::steady_clock::now();
  cfr_params params;
  if (p) params = *p; else cfr_default_params(&params);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(CFR_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(CFR_ERR_ARG, "device ordinal out of range");
  CUDA_TRY(cudaSetDevice(device));
  cfr_handle *h = new cfr_handle();
  h->device = device;
  h->params = params;
  std::string err;
  int st = h->file.load(idx_prefix, err);
  if (st) {
    delete h;
    return fail(st, err);
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
  auto bail = [&](int code) {
    cfr_close(h);
    return code;
  };
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(fail(CFR_ERR_CUDA, "cudaStreamCreate failed"));
  h->P.max_result = params.max_result;
  // -k <= 0 reports every best-scoring sequence (Classifier.hpp:620-623, :784-785): the id area then keeps
  // unlimited_cap slots per read (a read with more raises CFR_ERR_OVERFLOW, never a silently shortened list)
  h->P.ids_stride = params.max_result > 0 ? params.max_result : (params.unlimited_cap > 0 ? params.unlimited_cap : 64);
  h->P.min_hit_len = params.min_hit_len > 0 ? params.min_hit_len : infer_min_hit_len(h->file.n);
  h->P.hitk_factor = params.max_result_per_hit_factor;
  h->P.secondary_len = params.consider_secondary_hit_len;
  h->P.secondary_factor = params.consider_secondary_score_factor;
  h->P.quorum = 8;
  if (const char *e = getenv("CFR_B200_TRACE")) h->trace = atoi(e) != 0;
  if (const char *e = getenv("CFR_B200_QUORUM")) h->P.quorum = std::max(1, atoi(e));
  bool search_blocks_given = false;
  if (const char *e = getenv("CFR_B200_SEARCH_BLOCKS")) {
    h->search_blocks = std::max(1, atoi(e));
    search_blocks_given = true;
  }
  if (const char *e = getenv("CFR_B200_PAIR_SEARCH_BLOCKS")) h->pair_search_blocks = std::max(1, atoi(e));
  if (const char *e = getenv("CFR_B200_PAIR_FETCH")) h->pair_fetch = std::min(4, std::max(1, atoi(e)));
  if (const char *e = getenv("CFR_B200_LANES")) h->lanes = atoi(e) != 0;
  if (const char *e = getenv("CFR_B200_DUST_SCREEN")) h->dust_screen = atoi(e) != 0;
  if (const char *e = getenv("CFR_B200_DUST_QUORUM")) h->dust_quorum = std::max(0, atoi(e));
  if (const char *e = getenv("CFR_B200_DUST_LANES")) h->dust_lanes = std::min(32, std::max(1, atoi(e)));
  if (const char *e = getenv("CFR_B200_OCC_LOAD")) h->occ_load = atoi(e) == 0 ? 0 : 4;
  h->pos32 = h->file.n < CFR_POS32_MAX_N;
  if (const char *e = getenv("CFR_B200_POS64")) if (atoi(e) != 0) h->pos32 = false;
  if ((st = upload_index(h))) return bail(st);
  void *p2;
  if ((st = dev_alloc(h, &p2, (h->ix.node_cnt + 3) * 8))) return bail(st);
  h->d_taxon = (u64 *)p2;
  if ((st = dev_alloc(h, &p2, sizeof(DevCounters) * CFR_N_STAGES))) return bail(st);
  h->d_counters = (DevCounters *)p2;
  cudaMemset(h->d_taxon, 0, (h->ix.node_cnt + 3) * 8);
  cudaMemset(h->d_counters, 0, sizeof(DevCounters) * CFR_N_STAGES);
  h->layout = params.layout == CFR_LAYOUT_RUNBLOCK ? CFR_LAYOUT_RUNBLOCK : CFR_LAYOUT_OCCLINE;
  if (params.layout == CFR_LAYOUT_AUTO) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const u64 need = (h->ix.n / 64 + 1) * sizeof(OccLine);
    if ((u64)free_b < need + (8ull << 30)) h->layout = CFR_LAYOUT_RUNBLOCK;  // keep 8 GiB for work areas
  }
  if (h->layout == CFR_LAYOUT_OCCLINE && h->ix.n >= (1ull << 40)) {  // 40-bit sector counters
    if (params.layout == CFR_LAYOUT_OCCLINE) return bail(fail(CFR_ERR_UNSUPPORTED, "occ-sector layout needs n < 2^40"));
    h->layout = CFR_LAYOUT_RUNBLOCK;
  }
  if (h->layout == CFR_LAYOUT_OCCLINE) {
    if ((st = build_occ_lines(h))) return bail(st);
    h->occ_bytes = (h->ix.n / 64 + 1) * sizeof(OccLine);
    // sectors beyond L2: the walker is bound by DRAM requests, six resident blocks per SM issue as many as ten do (7.3 ms
    // per 1 M pairs either way on the 20 Gbp index) and leave room for the other stage lanes' kernels (14.3 -> 13.0 ms per batch)
    if (!search_blocks_given && h->occ_bytes > (96ull << 20)) h->search_blocks = 6;
    // the run-block arrays have served their purpose (k_transcode read them with the literal
    // Sequence_RunBlock::Rank / Access); at 140 Gbp they are ~43 GB that the sectors replace.
    // CFR_B200_KEEP_RUNBLOCK=1 keeps them (diagnostics).
    const char *keep = getenv("CFR_B200_KEEP_RUNBLOCK");
    if (!(keep && atoi(keep) != 0)) {
      const size_t before = h->hbm_bytes;
      DevBV *bvs[7] = {&h->ix.block_type, &h->ix.plain.node[0], &h->ix.plain.node[1], &h->ix.plain.node[2],
                       &h->ix.run.node[0], &h->ix.run.node[1], &h->ix.run.node[2]};
      for (DevBV *bv : bvs) {
        dev_release(h, bv->B);
        dev_release(h, bv->R);
        bv->B = bv->R = nullptr;
      }
      h->runblock_bytes = before - h->hbm_bytes;
    }
  }
  {
    // wide lookup table: on by default when the BWT does not fit L2 (there the W-mer table and the
    // first extends are L2 hits and a 1 GB table would only add DRAM misses); CFR_B200_WIDE_LOOKUP=WW
    // forces a width (0 = off)
    int ww = 0;
    const u64 bwt_bytes = h->layout == CFR_LAYOUT_OCCLINE ? (h->ix.n / 64 + 1) * sizeof(OccLine) : h->ix.n / 2;
    if (bwt_bytes > (96ull << 20) && h->ix.n < (1ull << 56)) {
      // a WW-mer of a random read occurs about n / 4^WW times: two bases short of log4(n) leaves
      // ranges of a few rows, which share a sector (measured on the 700 Mbp index: 12 / 13 / 14 ->
      // 1.98 / 1.85 / 1.73 ms per 1 M reads, 2.36 without the table)
      int lg = 0;
      while (lg < 32 && (1ull << (2 * lg)) < h->ix.n) ++lg;
      ww = std::min(15, std::max(12, lg - 2));
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      while (ww > h->ix.pre_width && ((16ull << (2 * ww)) + (8ull << 30)) > (u64)free_b) --ww;
    }
    if (const char *e = getenv("CFR_B200_WIDE_LOOKUP")) ww = atoi(e);
    if ((st = build_wide_lookup(h, ww))) return bail(st);
  }
  {
    // pair lines for the search kernel: when the occ sectors do not fit L2 (every rank is a DRAM line fill)
    // and 2 bytes per row fit next to everything else with 24 GiB to spare; CFR_B200_PAIRS=1 / 0 forces / forbids
    bool want = false;
    if (h->layout == CFR_LAYOUT_OCCLINE) {
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      const u64 need = (h->ix.n / 64 + 1) * sizeof(PairLine);
      want = h->occ_bytes > (96ull << 20) && (u64)free_b > need + (24ull << 30);
      if (const char *e = getenv("CFR_B200_PAIRS")) want = atoi(e) != 0;
    }
    if (want && (st = build_pair_lines(h))) return bail(st);
  }
  {
    // dense locate table: the densest spacing whose table (2 or 4 bytes per entry) stays below half of the
    // free HBM and 48 GB -- every row up to 20 Gbp, every 8th row at 140 Gbp; CFR_B200_DENSE_LOCATE=shift
    // forces a spacing (-1 = off).  Measured on configs[1]: every 8th / 4th / 2nd row -> 0.39 / 0.33 /
    // 0.24 ms per 1 M reads (0.49 with the stored samples only, every 16th row).
    int shift = 0;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const u64 budget = std::min<u64>((u64)free_b / 2, 48ull << 30);
    // 16-bit entries when every id a locate can return fits: the sampled SA's ids (sa_bits wide), the boundary
    // table's, and the id of row firstISA (CFR_B200_DENSE16=0 keeps 32-bit entries)
    bool e16 = h->file.sa_bits <= 16 && h->file.adjusted_sa0 < 65536;
    for (u64 i = 0; e16 && i < h->file.sel_cnt; ++i) e16 = load_u64(h->file.sel + i * 16 + 8) < 65536;
    if (const char *e = getenv("CFR_B200_DENSE16")) e16 = e16 && atoi(e) != 0;
    while (shift < 8 && ((h->ix.n >> shift) + 1) * (e16 ? 2 : 4) > budget) ++shift;
    if (const char *e = getenv("CFR_B200_DENSE_LOCATE")) shift = atoi(e);
    if ((st = build_dense_locate(h, shift, e16))) return bail(st);
  }
  h->file.map1.close();  // everything needed from .1.cfr now lives in HBM
  h->open_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_open0).count();
  *out = h;
  return CFR_OK;
}

void cfr_close(cfr_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (int i = 0; i < cfr_handle::NSLOT; ++i) h->slots[i].release();
  for (cudaStream_t &l : h->lane)
    if (l) cudaStreamDestroy(l);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  for (int i = 0; i < cfr_handle::NSLOT; ++i)
    if (h->s_comp[i]) cudaStreamDestroy(h->s_comp[i]);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  for (int i = 0; i < cfr_handle::NSLOT; ++i) {
    if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
    if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
    if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
  }
  if (h->pinned_scalars) cudaFreeHost(h->pinned_scalars);
  for (void *p : h->index_allocs) cudaFree(p);
  for (auto &e : h->ev_pending) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

void *cfr_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    fail(CFR_ERR_NOMEM, "cudaHostAlloc failed");
    return nullptr;
  }
  return p;
}

void cfr_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

uint64_t cfr_index_info(const cfr_handle *h, int which) {
  if (!h) return 0;
  switch (which) {
    case 0: return h->ix.n;
    case 1: return h->ix.b;
    case 2: return h->ix.block_cnt;
    case 3: return h->ix.first_isa;
    case 4: return (uint64_t)h->P.min_hit_len;
    case 5: return h->ix.node_cnt;
    case 6: return h->file.tax.seq_cnt + h->file.tax.extra_seq_cnt;
    case 7: return h->ix.root;
    case 8: return (uint64_t)h->layout;
    case 9: return (uint64_t)h->hbm_bytes;
    case 10: return (uint64_t)h->ix.sample_rate;
    case 11: return (uint64_t)h->ix.pre_width;
    case 12: return (uint64_t)h->P.max_result;
    case 13: return (uint64_t)h->h2d_bytes;
    case 14: return (uint64_t)h->d2h_bytes;
    case 15: return (uint64_t)h->occ_bytes;
    case 16: return (uint64_t)h->wide_bytes;
    case 17: return (uint64_t)h->dense_bytes;
    case 18: return (uint64_t)(h->open_seconds * 1e6);
    case 19: return (uint64_t)h->runblock_bytes;
    case 20: return (uint64_t)(h->ix.dense_shift < 0 ? 255 : h->ix.dense_shift);
    case 21: return (uint64_t)h->ix.wide_width;
    case 22: return h->pos32 ? 32u : 64u;
    case 23: return (uint64_t)h->pair_bytes;
    case 24: return (uint64_t)h->P.ids_stride;
    default: return 0;
  }
}

int cfr_batch_upload(cfr_handle *h, const cfr_read_batch *in, void *stream, cfr_device_batch **out) {
  if (!h || !in || !out) return fail(CFR_ERR_ARG, "null argument");
  if (in->n_reads && (!in->seq1 || !in->off1)) return fail(CFR_ERR_ARG, "seq1/off1 missing");
  CUDA_TRY(cudaSetDevice(h->device));
  cfr_device_batch *b = new cfr_device_batch();
  int st = upload_chunk(h, in, 0, in->n_reads, b, pick_stream(h, stream));
  if (st) {
    b->release();
    delete b;
    return st;
  }
  CUDA_TRY(cudaStreamSynchronize(pick_stream(h, stream)));
  *out = b;
  return CFR_OK;
}

int cfr_classify_resident(cfr_handle *h, cfr_device_batch *b, void *stream) {
  if (!h || !b) return fail(CFR_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = pick_stream(h, stream);
  h->host_bases += b->total_bases;
  b->classified = true;
  b->quant_counted = false;
  if (h->layout == CFR_LAYOUT_OCCLINE) {
    if (h->ix.pairs) {  // pair lines in the search kernel (CFR_B200_PAIR_FETCH picks how they are awaited)
      if (h->pair_fetch == 1) {
        if (h->pos32) return run_first<BwtOccLine, BwtOccLine32T<4>, BwtPairT<1>>(h, b, s);
        return run_first<BwtOccLine, BwtOccLineT<4>, BwtPairT<1>>(h, b, s);
      }
      if (h->pair_fetch == 2) {
        if (h->pos32) return run_first<BwtOccLine, BwtOccLine32T<4>, BwtPairT<2>>(h, b, s);
        return run_first<BwtOccLine, BwtOccLineT<4>, BwtPairT<2>>(h, b, s);
      }
      if (h->pair_fetch == 4) {  // bulk (TMA) copies, one per lane, counted on an mbarrier per warp: measured, not the default
        if (h->pos32) return run_first<BwtOccLine, BwtOccLine32T<4>, BwtPairT<4>>(h, b, s);
        return run_first<BwtOccLine, BwtOccLineT<4>, BwtPairT<4>>(h, b, s);
      }
      if (h->pos32) return run_first<BwtOccLine, BwtOccLine32T<4>, BwtPairT<3>>(h, b, s);
      return run_first<BwtOccLine, BwtOccLineT<4>, BwtPairT<3>>(h, b, s);
    }
    if (h->pos32) {
      if (h->occ_load == 0) return run_first<BwtOccLine, BwtOccLine32T<0>>(h, b, s);
      return run_first<BwtOccLine, BwtOccLine32T<4>>(h, b, s);
    }
    if (h->occ_load == 0) return run_first<BwtOccLine, BwtOccLine>(h, b, s);
    return run_first<BwtOccLine, BwtOccLineT<4>>(h, b, s);
  }
  return run_first<BwtRunBlock, BwtRunBlock>(h, b, s);
}

static int quant_coalesce_batch(cfr_handle *h, cfr_device_batch *b, cudaStream_t s);

int cfr_batch_fetch(cfr_handle *h, cfr_device_batch *b, cfr_result *results, uint64_t *ids, void *stream) {
  if (!h || !b || !results || !ids) return fail(CFR_ERR_ARG, "null argument");
  if (!b->classified) return fail(CFR_ERR_ARG, "batch was not classified");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = pick_stream(h, stream);
  int st;
  if (h->layout == CFR_LAYOUT_OCCLINE)
    st = finish_deferred<BwtOccLine, BwtOccLine>(h, b, s);
  else
    st = finish_deferred<BwtRunBlock, BwtRunBlock>(h, b, s);
  if (st) return st;
  static_assert(sizeof(cfr_result) == sizeof(DevResult), "result layout");
  if (b->n_reads) {
    CUDA_TRY(cudaMemcpyAsync(results, b->results.p, b->n_reads * sizeof(DevResult), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(ids, b->out_ids.p, b->n_reads * (u64)h->P.ids_stride * 8, cudaMemcpyDeviceToHost, s));
    h->d2h_bytes += b->n_reads * (sizeof(DevResult) + (u64)h->P.ids_stride * 8);
  }
  if ((st = check_device_errors(h, s))) return st;
  if (b->quant_counted) return CFR_OK;  // a resident batch counts once per classification, however often it is fetched
  b->quant_counted = true;
  return quant_coalesce_batch(h, b, s);
}

int cfr_batch_fetch_expanded(cfr_handle *h, cfr_device_batch *b, uint32_t *exp_cnt, uint64_t *exp_off,
                             uint64_t *exp_ids, uint64_t exp_cap, uint64_t *exp_n, void *stream) {
  if (!h || !b) return fail(CFR_ERR_ARG, "null argument");
  if (!b->classified) return fail(CFR_ERR_ARG, "batch was not classified");
  CUDA_TRY(cudaSetDevice(h->device));
  return fetch_expanded(h, b, exp_cnt, exp_off, exp_ids, exp_cap, exp_n, pick_stream(h, stream));
}

void cfr_batch_free(cfr_handle *h, cfr_device_batch *b) {
  if (!b) return;
  if (h) cudaSetDevice(h->device);
  b->release();
  delete b;
}

// Host <-> device copies overlap the kernels: the batch is cut into chunks that go
// through two slots; chunk c's H2D (stream s_in), kernels (one compute stream per
// slot, so the head of chunk c+1 fills the SMs the tail of chunk c leaves idle) and
// D2H (stream s_out) are chained with events.  Everything is ordered after the
// caller's stream, and the call returns when the results are on the host.  Pinned
// host buffers are needed for real copy/compute overlap.
static int pipeline_init(cfr_handle *h) {
  if (h->s_in) return CFR_OK;
  CUDA_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < cfr_handle::NSLOT; ++i) CUDA_TRY(cudaStreamCreateWithFlags(&h->s_comp[i], cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
  for (int i = 0; i < cfr_handle::NSLOT; ++i) {
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaHostAlloc((void **)&h->pinned_scalars, cfr_handle::NSLOT * sizeof(*h->pinned_scalars), cudaHostAllocDefault));
  return CFR_OK;
}

// Quantification, per finished batch: the reads' assignment records (k_quant_keys) are sorted by LSD radix passes
// over pairs of words, equal neighbours are counted, and only the distinct records with their counts cross to
// the host -- CoalesceAssignments (Quantifier.hpp:490-513) where the reads are.
static int quant_coalesce_batch(cfr_handle *h, cfr_device_batch *b, cudaStream_t s) {
  cfr_handle::Quant &q = h->quant;
  const u64 n = b->n_reads;
  if (!q.on || n == 0) return CFR_OK;
  if (n >= (1ull << 31)) return fail(CFR_ERR_UNSUPPORTED, "quantification: batch too large");
  const int K = h->P.ids_stride, W = K + 1;
  int st;
  if ((st = q.words.ensure(n * W * 4)) || (st = q.keys_a.ensure(n * 8)) || (st = q.keys_b.ensure(n * 8)) ||
      (st = q.idx_a.ensure(n * 4)) || (st = q.idx_b.ensure(n * 4)) || (st = q.flags.ensure(n)) || (st = q.heads.ensure(n * 4)) ||
      (st = q.num.ensure(16)) || (st = q.out.ensure(n * (W + 1) * 4)))
    return st;
  u32 *words = (u32 *)q.words.p;
  const int grid = grid_for(h, n, 256, 8);
  k_quant_keys<<<grid, 256, 0, s>>>(h->ix, (const DevResult *)b->results.p, (const u64 *)b->out_ids.p, n, K, q.min_score, q.min_hit, words);
  cub::DoubleBuffer<u64> dk((u64 *)q.keys_a.p, (u64 *)q.keys_b.p);
  cub::DoubleBuffer<u32> di((u32 *)q.idx_a.p, (u32 *)q.idx_b.p);
  size_t tb = 0, tb2 = 0;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, di, (int)n, 0, 64, s));
  CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, tb2, cub::CountingInputIterator<u32>(0), (unsigned char *)q.flags.p, (u32 *)q.heads.p,
                                      (u32 *)q.num.p, (int)n, s));
  if ((st = q.tmp.ensure(std::max(tb, tb2)))) return st;
  const int passes = (W + 1) / 2;
  for (int p = 0; p < passes; ++p) {
    k_quant_gather<<<grid, 256, 0, s>>>(words, p == 0 ? nullptr : di.Current(), n, W, p, dk.Current(), p == 0 ? di.Current() : nullptr);
    size_t t = q.tmp.bytes;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(q.tmp.p, t, dk, di, (int)n, 0, 64, s));
  }
  k_quant_heads<<<grid, 256, 0, s>>>(words, di.Current(), n, W, (unsigned char *)q.flags.p);
  {
    size_t t = q.tmp.bytes;
    CUDA_TRY(cub::DeviceSelect::Flagged(q.tmp.p, t, cub::CountingInputIterator<u32>(0), (unsigned char *)q.flags.p, (u32 *)q.heads.p,
                                        (u32 *)q.num.p, (int)n, s));
  }
  k_quant_emit<<<grid, 256, 0, s>>>(words, di.Current(), (const u32 *)q.heads.p, (const u32 *)q.num.p, n, W, (u32 *)q.out.p);
  h->launches += 3 + 2 * passes + 1;
  CUDA_TRY(cudaGetLastError());
  u32 num = 0;
  CUDA_TRY(cudaMemcpyAsync(&num, q.num.p, 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  std::vector<u32> host((size_t)num * (W + 1));
  if (num) {
    CUDA_TRY(cudaMemcpyAsync(host.data(), q.out.p, host.size() * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
  }
  for (u32 e = 0; e < num; ++e) {
    const u32 *o = host.data() + (size_t)e * (W + 1);
    q.table[std::vector<u32>(o, o + W)] += o[W];
  }
  ++q.batches;
  q.entries_moved += num;
  return CFR_OK;
}

// wait for a chunk's results; run the (rare) follow-up passes for reads that did not fit the arena
static int pipeline_drain(cfr_handle *h, int slot, cfr_result *results, uint64_t *ids, cudaStream_t sc) {
  cfr_device_batch *b = &h->slots[slot];
  CUDA_TRY(cudaEventSynchronize(h->ev_d2h[slot]));
  if (h->pinned_scalars[slot].n_def != 0) {
    int st = h->layout == CFR_LAYOUT_OCCLINE ? finish_deferred<BwtOccLine, BwtOccLine>(h, b, sc)
                                             : finish_deferred<BwtRunBlock, BwtRunBlock>(h, b, sc);
    if (st) return st;
    CUDA_TRY(cudaMemcpyAsync(results, b->results.p, b->n_reads * sizeof(DevResult), cudaMemcpyDeviceToHost, sc));
    CUDA_TRY(cudaMemcpyAsync(ids, b->out_ids.p, b->n_reads * (u64)h->P.ids_stride * 8, cudaMemcpyDeviceToHost, sc));
    CUDA_TRY(cudaStreamSynchronize(sc));
  }
  return quant_coalesce_batch(h, b, sc);
}

static int job_finish(cfr_handle *h, int slot) {
  cfr_handle::Job &j = h->jobs[slot];
  if (j.ticket < 0) return CFR_OK;
  int st = pipeline_drain(h, slot, j.results, j.ids, h->s_comp[slot]);
  j.ticket = -1;
  if (st) return st;
  return check_device_errors(h, h->s_comp[slot]);
}

int cfr_classify_batch(cfr_handle *h, const cfr_read_batch *in, cfr_result *results, uint64_t *ids, void *stream) {
  if (!h || !in || (in->n_reads && (!results || !ids))) return fail(CFR_ERR_ARG, "null argument");
  if (in->n_reads && (!in->seq1 || !in->off1)) return fail(CFR_ERR_ARG, "seq1/off1 missing");
  if (in->n_reads == 0) return CFR_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t sc = pick_stream(h, stream);
  int st = pipeline_init(h);
  if (st) return st;
  for (int i = 0; i < cfr_handle::NSLOT; ++i)
    if ((st = job_finish(h, i))) return st;  // complete streaming jobs first
  const u64 cap = h->params.max_batch_reads > 0 ? (u64)h->params.max_batch_reads : (1ull << 20);
  // Chunk plan.  Every chunk pays a fixed cost (the critical path of its slowest read in each
  // kernel), so few chunks are better for the GPU; but the first upload and the last download are
  // not overlapped with anything, so those two chunks are kept small: n/8, 3n/8, 3n/8, n/8.
  std::vector<u64> bounds;  // chunk start offsets + n_reads
  {
    const u64 n = in->n_reads;
    int plan = 1;  // 1 = tapered, N>1 = N equal chunks (CFR_B200_PIPE_CHUNKS)
    if (const char *e = getenv("CFR_B200_PIPE_CHUNKS")) plan = std::max(1, atoi(e));
    bounds.push_back(0);
    if (plan == 1 && n >= (1u << 18) && n <= 4 * cap) {
      bounds.push_back(n / 8);
      bounds.push_back(n / 2);
      bounds.push_back(n - n / 8);
    } else {
      u64 chunk = plan > 1 ? (n + plan - 1) / plan : cap;
      chunk = std::min<u64>(cap, std::max<u64>(chunk, 1));
      for (u64 r = chunk; r < n; r += chunk) bounds.push_back(r);
    }
    bounds.push_back(n);
    // respect the device chunk cap
    std::vector<u64> capped;
    for (size_t i = 0; i + 1 < bounds.size(); ++i)
      for (u64 r = bounds[i]; r < bounds[i + 1]; r += cap) capped.push_back(r);
    capped.push_back(n);
    bounds.swap(capped);
  }
  const u64 k = (u64)h->P.ids_stride;
  CUDA_TRY(cudaEventRecord(h->ev_start, sc));  // everything below is ordered after the caller's stream
  CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->ev_start, 0));
  CUDA_TRY(cudaStreamWaitEvent(h->s_comp[0], h->ev_start, 0));
  CUDA_TRY(cudaStreamWaitEvent(h->s_comp[1], h->ev_start, 0));
  u64 starts[2] = {0, 0};
  u64 c = 0;
  for (; c + 1 < bounds.size(); ++c) {
    const int slot = (int)(c & 1);
    const u64 r0 = bounds[c], r1 = bounds[c + 1];
    cfr_device_batch *b = &h->slots[slot];
    if (c >= 2) {  // the slot's previous chunk must have left the device before its buffers are reused
      if ((st = pipeline_drain(h, slot, results + starts[slot], ids + starts[slot] * k, h->s_comp[slot]))) return st;
    }
    starts[slot] = r0;
    b->want_masked = false;
    b->ticket = -1;  // the slot no longer holds a streaming batch
    if ((st = upload_chunk(h, in, r0, r1, b, h->s_in))) return st;
    CUDA_TRY(cudaEventRecord(h->ev_h2d[slot], h->s_in));
    CUDA_TRY(cudaStreamWaitEvent(h->s_comp[slot], h->ev_h2d[slot], 0));
    if ((st = cfr_classify_resident(h, b, h->s_comp[slot]))) return st;
    CUDA_TRY(cudaEventRecord(h->ev_comp[slot], h->s_comp[slot]));
    CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_comp[slot], 0));
    CUDA_TRY(cudaMemcpyAsync(results + r0, b->results.p, (r1 - r0) * sizeof(DevResult), cudaMemcpyDeviceToHost, h->s_out));
    CUDA_TRY(cudaMemcpyAsync(ids + r0 * k, b->out_ids.p, (r1 - r0) * k * 8, cudaMemcpyDeviceToHost, h->s_out));
    h->d2h_bytes += (r1 - r0) * (sizeof(DevResult) + k * 8);
    CUDA_TRY(cudaMemcpyAsync(&h->pinned_scalars[slot], b->scalars.p, 16, cudaMemcpyDeviceToHost, h->s_out));
    CUDA_TRY(cudaEventRecord(h->ev_d2h[slot], h->s_out));
    // the next upload into the OTHER slot may start at once; an upload into THIS slot (chunk c+2)
    // is issued only after pipeline_drain() above has seen ev_d2h[slot]
  }
  // drain the last (up to two) chunks in order
  for (u64 d = c >= 2 ? c - 2 : 0; d < c; ++d) {
    const int slot = (int)(d & 1);
    if ((st = pipeline_drain(h, slot, results + starts[slot], ids + starts[slot] * k, h->s_comp[slot]))) return st;
  }
  CUDA_TRY(cudaStreamSynchronize(h->s_comp[0]));
  CUDA_TRY(cudaStreamSynchronize(h->s_comp[1]));
  return check_device_errors(h, sc);
}

// ---- host side of the packed form: bytes -> 2-bit codes + N bits, exactly what k_encode (encode_stage) writes
uint64_t cfr_packed_words(const cfr_read_batch *in) {
  if (!in || !in->off1) return 2;
  const u64 n = in->n_reads;
  const u64 len1 = in->off1[n] - in->off1[0];
  const u64 len2 = (in->seq2 && in->off2) ? in->off2[n] - in->off2[0] : 0;
  return (((len1 + 31) & ~31ull) + len2) / 32 + 2;
}

int cfr_pack_reads(const cfr_read_batch *in, uint64_t *codes, uint32_t *nmask, uint64_t *off1, uint64_t *off2, int threads,
                   cfr_packed_batch *out) {
  if (!in || !codes || !nmask || !off1 || !out || !in->off1 || (in->n_reads && !in->seq1)) return fail(CFR_ERR_ARG, "null argument");
  const bool two = in->seq2 != nullptr;
  if (two && (!in->off2 || !off2)) return fail(CFR_ERR_ARG, "seq2 given without off2");
  const u64 n = in->n_reads;
  const u64 s1 = in->off1[0], len1 = in->off1[n] - s1;
  const u64 s2 = two ? in->off2[0] : 0, len2 = two ? in->off2[n] - s2 : 0;
  const u64 pos2 = (len1 + 31) & ~31ull;
  const u64 n_words = (pos2 + len2) / 32 + 2;
  static const struct Lut {
    unsigned char v[256];  // 0..3 = code, 4 = not one of "ACGT"
    Lut() {
      memset(v, 4, sizeof(v));
      v[(unsigned char)'A'] = 0;
      v[(unsigned char)'C'] = 1;
      v[(unsigned char)'G'] = 2;
      v[(unsigned char)'T'] = 3;
    }
  } lut;
  const unsigned char *a1 = reinterpret_cast<const unsigned char *>(in->seq1) + s1;
  const unsigned char *a2 = two ? reinterpret_cast<const unsigned char *>(in->seq2) + s2 : nullptr;
  auto words = [&](u64 w0, u64 w1) {
    for (u64 w = w0; w < w1; ++w) {
      const u64 p0 = w * 32;
      u64 c = 0;
      u32 m = 0;
      // pos2 is a multiple of 32, so a word lies in one mate's region (or in the padding)
      const unsigned char *src = nullptr;
      int cnt = 0;
      if (p0 < len1) {
        src = a1 + p0;
        cnt = (int)std::min<u64>(32, len1 - p0);
      } else if (p0 >= pos2 && p0 - pos2 < len2) {
        src = a2 + (p0 - pos2);
        cnt = (int)std::min<u64>(32, len2 - (p0 - pos2));
      }
      m = cnt >= 32 ? 0u : ~((1u << cnt) - 1u);  // padding reads as N
      int i = 0;
      for (; i + 8 <= cnt; i += 8) {  // eight bases per step, branch-free (SWAR)
        u64 x;
        memcpy(&x, src + i, 8);
        const u64 ONES = 0x0101010101010101ull, LOW7 = 0x7f7f7f7f7f7f7f7full, HIGH = 0x8080808080808080ull;
        // bytes equal to 'A', 'C', 'G' or 'T': exact per-byte zero test of x ^ letter
        u64 ok = 0;
        for (const u64 letter : {0x41ull, 0x43ull, 0x47ull, 0x54ull}) {
          const u64 z = x ^ (letter * ONES);
          ok |= ~(((z & LOW7) + LOW7) | z) & HIGH;
        }
        const u32 valid8 = (u32)((((ok >> 7) & ONES) * 0x0102040810204080ull) >> 56);  // byte j -> bit j (terms 8j + 56 - 7j, no two collide)
        // 2-bit codes: ((b >> 1) ^ (b >> 2)) & 3 maps A, C, G, T to 0, 1, 2, 3; bytes that are not ACGT give 0
        u64 q = ((x >> 1) ^ (x >> 2)) & (3ull * ONES) & ((ok >> 7) * 3ull);
        q = (q | (q >> 6)) & 0x000f000f000f000full;
        q = (q | (q >> 12)) & 0x000000ff000000ffull;
        q = (q | (q >> 24)) & 0xffffull;
        c |= q << (2 * i);
        m |= (~valid8 & 0xffu) << i;
      }
      for (; i < cnt; ++i) {
        const unsigned char xb = lut.v[src[i]];
        if (xb > 3) m |= 1u << i;
        else c |= (u64)xb << (2 * i);
      }
      codes[w] = c;
      nmask[w] = m;
    }
  };
  auto offsets = [&](u64 i0, u64 i1) {
    for (u64 i = i0; i < i1; ++i) {
      off1[i] = in->off1[i] - s1;
      if (two) off2[i] = in->off2[i] - s2 + pos2;
    }
  };
  const unsigned T = (unsigned)std::max(1, std::min(threads, 256));
  if (T == 1 || n_words < 4096) {
    words(0, n_words);
    offsets(0, n + 1);
  } else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; ++t)
      pool.emplace_back([&, t] {
        words(n_words * t / T, n_words * (t + 1) / T);
        offsets((n + 1) * t / T, (n + 1) * (t + 1) / T);
      });
    for (auto &th : pool) th.join();
  }
  out->n_reads = n;
  out->codes = codes;
  out->nmask = nmask;
  out->n_words = n_words;
  out->off1 = off1;
  out->off2 = two ? off2 : nullptr;
  return CFR_OK;
}

int cfr_submit_batch(cfr_handle *h, const cfr_read_batch *in, cfr_result *results, uint64_t *ids, void *stream,
                     int *ticket) {
  return cfr_submit_batch_masked(h, in, results, ids, nullptr, nullptr, stream, ticket);
}

static int submit_impl(cfr_handle *h, const cfr_read_batch *in, const cfr_packed_batch *pk, cfr_result *results, uint64_t *ids,
                       char *masked1, char *masked2, void *stream, int *ticket);

int cfr_submit_batch_masked(cfr_handle *h, const cfr_read_batch *in, cfr_result *results, uint64_t *ids,
                            char *masked1, char *masked2, void *stream, int *ticket) {
  if (!h || !in || !ticket || (in->n_reads && (!results || !ids))) return fail(CFR_ERR_ARG, "null argument");
  if (in->n_reads && (!in->seq1 || !in->off1)) return fail(CFR_ERR_ARG, "seq1/off1 missing");
  return submit_impl(h, in, nullptr, results, ids, masked1, masked2, stream, ticket);
}

int cfr_submit_packed(cfr_handle *h, const cfr_packed_batch *pk, cfr_result *results, uint64_t *ids, void *stream, int *ticket) {
  if (!h || !pk || !ticket || (pk->n_reads && (!results || !ids))) return fail(CFR_ERR_ARG, "null argument");
  if (pk->n_reads && (!pk->codes || !pk->nmask || !pk->off1)) return fail(CFR_ERR_ARG, "codes / nmask / off1 missing");
  cfr_read_batch view;  // the offsets only: the bases travel packed
  view.n_reads = pk->n_reads;
  view.seq1 = view.seq2 = nullptr;
  view.off1 = pk->off1;
  view.off2 = pk->off2;
  return submit_impl(h, &view, pk, results, ids, nullptr, nullptr, stream, ticket);
}

static int submit_impl(cfr_handle *h, const cfr_read_batch *in, const cfr_packed_batch *pk, cfr_result *results, uint64_t *ids,
                       char *masked1, char *masked2, void *stream, int *ticket) {
  const u64 cap = h->params.max_batch_reads > 0 ? (u64)h->params.max_batch_reads : (1ull << 20);
  if (in->n_reads > cap) return fail(CFR_ERR_ARG, "cfr_submit_batch: n_reads exceeds max_batch_reads");
  CUDA_TRY(cudaSetDevice(h->device));
  int st = pipeline_init(h);
  if (st) return st;
  const int slot = h->next_ticket % cfr_handle::NSLOT;
  if ((st = job_finish(h, slot))) return st;  // at most NSLOT batches in flight
  cfr_device_batch *b = &h->slots[slot];
  b->want_masked = masked1 != nullptr;
  cudaStream_t sc = pick_stream(h, stream);
  CUDA_TRY(cudaEventRecord(h->ev_start, sc));  // ordered after the caller's stream
  CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->ev_start, 0));
  CUDA_TRY(cudaStreamWaitEvent(h->s_comp[slot], h->ev_start, 0));
  const auto host_t0 = std::chrono::steady_clock::now();
  if (h->trace) {
    if (!h->tr_base) {
      cudaEventCreate(&h->tr_base);
      for (int a = 0; a < cfr_handle::NSLOT; ++a)
        for (int q = 0; q < 4; ++q) cudaEventCreate(&h->tr_ev[a][q]);
      cudaEventRecord(h->tr_base, h->s_in);
    }
    cudaEventRecord(h->tr_ev[slot][0], h->s_in);
  }
  if ((st = upload_chunk(h, in, 0, in->n_reads, b, h->s_in, pk))) return st;
  CUDA_TRY(cudaEventRecord(h->ev_h2d[slot], h->s_in));
  if (h->trace) cudaEventRecord(h->tr_ev[slot][1], h->s_in);
  CUDA_TRY(cudaStreamWaitEvent(h->s_comp[slot], h->ev_h2d[slot], 0));
  if ((st = cfr_classify_resident(h, b, h->s_comp[slot]))) return st;
  CUDA_TRY(cudaEventRecord(h->ev_comp[slot], h->s_comp[slot]));
  if (h->trace) {
    cudaEventRecord(h->tr_ev[slot][2], h->s_comp[slot]);
    h->tr_host_submit[slot] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
  }
  CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_comp[slot], 0));
  const u64 k = (u64)h->P.ids_stride;
  if (in->n_reads) {
    CUDA_TRY(cudaMemcpyAsync(results, b->results.p, in->n_reads * sizeof(DevResult), cudaMemcpyDeviceToHost, h->s_out));
    CUDA_TRY(cudaMemcpyAsync(ids, b->out_ids.p, in->n_reads * k * 8, cudaMemcpyDeviceToHost, h->s_out));
    h->d2h_bytes += in->n_reads * (sizeof(DevResult) + k * 8);
  }
  if (b->want_masked && in->n_reads) {  // (never with a packed batch: masked1 is null there)
    const u64 len1 = in->off1[in->n_reads] - in->off1[0];
    const u64 len2 = in->seq2 ? in->off2[in->n_reads] - in->off2[0] : 0;
    h->d2h_bytes += len1 + (masked2 ? len2 : 0);
    if (len1) CUDA_TRY(cudaMemcpyAsync(masked1, b->masked.p, len1, cudaMemcpyDeviceToHost, h->s_out));
    if (len2 && masked2)
      CUDA_TRY(cudaMemcpyAsync(masked2, (char *)b->masked.p + ((len1 + 31) & ~31ull), len2, cudaMemcpyDeviceToHost, h->s_out));
  }
  CUDA_TRY(cudaMemcpyAsync(&h->pinned_scalars[slot], b->scalars.p, 16, cudaMemcpyDeviceToHost, h->s_out));
  CUDA_TRY(cudaEventRecord(h->ev_d2h[slot], h->s_out));
  if (h->trace) cudaEventRecord(h->tr_ev[slot][3], h->s_out);
  b->ticket = h->next_ticket;
  h->jobs[slot].ticket = h->next_ticket;
  h->jobs[slot].results = results;
  h->jobs[slot].ids = ids;
  *ticket = h->next_ticket++;
  return CFR_OK;
}

int cfr_wait_batch(cfr_handle *h, int ticket) {
  if (!h || ticket < 0) return fail(CFR_ERR_ARG, "bad ticket");
  CUDA_TRY(cudaSetDevice(h->device));
  const int slot = ticket % cfr_handle::NSLOT;
  if (h->jobs[slot].ticket != ticket) {
    if (ticket < h->next_ticket) return CFR_OK;  // already completed (by a later submit)
    return fail(CFR_ERR_ARG, "unknown ticket");
  }
  const int rc = job_finish(h, slot);
  if (h->trace && rc == CFR_OK) {
    float t[4] = {0, 0, 0, 0};
    for (int q = 0; q < 4; ++q) cudaEventElapsedTime(&t[q], h->tr_base, h->tr_ev[slot][q]);
    fprintf(stderr, "[cfr trace] ticket %d: h2d %.2f..%.2f ms, kernels end %.2f, d2h end %.2f (host submit %.2f ms)\n", ticket,
            t[0], t[1], t[2], t[3], h->tr_host_submit[slot]);
  }
  return rc;
}

int cfr_fetch_expanded(cfr_handle *h, int ticket, uint32_t *exp_cnt, uint64_t *exp_off, uint64_t *exp_ids,
                       uint64_t exp_cap, uint64_t *exp_n) {
  if (!h || ticket < 0) return fail(CFR_ERR_ARG, "bad ticket");
  CUDA_TRY(cudaSetDevice(h->device));
  const int slot = ticket % cfr_handle::NSLOT;
  if (h->slots[slot].ticket != ticket) return fail(CFR_ERR_ARG, "the batch of this ticket has left the device (fetch before submitting three more)");
  if (h->jobs[slot].ticket == ticket) {  // not waited for yet
    const int st = job_finish(h, slot);
    if (st) return st;
  }
  return fetch_expanded(h, &h->slots[slot], exp_cnt, exp_off, exp_ids, exp_cap, exp_n, h->s_comp[slot]);
}

const char *cfr_seq_name(const cfr_handle *h, uint64_t seq_id) {
  if (!h || seq_id >= h->file.tax.seq_name.size()) return "";
  return h->file.tax.seq_name[seq_id].c_str();
}

const char *cfr_rank_name(const cfr_handle *h, uint64_t ctid) {
  if (!h) return "";
  const uint8_t r = ctid < h->file.tax.node_cnt ? h->file.tax.rank[ctid] : 0;  // GetTaxIdRank: unknown -> RANK_UNKNOWN
  return tax_rank_string(r);
}

uint64_t cfr_orig_taxid(const cfr_handle *h, uint64_t ctid) {
  if (!h) return 0;
  const TaxonomyHost &t = h->file.tax;
  if (ctid >= t.node_cnt) ctid = t.root;
  return ctid < t.orig_taxid.size() ? t.orig_taxid[ctid] : 0;
}

uint64_t cfr_seq_taxid(const cfr_handle *h, uint64_t seq_id) {
  if (!h) return 0;
  const TaxonomyHost &t = h->file.tax;
  return seq_id < t.seq_cnt ? t.seq_to_tax[seq_id] : t.node_cnt;
}

int cfr_format_tsv(const cfr_handle *h, const char *read_id, const cfr_result *r, const uint64_t *ids, char *buf,
                   size_t cap) {
  if (!h || !read_id || !r || !buf) return CFR_ERR_ARG;
  size_t off = 0;
  if (r->n_assign > 0) {
    const int m = std::min<int>(r->n_assign, h->P.ids_stride);
    for (int i = 0; i < m; ++i) {
      const char *name = r->by_rank ? cfr_rank_name(h, ids[i]) : cfr_seq_name(h, ids[i]);
      const uint64_t tax = r->by_rank ? cfr_orig_taxid(h, ids[i]) : cfr_orig_taxid(h, cfr_seq_taxid(h, ids[i]));
      int w = snprintf(buf + off, cap - off, "%s\t%s\t%lu\t%lu\t%lu\t%d\t%d\t%d\n", read_id, name, (unsigned long)tax,
                       (unsigned long)r->score, (unsigned long)r->secondary_score, r->hit_length, r->query_length,
                       r->n_assign);
      if (w < 0 || (size_t)w >= cap - off) return CFR_ERR_ARG;
      off += (size_t)w;
    }
  } else {
    int w = snprintf(buf + off, cap - off, "%s\tunclassified\t0\t0\t0\t0\t%d\t1\n", read_id, r->query_length);
    if (w < 0 || (size_t)w >= cap - off) return CFR_ERR_ARG;
    off += (size_t)w;
  }
  return (int)off;
}

int cfr_taxon_counts_device(cfr_handle *h, void **dev_ptr, uint64_t *n_entries) {
  if (!h || !dev_ptr || !n_entries) return fail(CFR_ERR_ARG, "null argument");
  *dev_ptr = h->d_taxon;
  *n_entries = h->ix.node_cnt + 3;
  return CFR_OK;
}

int cfr_taxon_counts_read(cfr_handle *h, uint64_t *out, uint64_t n_entries, void *stream) {
  if (!h || !out) return fail(CFR_ERR_ARG, "null argument");
  if (n_entries > h->ix.node_cnt + 3) n_entries = h->ix.node_cnt + 3;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = pick_stream(h, stream);
  CUDA_TRY(cudaMemcpyAsync(out, h->d_taxon, n_entries * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return CFR_OK;
}

int cfr_taxon_counts_reset(cfr_handle *h, void *stream) {
  if (!h) return fail(CFR_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemsetAsync(h->d_taxon, 0, (h->ix.node_cnt + 3) * 8, pick_stream(h, stream)));
  return CFR_OK;
}

// ---------------------------------------------------------------------------
// NCCL: the path's one collective -- SUM all-reduce of the per-taxon assignment counters, once, after the last
// batch.  libnccl is loaded at run time (dlopen): the library itself has no link-time
// dependency on it, and a process that never reduces never needs it.
// ---------------------------------------------------------------------------
namespace {
struct NcclApi {
  void *lib = nullptr;
  int (*CommInitAll)(void **, int, const int *) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi &nccl_api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    // NCCL writes its debug lines -- with NCCL_DEBUG=VERSION, which some images set, its version banner -- to STDOUT, where a
    // caller of this library (the CLI) has its TSV: a library keeps out of its caller's stdout
    if (!getenv("NCCL_DEBUG_FILE")) setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    if (const char *d = getenv("NCCL_DEBUG"))  // the version banner is a plain printf to stdout: ask for warnings instead
      if (strcasecmp(d, "VERSION") == 0) setenv("NCCL_DEBUG", "WARN", 1);
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) return;
    a.CommInitAll = (int (*)(void **, int, const int *))dlsym(a.lib, "ncclCommInitAll");
    a.CommDestroy = (int (*)(void *))dlsym(a.lib, "ncclCommDestroy");
    a.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(a.lib, "ncclAllReduce");
    a.GroupStart = (int (*)())dlsym(a.lib, "ncclGroupStart");
    a.GroupEnd = (int (*)())dlsym(a.lib, "ncclGroupEnd");
    a.GetErrorString = (const char *(*)(int))dlsym(a.lib, "ncclGetErrorString");
    a.ok = a.CommInitAll && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd;
  });
  return a;
}
enum { CFR_NCCL_UINT64 = 5, CFR_NCCL_SUM = 0 };  // ncclUint64, ncclSum (nccl.h)

int nccl_fail(int rc, const char *what) {
  NcclApi &a = nccl_api();
  return fail(CFR_ERR_CUDA, std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error"));
}

// snapshot of the live counters, taken after the handle's streams are idle: the copy is what gets reduced
int snapshot_counts(cfr_handle *h) {
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaDeviceSynchronize());
  const size_t bytes = (h->ix.node_cnt + 3) * 8;
  if (!h->d_taxon_reduced) {
    void *p;
    int st = dev_alloc(h, &p, bytes);
    if (st) return st;
    h->d_taxon_reduced = (u64 *)p;
  }
  CUDA_TRY(cudaMemcpyAsync(h->d_taxon_reduced, h->d_taxon, bytes, cudaMemcpyDeviceToDevice, h->stream));
  return CFR_OK;
}
}  // namespace

int cfr_counts_allreduce(cfr_handle *h, void *nccl_comm, uint64_t *out, uint64_t n_entries) {
  if (!h || !nccl_comm) return fail(CFR_ERR_ARG, "null argument");
  NcclApi &a = nccl_api();
  if (!a.ok) return fail(CFR_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  int st = snapshot_counts(h);
  if (st) return st;
  const size_t cnt = h->ix.node_cnt + 3;
  int rc = a.AllReduce(h->d_taxon_reduced, h->d_taxon_reduced, cnt, CFR_NCCL_UINT64, CFR_NCCL_SUM, nccl_comm, h->stream);
  if (rc != 0) return nccl_fail(rc, "ncclAllReduce");
  if (out) {
    if (n_entries > cnt) n_entries = cnt;
    CUDA_TRY(cudaMemcpyAsync(out, h->d_taxon_reduced, n_entries * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return CFR_OK;
}

int cfr_counts_allreduce_local(cfr_handle **handles, int n_handles, uint64_t *out, uint64_t n_entries) {
  if (!handles || n_handles < 1 || !handles[0]) return fail(CFR_ERR_ARG, "null argument");
  const size_t cnt = handles[0]->ix.node_cnt + 3;
  for (int i = 0; i < n_handles; ++i) {
    if (!handles[i] || handles[i]->ix.node_cnt + 3 != cnt) return fail(CFR_ERR_ARG, "handles do not share one index");
    int st = snapshot_counts(handles[i]);
    if (st) return st;
  }
  if (n_handles > 1) {
    NcclApi &a = nccl_api();
    if (!a.ok) return fail(CFR_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    std::vector<void *> comms(n_handles, nullptr);
    std::vector<int> devs(n_handles);
    for (int i = 0; i < n_handles; ++i) devs[i] = handles[i]->device;
    int rc = a.CommInitAll(comms.data(), n_handles, devs.data());
    if (rc != 0) return nccl_fail(rc, "ncclCommInitAll");
    rc = a.GroupStart();
    for (int i = 0; i < n_handles && rc == 0; ++i) {
      cudaSetDevice(handles[i]->device);
      rc = a.AllReduce(handles[i]->d_taxon_reduced, handles[i]->d_taxon_reduced, cnt, CFR_NCCL_UINT64, CFR_NCCL_SUM, comms[i],
                       handles[i]->stream);
    }
    const int rc2 = a.GroupEnd();
    for (int i = 0; i < n_handles; ++i) {
      cudaSetDevice(handles[i]->device);
      cudaStreamSynchronize(handles[i]->stream);
    }
    for (void *c : comms)
      if (c) a.CommDestroy(c);
    if (rc != 0) return nccl_fail(rc, "ncclAllReduce");
    if (rc2 != 0) return nccl_fail(rc2, "ncclGroupEnd");
  }
  if (out) {
    cfr_handle *h = handles[0];
    if (n_entries > cnt) n_entries = cnt;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(out, h->d_taxon_reduced, n_entries * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  return CFR_OK;
}

int cfr_quant_enable(cfr_handle *h, uint64_t min_score, uint64_t min_hit_length) {
  if (!h) return fail(CFR_ERR_ARG, "null argument");
  if (h->P.ids_stride > 64) return fail(CFR_ERR_UNSUPPORTED, "quantification keeps at most 64 targets per read");
  h->quant.on = true;
  h->quant.min_score = min_score;
  h->quant.min_hit = min_hit_length;
  return CFR_OK;
}

int cfr_quant_reset(cfr_handle *h) {
  if (!h) return fail(CFR_ERR_ARG, "null argument");
  h->quant.table.clear();
  h->quant.batches = h->quant.entries_moved = 0;
  return CFR_OK;
}

int cfr_quant_stats(cfr_handle *h, uint64_t *distinct, uint64_t *batches, uint64_t *entries_moved) {
  if (!h) return fail(CFR_ERR_ARG, "null argument");
  if (distinct) *distinct = h->quant.table.size();
  if (batches) *batches = h->quant.batches;
  if (entries_moved) *entries_moved = h->quant.entries_moved;
  return CFR_OK;
}

int cfr_quant_report(cfr_handle **handles, int n_handles, const char *idx_prefix, int format, const char *path) {
  if (!handles || n_handles < 1 || !handles[0] || !idx_prefix) return fail(CFR_ERR_ARG, "null argument");
  Quantifier q;
  std::string err;
  if (q.init(idx_prefix, err) != 0) return fail(CFR_ERR_IO, err);
  const int W = handles[0]->P.ids_stride + 1;
  std::map<std::vector<u32>, u64> all;  // the replicas' tables merged (one process, one handle per GPU)
  for (int g = 0; g < n_handles; ++g) {
    if (!handles[g] || handles[g]->P.ids_stride + 1 != W) return fail(CFR_ERR_ARG, "handles differ in -k");
    for (auto &kv : handles[g]->quant.table) all[kv.first] += kv.second;
  }
  for (auto &kv : all) {
    const std::vector<u32> &w = kv.first;
    if (w[0] == 0xffffffffu) {
      q.add_unclassified(kv.second);
      continue;
    }
    const size_t nt = w[0] & 0xffu;
    uint64_t targets[64];
    for (size_t j = 0; j < nt; ++j) targets[j] = w[1 + j];
    q.add(targets, nt, (int)((w[0] >> 8) & 0xffu), ((w[0] >> 16) & 1u) != 0, kv.second);
  }
  q.quantify();
  FILE *fp = (!path || !strcmp(path, "-")) ? stdout : fopen(path, "w");
  if (!fp) return fail(CFR_ERR_IO, std::string("cannot write ") + path);
  q.output(fp, format);
  if (fp != stdout) fclose(fp); else fflush(fp);
  return CFR_OK;
}

static int read_counters(cfr_handle *h, int stage, cfr_counters *c, cudaStream_t s) {
  DevCounters d[CFR_N_STAGES];
  CUDA_TRY(cudaMemcpyAsync(d, h->d_counters, sizeof(d), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  memset(c, 0, sizeof(*c));
  for (int i = 0; i < CFR_N_STAGES; ++i) {
    if (stage >= 0 && i != stage) continue;
    c->n_rank += d[i].n_rank;
    c->n_access += d[i].n_access;
    c->n_search += d[i].n_search;
    c->n_locate += d[i].n_locate;
    c->n_lf += d[i].n_lf;
    c->n_extend += d[i].n_extend;
  }
  c->n_bases = h->host_bases;
  c->n_reads = d[CFR_STAGE_SCORE].n_reads;
  c->n_launches = stage >= 0 ? h->stage_launches[stage] : h->launches;
  return CFR_OK;
}

int cfr_get_counters(cfr_handle *h, cfr_counters *c, void *stream) {
  if (!h || !c) return fail(CFR_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  return read_counters(h, -1, c, pick_stream(h, stream));
}

int cfr_get_stage_counters(cfr_handle *h, int stage, cfr_counters *c, void *stream) {
  if (!h || !c || stage < 0 || stage >= CFR_N_STAGES) return fail(CFR_ERR_ARG, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  return read_counters(h, stage, c, pick_stream(h, stream));
}

int cfr_reset_counters(cfr_handle *h, void *stream) {
  if (!h) return fail(CFR_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(DevCounters) * CFR_N_STAGES, pick_stream(h, stream)));
  h->launches = 0;
  h->host_bases = 0;
  return CFR_OK;
}

int cfr_set_profiling(cfr_handle *h, int on) {
  if (!h) return fail(CFR_ERR_ARG, "null argument");
  h->profile = on != 0;
  return CFR_OK;
}

int cfr_get_stage_times(cfr_handle *h, cfr_stage_times *out, int reset) {
  if (!h || !out) return fail(CFR_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaDeviceSynchronize());
  for (auto &e : h->ev_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) h->stage_ms[e.stage] += ms;
    h->ev_pool.push_back(e.a);
    h->ev_pool.push_back(e.b);
  }
  h->ev_pending.clear();
  for (int i = 0; i < CFR_N_STAGES; ++i) {
    out->ms[i] = h->stage_ms[i];
    out->launches[i] = h->stage_launches[i];
    if (reset) {
      h->stage_ms[i] = 0;
      h->stage_launches[i] = 0;
    }
  }
  return CFR_OK;
}

// ---- diagnostics -------------------------------------------------------------------
int cfr_debug_bwt_rank(cfr_handle *h, const uint8_t *codes, const uint64_t *pos, const int32_t *inclusive, uint64_t n,
                       uint64_t *out) {
  if (!h || !codes || !pos || !inclusive || !out) return fail(CFR_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  DevBuf dc, dp, di, dout;
  int st;
  if ((st = dc.ensure(n)) || (st = dp.ensure(n * 8)) || (st = di.ensure(n * 4)) || (st = dout.ensure(n * 8))) return st;
  CUDA_TRY(cudaMemcpy(dc.p, codes, n, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dp.p, pos, n * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(di.p, inclusive, n * 4, cudaMemcpyHostToDevice));
  const int grid = (int)((n + 127) / 128);
  if (h->layout == CFR_LAYOUT_OCCLINE)
    k_debug_rank<BwtOccLine><<<grid, 128, 0, h->stream>>>(h->ix, (const unsigned char *)dc.p, (const u64 *)dp.p,
                                                          (const int *)di.p, n, (u64 *)dout.p);
  else
    k_debug_rank<BwtRunBlock><<<grid, 128, 0, h->stream>>>(h->ix, (const unsigned char *)dc.p, (const u64 *)dp.p,
                                                           (const int *)di.p, n, (u64 *)dout.p);
  ++h->launches;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(out, dout.p, n * 8, cudaMemcpyDeviceToHost));
  dc.release(); dp.release(); di.release(); dout.release();
  return CFR_OK;
}

int cfr_debug_bwt_access(cfr_handle *h, const uint64_t *pos, uint64_t n, uint8_t *out) {
  if (!h || !pos || !out) return fail(CFR_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  DevBuf dp, dout;
  int st;
  if ((st = dp.ensure(n * 8)) || (st = dout.ensure(n))) return st;
  CUDA_TRY(cudaMemcpy(dp.p, pos, n * 8, cudaMemcpyHostToDevice));
  const int grid = (int)((n + 127) / 128);
  if (h->layout == CFR_LAYOUT_OCCLINE)
    k_debug_access<BwtOccLine><<<grid, 128, 0, h->stream>>>(h->ix, (const u64 *)dp.p, n, (unsigned char *)dout.p);
  else
    k_debug_access<BwtRunBlock><<<grid, 128, 0, h->stream>>>(h->ix, (const u64 *)dp.p, n, (unsigned char *)dout.p);
  ++h->launches;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(out, dout.p, n, cudaMemcpyDeviceToHost));
  dp.release(); dout.release();
  return CFR_OK;
}

int cfr_debug_locate(cfr_handle *h, const uint64_t *rows, uint64_t n, uint64_t *seq_ids) {
  if (!h || !rows || !seq_ids) return fail(CFR_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  DevBuf dp, dout;
  int st;
  if ((st = dp.ensure(n * 8)) || (st = dout.ensure(n * 8))) return st;
  CUDA_TRY(cudaMemcpy(dp.p, rows, n * 8, cudaMemcpyHostToDevice));
  const int grid = (int)((n + 127) / 128);
  if (h->layout == CFR_LAYOUT_OCCLINE)
    k_debug_locate<BwtOccLine><<<grid, 128, 0, h->stream>>>(h->ix, (const u64 *)dp.p, n, (u64 *)dout.p);
  else
    k_debug_locate<BwtRunBlock><<<grid, 128, 0, h->stream>>>(h->ix, (const u64 *)dp.p, n, (u64 *)dout.p);
  ++h->launches;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(seq_ids, dout.p, n * 8, cudaMemcpyDeviceToHost));
  dp.release(); dout.release();
  return CFR_OK;
}

int cfr_debug_dust(cfr_handle *h, const cfr_read_batch *in, char *masked1, char *masked2) {
  if (!h || !in || !masked1) return fail(CFR_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  cfr_device_batch b;
  int st = upload_chunk(h, in, 0, in->n_reads, &b, h->stream);
  DevBuf dbits, outb;
  if (!st) st = dbits.ensure((b.seq_bytes / 32 + 2) * 4);
  if (!st) st = outb.ensure(b.seq_bytes + 64);
  if (st) {
    b.release();
    dbits.release();
    outb.release();
    return st;
  }
  ChunkDev B;
  fill_chunk(h, &b, B);
  B.mask = (u32 *)b.mask.p;
  B.dust_bits = (u32 *)dbits.p;
  const u64 len1 = in->off1[in->n_reads] - in->off1[0];
  const u64 len2 = in->seq2 ? in->off2[in->n_reads] - in->off2[0] : 0;
  cudaMemsetAsync(b.scalars.p, 0, 64, h->stream);
  k_encode<<<grid_for(h, B.n_words, 256, 8), 256, 0, h->stream>>>(B, b.seq_bytes);
  if (in->n_reads) launch_dust(h, B, h->stream);
  k_apply_dust<<<grid_for(h, b.seq_bytes, 256, 8), 256, 0, h->stream>>>(B, (unsigned char *)outb.p, b.seq_bytes);
  h->launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && len1) e = cudaMemcpyAsync(masked1, outb.p, len1, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && masked2 && len2)
    e = cudaMemcpyAsync(masked2, (char *)outb.p + ((len1 + 31) & ~31ull), len2, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  b.release();
  dbits.release();
  outb.release();
  if (e != cudaSuccess) return fail(CFR_ERR_CUDA, cudaGetErrorString(e));
  return CFR_OK;
}

}  // extern "C"

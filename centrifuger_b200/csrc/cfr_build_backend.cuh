// cfr_build_backend.cuh -- the thin execution layer under the index builder (cfr_build.cu).
//
// The builder is written once as data-parallel passes (par_for over elements, radix sorts of
// key/value pairs, prefix scans, a few atomics).  nvcc compiles those passes as sm_100a kernels
// (extended __device__ lambdas, CUB device-wide sorts and scans: the builder is not the hot path,
// library primitives are fine here).  With -DCFR_HOSTSIM the same passes run as plain loops on the
// host: that twin exists only so the tests in this GPU-less container can compare the builder's
// output with the reference builder's files byte for byte; it is never part of the shipped library.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#if !defined(CFR_HOSTSIM)
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#endif

namespace cfrbuild {

typedef unsigned long long u64;
typedef unsigned int u32;

struct BuildError : std::runtime_error {
  int code;
  BuildError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#if defined(CFR_HOSTSIM)
// ------------------------------------------------------------------ host twin
#define BK_HD inline
#define BK_D inline
#define BK_LAMBDA

template <class F>
void par_for(u64 n, F f) {
  for (u64 i = 0; i < n; ++i) f(i);
}
inline void bk_sync() {}
inline void *bk_alloc(size_t bytes) {
  void *p = calloc(bytes ? bytes : 1, 1);
  if (!p) throw BuildError(-6, "host twin: out of memory");
  return p;
}
inline void bk_free(void *p) { free(p); }
inline void bk_zero(void *p, size_t bytes) { memset(p, 0, bytes); }
inline void bk_fill_ff(void *p, size_t bytes) { memset(p, 0xff, bytes); }
inline void bk_to_host(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
inline void bk_to_dev(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
inline size_t bk_free_bytes() { return (size_t)2 << 30; }
inline void a_or64(u64 *p, u64 v) { *p |= v; }
inline void a_and64(u64 *p, u64 v) { *p &= v; }
inline void a_min64(u64 *p, u64 v) { if (v < *p) *p = v; }
inline void a_max32(u32 *p, u32 v) { if (v > *p) *p = v; }
inline u64 a_add64(u64 *p, u64 v) { u64 o = *p; *p += v; return o; }
inline u64 brev64(u64 x) {
  x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
  return __builtin_bswap64(x);
}
inline int popc64(u64 x) { return __builtin_popcountll(x); }

// stable LSD order by the key bits [begin_bit, end_bit)
template <class K, class V>
void sort_pairs(K *&k, K *&k_alt, V *&v, V *&v_alt, u64 n, int begin_bit, int end_bit) {
  if (n == 0) return;
  const int bits = end_bit - begin_bit;
  const K mask = bits >= (int)(8 * sizeof(K)) ? (K)~(K)0 : (K)((((K)1) << bits) - 1);
  std::vector<u64> perm(n);
  for (u64 i = 0; i < n; ++i) perm[i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](u64 a, u64 b) {
    return ((k[a] >> begin_bit) & mask) < ((k[b] >> begin_bit) & mask);
  });
  for (u64 i = 0; i < n; ++i) {
    k_alt[i] = k[perm[i]];
    v_alt[i] = v[perm[i]];
  }
  std::swap(k, k_alt);
  std::swap(v, v_alt);
}
// out[i] = sum of in[0..i) ; returns the total.  in == out allowed.
template <class T>
u64 exclusive_sum(const T *in, T *out, u64 n) {
  u64 s = 0;
  for (u64 i = 0; i < n; ++i) {
    const T x = in[i];
    out[i] = (T)s;
    s += x;
  }
  return s;
}
inline void inclusive_max_u32(const u32 *in, u32 *out, u64 n) {
  u32 m = 0;
  for (u64 i = 0; i < n; ++i) {
    m = in[i] > m ? in[i] : m;
    out[i] = m;
  }
}
inline void bk_release_temp() {}

#else
// ------------------------------------------------------------------ sm_100a
#define BK_HD __host__ __device__ __forceinline__
#define BK_D __device__ __forceinline__
#define BK_LAMBDA __device__

#define BK_CUDA(x)                                                                                     \
  do {                                                                                                 \
    cudaError_t e_ = (x);                                                                              \
    if (e_ != cudaSuccess) throw BuildError(-5, std::string(#x ": ") + cudaGetErrorString(e_));        \
  } while (0)

template <class F>
__global__ void __launch_bounds__(256) k_build_pass(u64 n, F f) {
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) f(i);
}
template <class F>
void par_for(u64 n, F f) {
  if (n == 0) return;
  const u64 blocks = (n + 255) / 256;
  const int grid = (int)std::min<u64>(blocks, 148ull * 32);  // grid-stride: a multiple of the SM count
  k_build_pass<<<grid, 256>>>(n, f);
  BK_CUDA(cudaGetLastError());
}
inline void bk_sync() { BK_CUDA(cudaDeviceSynchronize()); }
inline void *bk_alloc(size_t bytes) {
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
  if (e != cudaSuccess) throw BuildError(-6, std::string("cudaMalloc(builder, ") + std::to_string(bytes) + " B): " + cudaGetErrorString(e));
  return p;
}
inline void bk_free(void *p) {
  if (p) cudaFree(p);
}
inline void bk_zero(void *p, size_t bytes) { BK_CUDA(cudaMemset(p, 0, bytes)); }
inline void bk_fill_ff(void *p, size_t bytes) { BK_CUDA(cudaMemset(p, 0xff, bytes)); }
inline void bk_to_host(void *dst, const void *src, size_t bytes) { BK_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); }
inline void bk_to_dev(void *dst, const void *src, size_t bytes) { BK_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)); }
inline size_t bk_free_bytes() {
  size_t f = 0, t = 0;
  BK_CUDA(cudaMemGetInfo(&f, &t));
  return f;
}
BK_D void a_or64(u64 *p, u64 v) { atomicOr(p, v); }
BK_D void a_and64(u64 *p, u64 v) { atomicAnd(p, v); }
BK_D void a_min64(u64 *p, u64 v) { atomicMin(p, v); }
BK_D void a_max32(u32 *p, u32 v) { atomicMax(p, v); }
BK_D u64 a_add64(u64 *p, u64 v) { return atomicAdd(p, v); }
BK_D u64 brev64(u64 x) { return __brevll(x); }
BK_D int popc64(u64 x) { return __popcll(x); }

// one scratch area for the CUB calls, grown on demand
struct Temp {
  void *p = nullptr;
  size_t bytes = 0;
  void *need(size_t b) {
    if (b > bytes) {
      bk_free(p);
      p = bk_alloc(b + (b >> 3));
      bytes = b + (b >> 3);
    }
    return p;
  }
};
inline Temp &bk_temp() {
  static Temp t;
  return t;
}
inline void bk_release_temp() {
  Temp &t = bk_temp();
  bk_free(t.p);
  t.p = nullptr;
  t.bytes = 0;
}

template <class K, class V>
void sort_pairs(K *&k, K *&k_alt, V *&v, V *&v_alt, u64 n, int begin_bit, int end_bit) {
  if (n == 0) return;
  if (n >= (1ull << 31)) throw BuildError(-1, "builder: more than 2^31 - 1 suffixes in one batch");
  cub::DoubleBuffer<K> dk(k, k_alt);
  cub::DoubleBuffer<V> dv(v, v_alt);
  size_t tb = 0;
  BK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)n, begin_bit, end_bit));
  void *tp = bk_temp().need(tb);
  BK_CUDA(cub::DeviceRadixSort::SortPairs(tp, tb, dk, dv, (int)n, begin_bit, end_bit));
  k = dk.Current();
  k_alt = dk.Alternate();
  v = dv.Current();
  v_alt = dv.Alternate();
}
template <class T>
u64 exclusive_sum(const T *in, T *out, u64 n) {
  if (n == 0) return 0;
  if (n >= (1ull << 31)) throw BuildError(-1, "builder: scan longer than 2^31 - 1");
  T last_in = 0, last_out = 0;
  bk_to_host(&last_in, in + (n - 1), sizeof(T));
  size_t tb = 0;
  BK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int)n));
  void *tp = bk_temp().need(tb);
  BK_CUDA(cub::DeviceScan::ExclusiveSum(tp, tb, in, out, (int)n));
  bk_to_host(&last_out, out + (n - 1), sizeof(T));
  return (u64)last_in + (u64)last_out;
}
struct MaxU32 {
  __host__ __device__ u32 operator()(u32 a, u32 b) const { return a > b ? a : b; }
};
inline void inclusive_max_u32(const u32 *in, u32 *out, u64 n) {
  if (n == 0) return;
  size_t tb = 0;
  BK_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb, in, out, MaxU32(), (int)n));
  void *tp = bk_temp().need(tb);
  BK_CUDA(cub::DeviceScan::InclusiveScan(tp, tb, in, out, MaxU32(), (int)n));
}
#endif

// owning device (or host-twin) array
template <class T>
struct Buf {
  T *p = nullptr;
  u64 n = 0;
  Buf() {}
  explicit Buf(u64 count) { alloc(count); }
  Buf(const Buf &) = delete;
  Buf &operator=(const Buf &) = delete;
  ~Buf() { release(); }
  void alloc(u64 count) {
    release();
    p = (T *)bk_alloc(count * sizeof(T));
    n = count;
  }
  void release() {
    if (p) bk_free(p);
    p = nullptr;
    n = 0;
  }
  void zero() { bk_zero(p, n * sizeof(T)); }
  std::vector<T> host() const {
    std::vector<T> h(n);
    if (n) bk_to_host(h.data(), p, n * sizeof(T));
    return h;
  }
};

}  // namespace cfrbuild

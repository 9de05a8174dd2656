// membench.cu -- what does B200's memory system deliver for the index access pattern?
// Independent random reads of S bytes (S = 32, 64, 128; aligned) over an array of A bytes.
// Prints accesses/s and useful GB/s; run under ncu for dram__bytes_read.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o membench membench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

__device__ __forceinline__ u64 mix(u64 x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}

template <int S, int DEP>
__global__ void __launch_bounds__(128) k_read(const u64 *a, u64 n_units, int iters, u64 seed, u64 *out) {
  u64 acc = 0;
  u64 r = mix(seed + blockIdx.x * 1315423911ull + threadIdx.x);
  for (int it = 0; it < iters; ++it) {
    r = mix(r + it + (DEP ? (acc & 1) : 0));   // DEP: next address depends on the loaded data (a dependent chain)
    const u64 unit = r % n_units;
    const u64 *p = a + unit * (S / 8);
#pragma unroll
    for (int q = 0; q < S / 32; ++q) {
      u64 x0, x1, x2, x3;
      asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x0), "=l"(x1), "=l"(x2), "=l"(x3) : "l"(p + 4 * q));
      acc += x0 ^ x1 ^ x2 ^ x3;
    }
  }
  if (acc == 0x1234567) out[0] = acc;
}

// G adjacent lanes fetch ONE random 32*G-byte unit together (lane g its g-th sector): a single
// coalesced request per unit instead of G separate ones
template <int G>
__global__ void __launch_bounds__(128) k_read_coop(const u64 *a, u64 n_units, int iters, u64 seed, u64 *out) {
  u64 acc = 0;
  const int g = threadIdx.x % G;
  u64 r = mix(seed + blockIdx.x * 1315423911ull + threadIdx.x / G);
  for (int it = 0; it < iters; ++it) {
    r = mix(r + it);
    const u64 unit = r % n_units;
    const u64 *p = a + unit * (4 * G) + 4 * g;
    u64 x0, x1, x2, x3;
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x0), "=l"(x1), "=l"(x2), "=l"(x3) : "l"(p));
    acc += x0 ^ x1 ^ x2 ^ x3;
  }
  if (acc == 0x1234567) out[0] = acc;
}

template <int G>
void run_coop(const u64 *a, size_t bytes, u64 *out, int blocks_per_sm) {
  const u64 n_units = bytes / (32 * G);
  const int iters = 2000;
  const int grid = 148 * blocks_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_read_coop<G><<<grid, 128>>>(a, n_units, 200, 1, out);
  cudaEventRecord(e0);
  k_read_coop<G><<<grid, 128>>>(a, n_units, iters, 7, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double units = (double)grid * 128 / G * iters;
  printf("array %6.0f MB  %d lanes share one %3d-B unit  blocks/SM %2d : %7.2f G units/s  %7.1f GB/s useful  (%.2f ms)\n",
         bytes / 1e6, G, 32 * G, blocks_per_sm, units / ms / 1e6, units * 32 * G / ms / 1e6, ms);
}

template <int S, int DEP>
void run(const u64 *a, size_t bytes, u64 *out, int blocks_per_sm) {
  const u64 n_units = bytes / S;
  const int iters = 2000;
  const int grid = 148 * blocks_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_read<S, DEP><<<grid, 128>>>(a, n_units, 200, 1, out);
  cudaEventRecord(e0);
  k_read<S, DEP><<<grid, 128>>>(a, n_units, iters, 7, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double acc = (double)grid * 128 * iters;
  printf("array %6.0f MB  read %3d B  dep %d  blocks/SM %2d : %7.2f G access/s  %7.1f GB/s useful  (%.2f ms)\n",
         bytes / 1e6, S, DEP, blocks_per_sm, acc / ms / 1e6, acc * S / ms / 1e6, ms);
}

int main(int argc, char **argv) {
  size_t gran = argc > 1 ? (size_t)atoi(argv[1]) : 0;
  if (gran) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("set L2 fetch granularity %zu -> %s, now %zu\n", gran, cudaGetErrorString(e), got);
  } else {
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("default L2 fetch granularity %zu\n", got);
  }
  u64 *out; cudaMalloc(&out, 64);
  const size_t sizes[] = {50ull << 20, 350ull << 20, 4096ull << 20, 40960ull << 20};
  for (size_t bytes : sizes) {
    u64 *a;
    if (cudaMalloc(&a, bytes) != cudaSuccess) { printf("alloc %zu failed\n", bytes); continue; }
    cudaMemset(a, 1, bytes);
    run<32, 0>(a, bytes, out, 16);
    run<64, 0>(a, bytes, out, 16);
    run<128, 0>(a, bytes, out, 16);
    run<32, 1>(a, bytes, out, 10);
    run<64, 1>(a, bytes, out, 10);
    run<32, 1>(a, bytes, out, 16);
    run_coop<2>(a, bytes, out, 16);
    run_coop<4>(a, bytes, out, 16);
    cudaFree(a);
  }
  return 0;
}

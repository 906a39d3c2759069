// scan.cuh -- device-wide exclusive prefix sum over int32 (hand-written, three-phase: per-tile reduce,
// recursive scan of the tile sums, per-tile scan + offset).  In-place operation (out == in) is allowed.
#pragma once
#include "common.cuh"

namespace gi {

constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 16;
constexpr int kScanTile = kScanThreads * kScanPerThread;  // 4096

__device__ __forceinline__ int block_exclusive_scan_256(int v, int* total) {
  __shared__ int warp_sums[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  int wbase = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s = warp_sums[i];
    if (i < wid) wbase += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return wbase + inc - v;
}

static __global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const int* __restrict__ in, int64_t n, int* __restrict__ sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int acc = 0;
#pragma unroll 4
  for (int i = threadIdx.x; i < kScanTile; i += kScanThreads) {
    const int64_t g = base + i;
    if (g < n) acc += in[g];
  }
  int total;
  block_exclusive_scan_256(acc, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

static __global__ void __launch_bounds__(kScanThreads)
scan_tile_kernel(const int* in, int* out, int64_t n, const int* __restrict__ offsets) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanPerThread;
  int v[kScanPerThread];
  int acc = 0;
#pragma unroll
  for (int i = 0; i < kScanPerThread; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    acc += v[i];
  }
  int total;
  int run = block_exclusive_scan_256(acc, &total) + (offsets ? offsets[blockIdx.x] : 0);
#pragma unroll
  for (int i = 0; i < kScanPerThread; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

// up to kScanSmallTiles tiles by ONE block with a running carry: one launch instead of three for small inputs,
// which are launch-latency-bound (the k-d build scans once per tree level)
constexpr int kScanSmallTiles = 16;
static __global__ void __launch_bounds__(kScanThreads)
scan_small_kernel(const int* in, int* out, int64_t n) {
  int carry = 0;
  for (int64_t t0 = 0; t0 < n; t0 += kScanTile) {
    const int64_t base = t0 + (int64_t)threadIdx.x * kScanPerThread;
    int v[kScanPerThread];
    int acc = 0;
#pragma unroll
    for (int i = 0; i < kScanPerThread; ++i) {
      v[i] = (base + i < n) ? in[base + i] : 0;
      acc += v[i];
    }
    int total;
    int run = block_exclusive_scan_256(acc, &total) + carry;
#pragma unroll
    for (int i = 0; i < kScanPerThread; ++i) {
      if (base + i < n) out[base + i] = run;
      run += v[i];
    }
    carry += total;
  }
}

// exclusive scan of in[0..n) into out[0..n); tile sums live in scratch memory
inline int exclusive_scan_i32(Scratch& scr, const int* in, int* out, int64_t n) {
  cudaStream_t st = scr.stream();
  if (n <= 0) return GI_OK;
  const int64_t tiles = (n + kScanTile - 1) / kScanTile;
  if (tiles == 1) {
    scan_tile_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr);
    GI_CUDA(cudaGetLastError());
    return GI_OK;
  }
  if (tiles <= kScanSmallTiles) {
    scan_small_kernel<<<1, kScanThreads, 0, st>>>(in, out, n);
    GI_CUDA(cudaGetLastError());
    return GI_OK;
  }
  int* sums;
  GI_TRY(scr.alloc(&sums, (size_t)tiles));
  scan_reduce_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, n, sums);
  GI_CUDA(cudaGetLastError());
  GI_TRY(exclusive_scan_i32(scr, sums, sums, tiles));
  scan_tile_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, out, n, sums);
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}

}  // namespace gi

// radix.cuh -- stable LSD radix sort of (key, id) pairs, ascending: passes of 8 bits, 3 launches per pass
// (per-tile digit histogram, per-digit scan over the tiles, stable scatter).  Shared by the k-d tree build (64-bit
// coordinate keys) and the esrg cell binning (32-bit cell keys).
#pragma once
#include <utility>

#include "common.cuh"

namespace gi {

constexpr int kRadix = 256;
constexpr int kSortThreads = 256;
// elements per block: the scatter pass stages a whole tile (keys + ids) in shared memory
template <class K> struct RadixTile { static constexpr int value = sizeof(K) == 8 ? 2048 : 4096; };

// hist[d * tiles + tile] = number of keys of the tile with digit d
template <class K>
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const K* __restrict__ keys, int n, int shift, int* __restrict__ hist, int tiles) {
  __shared__ int h[kRadix];
  h[threadIdx.x] = 0;
  __syncthreads();
  constexpr int kTile = RadixTile<K>::value;
  const int base = blockIdx.x * kTile;
#pragma unroll 4
  for (int c = 0; c < kTile / kSortThreads; ++c) {
    const int i = base + c * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(int)((keys[i] >> shift) & (kRadix - 1))], 1);
  }
  __syncthreads();
  hist[threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

// one block per digit: exclusive scan of its row hist[d][0..tiles) in place, row total to digit_total[d]
static __global__ void __launch_bounds__(256)
radix_scan_kernel(int* __restrict__ hist, int tiles, int* __restrict__ digit_total) {
  __shared__ int warp_sums[8];
  int* row = hist + (size_t)blockIdx.x * tiles;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int carry = 0;
  for (int t0 = 0; t0 < tiles; t0 += 256) {
    const int t = t0 + threadIdx.x;
    const int v = t < tiles ? row[t] : 0;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += u;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int x = warp_sums[w];
      if (w < wid) wbase += x;
      tot += x;
    }
    if (t < tiles) row[t] = carry + wbase + inc - v;
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = carry;
}

// Stable scatter.  An element's position = global offset of (digit, tile) + its rank among the tile's earlier
// elements with the same digit.  Warp w owns the contiguous sub-tile [w*T/8, (w+1)*T/8) and walks it in chunks of
// 32 (lanes in order), so the ranks inside a warp need only warp-level synchronisation: pass 1 counts the warp's
// digits, one block-wide exclusive scan over (digit, warp) turns the counts into running TILE-LOCAL positions,
// pass 2 places the elements in shared memory in sorted order, pass 3 copies them out -- consecutive threads
// write consecutive addresses of a digit's run.  (Writing straight from pass 2 scatters every warp store over up
// to 32 runs: 142 us per pass for 10 M keys.)
template <class K>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const K* __restrict__ keys_in, const int* __restrict__ ids_in,
                     K* __restrict__ keys_out, int* __restrict__ ids_out, int n, int shift,
                     const int* __restrict__ offs, const int* __restrict__ digit_total, int tiles) {
  constexpr int kTile = RadixTile<K>::value;
  constexpr int kWarps = kSortThreads / 32;
  constexpr int kPerWarp = kTile / kWarps;
  constexpr int kChunks = kPerWarp / 32;
  __shared__ int wpos[kWarps][kRadix];  // per-warp digit counts, then running tile-local positions
  __shared__ int gdelta[kRadix];        // global position - tile-local position, per digit
  __shared__ int wsum[kWarps];
  __shared__ K skey[kTile];
  __shared__ int sid[kTile];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) wpos[w][threadIdx.x] = 0;
  int digit_base;  // global offset of (digit = threadIdx.x, this tile)
  {
    const int v = digit_total[threadIdx.x];
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += u;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w)
      if (w < wid) wbase += wsum[w];
    digit_base = wbase + inc - v + offs[threadIdx.x * tiles + blockIdx.x];
  }
  const int tile0 = blockIdx.x * kTile, my0 = tile0 + wid * kPerWarp;
  K key[kChunks];
  unsigned peers[kChunks];  // lanes of the warp with the same digit (0 for lanes past the end of the array)
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int i = my0 + c * 32 + lane;
    const bool valid = i < n;
    key[c] = valid ? keys_in[i] : K(0);
    const int d = (int)((key[c] >> shift) & (kRadix - 1));
    const unsigned same = __match_any_sync(0xffffffffu, valid ? d : kRadix + lane) & __ballot_sync(0xffffffffu, valid);
    peers[c] = valid ? same : 0u;
    if (valid && (same & ((1u << lane) - 1u)) == 0u) wpos[wid][d] += __popc(same);  // lowest lane of its group
    __syncwarp();
  }
  __syncthreads();
  {  // digit = threadIdx.x: tile count, tile-local start (block scan over the digits), per-warp starts
    int cnt = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) cnt += wpos[w][threadIdx.x];
    int inc = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += u;
    }
    __syncthreads();  // (wsum is reused)
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w)
      if (w < wid) wbase += wsum[w];
    int run = wbase + inc - cnt;  // tile-local start of this digit
    gdelta[threadIdx.x] = digit_base - run;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const int c = wpos[w][threadIdx.x];
      wpos[w][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int i = my0 + c * 32 + lane;
    const unsigned same = peers[c];
    const int d = (int)((key[c] >> shift) & (kRadix - 1));
    int pos = 0;
    if (same) pos = wpos[wid][d] + __popc(same & ((1u << lane) - 1u));
    __syncwarp();
    if (same && (same & ((1u << lane) - 1u)) == 0u) wpos[wid][d] += __popc(same);
    __syncwarp();
    if (same) { skey[pos] = key[c]; sid[pos] = ids_in[i]; }
  }
  __syncthreads();
  const int m = min(kTile, n - tile0);
  for (int e = threadIdx.x; e < m; e += kSortThreads) {
    const K k = skey[e];
    const int g = e + gdelta[(int)((k >> shift) & (kRadix - 1))];
    keys_out[g] = k;
    ids_out[g] = sid[e];
  }
}

// Sorts (keys, ids) by the low 8 * passes bits of the keys, stable.  keys_tmp / ids_tmp are scratch of the same
// size, hist holds 256 * (tiles + 1) ints.  The result is in *keys_out / *ids_out (the caller's arrays for an
// even number of passes, the scratch arrays otherwise).
template <class K>
int radix_sort(K* keys, int* ids, K* keys_tmp, int* ids_tmp, int* hist, int n, int passes, cudaStream_t st,
               K** keys_out, int** ids_out) {
  const int tiles = div_up(n, RadixTile<K>::value);
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = 8 * pass;
    radix_hist_kernel<K><<<tiles, kSortThreads, 0, st>>>(keys, n, shift, hist, tiles);
    radix_scan_kernel<<<kRadix, 256, 0, st>>>(hist, tiles, hist + (size_t)kRadix * tiles);
    radix_scatter_kernel<K><<<tiles, kSortThreads, 0, st>>>(keys, ids, keys_tmp, ids_tmp, n, shift, hist,
                                                           hist + (size_t)kRadix * tiles, tiles);
    std::swap(keys, keys_tmp);
    std::swap(ids, ids_tmp);
  }
  GI_CUDA(cudaGetLastError());
  *keys_out = keys;
  *ids_out = ids;
  return GI_OK;
}
template <class K> inline size_t radix_hist_size(int n) { return (size_t)kRadix * (div_up(n, RadixTile<K>::value) + 1); }

}  // namespace gi

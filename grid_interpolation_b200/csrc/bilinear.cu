// bilinear.cu -- regular grid -> scattered points (interpolation.f90:20-100).
//
// One thread handles kPts points strided by the CTA size, so the coordinate loads and the result store
// are fully coalesced and the 4 x kPts dependent field gathers of a thread are all in flight together
// (the kernel is bound by random gathers from a field far larger than L2: ncu shows 8.7 DRAM sectors read per
// point -- L2 fetches more than the touched sector -- at 46 % of peak DRAM throughput;
// cudaLimitMaxL2FetchGranularity = 32 had no effect on B200).  The field is
// addressed through (stride_y, stride_x), so the reference's column-major data_in(len_y,len_x) and numpy's
// row-major layout are read in place.
//
// Bit parity: dx_inv / dy_inv are the reference's SERIAL running sums of axis differences
// (interpolation.f90:49-58) -- floating-point addition is not associative, so the sum is done by a single
// thread per axis in the reference's order; weights and the final sum follow interpolation.f90:81-93 with
// FMA contraction disabled.
#include <algorithm>

#include "kernels.cuh"
#include "radix.cuh"

namespace gi {
namespace {

constexpr int kThreads = 256;
constexpr int kPts = 4;

// interpolation.f90:49-58.  Block 0: x axis, block 1: y axis.  The CTA stages the axis through shared
// memory in coalesced chunks; thread 0 accumulates in the reference's order.
template <class R>
__global__ void __launch_bounds__(256) axis_inv_spacing_kernel(const R* __restrict__ x_axis, int len_x,
                                                               const R* __restrict__ y_axis, int len_y,
                                                               R* __restrict__ inv) {
  __shared__ R buf[2049];
  const R* axis = blockIdx.x == 0 ? x_axis : y_axis;
  const int len = blockIdx.x == 0 ? len_x : len_y;
  R delta = R(0);
  for (int base = 0; base < len - 1; base += 2048) {
    const int cnt = min(2048, len - 1 - base);  // differences in this chunk; needs cnt + 1 values
    __syncthreads();
    for (int i = threadIdx.x; i <= cnt; i += blockDim.x) buf[i] = axis[base + i];
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll 8
      for (int i = 0; i < cnt; ++i) delta = delta + (buf[i + 1] - buf[i]);
    }
  }
  if (threadIdx.x == 0) inv[blockIdx.x] = R(1) / (delta / R(len - 1));
}

// GI_BILINEAR_SEARCH_AXES (extension): the cell is found by bisection on the axis values instead of the
// reference's mean-spacing formula (interpolation.f90:49-58,67,73), which is only right for equally spaced axes.
// Returns the largest i in [0, len-2] with axis[i] <= p, or -1 if p is outside [axis[0], axis[len-1]).
template <class R> __device__ __forceinline__ int axis_cell(const R* __restrict__ axis, int len, R p) {
  if (!(p >= __ldg(axis)) || !(p < __ldg(axis + len - 1))) return -1;  // also NaN
  int lo = 0, hi = len - 1;  // axis[lo] <= p < axis[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(axis + mid) <= p) lo = mid; else hi = mid;
  }
  return lo;
}

template <class R, bool SEARCH>
__global__ void __launch_bounds__(kThreads)
bilinear_kernel(const R* __restrict__ x_axis, int len_x, const R* __restrict__ y_axis, int len_y,
                const R* __restrict__ data, int64_t sy, int64_t sx, PointsView<R> pts, int64_t n,
                const R* __restrict__ inv, R* __restrict__ out) {
  const R dx_inv = __ldg(inv), dy_inv = __ldg(inv + 1);
  const R x_first = __ldg(x_axis), y_first = __ldg(y_axis);
  const R scale = dx_inv * dy_inv;                                        // :93
  const int64_t base = (int64_t)blockIdx.x * (kThreads * kPts) + threadIdx.x;
  R px[kPts], py[kPts];
  int ix[kPts], iy[kPts];
  bool ok[kPts];
#pragma unroll
  for (int j = 0; j < kPts; ++j) {
    const int64_t i = base + (int64_t)j * kThreads;
    if (i < n) { px[j] = pts.x(i); py[j] = pts.y(i); }
    else { px[j] = x_first; py[j] = y_first; }
  }
#pragma unroll
  for (int j = 0; j < kPts; ++j) {
    if (SEARCH) {
      const int cx = axis_cell(x_axis, len_x, px[j]), cy = axis_cell(y_axis, len_y, py[j]);
      ok[j] = cx >= 0 && cy >= 0;
      ix[j] = ok[j] ? cx : 0;
      iy[j] = ok[j] ? cy : 0;
    } else {
      const R fx = floor((px[j] - x_first) * dx_inv);                     // :67
      const R fy = floor((py[j] - y_first) * dy_inv);                     // :73
      // 1 <= ind < len in the reference's 1-based terms (:68,:74); NaN fails every comparison
      ok[j] = fx >= R(0) && fx < R(len_x - 1) && fy >= R(0) && fy < R(len_y - 1);
      ix[j] = ok[j] ? (int)fx : 0;
      iy[j] = ok[j] ? (int)fy : 0;
    }
  }
  R d11[kPts], d12[kPts], d21[kPts], d22[kPts], xa[kPts], xb[kPts], ya[kPts], yb[kPts];
#pragma unroll
  for (int j = 0; j < kPts; ++j) {
    const R* p = data + (int64_t)iy[j] * sy + (int64_t)ix[j] * sx;       // ix <= len_x-2, iy <= len_y-2
    d11[j] = __ldg(p);                                                    // data_in(ind_y,   ind_x)
    d12[j] = __ldg(p + sy);                                               // data_in(ind_y+1, ind_x)
    d21[j] = __ldg(p + sx);                                               // data_in(ind_y,   ind_x+1)
    d22[j] = __ldg(p + sy + sx);                                          // data_in(ind_y+1, ind_x+1)
    xa[j] = __ldg(x_axis + ix[j]); xb[j] = __ldg(x_axis + ix[j] + 1);
    ya[j] = __ldg(y_axis + iy[j]); yb[j] = __ldg(y_axis + iy[j] + 1);
  }
#pragma unroll
  for (int j = 0; j < kPts; ++j) {
    const int64_t i = base + (int64_t)j * kThreads;
    if (i >= n) continue;
    const R w11 = (xb[j] - px[j]) * (yb[j] - py[j]);                      // :81-88
    const R w12 = (xb[j] - px[j]) * (py[j] - ya[j]);
    const R w21 = (px[j] - xa[j]) * (yb[j] - py[j]);
    const R w22 = (px[j] - xa[j]) * (py[j] - ya[j]);
    const R v = (w11 * d11[j] + w12 * d12[j] + w21 * d21[j] + w22 * d22[j]) * scale;  // :90-93
    out[i] = ok[j] ? v : Consts<R>::ofb();                                // :44,:69,:75
  }
}

// ---------------------------------------------------------------------------------------------
// Sector-load variant (f64).  A micro-benchmark (bench/gather_probe.cu) shows that random gathers from a field
// far larger than L2 run at a fixed ~50 G load requests/s on B200 whatever the load flavour or row pitch --
// the limit is the number of outstanding L1 misses per SM, not DRAM bytes.  So the kernel minimises REQUESTS:
// the two field values that are adjacent in memory come from ONE aligned 256-bit load of their 32 B sector
// (plus one scalar load in the 1-in-4 case that the pair straddles two sectors), and the axis pairs likewise:
// 5 requests per point instead of 8.  Needs a 32 B aligned field whose contiguous extent is a multiple of 4.
// `fast` = the unit-stride axis of the field (x for row-major, y for column-major).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld_sector(const double* p, double (&v)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
// v = the aligned sector holding element i; returns elements i and i+1 (the latter from `next` if i % 4 == 3)
__device__ __forceinline__ void pick_pair(const double (&v)[4], int o, double next, double& e0, double& e1) {
  e0 = o == 0 ? v[0] : (o == 1 ? v[1] : (o == 2 ? v[2] : v[3]));
  e1 = o == 0 ? v[1] : (o == 1 ? v[2] : (o == 2 ? v[3] : next));
}

constexpr int kVecPts = 2;

template <bool ROWMAJOR, bool SEARCH>
__global__ void __launch_bounds__(kThreads)
bilinear_sector_kernel(const double* __restrict__ x_axis, int len_x, const double* __restrict__ y_axis, int len_y,
                       const double* __restrict__ data, PointsView<double> pts, int64_t n,
                       const double* __restrict__ inv, double* __restrict__ out) {
  const double dx_inv = __ldg(inv), dy_inv = __ldg(inv + 1);
  const double x_first = __ldg(x_axis), y_first = __ldg(y_axis);
  const double scale = dx_inv * dy_inv;                                   // :93
  const int64_t pitch = ROWMAJOR ? len_x : len_y;                         // elements between two fast-axis runs
  const int64_t base = (int64_t)blockIdx.x * (kThreads * kVecPts) + threadIdx.x;
  double px[kVecPts], py[kVecPts];
  int ix[kVecPts], iy[kVecPts];
  bool ok[kVecPts];
#pragma unroll
  for (int j = 0; j < kVecPts; ++j) {
    const int64_t i = base + (int64_t)j * kThreads;
    if (i < n) { px[j] = pts.x(i); py[j] = pts.y(i); }
    else { px[j] = x_first; py[j] = y_first; }
    if (SEARCH) {
      const int cx = axis_cell(x_axis, len_x, px[j]), cy = axis_cell(y_axis, len_y, py[j]);
      ok[j] = cx >= 0 && cy >= 0;
      ix[j] = ok[j] ? cx : 0;
      iy[j] = ok[j] ? cy : 0;
    } else {
      const double fx = floor((px[j] - x_first) * dx_inv);                // :67
      const double fy = floor((py[j] - y_first) * dy_inv);                // :73
      ok[j] = fx >= 0.0 && fx < double(len_x - 1) && fy >= 0.0 && fy < double(len_y - 1);
      ix[j] = ok[j] ? (int)fx : 0;
      iy[j] = ok[j] ? (int)fy : 0;
    }
  }
  double fa[kVecPts][4], fb[kVecPts][4], xs[kVecPts][4], ys[kVecPts][4], fan[kVecPts], fbn[kVecPts], xn[kVecPts],
      yn[kVecPts];
#pragma unroll
  for (int j = 0; j < kVecPts; ++j) {  // issue every load of every point before the first use
    const int fi = ROWMAJOR ? ix[j] : iy[j], si = ROWMAJOR ? iy[j] : ix[j];
    const double* p = data + (int64_t)si * pitch + (fi & ~3);
    ld_sector(p, fa[j]);
    ld_sector(p + pitch, fb[j]);
    ld_sector(x_axis + (ix[j] & ~3), xs[j]);
    ld_sector(y_axis + (iy[j] & ~3), ys[j]);
    fan[j] = (fi & 3) == 3 ? __ldg(p + 4) : 0.0;
    fbn[j] = (fi & 3) == 3 ? __ldg(p + pitch + 4) : 0.0;
    xn[j] = (ix[j] & 3) == 3 ? __ldg(x_axis + ix[j] + 1) : 0.0;
    yn[j] = (iy[j] & 3) == 3 ? __ldg(y_axis + iy[j] + 1) : 0.0;
  }
#pragma unroll
  for (int j = 0; j < kVecPts; ++j) {
    const int64_t i = base + (int64_t)j * kThreads;
    if (i >= n) continue;
    const int fi = ROWMAJOR ? ix[j] : iy[j];
    double a0, a1, b0, b1, xa, xb, ya, yb;
    pick_pair(fa[j], fi & 3, fan[j], a0, a1);
    pick_pair(fb[j], fi & 3, fbn[j], b0, b1);
    pick_pair(xs[j], ix[j] & 3, xn[j], xa, xb);
    pick_pair(ys[j], iy[j] & 3, yn[j], ya, yb);
    // data_in(ind_y, ind_x), (ind_y+1, ind_x), (ind_y, ind_x+1), (ind_y+1, ind_x+1)
    const double d11 = a0, d12 = ROWMAJOR ? b0 : a1, d21 = ROWMAJOR ? a1 : b0, d22 = b1;
    const double w11 = (xb - px[j]) * (yb - py[j]);                       // :81-88
    const double w12 = (xb - px[j]) * (py[j] - ya);
    const double w21 = (px[j] - xa) * (yb - py[j]);
    const double w22 = (px[j] - xa) * (py[j] - ya);
    const double v = (w11 * d11 + w12 * d12 + w21 * d21 + w22 * d22) * scale;  // :90-93
    out[i] = ok[j] ? v : Consts<double>::ofb();                           // :44,:69,:75
  }
}

template <class R> struct SectorPath {
  static bool launch(const R*, int, const R*, int, const R*, PointsView<R>, int64_t, const R*, R*, bool, bool,
                     cudaStream_t) { return false; }
};
template <> struct SectorPath<double> {
  static bool launch(const double* x_axis, int len_x, const double* y_axis, int len_y, const double* data,
                     PointsView<double> pv, int64_t n, const double* inv, double* out, bool rowmajor, bool search,
                     cudaStream_t st) {
    const int64_t pitch = rowmajor ? len_x : len_y;
    // aligned sector loads must stay inside the arrays: 32 B aligned bases, extents that are multiples of 4
    if ((reinterpret_cast<uintptr_t>(data) & 31) || (reinterpret_cast<uintptr_t>(x_axis) & 31) ||
        (reinterpret_cast<uintptr_t>(y_axis) & 31) || (pitch & 3) || (len_x & 3) || (len_y & 3))
      return false;
    const int64_t blocks = (n + kThreads * kVecPts - 1) / (kThreads * kVecPts);
    if (blocks > 0x7fffffffLL) return false;
#define GI_LAUNCH_SECTOR(RM, SE)                                                                              \
  bilinear_sector_kernel<RM, SE><<<(unsigned)blocks, kThreads, 0, st>>>(x_axis, len_x, y_axis, len_y, data, pv, n, \
                                                                      inv, out)
    if (rowmajor) { if (search) GI_LAUNCH_SECTOR(true, true); else GI_LAUNCH_SECTOR(true, false); }
    else { if (search) GI_LAUNCH_SECTOR(false, true); else GI_LAUNCH_SECTOR(false, false); }
#undef GI_LAUNCH_SECTOR
    return true;
  }
};

}  // namespace

namespace {

// the per-point part: `pv` / `out` address the chunk, `inv` holds the two mean-spacing inverses
template <class R>
int bilinear_points(const R* x_axis, int len_x, const R* y_axis, int len_y, const R* data_in, PointsView<R> pv,
                    int64_t num_points, const R* inv, R* out, int flags, cudaStream_t st) {
  if (num_points == 0) return GI_OK;
  const bool rowmajor = flags & GI_FIELD_ROWMAJOR;
  const int64_t sy = rowmajor ? len_x : 1, sx = rowmajor ? 1 : len_y;
  const bool search = flags & GI_BILINEAR_SEARCH_AXES;
  if (!SectorPath<R>::launch(x_axis, len_x, y_axis, len_y, data_in, pv, num_points, inv, out, rowmajor, search, st)) {
    const int64_t blocks = (num_points + kThreads * kPts - 1) / (kThreads * kPts);
    GI_REQUIRE(blocks <= 0x7fffffffLL, "too many points for one launch");
    if (search)
      bilinear_kernel<R, true><<<(unsigned)blocks, kThreads, 0, st>>>(x_axis, len_x, y_axis, len_y, data_in, sy, sx,
                                                                     pv, num_points, inv, out);
    else
      bilinear_kernel<R, false><<<(unsigned)blocks, kThreads, 0, st>>>(x_axis, len_x, y_axis, len_y, data_in, sy, sx,
                                                                      pv, num_points, inv, out);
  }
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}

}  // namespace

template <class R>
int bilinear_device(const R* x_axis, int len_x, const R* y_axis, int len_y, const R* data_in, const R* points,
                    int64_t num_points, R* data_ip, int flags, cudaStream_t st) {
  GI_REQUIRE(len_x >= 2 && len_y >= 2, "axes need at least 2 nodes");
  GI_REQUIRE(num_points >= 0, "negative num_points");
  GI_TRY(ensure_device());
  if (num_points == 0) return GI_OK;
  Scratch scr(st);
  Phases ph(st);
  ph.mark("axis_spacing");
  R* inv;
  GI_TRY(scr.alloc(&inv, 2));
  axis_inv_spacing_kernel<R><<<2, 256, 0, st>>>(x_axis, len_x, y_axis, len_y, inv);
  GI_CUDA(cudaGetLastError());
  ph.mark("bilinear");
  GI_TRY(bilinear_points<R>(x_axis, len_x, y_axis, len_y, data_in, points_n2(points, num_points, flags & GI_POINTS_ROWMAJOR),
                            num_points, inv, data_ip, flags, st));
  ph.finish();
  return GI_OK;
}

// GI_HOST form: the field goes in first; the points then stream through in chunks -- chunk c+1 is copied in and
// chunk c-1's results are copied out (opposite PCIe direction) while chunk c is interpolated (common.cuh).
template <class R>
int bilinear_host(const R* x_axis, int len_x, const R* y_axis, int len_y, const R* data_in, const R* points,
                  int64_t num_points, R* data_ip, int flags) {
  GI_REQUIRE(len_x >= 2 && len_y >= 2, "axes need at least 2 nodes");
  GI_REQUIRE(num_points >= 0, "negative num_points");
  GI_TRY(ensure_device());
  if (num_points == 0) return GI_OK;
  HostPipe* pipe;
  GI_TRY(host_pipe(&pipe));
  cudaStream_t st = pipe->compute;
  Scratch scr(st);
  Phases ph(st);
  ph.mark("h2d");
  R *dx, *dy, *dd, *dp, *dout, *inv;
  GI_TRY(to_device(scr, x_axis, (size_t)len_x, &dx));
  GI_TRY(to_device(scr, y_axis, (size_t)len_y, &dy));
  GI_TRY(scr.alloc(&dp, (size_t)2 * num_points));
  GI_TRY(scr.alloc(&dout, (size_t)num_points));
  GI_TRY(scr.alloc(&inv, 2));
  axis_inv_spacing_kernel<R><<<2, 256, 0, st>>>(dx, len_x, dy, len_y, inv);
  GI_CUDA(cudaGetLastError());
  ph.mark("bilinear");
  const bool interleaved = flags & GI_POINTS_ROWMAJOR;
  const int nc = plan_bands(num_points, (int64_t)num_points * 3 * sizeof(R));
  // a single chunk needs no extra streams (small calls are latency-bound)
  cudaStream_t cin = nc > 1 ? pipe->copy_in : st, cout = nc > 1 ? pipe->copy_out : st;
  if (nc > 1) {  // the allocations above are ordered on the compute stream: the copy-in stream waits for them
    GI_CUDA(cudaEventRecord(pipe->computed[0], st));
    GI_CUDA(cudaStreamWaitEvent(cin, pipe->computed[0], 0));
  }
  GI_TRY(to_device(scr, data_in, (size_t)len_x * len_y, &dd));  // (compute stream, overlaps the first point chunks)
  for (int c = 0; c < nc; ++c) {
    int64_t c0, c1;
    band_range(0, num_points, nc, c, 4096, &c0, &c1);
    if (c0 >= c1) continue;
    const size_t cnt = (size_t)(c1 - c0);
    if (interleaved) {
      GI_CUDA(cudaMemcpyAsync(dp + 2 * c0, points + 2 * c0, 2 * cnt * sizeof(R), cudaMemcpyHostToDevice, cin));
    } else {
      GI_CUDA(cudaMemcpyAsync(dp + c0, points + c0, cnt * sizeof(R), cudaMemcpyHostToDevice, cin));
      GI_CUDA(cudaMemcpyAsync(dp + num_points + c0, points + num_points + c0, cnt * sizeof(R), cudaMemcpyHostToDevice,
                              cin));
    }
    if (nc > 1) {
      GI_CUDA(cudaEventRecord(pipe->arrived[c], cin));
      GI_CUDA(cudaStreamWaitEvent(st, pipe->arrived[c], 0));
    }
    const PointsView<R> pv = interleaved ? PointsView<R>{dp + 2 * c0, 2, 1} : PointsView<R>{dp + c0, 1, num_points};
    GI_TRY(bilinear_points<R>(dx, len_x, dy, len_y, dd, pv, (int64_t)cnt, inv, dout + c0, flags, st));
    if (nc > 1) {
      GI_CUDA(cudaEventRecord(pipe->computed[c], st));
      GI_CUDA(cudaStreamWaitEvent(cout, pipe->computed[c], 0));
    }
    GI_TRY(to_host(cout, data_ip + c0, dout + c0, cnt));
  }
  ph.finish();
  if (nc > 1) {
    GI_CUDA(cudaStreamSynchronize(pipe->copy_out));
    GI_CUDA(cudaStreamSynchronize(pipe->copy_in));
  }
  GI_CUDA(cudaStreamSynchronize(st));
  return GI_OK;
}

// ---------------------------------------------------------------------------------------------
// Morton order of the points (north_star: "Morton-ordered targets", "Morton ranges" per GPU).  bilinear's gathers
// are random reads from a field far larger than L2 (8.7 DRAM sectors per point, see above); points that arrive
// in the Morton order of their cells read each part of the field once, and a contiguous range of that order
// touches a compact part of the field -- the natural shard for several GPUs.  This entry point only computes
// the order (cell as interpolation.f90:67,73; 16 x 16-node blocks along a Z curve; stable radix sort): whether
// the points are moved, and when, is the caller's business (an ingest-time decision; the un-permutation of
// 1e8 results costs as much as the gathers it saves, profiles/r01_kd_esrg_variants.md).
// ---------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ unsigned spread_bits16(unsigned v) {  // 16 bits -> every other bit of 32
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
template <class R>
__global__ void __launch_bounds__(256)
morton_key_kernel(const R* __restrict__ x_axis, int len_x, const R* __restrict__ y_axis, int len_y, PointsView<R> pts,
                  int n, const R* __restrict__ inv, int shift, unsigned key_oob, unsigned* __restrict__ keys,
                  int* __restrict__ ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const R fx = floor((pts.x(i) - __ldg(x_axis)) * __ldg(inv));            // interpolation.f90:67
  const R fy = floor((pts.y(i) - __ldg(y_axis)) * __ldg(inv + 1));        // :73
  const bool ok = fx >= R(0) && fx < R(len_x - 1) && fy >= R(0) && fy < R(len_y - 1);
  keys[i] = ok ? (spread_bits16((unsigned)fx >> shift) | (spread_bits16((unsigned)fy >> shift) << 1)) : key_oob;
  ids[i] = i;
}
}  // namespace

template <class R>
int bilinear_morton_order_device(const R* x_axis, int len_x, const R* y_axis, int len_y, const R* points,
                                 int64_t num_points, int32_t* order, int flags, cudaStream_t st) {
  GI_REQUIRE(len_x >= 2 && len_y >= 2, "axes need at least 2 nodes");
  GI_REQUIRE(num_points >= 0 && num_points < (int64_t)0x7fffffff, "num_points must be below 2^31");
  GI_TRY(ensure_device());
  if (num_points == 0) return GI_OK;
  const int n = (int)num_points;
  Scratch scr(st);
  R* inv;
  GI_TRY(scr.alloc(&inv, 2));
  axis_inv_spacing_kernel<R><<<2, 256, 0, st>>>(x_axis, len_x, y_axis, len_y, inv);
  // 16 x 16-node blocks, coarser if the Z value of the last block would not fit 31 bits
  int shift = 4;
  while ((std::max(len_x, len_y) >> shift) >= (1 << 15)) ++shift;
  int bits = 0;
  while (((std::max(len_x, len_y) - 1) >> shift) >> bits) ++bits;
  const unsigned key_oob = 1u << (2 * bits);              // points outside the grid sort last
  const int passes = (2 * bits + 1 + 7) / 8;
  unsigned *keys, *keys_tmp, *keys_sorted;
  int *ids, *ids_tmp, *ids_sorted, *hist;
  GI_TRY(scr.alloc(&keys, (size_t)n));
  GI_TRY(scr.alloc(&keys_tmp, (size_t)n));
  GI_TRY(scr.alloc(&ids, (size_t)n));
  GI_TRY(scr.alloc(&ids_tmp, (size_t)n));
  GI_TRY(scr.alloc(&hist, radix_hist_size<unsigned>(n)));
  morton_key_kernel<R><<<div_up(n, 256), 256, 0, st>>>(x_axis, len_x, y_axis, len_y,
                                                     points_n2(points, num_points, flags & GI_POINTS_ROWMAJOR), n, inv,
                                                     shift, key_oob, keys, ids);
  GI_CUDA(cudaGetLastError());
  GI_TRY(radix_sort<unsigned>(keys, ids, keys_tmp, ids_tmp, hist, n, passes, st, &keys_sorted, &ids_sorted));
  GI_CUDA(cudaMemcpyAsync(order, ids_sorted, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  return GI_OK;
}
template int bilinear_morton_order_device<double>(const double*, int, const double*, int, const double*, int64_t,
                                                  int32_t*, int, cudaStream_t);
template int bilinear_morton_order_device<float>(const float*, int, const float*, int, const float*, int64_t, int32_t*,
                                                 int, cudaStream_t);

template int bilinear_device<double>(const double*, int, const double*, int, const double*, const double*, int64_t,
                                     double*, int, cudaStream_t);
template int bilinear_device<float>(const float*, int, const float*, int, const float*, const float*, int64_t,
                                    float*, int, cudaStream_t);
template int bilinear_host<double>(const double*, int, const double*, int, const double*, const double*, int64_t,
                                   double*, int);
template int bilinear_host<float>(const float*, int, const float*, int, const float*, const float*, int64_t, float*,
                                  int);

}  // namespace gi

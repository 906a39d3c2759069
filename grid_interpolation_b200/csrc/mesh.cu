// mesh.cu -- mesh ingest: spatial ordering of a triangle mesh and of its points.
//
// The reference takes the mesh as scipy / Qhull deliver it (test_interpolation.py:64-67: simplices and neighbours
// in Qhull's facet order, points in the caller's order), and neither order has anything to do with position.  Every
// kernel of the barycentric path that follows a link -- the pack pass gathering three vertices per triangle, the walk
// moving to a neighbour -- then reads from a random place (ncu on the pack pass: 1.7 GB fetched for 120 MB of vertex
// data, profiles/r01_pack_mesh_ncu.csv).  gi_mesh_reorder numbers triangles and (optionally) points along the Z
// curve of their position once, at ingest, so that neighbours in space are neighbours in memory: the same mesh,
// the same fields (vertex order inside a triangle and the neighbour columns are kept, so every value is computed
// from the same operands), other ids -- and the returned orders map ids back.
//
// No counterpart in the reference ("mesh ingest"); the definition is restated in numpy by
// tests/test_gpu_mesh.py.
#include "kernels.cuh"
#include "radix.cuh"

namespace gi {
namespace {

__device__ __forceinline__ unsigned long long box_key(double v) {  // monotone map double -> uint64
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double box_value(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// box[0..3] = keys of min x, min y, max x, max y (initialised to ~0, ~0, 0, 0)
template <class R>
__global__ void __launch_bounds__(256) mesh_box_kernel(PointsView<R> pts, int n, unsigned long long* box) {
  double x0 = 1.7976931348623157e308, y0 = x0, x1 = -x0, y1 = -x0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    R x, y;
    pts.xy(i, x, y);
    x0 = fmin(x0, (double)x); x1 = fmax(x1, (double)x); y0 = fmin(y0, (double)y); y1 = fmax(y1, (double)y);
  }
  for (int off = 16; off > 0; off >>= 1) {
    x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, off)); y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, off));
    x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, off)); y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, off));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(box + 0, box_key(x0)); atomicMin(box + 1, box_key(y0));
    atomicMax(box + 2, box_key(x1)); atomicMax(box + 3, box_key(y1));
  }
}

// The Z value of a position: 16 bits per axis over the bounding box of the points,
//   q = trunc((v - v_min) * (65535 / (v_max - v_min)))  clamped to 0..65535, evaluated in R,
// x in the even bits, y in the odd ones.
template <class R> struct ZScale { R x0, y0, sx, sy; };
template <class R> __global__ void mesh_scale_kernel(const unsigned long long* box, ZScale<R>* out) {
  const R x0 = (R)box_value(box[0]), y0 = (R)box_value(box[1]), x1 = (R)box_value(box[2]), y1 = (R)box_value(box[3]);
  ZScale<R> s;
  s.x0 = x0; s.y0 = y0;
  s.sx = x1 > x0 ? R(65535) / (x1 - x0) : R(0);
  s.sy = y1 > y0 ? R(65535) / (y1 - y0) : R(0);
  *out = s;
}
__device__ __forceinline__ unsigned spread16(unsigned v) {
  v &= 0xffffu;
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
template <class R> __device__ __forceinline__ unsigned quant16(R v, R v0, R scale) {
  const R f = (v - v0) * scale;
  if (!(f >= R(0))) return 0u;   // also NaN
  if (f >= R(65535)) return 65535u;
  return (unsigned)f;
}
template <class R> __device__ __forceinline__ unsigned z_value(const ZScale<R>& s, R x, R y) {
  return spread16(quant16(x, s.x0, s.sx)) | (spread16(quant16(y, s.y0, s.sy)) << 1);
}

template <class R>
__global__ void __launch_bounds__(256)
point_key_kernel(PointsView<R> pts, int n, const ZScale<R>* __restrict__ scale, unsigned* __restrict__ keys,
                 int* __restrict__ ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ZScale<R> s = *scale;
  R x, y;
  pts.xy(i, x, y);
  keys[i] = z_value(s, x, y);
  ids[i] = i;
}

// a triangle sorts by its centroid ((x1 + x2 + x3) / 3, (y1 + y2 + y3) / 3); a vertex id out of range: last
template <class R>
__global__ void __launch_bounds__(256)
tri_key_kernel(PointsView<R> pts, int num_points, TopoView simp, int num_tri, const ZScale<R>* __restrict__ scale,
               unsigned* __restrict__ keys, int* __restrict__ ids, int* status) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tri) return;
  const ZScale<R> s = *scale;
  const int a = simp(t, 0), b = simp(t, 1), c = simp(t, 2);
  unsigned key = 0xffffffffu;
  if (a < 1 || a > num_points || b < 1 || b > num_points || c < 1 || c > num_points) {
    raise_status(status, GI_ERR_BAD_INDEX);
  } else {
    R x1, y1, x2, y2, x3, y3;
    pts.xy(a - 1, x1, y1); pts.xy(b - 1, x2, y2); pts.xy(c - 1, x3, y3);
    key = z_value(s, (x1 + x2 + x3) / R(3), (y1 + y2 + y3) / R(3));
  }
  keys[t] = key;
  ids[t] = t;
}

// inv[order[j]] = j, order_out[j] = order[j] + 1
__global__ void __launch_bounds__(256)
invert_kernel(const int* __restrict__ order, int n, int* __restrict__ inv, int32_t* __restrict__ order_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int i = order[j];
  inv[i] = j;
  if (order_out) order_out[j] = i + 1;
}

template <class R>
__global__ void __launch_bounds__(256)
move_points_kernel(PointsView<R> pts, const int* __restrict__ order, int n, R* __restrict__ out, int rowmajor) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  R x, y;
  pts.xy(order[j], x, y);
  if (rowmajor) { out[2 * (int64_t)j] = x; out[2 * (int64_t)j + 1] = y; }
  else { out[j] = x; out[(int64_t)n + j] = y; }
}

// triangle j of the new numbering = triangle order[j] of the old one, vertex ids and links renumbered
__global__ void __launch_bounds__(256)
move_topology_kernel(TopoView simp, TopoView nbr, const int* __restrict__ tri_order, const int* __restrict__ tri_inv,
                     const int* __restrict__ point_inv, int num_points, int num_tri, int32_t* __restrict__ simp_out,
                     int32_t* __restrict__ nbr_out, int rowmajor, int* status) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= num_tri) return;
  const int t = tri_order[j];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int v = simp(t, c), nb = nbr(t, c);
    if (v < 1 || v > num_points) { raise_status(status, GI_ERR_BAD_INDEX); v = 0; }
    else if (point_inv) v = point_inv[v - 1] + 1;
    if (nb > num_tri) { raise_status(status, GI_ERR_BAD_INDEX); nb = 0; }
    else if (nb >= 1) nb = tri_inv[nb - 1] + 1;
    else nb = 0;   // no neighbour (0; scipy's -1 + 1)
    const int64_t o = rowmajor ? (int64_t)j * 3 + c : (int64_t)j + (int64_t)c * num_tri;
    simp_out[o] = v;
    nbr_out[o] = nb;
  }
}

// ids 0..n-1 sorted by key (stable): *order lives in scr
int sort_by_z(Scratch& scr, unsigned* keys, int* ids, int n, cudaStream_t st, int** order) {
  unsigned *keys_tmp, *keys_sorted;
  int *ids_tmp, *hist;
  GI_TRY(scr.alloc(&keys_tmp, (size_t)n));
  GI_TRY(scr.alloc(&ids_tmp, (size_t)n));
  GI_TRY(scr.alloc(&hist, radix_hist_size<unsigned>(n)));
  return radix_sort<unsigned>(keys, ids, keys_tmp, ids_tmp, hist, n, 4, st, &keys_sorted, order);
}
}  // namespace

template <class R>
int mesh_reorder_device(const R* points, int num_points, const int32_t* simplices, const int32_t* neighbours,
                        int num_tri, R* points_out, int32_t* point_order, int32_t* simplices_out,
                        int32_t* neighbours_out, int32_t* triangle_order, int flags, cudaStream_t st) {
  GI_REQUIRE(num_points >= 3 && num_tri >= 1, "need at least 3 points and 1 triangle");
  GI_REQUIRE((points_out == nullptr) == (point_order == nullptr), "points_out and point_order come together");
  GI_TRY(ensure_device());
  StatusSlot slot;
  GI_TRY(get_status_slot(st, &slot));
  Scratch scr(st);
  Phases ph(st);
  ph.mark("mesh_reorder");
  const PointsView<R> pv = points_n2(points, num_points, flags & GI_POINTS_ROWMAJOR);
  const TopoView sv = topo_t3(simplices, num_tri, flags & GI_TOPO_ROWMAJOR);
  const TopoView nv = topo_t3(neighbours, num_tri, flags & GI_TOPO_ROWMAJOR);
  unsigned long long* box;
  ZScale<R>* scale;
  GI_TRY(scr.alloc(&box, 4));
  GI_TRY(scr.alloc(&scale, 1));
  GI_CUDA(cudaMemsetAsync(box, 0xff, sizeof(unsigned long long) * 2, st));
  GI_CUDA(cudaMemsetAsync(box + 2, 0, sizeof(unsigned long long) * 2, st));
  mesh_box_kernel<R><<<std::min(div_up(num_points, 256), kNumSMs * 8), 256, 0, st>>>(pv, num_points, box);
  mesh_scale_kernel<R><<<1, 1, 0, st>>>(box, scale);
  GI_CUDA(cudaGetLastError());
  const int nmax = std::max(num_points, num_tri);
  unsigned* keys;
  int* ids;
  GI_TRY(scr.alloc(&keys, (size_t)nmax));
  GI_TRY(scr.alloc(&ids, (size_t)nmax));
  int* point_inv = nullptr;
  if (points_out) {
    int* order;
    point_key_kernel<R><<<div_up(num_points, 256), 256, 0, st>>>(pv, num_points, scale, keys, ids);
    GI_TRY(sort_by_z(scr, keys, ids, num_points, st, &order));
    GI_TRY(scr.alloc(&point_inv, (size_t)num_points));
    invert_kernel<<<div_up(num_points, 256), 256, 0, st>>>(order, num_points, point_inv, point_order);
    move_points_kernel<R><<<div_up(num_points, 256), 256, 0, st>>>(pv, order, num_points, points_out,
                                                                  (flags & GI_POINTS_ROWMAJOR) ? 1 : 0);
    GI_CUDA(cudaGetLastError());
  }
  int *tri_order, *tri_inv;
  tri_key_kernel<R><<<div_up(num_tri, 256), 256, 0, st>>>(pv, num_points, sv, num_tri, scale, keys, ids, slot.dev);
  GI_TRY(sort_by_z(scr, keys, ids, num_tri, st, &tri_order));
  GI_TRY(scr.alloc(&tri_inv, (size_t)num_tri));
  invert_kernel<<<div_up(num_tri, 256), 256, 0, st>>>(tri_order, num_tri, tri_inv, triangle_order);
  move_topology_kernel<<<div_up(num_tri, 256), 256, 0, st>>>(sv, nv, tri_order, tri_inv, point_inv, num_points, num_tri,
                                                            simplices_out, neighbours_out,
                                                            (flags & GI_TOPO_ROWMAJOR) ? 1 : 0, slot.dev);
  GI_CUDA(cudaGetLastError());
  ph.finish();
  GI_TRY(flush_status(st, slot));
  return GI_OK;
}

template int mesh_reorder_device<double>(const double*, int, const int32_t*, const int32_t*, int, double*, int32_t*,
                                         int32_t*, int32_t*, int32_t*, int, cudaStream_t);
template int mesh_reorder_device<float>(const float*, int, const int32_t*, const int32_t*, int, float*, int32_t*,
                                        int32_t*, int32_t*, int32_t*, int, cudaStream_t);

}  // namespace gi

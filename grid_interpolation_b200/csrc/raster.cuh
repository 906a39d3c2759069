// raster.cuh -- tile rasteriser for the barycentric path (included by barycentric.cu inside gi::<anonymous>).
//
// Replaces the per-target visibility walk (triangle_walk.f90:24-84, called from interpolation.f90:365-373) for dense
// meshes.  The triangle the reference finds for a target is the one whose three orientation tests pass
// (triangle_walk.f90:140-153); outside the margin band of an edge that triangle is unique and does not depend on
// the walk, so it can be found from the triangle's side instead of the target's:
//
//   pack_mesh_kernel +  (two passes: count -- part of the pack pass, which has the vertices in registers --
//   bin_fill_kernel     and fill) every packed triangle is entered in the list of each 128 x 32-target tile its
//                       bounding box meets; bounds come from the axes themselves
//                       (searched, not assumed uniform), inflated by the reach of the -1e-10 margin.
//   raster_eval_kernel  one CTA per tile.  Phase A: one thread per listed triangle tests the grid nodes of its
//                       (tile-clipped) bounding box with the reference's orientation expression -- same operations,
//                       same order, the row-invariant product hoisted -- and writes its id into a shared-memory
//                       tile for the nodes it accepts.  No neighbour links, no dependent loads, no walk.
//                       Nodes inside the margin band of an edge (|min cross| below the triangle's threshold, see
//                       amb_threshold) are accepted by two triangles or by a hull triangle only through the margin:
//                       they go to the fix-up list and are re-located along the reference's own walk path by
//                       fixup_kernel, exactly as with the walk kernel.  Phase B = evaluate_strip, shared with it.
//
// The rasteriser needs ascending axes and a tile list that fits its buffer; both are checked on the device
// (TileBins::bad) and the walk kernel runs instead when the flag is raised.

// min / max by compare-and-select (fmin / fmax on doubles expand to ~7 instructions of NaN handling each)
template <class R> __device__ __forceinline__ R rmin(R a, R b) { return a < b ? a : b; }
template <class R> __device__ __forceinline__ R rmax(R a, R b) { return a > b ? a : b; }

__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(unsigned addr, int v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <class R> struct Inf;
template <> struct Inf<double> { __device__ static double v() { return __longlong_as_double(0x7ff0000000000000ll); } };
template <> struct Inf<float> { __device__ static float v() { return __int_as_float(0x7f800000); } };

// First index i in [0, n] with !(ax(i) < v) (UPPER = false) or with ax(i) > v (UPPER = true) on an ascending
// sequence; starts at `guess` (exact for uniform axes), gallops, then bisects.
template <bool UPPER, class R, class Ax>
__device__ __forceinline__ int bound_from(Ax ax, int n, R v, int guess) {
  auto before = [&](int i) { return UPPER ? !(ax(i) > v) : (ax(i) < v); };  // i lies before the bound
  int g = min(max(guess, 0), n), lo, hi;
  if (g < n && before(g)) {
    lo = g + 1; hi = n;
    for (int step = 1; lo < n; step <<= 1) {
      const int p = min(lo + step - 1, n - 1);
      if (before(p)) lo = p + 1;
      else { hi = p; break; }
    }
  } else {
    hi = g; lo = 0;
    for (int step = 1; hi > 0; step <<= 1) {
      const int p = max(hi - step, 0);
      if (!before(p)) hi = p;
      else { lo = p + 1; break; }
    }
  }
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (before(mid)) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

template <class R> __device__ __forceinline__ int guess_index(R v, R first, R inv_step, int n) {
  const R g = (v - first) * inv_step;
  return (int)rmin(rmax(g, R(-1)), R(n + 1));  // (NaN -> n + 1: the search starts at the end)
}

// Bounds of the nodes a triangle can accept, along the fast and the slow axis.  The orientation test passes down
// to cross >= -1e-10 (triangle_walk.f90:148), i.e. up to 1e-10 / |edge| beyond an edge, and rounding adds a few
// ulps of the coordinates: the bounding box is inflated by both (capped by its own extent for degenerate
// triangles).
template <class R>
__device__ __forceinline__ void tri_bounds(const TriGeom<R>& g, bool fy, R& flo, R& fhi, R& slo, R& shi) {
  const R xmin = rmin(rmin(g.x1, g.x2), g.x3), xmax = rmax(rmax(g.x1, g.x2), g.x3);
  const R ymin = rmin(rmin(g.y1, g.y2), g.y3), ymax = rmax(rmax(g.y1, g.y2), g.y3);
  const R l1 = rmax(fabs(g.x2 - g.x1), fabs(g.y2 - g.y1));
  const R l2 = rmax(fabs(g.x3 - g.x2), fabs(g.y3 - g.y2));
  const R l3 = rmax(fabs(g.x1 - g.x3), fabs(g.y1 - g.y3));
  const R lmin = rmin(rmin(l1, l2), l3), ext = rmax(xmax - xmin, ymax - ymin);
  const R eps = sizeof(R) == 8 ? R(1.1102230246251565e-16) : R(5.9604645e-8);
  const R mag = rmax(rmax(fabs(xmin), fabs(xmax)), rmax(fabs(ymin), fabs(ymax)));
  R d = lmin > R(0) ? rmin(R(8e-10) / lmin, ext) : ext;
  d = d + R(64) * eps * (ext + mag);
  flo = (fy ? ymin : xmin) - d; fhi = (fy ? ymax : xmax) + d;
  slo = (fy ? xmin : ymin) - d; shi = (fy ? xmax : ymax) + d;
}

// Inclusive range of the tiles of window `w` holding nodes inside the triangle's bounds; false if there are none.
template <class R>
__device__ __forceinline__ bool tri_tile_range(const TriGeom<R>& g, const TargetWindow<R>& w, int& tx0, int& tx1, int& ty0,
                                               int& ty1) {
  R flo, fhi, slo, shi;
  tri_bounds(g, w.fast_is_y != 0, flo, fhi, slo, shi);
  const int nf = w.nf(), ns = w.ns();
  const R* fa = w.fast_axis + w.f0;
  const R* sa = w.slow_axis + w.s0;
  const R f_first = __ldg(fa), s_first = __ldg(sa);
  const R f_span = __ldg(fa + nf - 1) - f_first, s_span = __ldg(sa + ns - 1) - s_first;
  const R inv_f = f_span > R(0) ? R(nf - 1) / f_span : R(0), inv_s = s_span > R(0) ? R(ns - 1) / s_span : R(0);
  auto fax = [&](int i) { return __ldg(fa + i); };
  auto sax = [&](int i) { return __ldg(sa + i); };
  // nodes [c0, c1] x [r0, r1] of the window lie inside the bounds
  const int c0 = bound_from<false>(fax, nf, flo, guess_index(flo, f_first, inv_f, nf));
  const int c1 = bound_from<true>(fax, nf, fhi, guess_index(fhi, f_first, inv_f, nf) + 1) - 1;
  if (c0 > c1) return false;
  const int r0 = bound_from<false>(sax, ns, slo, guess_index(slo, s_first, inv_s, ns));
  const int r1 = bound_from<true>(sax, ns, shi, guess_index(shi, s_first, inv_s, ns) + 1) - 1;
  if (r0 > r1) return false;
  tx0 = c0 / kTileW; tx1 = c1 / kTileW; ty0 = r0 / kTileH; ty1 = r1 / kTileH;
  return true;
}

// Ascending, finite axes?  (thread t looks at nodes t, t + 1 of both axes; part of the pack pass)
template <class R> __device__ __forceinline__ void check_axes(const TargetWindow<R>& w, int t, int* bad) {
  if (t + 1 < w.nf_full && !(__ldg(w.fast_axis + t) <= __ldg(w.fast_axis + t + 1))) *bad = 1;
  if (t + 1 < w.ns_full && !(__ldg(w.slow_axis + t) <= __ldg(w.slow_axis + t + 1))) *bad = 1;
  if (t == 0 && !(fabs(__ldg(w.fast_axis)) < Inf<R>::v() && fabs(__ldg(w.fast_axis + w.nf_full - 1)) < Inf<R>::v() &&
                  fabs(__ldg(w.slow_axis)) < Inf<R>::v() && fabs(__ldg(w.slow_axis + w.ns_full - 1)) < Inf<R>::v()))
    *bad = 1;
}

// Second binning pass (the first, counting, is part of pack_mesh_kernel): every triangle enters the lists of its
// tiles.  The order inside a list is whatever the atomics make it; nothing depends on it.
__global__ void __launch_bounds__(256) bin_fill_kernel(int num_tri, const TileBins b) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tri) return;
  const uint2 rg = b.range[t];
  const int tx0 = rg.x & 0xffff, tx1 = rg.x >> 16, ty0 = rg.y & 0xffff, ty1 = rg.y >> 16;
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx) {
      const int tile = ty * b.ntx + tx;
      const int pos = __ldg(b.start + tile) + atomicAdd(b.fill + tile, 1);
      if (pos < b.cap) b.list[pos] = t;
      else *b.bad = 1;
    }
}

// Phase A works in double precision whatever R is, on coordinates relative to the tile's first node, with one
// fused expression per edge:  y_i = cross_i + 1e-10 = af_i * fast' + as_i * slow' + C_i.  That is not the reference's
// rounding sequence, and does not have to be: a node whose smallest y_i is at least T_r away from zero gets the same
// verdict from every evaluation that is accurate to T_r, the reference's included, and the nodes closer than T_r are
// exactly the ones handed to the fix-up pass.  T_r = the record's threshold (amb_threshold: the margin band plus
// the reference's own rounding) + 1e-10 (the shift) + 32 eps M Q, where M bounds |ex| + |ey| of the edges and Q the
// local coordinates that enter the expression (the error of the two FMAs and of C_i is below 10 eps M Q).
constexpr int kRasterChunk = 384;   // triangles of a tile processed per pass (a C4 tile lists ~210)
#ifndef GI_RASTER_MINB
#define GI_RASTER_MINB 8
#endif
constexpr int kRasterMinBlocks = GI_RASTER_MINB;
constexpr int kRasterBuckets = 64;  // size classes of the per-tile counting sort

template <class R, bool FY, bool PEERS>
__global__ void __launch_bounds__(kWalkThreads, kRasterMinBlocks)
raster_eval_kernel(const WalkArgs<R> a, const TileBins b, int tile_row0) {
  __shared__ __align__(16) int stri[kTileH][kWalkThreads];
  __shared__ R sfast[kTileW];
  __shared__ R saxis[kTileH];
  __shared__ double sfx[kTileW], sfy[kTileH];  // the same nodes relative to the tile's first node
  __shared__ unsigned s_item[kRasterChunk];    // c0 | c1 << 7 | r0 << 14 | r1 << 19 | size class << 24
  __shared__ int s_id[kRasterChunk];
  __shared__ unsigned short s_sorted[kRasterChunk];
  __shared__ int s_hist[kRasterBuckets];
  __shared__ int s_n;
  if (*b.bad) return;
  const TargetWindow<R>& w = a.win;
  const int tid = threadIdx.x;
  const int lane_f = blockIdx.x * kWalkThreads + tid;
  const int f = w.f0 + lane_f;
  const bool active = f < w.f1;
  const int srow0 = w.s0 + blockIdx.y * kTileH;
  const int nrows = min(kTileH, w.s1 - srow0);
  const int ncols = min(kTileW, w.nf() - blockIdx.x * kWalkThreads);
  const R cf = active ? __ldg(w.fast_axis + f) : Inf<R>::v();
  // first node of the tile along the fast (X0) and the slow (Y0) axis
  const double X0 = (double)__ldg(w.fast_axis + w.f0 + blockIdx.x * kWalkThreads), Y0 = (double)__ldg(w.slow_axis + srow0);
  sfast[tid] = cf;  // nodes beyond the window: +inf, never inside a bound
  sfx[tid] = (double)cf - X0;
  if (tid < kTileH) {
    const R sv = tid < nrows ? __ldg(w.slow_axis + srow0 + tid) : Inf<R>::v();
    saxis[tid] = sv;
    sfy[tid] = (double)sv - Y0;
  }
  for (int i = tid; i < kTileH * kWalkThreads / 4; i += kWalkThreads)
    reinterpret_cast<int4*>(&stri[0][0])[i] = make_int4(-1, -1, -1, -1);
  if (tid < kRasterBuckets) s_hist[tid] = 0;
  if (tid == 0) s_n = 0;
  __syncthreads();

  // ---- phase A: rasterise the tile's triangles ---------------------------------------------------
  const int tile = (tile_row0 + blockIdx.y) * b.ntx + blockIdx.x;
  const int beg = __ldg(b.start + tile), end = __ldg(b.start + tile + 1);
  const R f_first = sfast[0], s_first = saxis[0];
  const R f_span = sfast[ncols - 1] - f_first, s_span = saxis[nrows - 1] - s_first;
  const R inv_f = f_span > R(0) ? R(ncols - 1) / f_span : R(0), inv_s = s_span > R(0) ? R(nrows - 1) / s_span : R(0);
  const double wq = sfx[ncols - 1] + sfy[nrows - 1];
  const unsigned sfx_s = (unsigned)__cvta_generic_to_shared(sfx), sfy_s = (unsigned)__cvta_generic_to_shared(sfy);
  const unsigned stri_s = (unsigned)__cvta_generic_to_shared(&stri[0][0]);
  auto fax = [&](int i) { return sfast[i]; };
  auto sax = [&](int i) { return saxis[i]; };
  long long tests = 0;
  for (int chunk = beg; chunk < end; chunk += kRasterChunk) {
    const int cn = min(kRasterChunk, end - chunk);
    // A1: clip every triangle's bounds to the tile; the non-empty ones become work items with a size class
    for (int k = tid; k < cn; k += kWalkThreads) {
      const int id = __ldg(b.list + chunk + k);
      TriGeom<R> g;
      load_walk(a.geom + id, g);
      R flo, fhi, slo, shi;
      tri_bounds(g, FY, flo, fhi, slo, shi);
      const int c0 = bound_from<false>(fax, kTileW, flo, guess_index(flo, f_first, inv_f, kTileW));
      const int c1 = bound_from<true>(fax, kTileW, fhi, guess_index(fhi, f_first, inv_f, kTileW) + 1) - 1;
      const int r0 = bound_from<false>(sax, kTileH, slo, guess_index(slo, s_first, inv_s, kTileH));
      const int r1 = bound_from<true>(sax, kTileH, shi, guess_index(shi, s_first, inv_s, kTileH) + 1) - 1;
      if (c0 > c1 || r0 > r1) continue;
      const int n = (c1 - c0 + 1) * (r1 - r0 + 1);
      const int cls = kRasterBuckets - 1 - (n < 192 ? n >> 2 : min(48 + ((n - 192) >> 8), kRasterBuckets - 1));  // large first
      const int slot = atomicAdd(&s_n, 1);
      s_item[slot] = (unsigned)c0 | (unsigned)c1 << 7 | (unsigned)r0 << 14 | (unsigned)r1 << 19 | (unsigned)cls << 24;
      s_id[slot] = id;
      atomicAdd(&s_hist[cls], 1);
    }
    __syncthreads();
    // A2: counting sort of the items by size class, so that the lanes of a warp get boxes of similar size
    if (tid < 32) {
      const int h0 = s_hist[2 * tid], h1 = s_hist[2 * tid + 1];
      int inc = h0 + h1;
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (tid >= off) inc += t;
      }
      s_hist[2 * tid] = inc - h0 - h1;
      s_hist[2 * tid + 1] = inc - h1;
    }
    __syncthreads();
    const int n_items = s_n;
    for (int k = tid; k < n_items; k += kWalkThreads) s_sorted[atomicAdd(&s_hist[s_item[k] >> 24], 1)] = (unsigned short)k;
    __syncthreads();
    // A3: one thread per item, largest first, every other round in reverse (snake) to even out the warps
    for (int q = 0; q * kWalkThreads < n_items; ++q) {
      const int idx = q * kWalkThreads + ((q & 1) ? kWalkThreads - 1 - tid : tid);
      if (idx >= n_items) continue;
      const int slot = s_sorted[idx];
      const unsigned item = s_item[slot];
      const int id = s_id[slot];
      const int c0 = item & 127, c1 = (item >> 7) & 127, r0 = (item >> 14) & 31, r1 = (item >> 19) & 31;
      TriGeom<R> g;
      load_walk(a.geom + id, g);
      // vertices relative to the tile's first node (fast axis = y when FY)
      const double OX = FY ? Y0 : X0, OY = FY ? X0 : Y0;
      const double vx1 = (double)g.x1 - OX, vy1 = (double)g.y1 - OY;
      const double vx2 = (double)g.x2 - OX, vy2 = (double)g.y2 - OY;
      const double vx3 = (double)g.x3 - OX, vy3 = (double)g.y3 - OY;
      const double ex1 = vx2 - vx1, ey1 = vy2 - vy1;   // edge (v1,v2)
      const double ex2 = vx3 - vx2, ey2 = vy3 - vy2;   // edge (v2,v3)
      const double ex3 = vx1 - vx3, ey3 = vy1 - vy3;   // edge (v3,v1)
      // cross_i = ex_i*(ty - ya_i) - ey_i*(tx - xa_i)   (triangle_walk.f90:149-150)
      const double C1 = __fma_rn(ey1, vx1, __fma_rn(-ex1, vy1, 1e-10));
      const double C2 = __fma_rn(ey2, vx2, __fma_rn(-ex2, vy2, 1e-10));
      const double C3 = __fma_rn(ey3, vx3, __fma_rn(-ex3, vy3, 1e-10));
      const double af1 = FY ? ex1 : -ey1, af2 = FY ? ex2 : -ey2, af3 = FY ? ex3 : -ey3;   // factor of the fast coordinate
      const double as1 = FY ? -ey1 : ex1, as2 = FY ? -ey2 : ex2, as3 = FY ? -ey3 : ex3;   // factor of the slow coordinate
      // (sums instead of maxima: cheaper, and a larger threshold only sends a few more nodes to the fix-up)
      const double M = fabs(ex1) + fabs(ey1) + fabs(ex2) + fabs(ey2) + fabs(ex3) + fabs(ey3);
      const double Q = fabs(vx1) + fabs(vy1) + fabs(vx2) + fabs(vy2) + fabs(vx3) + fabs(vy3) + wq;
      const double t_rec = sizeof(R) == 8 ? __hiloint2double(g.pad & kThrMask, 0)
                                           : (double)__int_as_float(g.pad & kThrMask);
      const double Tr = t_rec + 1e-10 + 32.0 * 1.1102230246251565e-16 * M * Q;
      const int thi = Tr < 1e300 ? __double2hiint(Tr) + 1 : 0x7fe00000;  // |y| < Tr  =>  high word of |y| < thi
      const int sliver = (g.pad & kSliver) ? (int)0x80000000 : 0;  // a hull sliver keeps no node (sliver_guard)
      const int ncol = c1 - c0 + 1, n = ncol * (r1 - r0 + 1);
      tests += n;
      // the node loop, flattened over the box (lanes with boxes of different shapes stay together), on 32-bit
      // shared-memory addresses
      const unsigned fa0 = sfx_s + 8u * c0, fa_end = sfx_s + 8u * c1;
      unsigned fa = fa0, sa = sfy_s + 8u * r0, ta = stri_s + 4u * (r0 * kWalkThreads + c0);
      const unsigned ta_wrap = 4u * (kWalkThreads - ncol) + 4u;
      for (int i = 0; i < n; ++i) {
        const double fv = lds_f64(fa), sv = lds_f64(sa);
        const double y1 = __fma_rn(as1, sv, __fma_rn(af1, fv, C1));
        const double y2 = __fma_rn(as2, sv, __fma_rn(af2, fv, C2));
        const double y3 = __fma_rn(as3, sv, __fma_rn(af3, fv, C3));
        const int h1 = __double2hiint(y1), h2 = __double2hiint(y2), h3 = __double2hiint(y3);
        // every orientation test passes (:57-60; all three y_i >= 0, i.e. no sign bit): the node lies in this triangle
        if (((h1 | h2 | h3) | sliver) >= 0) sts_u32(ta, id);
        // some |y_i| below T_r: too close to an edge('s line) to call, the reference's own walk decides
        if (min(min(h1 & 0x7fffffff, h2 & 0x7fffffff), h3 & 0x7fffffff) < thi) {
          const int cell = (int)((ta - stri_s) >> 2);  // r * 128 + c
          const unsigned long long slot2 = atomicAdd((unsigned long long*)(a.counters + 5), 1ull);
          if (slot2 < (unsigned long long)a.amb_cap)
            a.amb_list[slot2] = (srow0 + (cell >> 7) - w.s0) * w.nf() + blockIdx.x * kWalkThreads + (cell & 127);
        }
        const bool wrap = fa == fa_end;
        fa = wrap ? fa0 : fa + 8u;
        sa += wrap ? 8u : 0u;
        ta += wrap ? ta_wrap : 4u;
      }
    }
    __syncthreads();
    if (tid < kRasterBuckets) s_hist[tid] = 0;
    if (tid == 0) s_n = 0;
    __syncthreads();
  }

  // ---- phase B: evaluate ---------------------------------------------------------------------------
  evaluate_strip<R, FY, PEERS, kTileH>(a, stri, saxis, srow0, nrows, lane_f, active, cf);
  for (int off = 16; off > 0; off >>= 1) tests += __shfl_down_sync(0xffffffffu, tests, off);
  if ((tid & 31) == 0 && tests) atomicAdd((unsigned long long*)a.counters, (unsigned long long)tests);
}

// Tile lists of window `w` (whole tiles from its origin): buffers for the pack pass to count into ...
template <class R>
int alloc_tile_bins(Scratch& scr, int num_tri, const TargetWindow<R>& w, TileBins* bins) {
  cudaStream_t st = scr.stream();
  TileBins b;
  b.ntx = div_up(w.nf(), kTileW);
  b.nty = div_up(w.ns(), kTileH);
  const int64_t ntiles = (int64_t)b.ntx * b.nty;
  const int64_t cap = 2 * (int64_t)num_tri + 4 * ntiles + 1024;
  GI_REQUIRE(cap < (int64_t)0x7fffffff && b.ntx < 65536 && b.nty < 65536, "tile lists too large");
  b.cap = (int)cap;
  GI_TRY(scr.alloc(&b.start, (size_t)ntiles + 1));
  GI_TRY(scr.alloc(&b.fill, (size_t)ntiles + 1));
  GI_TRY(scr.alloc(&b.list, (size_t)cap));
  GI_TRY(scr.alloc(&b.bad, 1));
  GI_TRY(scr.alloc(&b.range, (size_t)num_tri));
  GI_CUDA(cudaMemsetAsync(b.start, 0, sizeof(int) * ((size_t)ntiles + 1), st));
  GI_CUDA(cudaMemsetAsync(b.fill, 0, sizeof(int) * ((size_t)ntiles + 1), st));
  GI_CUDA(cudaMemsetAsync(b.bad, 0, sizeof(int), st));
  *bins = b;
  return GI_OK;
}
// ... and, after the pack pass: offsets, then the lists
inline int finish_tile_bins(Scratch& scr, int num_tri, const TileBins& b) {
  cudaStream_t st = scr.stream();
  GI_TRY(exclusive_scan_i32(scr, b.start, b.start, (int64_t)b.ntx * b.nty + 1));
  bin_fill_kernel<<<div_up(num_tri, 256), 256, 0, st>>>(num_tri, b);
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}

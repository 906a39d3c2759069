// barycentric.cu -- triangle mesh -> regular grid: visibility walk + barycentric weights + hull fill.
//
// Replaces (reference file:line):
//   interpolation.f90:298-416   barycentric_interpolation (operator)
//   triangle_walk.f90:24-84     find_triangle (visibility walk)
//   triangle_walk.f90:140-153   triangle_point_outside (orientation predicate, margin -1e-10)
//   triangle_walk.f90:90-132    fill_nearest_neighbour (nearest unmasked cell in index space)
//
// B200 design (not a port):
//   1. pack_mesh_kernel    gathers every triangle into ONE aligned 96 B record: sectors 0-1 hold the 3 vertices +
//      3 neighbour ids (a walk step is two 256-bit loads instead of 3 index loads + 6 scattered coordinate
//      loads), sector 2 the 3 nodal values.  The same pass bins one triangle id per 16x16-target cell
//      (atomicMin -> deterministic) into a seed grid.
//   2. locate_eval_kernel  a CTA owns 128 columns x 32 rows of targets.  Phase 1 (locate): one thread per
//      column walks down the rows; the triangle found for row r seeds row r+1 (the reference chains along a
//      grid row in the same way, interpolation.f90:366-373, but restarts every row from one global start
//      triangle; here the first row of a strip is seeded from the seed grid).  The found triangle is
//      path-independent outside the 1e-10 margin band, so the seeding changes the iteration
//      count, not the result.  Lanes are decoupled (each runs its own test / move loop) and keep the
//      (triangle, column)-invariant half of every orientation product in registers, so a target in the same
//      triangle as the one above it costs no memory access.  Phase 2 (evaluate): row-synchronous; the record
//      (L1-resident after phase 1) is re-read only when the column enters another triangle, the two
//      quotients of interpolation.f90:387-394 share one refined reciprocal, 128 consecutive doubles per row
//      are stored coalesced; outside-hull flags go into a 1-bit-per-target mask (one ballot per warp row)
//      and the non-zero mask words into a compact work list.
//   3. fill_kernel         one warp per non-zero mask word, one lane per masked cell: exact ring search over
//      the bit mask with the reference's tie-break (smallest squared index distance, then smallest row,
//      then smallest column).
//
// All index decisions (orientation signs, first failing edge) and all values use the reference's operations
// in the reference's order with FMA contraction disabled (-fmad=false): bit-identical to the oracle.
#include <algorithm>
#include <atomic>
#include <memory>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ieee_div.cuh"
#include "kernels.cuh"
#include "scan.cuh"

namespace gi {

namespace {

constexpr int kSeedCell = 16;      // targets per seed-grid cell edge
constexpr int kWalkThreads = 128;  // columns per CTA

// tile lists of the rasteriser (raster.cuh); counted by the pack pass
constexpr int kTileW = kWalkThreads;  // fast-axis nodes per tile
constexpr int kTileH = 32;            // slow-axis nodes per tile

struct TileBins {
  int* start;  // [ntx*nty + 1] exclusive offsets into list
  int* fill;   // [ntx*nty] fill cursors (zeroed between the passes)
  int* list;   // [cap] triangle ids, tile by tile
  int* bad;    // device flag: != 0 -> axes not ascending or list overflow: the walk kernel takes over
  uint2* range;  // [num_tri] tile range of every triangle {tx0 | tx1 << 16, ty0 | ty1 << 16}, written by the pack pass
  int ntx, nty, cap;
};
template <class R> struct TriGeom;
template <class R>
__device__ __forceinline__ bool tri_tile_range(const TriGeom<R>& g, const TargetWindow<R>& w, int& tx0, int& tx1, int& ty0, int& ty1);
template <class R> __device__ __forceinline__ void check_axes(const TargetWindow<R>& w, int t, int* bad);
template <class R> __device__ __forceinline__ void plane_of(TriGeom<R>& g);



// One record per triangle, f64: 96 B = 3 sectors.
//   sectors 0-1  what a walk step needs: 3 vertices + 3 neighbour ids + the ambiguity threshold / flags word
//   sector  2    denom = (y2-y3)*(x1-x3) + (x3-x2)*(y1-y3) (interpolation.f90:385-386) and the nodal values d1, d2, d3
// Measured alternatives (same box, profiles/r02_walk_variants.md): round 1's record without denom (phase 2 re-derives
// it: 4 FP64 more per triangle change) and a 128 B record that also holds A = y2-y3 .. D = x1-x3 (8 FP64 fewer per
// triangle change, but a fourth sector per triangle to fetch and one more load to warm it: the kernel is bound by
// latency and L1 wavefronts, not by FP64 -- 2.99 ms against 2.84 ms, and the pack pass writes 1.28 GB instead of 0.96).
template <class R> struct alignas(sizeof(R) == 8 ? 32 : 16) TriGeom {
  R x1, y1, x2, y2, x3, y3;
  int n1, n2, n3;  // neighbour opposite vertex 1/2/3, 0-based, -1 = hull edge, -2 = not packed (band-limited pack)
  int pad;         // bits 2-30: ambiguity threshold of this triangle (bit pattern, see amb_threshold); bit 31:
                   // long triangle (kLongTriangle); bit 0: hull sliver, not to be interpolated in (kSliver);
                   // bit 1: sector 2 holds the triangle's plane instead (kPlane, GI_BARY_PLANE_VALUES):
                   // denom = gx, d1 = gy, d2 = d3 of the text, d3 unused
  R denom, d1, d2, d3;
};
static_assert(sizeof(TriGeom<double>) == 96 && sizeof(TriGeom<float>) == 64, "record layout");
template <class R> struct SeedGrid {
  R x0, y0, inv_cell_x, inv_cell_y;
  int ncx, ncy;
};

// Record loads.  The kernels below are bound by L1 data-pipe wavefronts (ncu: l1tex__data_pipe_lsu_wavefronts
// 75 % of peak with 128-bit loads): every lane fetches its own record, so one warp-wide load costs one wavefront
// per distinct cache line.  sm_100 has 256-bit global loads (LDG.E.ENL2.256): the walk part of a record is 2 loads
// instead of 4, the evaluation part 3 (x3 / y3 come with sector 1).
__device__ __forceinline__ void ld256(const void* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// walk part (sectors 0-1) of a record
__device__ __forceinline__ void load_walk(const TriGeom<double>* p, TriGeom<double>& g) {
  double n12, w;
  ld256(p, g.x1, g.y1, g.x2, g.y2);
  ld256(reinterpret_cast<const char*>(p) + 32, g.x3, g.y3, n12, w);
  g.n1 = __double2loint(n12); g.n2 = __double2hiint(n12);
  g.n3 = __double2loint(w); g.pad = __double2hiint(w);
}
__device__ __forceinline__ void load_walk(const TriGeom<float>* p, TriGeom<float>& g) {
  g.x1 = p->x1; g.y1 = p->y1; g.x2 = p->x2; g.y2 = p->y2; g.x3 = p->x3; g.y3 = p->y3;
  g.n1 = p->n1; g.n2 = p->n2; g.n3 = p->n3; g.pad = p->pad;
}
// evaluation part: all three sectors (A..D of interpolation.f90:385-394 are differences of the vertices)
__device__ __forceinline__ void load_eval(const TriGeom<double>* p, TriGeom<double>& g) {
  load_walk(p, g);
  ld256(reinterpret_cast<const char*>(p) + 64, g.denom, g.d1, g.d2, g.d3);
}
__device__ __forceinline__ void load_eval(const TriGeom<float>* p, TriGeom<float>& g) { g = *p; }
// plane evaluation (GI_BARY_PLANE_VALUES): sectors 1 and 2 only
__device__ __forceinline__ void load_plane(const TriGeom<double>* p, double& x3, double& y3, int& pad, double& s0,
                                           double& s1, double& s2, double& s3) {
  double n12, w;
  ld256(reinterpret_cast<const char*>(p) + 32, x3, y3, n12, w);
  ld256(reinterpret_cast<const char*>(p) + 64, s0, s1, s2, s3);
  pad = __double2hiint(w);
}
__device__ __forceinline__ void load_plane(const TriGeom<float>* p, float& x3, float& y3, int& pad, float& s0, float& s1,
                                           float& s2, float& s3) {
  x3 = p->x3; y3 = p->y3; pad = p->pad; s0 = p->denom; s1 = p->d1; s2 = p->d2; s3 = p->d3;
}
// new nodal values (plans): sector 2 as ONE full-sector store; denom comes from the plan's own copy (reading it
// back from the record is a 32 B read-modify-write per 96 B record: 0.33 ms instead of 0.2 ms at C4)
__device__ __forceinline__ void store_values(TriGeom<double>* p, double denom, double d1, double d2, double d3) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(reinterpret_cast<char*>(p) + 64), "d"(denom), "d"(d1), "d"(d2),
               "d"(d3)
               : "memory");
}
__device__ __forceinline__ void store_values(TriGeom<float>* p, float denom, float d1, float d2, float d3) {
  p->denom = denom; p->d1 = d1; p->d2 = d2; p->d3 = d3;
}
// a whole record, written once and not read again by the writing pass: three full-sector stores, evict_first in L2
__device__ __forceinline__ void store_record(TriGeom<double>* p, const TriGeom<double>& g, unsigned long long pol) {
  const double n12 = __hiloint2double(g.n2, g.n1), w = __hiloint2double(g.pad, g.n3);
  asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(g.x1), "d"(g.y1), "d"(g.x2),
               "d"(g.y2), "l"(pol) : "memory");
  asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(reinterpret_cast<char*>(p) + 32),
               "d"(g.x3), "d"(g.y3), "d"(n12), "d"(w), "l"(pol) : "memory");
  asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(reinterpret_cast<char*>(p) + 64),
               "d"(g.denom), "d"(g.d1), "d"(g.d2), "d"(g.d3), "l"(pol) : "memory");
}
__device__ __forceinline__ void store_record(TriGeom<float>* p, const TriGeom<float>& g, unsigned long long) { *p = g; }
// Fire-and-forget load of sector 2: the registers are never read, so nothing waits on them, but the sector
// is in L1 when phase 2 asks for it a few thousand cycles later (prefetch.global.L1 had no measurable effect).
__device__ __forceinline__ void touch_values(const TriGeom<double>* p) {
  double a, b, c, d;
  ld256(reinterpret_cast<const char*>(p) + 64, a, b, c, d);
}
__device__ __forceinline__ void touch_values(const TriGeom<float>*) {}

// Warm the walk part (sectors 0-1) of a record the walk will probably move to: two one-word loads whose results are
// never read, so no register waits on them.  (A/B on C4, same box: this form 2.844 -> 2.800 ms; prefetch.global.L1
// 2.996 ms; cp.async.ca into a dummy shared slot 3.071 ms -- profiles/r02_walk_variants.md.)  GI_WALK_PREFETCH=0
// compiles the prediction out.
#ifndef GI_WALK_PREFETCH
#define GI_WALK_PREFETCH 1
#endif
__device__ __forceinline__ void touch_walk(const TriGeom<double>* p) {
  const char* c = reinterpret_cast<const char*>(p);
  int u, v;
  asm volatile("ld.global.b32 %0, [%1];" : "=r"(u) : "l"(c));
  asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(c + 32));
}
__device__ __forceinline__ void touch_walk(const TriGeom<float>*) {}

// seed-grid geometry from the axes (a uniform-axis approximation: seeds only have to be near)
template <class R>
__device__ __forceinline__ SeedGrid<R> seed_grid_of(const R* __restrict__ x_axis, int len_x,
                                                    const R* __restrict__ y_axis, int len_y) {
  SeedGrid<R> s;
  s.ncx = (len_x + kSeedCell - 1) / kSeedCell;
  s.ncy = (len_y + kSeedCell - 1) / kSeedCell;
  s.x0 = __ldg(x_axis);
  s.y0 = __ldg(y_axis);
  const R ex = __ldg(x_axis + len_x - 1) - s.x0, ey = __ldg(y_axis + len_y - 1) - s.y0;
  s.inv_cell_x = (len_x > 1 && ex > R(0)) ? R(len_x - 1) / ex / R(kSeedCell) : R(0);
  s.inv_cell_y = (len_y > 1 && ey > R(0)) ? R(len_y - 1) / ey / R(kSeedCell) : R(0);
  return s;
}

// Row-band calls (multi-GPU sharding) need only the triangles near their band.  need_kernel marks the triangles
// whose y-range meets [y_lo, y_hi] (the walked rows +- a margin): 3 y gathers per triangle from an array that stays
// L2-resident, no record traffic.  pack_mesh_kernel then skips the others, and a link from a packed triangle to a
// skipped one becomes kNotPacked: a walk that wants to cross it raises GI_ERR_PACK_WINDOW instead of reading a
// record that was never written (the caller retries with the whole mesh packed).
constexpr int kNotPacked = -2;
constexpr int kLongTriangle = (int)0x80000000;  // sign bit of TriGeom::pad, see farthest_failing_edge
constexpr int kSliver = 1;                      // bit 0 of TriGeom::pad, see sliver_guard
constexpr int kPlane = 2;                       // bit 1 of TriGeom::pad, see plane_of
constexpr int kThrMask = 0x7ffffffc;            // the threshold bits of TriGeom::pad
template <class R>
__global__ void __launch_bounds__(256)
need_kernel(PointsView<R> pts, TopoView simp, int num_points, int num_tri, const R* __restrict__ y_axis, int row_lo,
            int row_hi, int len_y, uint8_t* __restrict__ need) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tri) return;
  // y-range of the packed rows; open towards a grid border so that walks to targets of the first / last rows,
  // which may lie outside the hull, never meet a skipped triangle on that side
  const R ya0 = __ldg(y_axis + row_lo), ya1 = __ldg(y_axis + row_hi);
  const bool up = ya1 >= ya0;
  const bool open_lo = up ? row_lo == 0 : row_hi == len_y - 1, open_hi = up ? row_hi == len_y - 1 : row_lo == 0;
  const R y_lo = open_lo ? -Consts<R>::huge() : fmin(ya0, ya1), y_hi = open_hi ? Consts<R>::huge() : fmax(ya0, ya1);
  const int a = simp(t, 0), b = simp(t, 1), c = simp(t, 2);
  if ((unsigned)(a - 1) >= (unsigned)num_points || (unsigned)(b - 1) >= (unsigned)num_points ||
      (unsigned)(c - 1) >= (unsigned)num_points) {
    need[t] = 1;  // pack_mesh_kernel reports the bad id
    return;
  }
  const R ya = pts.y(a - 1), yb = pts.y(b - 1), yc = pts.y(c - 1);
  need[t] = (fmax(fmax(ya, yb), yc) >= y_lo && fmin(fmin(ya, yb), yc) <= y_hi) ? 1 : 0;
}

// Sliver guard (the reference's README to-do: "prohibiting barycentric interpolation for sliver triangles close to
// the edge (and filling the subsequent missing values from the nearest neighbours)", README.md:41-43).  Opt-in
// (gi_set_sliver_guard): a triangle with a hull edge whose quality
//     q = |denom| / max |edge|^2        (denom = twice the signed area, interpolation.f90:385-386; equilateral: 0.87)
// is below the threshold keeps its place in the mesh -- walks pass through it as before -- but a target found in
// it is treated like a target outside the hull: masked, then filled by fill_nearest_neighbour.  The oracle has the
// same switch (orc_set_sliver_guard) with the same expression.
template <class R> __device__ __forceinline__ R sliver_quality(const TriGeom<R>& g) {
  const R l1 = (g.x2 - g.x1) * (g.x2 - g.x1) + (g.y2 - g.y1) * (g.y2 - g.y1);
  const R l2 = (g.x3 - g.x2) * (g.x3 - g.x2) + (g.y3 - g.y2) * (g.y3 - g.y2);
  const R l3 = (g.x1 - g.x3) * (g.x1 - g.x3) + (g.y1 - g.y3) * (g.y1 - g.y3);
  return fabs(g.denom) / fmax(fmax(l1, l2), l3);
}

template <class R, bool BINS>
__global__ void __launch_bounds__(256)
pack_mesh_kernel(PointsView<R> pts, const R* __restrict__ data, TopoView simp, TopoView nbr, int num_points,
                 int num_tri, TriGeom<R>* __restrict__ geom, int* __restrict__ seed,
                 const R* __restrict__ x_axis, int len_x, const R* __restrict__ y_axis, int len_y,
                 const uint8_t* __restrict__ need, const TargetWindow<R> bw, const TileBins bb, R sliver_q,
                 int plane, int* __restrict__ status) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool packed = t < num_tri && (!need || need[t]);
  if (BINS) {  // rasteriser (raster.cuh): axes usable?  A triangle without a record is in no tile.
    check_axes(bw, t, bb.bad);
    if (t < num_tri && !packed) bb.range[t] = make_uint2(1u, 1u);  // tx0 = 1 > tx1 = 0: empty
  }
  // fallback seed (seed[ncx*ncy]): the smallest packed triangle id.  CTAs start roughly in id order, so after the
  // first few only a plain load is paid (39 000 atomics on one address cost 0.5 ms)
  const unsigned pm = __ballot_sync(0xffffffffu, packed);
  if (pm && (threadIdx.x & 31) == __ffs(pm) - 1) {
    int* fb = seed + ((len_x + kSeedCell - 1) / kSeedCell) * ((len_y + kSeedCell - 1) / kSeedCell);
    if (t < *reinterpret_cast<volatile int*>(fb)) atomicMin(fb, t);
  }
  if (!packed) return;
  const SeedGrid<R> sg = seed_grid_of(x_axis, len_x, y_axis, len_y);
  // L2 residency (common.cuh): ids and records stream through once (evict_first), vertices and nodal values are
  // gathered three to six times each from arrays that fit the L2 (evict_last)
  const unsigned long long keep = l2_evict_last(), once = l2_evict_first();
  int a = simp(t, 0, once), b = simp(t, 1, once), c = simp(t, 2, once);
  int n1 = nbr(t, 0, once), n2 = nbr(t, 1, once), n3 = nbr(t, 2, once);
  const bool bad = (unsigned)(a - 1) >= (unsigned)num_points || (unsigned)(b - 1) >= (unsigned)num_points ||
                   (unsigned)(c - 1) >= (unsigned)num_points || (unsigned)n1 > (unsigned)num_tri ||
                   (unsigned)n2 > (unsigned)num_tri || (unsigned)n3 > (unsigned)num_tri;
  if (bad) {  // reference: out-of-bounds read.  Here: error status, record made harmless.
    raise_status(status, GI_ERR_BAD_INDEX);
    a = b = c = 1;
    n1 = n2 = n3 = 0;
  }
  TriGeom<R> g;
  pts.xy(a - 1, g.x1, g.y1, keep);
  pts.xy(b - 1, g.x2, g.y2, keep);
  pts.xy(c - 1, g.x3, g.y3, keep);
  g.n1 = n1 - 1; g.n2 = n2 - 1; g.n3 = n3 - 1;
  if (need) {
    if (n1 > 0 && !need[n1 - 1]) g.n1 = kNotPacked;
    if (n2 > 0 && !need[n2 - 1]) g.n2 = kNotPacked;
    if (n3 > 0 && !need[n3 - 1]) g.n3 = kNotPacked;
  }
  g.pad = amb_threshold(g);
  {  // long triangle (an edge spans more than two seed cells): careful exit-edge choice, see farthest_failing_edge
    const R lx = fmax(fmax(fabs(g.x2 - g.x1), fabs(g.x3 - g.x2)), fabs(g.x1 - g.x3)) * sg.inv_cell_x;
    const R ly = fmax(fmax(fabs(g.y2 - g.y1), fabs(g.y3 - g.y2)), fabs(g.y1 - g.y3)) * sg.inv_cell_y;
    if (fmax(lx, ly) > R(2)) g.pad |= kLongTriangle;
  }
  {
    const R A = g.y2 - g.y3, B = g.x3 - g.x2, D = g.x1 - g.x3;                     // interpolation.f90:385-394
    g.denom = A * D + B * (g.y1 - g.y3);                                           // :385-386
  }
  if (sliver_q > R(0) && (n1 == 0 || n2 == 0 || n3 == 0) && sliver_quality(g) < sliver_q) g.pad |= kSliver;
  g.d1 = ldg_hint(data + (a - 1), keep); g.d2 = ldg_hint(data + (b - 1), keep); g.d3 = ldg_hint(data + (c - 1), keep);
  if (plane) plane_of(g);
  store_record(geom + t, g, once);
  if (BINS) {  // first binning pass: count this triangle in the tiles its bounds meet
    int tx0 = 1, tx1 = 0, ty0 = 1, ty1 = 0;
    if (!bad && tri_tile_range(g, bw, tx0, tx1, ty0, ty1))
      for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(bb.start + ty * bb.ntx + tx, 1);
    bb.range[t] = make_uint2((unsigned)tx0 | (unsigned)tx1 << 16, (unsigned)ty0 | (unsigned)ty1 << 16);
  }
  if (bad) return;
  // seed grid: cell of the centroid under a uniform-axis approximation (a heuristic: it only has to be near)
  const R cx = (g.x1 + g.x2 + g.x3) / R(3), cy = (g.y1 + g.y2 + g.y3) / R(3);
  const R fx = (cx - sg.x0) * sg.inv_cell_x, fy = (cy - sg.y0) * sg.inv_cell_y;
  if (fx >= R(0) && fy >= R(0) && fx < R(sg.ncx) && fy < R(sg.ncy))
    atomicMin(seed + (int)fy * sg.ncx + (int)fx, t);
}

// plans: new nodal values into sector 2 of the existing records (ids were validated by pack_mesh_kernel).
// plane != 0 (GI_BARY_PLANE_VALUES): the record's vertices are read back and the plane recomputed where the record
// is a plane record (the choice is geometric, see plane_of).
template <class R>
__global__ void __launch_bounds__(256)
pack_values_kernel(const R* __restrict__ data, TopoView simp, int num_points, int num_tri, const R* __restrict__ denom,
                   TriGeom<R>* __restrict__ geom, int plane) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tri) return;
  int a = simp(t, 0), b = simp(t, 1), c = simp(t, 2);
  if ((unsigned)(a - 1) >= (unsigned)num_points || (unsigned)(b - 1) >= (unsigned)num_points ||
      (unsigned)(c - 1) >= (unsigned)num_points)
    a = b = c = 1;
  const R d1 = __ldg(data + (a - 1)), d2 = __ldg(data + (b - 1)), d3 = __ldg(data + (c - 1));
  if (plane) {
    TriGeom<R> g;
    load_walk(geom + t, g);
    if (g.pad & kPlane) {
      const R A = g.y2 - g.y3, B = g.x3 - g.x2, D = g.x1 - g.x3;
      g.denom = A * D + B * (g.y1 - g.y3);
      g.d1 = d1; g.d2 = d2; g.d3 = d3;
      g.pad &= ~kPlane;
      plane_of(g);   // same geometry, same answer: a plane record again
      store_values(geom + t, g.denom, g.d1, g.d2, g.d3);
      return;
    }
  }
  store_values(geom + t, __ldg(denom + t), d1, d2, d3);
}

// plans: the records' denominators as an array of their own (see store_values)
template <class R>
__global__ void __launch_bounds__(256) copy_denom_kernel(const TriGeom<R>* __restrict__ geom, int num_tri, R* __restrict__ denom) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < num_tri) denom[t] = geom[t].denom;
}

// Seed triangle for grid target (ix, iy): the cell's own entry, else the nearest non-empty cell within 3
// rings, else triangle 0.  Any valid triangle is a correct seed.
__device__ __forceinline__ int seed_lookup(const int* __restrict__ seed, int ncx, int ncy, int ix, int iy,
                                           int num_tri) {
  const int cx = ix / kSeedCell, cy = iy / kSeedCell;
  int s = __ldg(seed + cy * ncx + cx);
  if (s < num_tri) return s;
  for (int r = 1; r <= 3; ++r)
    for (int dy = -r; dy <= r; ++dy)
      for (int dx = -r; dx <= r; ++dx) {
        if (max(abs(dx), abs(dy)) != r) continue;
        const int x = cx + dx, y = cy + dy;
        if (x < 0 || y < 0 || x >= ncx || y >= ncy) continue;
        s = __ldg(seed + y * ncx + x);
        if (s < num_tri) return s;
      }
  s = __ldg(seed + ncx * ncy);  // the smallest packed triangle id
  return s < num_tri ? s : 0;
}

// Which failing edge to leave through.  The reference takes the FIRST failing edge (triangle_walk.f90:57-60), and so
// does this walk -- except in LONG triangles (pad sign bit, set by the pack pass: an edge spans more than two seed
// cells).  Those are the flat slivers that tile the strip along the convex hull (a hull edge of a uniform cloud is
// hundreds of grid cells long); there the first-failing rule slides sideways from sliver to sliver: on C4 the
// targets of grid rows 1..8 needed up to 1 220 dependent moves (mean 10.6; interior rows: mean 4.0, max 14), a
// serial chain of ~0.6 ms in a kernel that otherwise takes 0.45 ms per 2048-row band (profiles/r02_band_*).
// Leaving through the failing edge the target is FARTHEST beyond (violation / max-norm length of the edge) brings
// that to max 15, mean 3.7 moves.  The choice only changes the path: the triangle a walk ends in does not depend
// on it (targets inside the margin band are re-located along the reference's own path by
// fixup_kernel), and a target that fails a hull edge is outside the hull whichever edge reports it.
template <class R>
__device__ __forceinline__ int farthest_failing_edge(const TriGeom<R>& g, R c1, R c2, R c3, R margin) {
  const R l1 = fmax(fabs(g.x2 - g.x1), fabs(g.y2 - g.y1));
  const R l2 = fmax(fabs(g.x3 - g.x2), fabs(g.y3 - g.y2));
  const R l3 = fmax(fabs(g.x1 - g.x3), fabs(g.y1 - g.y3));
  // normalised violation of the failing edges (negative; passing edges: +huge).  Only compared with each other.
  const R v1 = c1 < margin ? c1 / l1 : Consts<R>::huge();
  const R v2 = c2 < margin ? c2 / l2 : Consts<R>::huge();
  const R v3 = c3 < margin ? c3 / l3 : Consts<R>::huge();
  // edge 1 -> n3, edge 2 -> n1, edge 3 -> n2 (triangle_walk.f90:66-71); ties / NaN: the reference's order
  if (v1 <= v2 && v1 <= v3) return g.n3;
  if (v2 <= v3) return g.n1;
  return g.n2;
}

// Band-limited pack: the first failing edge leads to a triangle that was not packed.  Any failing edge is a valid
// move of a visibility walk and the triangle found does not depend on the path, so take a LATER failing
// edge whose neighbour is packed if there is one; otherwise the caller reports GI_ERR_PACK_WINDOW.  (A failing hull
// edge that comes FIRST still ends the walk as in the reference; this only replaces a move that cannot be made.)
__device__ __forceinline__ int packed_alternative(bool o1, bool o2, bool o3, int n1, int n2, int n3) {
  // edge 1 -> n3, edge 2 -> n1, edge 3 -> n2 (triangle_walk.f90:66-71)
  if (o1 && n3 >= 0) return n3;
  if (o2 && n1 >= 0) return n1;
  if (o3 && n2 >= 0) return n2;
  return kNotPacked;
}

// triangle_walk.f90:24-84 + :140-153.  Returns the triangle the walk ends in and leaves its record in g;
// inside=false if the walk left through a hull edge (tri is then the last valid triangle, the reference's
// ind_tri_pre).
template <class R>
__device__ __forceinline__ int walk(const TriGeom<R>* __restrict__ geom, int tri, R tx, R ty, int iter_cap,
                                    bool& inside, int& iters, TriGeom<R>& g, int* status) {
  const R margin = Consts<R>::margin();
  inside = true;
  int it = 0;
  for (;;) {
    ++it;
    g = geom[tri];
    // edges (v1,v2), (v2,v3), (v3,v1); cross = (x2-x1)*(py-y1) - (y2-y1)*(px-x1)   (:149-150)
    const R c1 = (g.x2 - g.x1) * (ty - g.y1) - (g.y2 - g.y1) * (tx - g.x1);
    const R c2 = (g.x3 - g.x2) * (ty - g.y2) - (g.y3 - g.y2) * (tx - g.x2);
    const R c3 = (g.x1 - g.x3) * (ty - g.y3) - (g.y1 - g.y3) * (tx - g.x3);
    int nxt;
    // first failing edge wins (:57-60); edge i -> neighbour opposite vertex i-1 (wrap) (:66-71)
    if (c1 < margin) nxt = g.n3;
    else if (c2 < margin) nxt = g.n1;
    else if (c3 < margin) nxt = g.n2;
    else break;
    if (g.pad < 0 && (int)(c1 < margin) + (int)(c2 < margin) + (int)(c3 < margin) > 1)
      nxt = farthest_failing_edge(g, c1, c2, c3, margin);
    if (nxt == kNotPacked) nxt = packed_alternative(c1 < margin, c2 < margin, c3 < margin, g.n1, g.n2, g.n3);
    if (nxt < 0) {  // :74-80
      if (nxt == kNotPacked) raise_status(status, GI_ERR_PACK_WINDOW);
      inside = false;
      break;
    }
    tri = nxt;
    if (it >= iter_cap) { raise_status(status, GI_ERR_ITER_CAP); inside = false; break; }
  }
  iters = it;
  return tri;
}

// The rest of one target's walk with the farthest-failing-edge rule in EVERY triangle, out of line: phase 1 of
// locate_eval_kernel calls it when its current triangle is a long one and more than one edge fails, then re-enters
// the triangle returned (>= 0: every edge passes there; ~id: left the hull / the packed set from triangle id).
// Everything the caller keeps per triangle is reloaded after the call, so almost nothing is live across it.
template <class R>
__device__ __noinline__ int careful_locate(const TriGeom<R>* __restrict__ geom, int tri, R tx, R ty, int iter_cap,
                                           int* status) {
  const R margin = Consts<R>::margin();
  for (int it = 0;; ++it) {
    TriGeom<R> g;
    load_walk(geom + tri, g);
    const R c1 = (g.x2 - g.x1) * (ty - g.y1) - (g.y2 - g.y1) * (tx - g.x1);
    const R c2 = (g.x3 - g.x2) * (ty - g.y2) - (g.y3 - g.y2) * (tx - g.x2);
    const R c3 = (g.x1 - g.x3) * (ty - g.y3) - (g.y1 - g.y3) * (tx - g.x3);
    if (!(c1 < margin) && !(c2 < margin) && !(c3 < margin)) return tri;
    int nxt = farthest_failing_edge(g, c1, c2, c3, margin);
    if (nxt == kNotPacked) nxt = packed_alternative(c1 < margin, c2 < margin, c3 < margin, g.n1, g.n2, g.n3);
    if (nxt < 0) {
      if (nxt == kNotPacked && status) raise_status(status, GI_ERR_PACK_WINDOW);
      return ~tri;
    }
    if (it >= iter_cap) { if (status) raise_status(status, GI_ERR_ITER_CAP); return ~tri; }
    tri = nxt;
  }
}

// interpolation.f90:385-398, same operation order; the two divisions share the reciprocal of denom.
template <class R>
__device__ __forceinline__ R eval_value(const TriGeom<R>& g, R tx, R ty) {
  DivRcp<R> dv;
  dv.set(g.denom);
  const R u = tx - g.x3, t = ty - g.y3;
  const R A = g.y2 - g.y3, B = g.x3 - g.x2, C = g.y3 - g.y1, D = g.x1 - g.x3;
  const R w1 = dv.quotient(A * u + B * t);
  const R w2 = dv.quotient(C * u + D * t);
  const R w3 = R(1) - w1 - w2;
  return g.d1 * w1 + g.d2 * w2 + g.d3 * w3;
}

// GI_BARY_PLANE_VALUES (opt-in): the value as the triangle's plane through its three nodal values,
//     v = d3 + gx*(tx - x3) + gy*(ty - y3),   gx = ((d1-d3)*A + (d2-d3)*C)/denom,  gy = ((d1-d3)*B + (d2-d3)*D)/denom
// (A..D, denom of interpolation.f90:385-394; the same affine function the reference evaluates through its weights),
// computed as fma(gy, ty - y3, d3 + gx*(tx - x3)) by every path that produces a value, so that bands, fix-ups and the
// hull fill agree bit for bit among themselves.  Against the reference's expression the two roundings differ by at
// most ~16 eps * cond * max|d| with cond = max|edge|^2 / |denom|; triangles with cond > kPlaneCond (needles: a few in
// 10^5 of a random Delaunay mesh, and the slivers along the hull) keep the exact record and the reference's
// expression, which bounds the difference by ~1e-12 * max|d| in double precision (north_star's tolerance; tested in
// tests/test_gpu_plane.py).  What it saves: 18 FP64 operations + the division guards per target become 2.
constexpr int kPlaneCond = 128;
template <class R> __device__ __forceinline__ void plane_of(TriGeom<R>& g) {
  const R A = g.y2 - g.y3, B = g.x3 - g.x2, C = g.y3 - g.y1, D = g.x1 - g.x3;
  const R l1 = (g.x2 - g.x1) * (g.x2 - g.x1) + (g.y2 - g.y1) * (g.y2 - g.y1);
  const R l2 = B * B + A * A;
  const R l3 = D * D + C * C;
  const R lmax = fmax(fmax(l1, l2), l3);
  // (written so that a NaN or zero denom, or an overflowing edge length, fails the test)
  if (!(fabs(g.denom) * R(kPlaneCond) >= lmax) || !(lmax < Consts<R>::huge()) || !(fabs(g.denom) > R(0))) return;
  const R e1 = g.d1 - g.d3, e2 = g.d2 - g.d3;
  const R gx = (e1 * A + e2 * C) / g.denom, gy = (e1 * B + e2 * D) / g.denom;
  g.denom = gx; g.d1 = gy; g.d2 = g.d3; g.d3 = R(0);
  g.pad |= kPlane;
}
template <class R> __device__ __forceinline__ R eval_plane(R gx, R gy, R d3, R x3, R y3, R tx, R ty) {
  return fma(gy, ty - y3, d3 + gx * (tx - x3));
}
// the value of a target in the triangle of record g, whichever form the record has
template <class R> __device__ __forceinline__ R eval_any(const TriGeom<R>& g, R tx, R ty) {
  if (g.pad & kPlane) return eval_plane(g.denom, g.d1, g.d2, g.x3, g.y3, tx, ty);
  return eval_value(g, tx, ty);
}
// out of line: the rare exact record met by the plane instantiation of evaluate_strip
template <class R> __device__ __noinline__ R eval_exact_record(const TriGeom<R>* p, R tx, R ty) {
  TriGeom<R> g;
  load_eval(p, g);
  return eval_value(g, tx, ty);
}

// When is a located target path-dependent?  A target accepted by this triangle (every cross >= -1e-10,
// triangle_walk.f90:148-151) is ALSO accepted by the neighbour behind edge i iff that neighbour's own evaluation of
// the shared edge, cross' = -cross + delta, is >= -1e-10, i.e. iff cross <= 1e-10 + delta.  delta is the rounding
// difference of the two evaluations (different base vertex): each is a difference of two products of coordinate
// differences, so |delta| <= ~12 eps D^2 with D the longest edge (the target is within D of every vertex).  With
// double precision and O(1) triangles that is far below the margin; in single precision, or with coordinates in
// the millions, it is the larger term.  A cross below  T = 2e-10 + 32 eps D^2  therefore marks the target for the
// fix-up (a superset of the truly ambiguous cases); T depends on the triangle only and travels in the record as
// the bit pattern the hot loop compares against: the high word of T for double (rounded up), the word for float.
template <class R> __device__ __forceinline__ int amb_threshold(const TriGeom<R>& g);
template <> __device__ __forceinline__ int amb_threshold<double>(const TriGeom<double>& g) {
  const double a = (g.x2 - g.x1) * (g.x2 - g.x1) + (g.y2 - g.y1) * (g.y2 - g.y1);
  const double b = (g.x3 - g.x2) * (g.x3 - g.x2) + (g.y3 - g.y2) * (g.y3 - g.y2);
  const double c = (g.x1 - g.x3) * (g.x1 - g.x3) + (g.y1 - g.y3) * (g.y1 - g.y3);
  const double t = 2e-10 + 32.0 * 1.1102230246251565e-16 * fmax(fmax(a, b), c);
  return t < 1e300 ? (__double2hiint(t) + 4) & kThrMask : 0x7fe00000;  // rounded up, bits 0-1 left free
}
template <> __device__ __forceinline__ int amb_threshold<float>(const TriGeom<float>& g) {
  const float a = (g.x2 - g.x1) * (g.x2 - g.x1) + (g.y2 - g.y1) * (g.y2 - g.y1);
  const float b = (g.x3 - g.x2) * (g.x3 - g.x2) + (g.y3 - g.y2) * (g.y3 - g.y2);
  const float c = (g.x1 - g.x3) * (g.x1 - g.x3) + (g.y1 - g.y3) * (g.y1 - g.y3);
  const float t = 2e-10f + 32.0f * 5.9604645e-8f * fmaxf(fmaxf(a, b), c);
  return t < 1e37f ? (__float_as_int(t) + 4) & kThrMask : 0x7f000000;
}
// |x| below the triangle's threshold?  Integer compares on the (high) word: two instructions per cross.
// (thr = the record's pad without its kLongTriangle bit)
__device__ __forceinline__ bool in_amb_band(double x, int thr) { return (__double2hiint(x) & 0x7fffffff) < thr; }
__device__ __forceinline__ bool in_amb_band(float x, int thr) { return (__float_as_int(x) & 0x7fffffff) < thr; }

// walk() that also reports whether the accepting (or hull-exit) test was inside the margin band.
// `g` / `have`: the walk part of record `have` (-1: none) kept by the caller from one walk to the next -- a chain of
// walks along a grid row (fix-up, row walk) mostly starts in the triangle the previous one ended in, and the record
// load is the longest link of its dependent chain.
template <class R>
__device__ __forceinline__ int walk_checked_cached(const TriGeom<R>* __restrict__ geom, int tri, TriGeom<R>& g, int& have,
                                                   R tx, R ty, int iter_cap, bool& inside, bool& amb, int* status,
                                                   int* iters_out = nullptr) {
  const R margin = Consts<R>::margin();
  inside = true;
  int it = 0;
  for (;;) {
    ++it;
    if (tri != have) {
      load_walk(geom + tri, g);
      g.pad &= kThrMask;  // the reference's path: first failing edge, long triangle or not
      have = tri;
    }
    const R c1 = (g.x2 - g.x1) * (ty - g.y1) - (g.y2 - g.y1) * (tx - g.x1);
    const R c2 = (g.x3 - g.x2) * (ty - g.y2) - (g.y3 - g.y2) * (tx - g.x2);
    const R c3 = (g.x1 - g.x3) * (ty - g.y3) - (g.y1 - g.y3) * (tx - g.x3);
    int nxt;
    R cf;
    if (c1 < margin) { nxt = g.n3; cf = c1; }
    else if (c2 < margin) { nxt = g.n1; cf = c2; }
    else if (c3 < margin) { nxt = g.n2; cf = c3; }
    else { amb = in_amb_band(c1, g.pad) | in_amb_band(c2, g.pad) | in_amb_band(c3, g.pad); break; }
    if (nxt == kNotPacked) nxt = packed_alternative(c1 < margin, c2 < margin, c3 < margin, g.n1, g.n2, g.n3);
    if (nxt < 0) {
      if (nxt == kNotPacked && status) raise_status(status, GI_ERR_PACK_WINDOW);
      inside = false; amb = in_amb_band(cf, g.pad); break;
    }
    tri = nxt;
    if (it >= iter_cap) { if (status) raise_status(status, GI_ERR_ITER_CAP); inside = false; amb = false; break; }
  }
  if (iters_out) *iters_out += it;
  return tri;
}
template <class R>
__device__ __forceinline__ int walk_checked(const TriGeom<R>* __restrict__ geom, int tri, R tx, R ty, int iter_cap,
                                            bool& inside, bool& amb, int* status, int* iters_out = nullptr) {
  TriGeom<R> g;
  int have = -1;
  return walk_checked_cached(geom, tri, g, have, tx, ty, iter_cap, inside, amb, status, iters_out);
}
// The same result for a walk that starts at a seed triangle, where no path is prescribed: the located triangle does
// not depend on the path (outside the margin band, which `amb` reports), so the moves are made with the careful
// exit-edge choice (at most ~15 of them through the sliver triangles along the hull, where the reference's
// first-failing-edge rule needs up to 1 220) and only the final test is the reference's.
template <class R>
__device__ __forceinline__ int locate_checked(const TriGeom<R>* __restrict__ geom, int seed_tri, R tx, R ty, int iter_cap,
                                              bool& inside, bool& amb, int* status) {
  const int res = careful_locate(geom, seed_tri, tx, ty, iter_cap, status);
  if (res < 0) {  // left through a hull edge of triangle ~res (callers want certain INSIDE starts, or a guess)
    inside = false; amb = false;
    return ~res;
  }
  return walk_checked(geom, res, tx, ty, iter_cap, inside, amb, status);   // one test: every edge passes in `res`
}

// Multi-GPU gather fused into the kernels (north_star "final gather of the output field"): every store
// of a result -- phase 2 of locate_eval_kernel, the fix-up paths, the hull fill -- goes to the local field and to the
// same element of up to kMaxPeers peer fields over NVLink, so the band arrives at its owners while the rest of
// the band is still being computed and no separate collective runs afterwards.
constexpr int kMaxPeers = GI_MAX_GATHER_FIELDS - 1;
template <class R> struct GatherDest {
  R* peer[kMaxPeers];
  int n_peer;
  int64_t ld;  // 0 = band-shaped contiguous
};

// Values of halo targets (walked rows outside the band) that a fix-up re-located.  The hull fill re-evaluates a
// source cell that lies in the halo on the fly (nothing is stored for halo rows), which finds the right triangle
// for every target except those inside the margin band of an edge, where the reference's choice depends on its
// walk path: those were re-located by the fix-up paths, and publish_target leaves their value here so that a band
// fills from exactly the value a one-shot run of the whole grid would have read.  Dense over the halo rows
// (2 * halo * len_x values), initialised to an all-ones bit pattern = "no fix-up here".
template <class R> struct HaloFix {
  R* vals;              // nullptr: no halo
  int y_e0, y_b0, y_b1; // walked rows start at y_e0, the band is [y_b0, y_b1)
  int len_x;
  __device__ __forceinline__ int64_t index(int ix, int iy) const {
    if (!vals || (iy >= y_b0 && iy < y_b1)) return -1;
    const int hr = iy < y_b0 ? iy - y_e0 : (iy - y_b1) + (y_b0 - y_e0);
    return (int64_t)hr * len_x + ix;
  }
};
__device__ __forceinline__ bool is_no_fix(double v) { return __double_as_longlong(v) == -1ll; }
__device__ __forceinline__ bool is_no_fix(float v) { return __float_as_int(v) == -1; }

template <class R> struct WalkArgs {
  TargetWindow<R> win;   // walked window E (band + halo)
  int bf0, bf1, bs0, bs1;  // band B inside E: the part that is written to `out`
  const TriGeom<R>* geom;
  const int* seed;
  int ncx, ncy, num_tri, iter_cap;
  R* out;               // band-shaped: (bs1-bs0) slow rows of (bf1-bf0), `out_ld` elements from one slow row to the next
  int64_t out_ld;
  R* peer_out[kMaxPeers];  // the same band in the fields of other GPUs (peer-mapped device memory), see GatherDest
  int n_peer;
  int32_t* tri_out;     // optional, band-shaped
  uint32_t* mask_words; // E-shaped bit mask: ((nf+31)/32) words per slow row
  int* nz_words;        // compact list of the mask words that are non-zero (work list of the hull fill)
  int64_t* counters;    // [0] walk iterations [1] masked cells [2] max fill level [3] fix-ups [4] #nz_words
                        // [5] #ambiguous targets flagged [6] a fix-up found no certain column to start from
                        // [7] the reference's start triangle
  int* amb_list;        // E-linear ids of the targets inside the margin band of an edge (see fixup_kernel)
  int amb_cap;
  unsigned long long* start_part;  // start_tri_kernel scratch: 2 * kStartBlocks partial results + ticket (zeroed)
  int reference_rows;   // GI_BARY_REFERENCE_ORDER: locate every target by the reference's row-sequential walk
  int* row_todo;        // [len_y], zeroed: grid rows a fix-up handed to row_walk_kernel
  const int* only_if;   // locate_eval_kernel: run only if this device flag is set (nullptr: always), see raster.cuh
  HaloFix<R> hfix;      // where fix-ups leave the values of halo targets (see HaloFix)
  const uint8_t* need;  // band-limited pack: which triangles have records (nullptr: all); set by bary_walk
  int* status;
};

// The reference's start triangle (ind_tri_start, counters[7]) as the start of a walk.  After a band-limited pack it
// may have no record (it lies near the first column at the MIDDLE row of the grid, wherever the band is): the call
// is then repeated with the whole mesh, like any walk that would leave the packed set (GI_ERR_PACK_WINDOW).
template <class R> __device__ __forceinline__ int start_triangle(const WalkArgs<R>& a) {
  const int t = (int)a.counters[7];
  if (a.need && !a.need[t]) {
    raise_status(a.status, GI_ERR_PACK_WINDOW);
    return -1;
  }
  return t;
}

// Phase 2 of the locate / raster kernels: evaluate one strip of ROWS x kWalkThreads targets whose triangles are in
// the shared-memory tile `stri` (>= 0: inside that triangle; < 0: outside the hull, ~id of the last valid triangle
// or -1), store the values and the outside-hull mask words.
template <class R, bool FY, bool PEERS, int ROWS, bool PLANE = false, bool LEAN = false>
__device__ __forceinline__ void evaluate_strip(const WalkArgs<R>& a, const int (*stri)[kWalkThreads], const R* saxis,
                                               int srow0, int nrows, int lane_f, bool active, R cf) {
  const TargetWindow<R>& w = a.win;
  const int tid = threadIdx.x, lane = tid & 31;
  const int f = w.f0 + lane_f;
  // ---- phase 2: evaluate -----------------------------------------------------------------------
  // Row-synchronous.  When the column enters another triangle the lane fetches the record's evaluation sectors
  // (L1-resident after phase 1) -- A..D, denom and the nodal values were computed by the pack pass -- and derives
  // what depends on (triangle, column): A*(tx-x3), C*(tx-x3) for fast axis x and the refined reciprocal of denom;
  // the row-dependent half of interpolation.f90:387-394 is evaluated per target.
  const bool f_in_band = active && f >= a.bf0 && f < a.bf1;
  const int wpr = (w.nf() + 31) >> 5;
  const int bnf = a.bf1 - a.bf0;
  static_assert(ROWS <= 32, "one lane per row keeps that row's mask word");
  const int rb0 = max(0, a.bs0 - srow0), rb1 = min(nrows, a.bs1 - srow0);   // rows of the strip inside the band
  int cur = -1;
  bool cur_exact = false;
  unsigned my_mask = 0;
  R c1 = 0, q1 = 0, c2 = 0, q2 = 0, p3 = 0, d1 = 0, d2 = 0, d3 = 0;
  DivRcp<R> dv;
  dv.d = R(1); dv.r2 = R(1);
  // (LEAN: the 56-register instantiation, see walk_lean(); two rows in flight cost it registers it does not have)
#pragma unroll(LEAN ? 1 : 2)
  for (int r = 0; r < nrows; ++r) {
    const int t = active ? stri[r][tid] : 0;
    const bool inside = t >= 0;
    if (f_in_band && r >= rb0 && r < rb1) {
      if (inside) {
        R v;
        if (PLANE) {
          // GI_BARY_PLANE_VALUES (see plane_of): two record sectors per triangle change, 2 (fast axis x) or 4 FP64
          // operations per target; c1 = the row-independent part, q1 / q2 = gx / gy, d3 = the text's d3.
          // cur_exact: this triangle kept the exact record (a needle): the reference's expression, out of line.
          if (t != cur) {
            cur = t;
            R x3, y3, gx, gy, g3, unused;
            int plane_pad;
            load_plane(a.geom + t, x3, y3, plane_pad, gx, gy, g3, unused);
            cur_exact = !(plane_pad & kPlane);
            if (!FY) { c1 = g3 + gx * (cf - x3); q2 = gy; p3 = y3; }
            else { c1 = cf - y3; q1 = gx; q2 = gy; d3 = g3; p3 = x3; }
          }
          if (cur_exact) v = eval_exact_record(a.geom + t, FY ? saxis[r] : cf, FY ? cf : saxis[r]);
          else if (!FY) v = fma(q2, saxis[r] - p3, c1);
          else v = fma(q2, c1, d3 + q1 * (saxis[r] - p3));
        } else {
        if (t != cur) {
          cur = t;
          TriGeom<R> g;
          load_eval(a.geom + t, g);
          dv.set(g.denom);
          d1 = g.d1; d2 = g.d2; d3 = g.d3;
          const R A = g.y2 - g.y3, B = g.x3 - g.x2, C = g.y3 - g.y1, D = g.x1 - g.x3;
          if (!FY) { const R u = cf - g.x3; c1 = A * u; q1 = B; c2 = C * u; q2 = D; p3 = g.y3; }
          else { const R u = cf - g.y3; c1 = B * u; q1 = A; c2 = D * u; q2 = C; p3 = g.x3; }
        }
        // w1 = (A*(tx-x3) + B*(ty-y3))/denom, w2 = (C*(tx-x3) + D*(ty-y3))/denom   (:387-394); the sum commutes
        const R t3 = saxis[r] - p3;
        R w1, w2;
        quotient_pair(dv, FY ? q1 * t3 + c1 : c1 + q1 * t3, FY ? q2 * t3 + c2 : c2 + q2 * t3, w1, w2);
        const R w3 = R(1) - w1 - w2;
        v = d1 * w1 + d2 * w2 + d3 * w3;                           // :396-398
        }
        // band-relative element (computed per row: pointers carried through the loop cost registers, and spills)
        const int64_t oo = (int64_t)(srow0 + r - a.bs0) * a.out_ld + (f - a.bf0);
        a.out[oo] = v;
        if (PEERS) {  // the fused gather: the same element of every peer field, over NVLink
          for (int d = 0; d < a.n_peer; ++d) a.peer_out[d][oo] = v;
        }
      }
      if (a.tri_out) a.tri_out[(int64_t)(srow0 + r - a.bs0) * bnf + (f - a.bf0)] = (inside ? t : ~t) + 1;
    }
    const unsigned m = __ballot_sync(0xffffffffu, active && !inside);
    if (lane == r) my_mask = m;  // lane r keeps the word of row r (ROWS <= 32)
  }
  if (lane < nrows && (lane_f >> 5) < wpr) {
    const int64_t wi = (int64_t)(srow0 + lane - w.s0) * wpr + (lane_f >> 5);
    a.mask_words[wi] = my_mask;
    if (my_mask) a.nz_words[atomicAdd((unsigned long long*)(a.counters + 4), 1ull)] = (int)wi;
  }
}

// ---------------------------------------------------------------------------------------------
// The production kernel: locate, then evaluate.
//
// ncu on the first two versions (profiles/r01_*): a fused "walk + interpolate per target" loop has three
// divergent regions (orientation test / move to the neighbour / barycentric value) and ran with 18-21 of 32
// lanes active, 6 warps per scheduler and 53-63 % issue utilisation.  Splitting it makes both halves SIMT
// friendly:
//   phase 1 (locate)   every lane walks its own column down the strip, decoupled from the other lanes: one
//                      loop whose body is "test the cached triangle; then either record the triangle id for
//                      this row and go to the next row, or load the neighbouring triangle's 64 B record".
//                      Only 9 doubles of per-(triangle, column) state stay in registers.  The ids go into a
//                      shared-memory tile (int32, ~id = left through a hull edge).
//   phase 2 (evaluate) row-synchronous and uniform: every lane reads its triangle's 96 B evaluation record
//                      (L1/L2 resident, written by pack_mesh_kernel) and computes the value with two
//                      shared-reciprocal divisions; 128 consecutive doubles per row are stored coalesced and
//                      the outside-hull bits come from one ballot per warp row.
// ---------------------------------------------------------------------------------------------
// The full-grid kernel of the headline layout runs at 9 CTAs per SM instead of 8 (+3.6 %, profiles/r02_walk_variants.md)
// if it fits 56 registers without spilling into the walk loop, which it does once the iteration counter (iters_tot,
// a diagnostic) and the second in-flight row of phase 2 (-1.6 %) are gone.  The y-fast instantiation still spills at
// 56 registers and loses (2.78 -> 2.92 ms), so only x-fast, whole-mesh, many-wave windows are "lean" (and not the plane form, which spills there too).
constexpr bool walk_lean(bool fy, bool peers, bool bp, bool cf) { return !fy && !peers && !bp && !cf; }
template <class R, bool FY, bool PEERS, int ROWS, bool BP, bool CF, bool PLANE = false, bool LEAN = false>
__device__ __forceinline__ void locate_eval_tile(const WalkArgs<R>& a, int bx, int by, int (*stri)[kWalkThreads], R* saxis) {
  const TargetWindow<R>& w = a.win;
  const int tid = threadIdx.x, lane = tid & 31;
  const int lane_f = bx * kWalkThreads + tid;
  const int f = w.f0 + lane_f;
  const bool active = f < w.f1;
  const int srow0 = w.s0 + by * ROWS;
  const int nrows = min(ROWS, w.s1 - srow0);
  if (tid < nrows) saxis[tid] = __ldg(w.slow_axis + srow0 + tid);
  __syncthreads();
  const R cf = active ? __ldg(w.fast_axis + f) : R(0);
  const R margin = Consts<R>::margin();
#if GI_WALK_PREFETCH
  const bool descending = saxis[nrows - 1] < saxis[0];
#endif

  // ---- phase 1: locate -------------------------------------------------------------------------
  long long iters_sum = 0;
  if (active) {
    int tri, n1, n2, n3, pad;  // pad: the record's threshold + flags word (kLongTriangle, kSliver), kept whole
    R m1, m2, m3, p1, p2, p3, k1, k2, k3;
    bool fresh;    // no target resolved in `tri` yet
    bool forced_out = false;  // careful_locate found this row's target outside the hull at `tri`
    auto enter = [&](int t) {  // make `t` the current triangle: one 64 B load + 12 flops
      tri = t;
      fresh = true;
      TriGeom<R> g;
      load_walk(a.geom + t, g);
      n1 = g.n1; n2 = g.n2; n3 = g.n3; pad = g.pad;
      const R ex1 = g.x2 - g.x1, ey1 = g.y2 - g.y1;   // edge (v1,v2)
      const R ex2 = g.x3 - g.x2, ey2 = g.y3 - g.y2;   // edge (v2,v3)
      const R ex3 = g.x1 - g.x3, ey3 = g.y1 - g.y3;   // edge (v3,v1)
      if (!FY) {
        m1 = ex1; m2 = ex2; m3 = ex3; p1 = g.y1; p2 = g.y2; p3 = g.y3;
        k1 = ey1 * (cf - g.x1); k2 = ey2 * (cf - g.x2); k3 = ey3 * (cf - g.x3);
      } else {
        m1 = ey1; m2 = ey2; m3 = ey3; p1 = g.x1; p2 = g.x2; p3 = g.x3;
        k1 = ex1 * (cf - g.y1); k2 = ex2 * (cf - g.y2); k3 = ex3 * (cf - g.y3);
      }
#if GI_WALK_PREFETCH
      {
        // The edge this column will leave the triangle through on its way along the slow axis: the cross product of
        // edge i changes by m_i (fast axis x) or -m_i (fast axis y) per unit of the slow coordinate, so only edges
        // with a falling cross can be crossed; two of them meet in the vertex between them, and the column passes
        // that vertex on one side.  A prediction only (margins, descending axes): the neighbour's record is warmed.
        const bool s1 = (FY ? m1 > R(0) : m1 < R(0)) != descending, s2 = (FY ? m2 > R(0) : m2 < R(0)) != descending,
                   s3 = (FY ? m3 > R(0) : m3 < R(0)) != descending;
        int pn;
        if (s1 && s2) pn = ((FY ? cf < g.y2 : cf > g.x2) != descending) ? n3 : n1;       // edge 1 -> n3, edge 2 -> n1
        else if (s2 && s3) pn = ((FY ? cf < g.y3 : cf > g.x3) != descending) ? n1 : n2;  // edge 3 -> n2
        else if (s3 && s1) pn = ((FY ? cf < g.y1 : cf > g.x1) != descending) ? n2 : n3;
        else pn = s1 ? n3 : (s2 ? n1 : n2);
        if (pn >= 0) touch_walk(a.geom + pn);
      }
#endif
    };
    const int ix = FY ? srow0 : f, iy = FY ? f : srow0;
    enter(seed_lookup(a.seed, a.ncx, a.ncy, ix, iy, a.num_tri));
    int r = 0, it = 0, iters = 0;
    while (r < nrows) {
      const R cv = saxis[r];
      // cross_i = (x_b-x_a)*(ty-y_a) - (y_b-y_a)*(tx-x_a)   (triangle_walk.f90:149-150) evaluated as
      // m_i*(cv-p_i) - k_i (fast axis x: m = x_b-x_a, p = y_a, k = (y_b-y_a)*(tx-x_a)) or k_i - m_i*(cv-p_i)
      // (fast axis y): the reference's operations in the reference's order, with the product that does not
      // depend on the row computed once per (triangle, column) -- the rounded results are identical
      const R t1 = cv - p1, t2 = cv - p2, t3 = cv - p3;
      const R x1 = FY ? k1 - m1 * t1 : m1 * t1 - k1;
      const R x2 = FY ? k2 - m2 * t2 : m2 * t2 - k2;
      const R x3 = FY ? k3 - m3 * t3 : m3 * t3 - k3;
      ++it;
      // first failing edge wins (:57-60); edge i -> neighbour opposite vertex i-1 (wrap) (:66-71)
      const bool o1 = x1 < margin, o2 = x2 < margin, o3 = x3 < margin;
      const bool outside = o1 | o2 | o3;
      int nxt = o1 ? n3 : (o2 ? n1 : n2);
      // (BP: the pack was band-limited, links may lead to triangles without a record)
      if (BP && outside && nxt == kNotPacked) nxt = packed_alternative(o1, o2, o3, n1, n2, n3);
      if (outside && nxt >= 0 && !(CF && forced_out)) {
        if (it < a.iter_cap) {
          // (CF: careful exit-edge choice in long triangles, see farthest_failing_edge; compiled in for windows of few
          // strips, where the hull strip's walks are the critical path)
          if (CF && pad < 0 && ((o1 && o2) || (o3 && (o1 || o2)))) {   // long triangle, more than one edge fails (rare)
            const int res = careful_locate(a.geom, tri, FY ? cv : cf, FY ? cf : cv, a.iter_cap, a.status);
            forced_out = res < 0;   // a failing hull edge was seen: outside, whichever edge fails first there
            nxt = forced_out ? ~res : res;
          }
          enter(nxt);   // (after careful_locate: the test above is repeated in the triangle that walk ended in)
          continue;
        }
        raise_status(a.status, GI_ERR_ITER_CAP);
      }
      if (CF) forced_out = false;
      if (BP && outside && nxt == kNotPacked) raise_status(a.status, GI_ERR_PACK_WINDOW);  // band-limited pack too tight
      // resolved: inside `tri`, or outside the hull (:74-80: the last valid triangle is kept).
      // A target within the margin band of an edge is accepted by both adjacent triangles and which one the
      // reference returns depends on its walk path: flag it for fixup_kernel.
      // (any |cross| below twice the margin: a superset of the truly ambiguous cases, three compares)
      const int thr = pad & kThrMask;
      const bool sliver = pad & kSliver;   // hull sliver: targets found in it are masked (sliver_guard)
      if (in_amb_band(x1, thr) | in_amb_band(x2, thr) | in_amb_band(x3, thr)) {
        const unsigned long long slot = atomicAdd((unsigned long long*)(a.counters + 5), 1ull);
        if (slot < (unsigned long long)a.amb_cap) a.amb_list[slot] = (srow0 + r - w.s0) * w.nf() + lane_f;
      }
      if (fresh && !outside && !sliver) {  // phase 2 will want this triangle's nodal values: pull the sector into L1 now
        touch_values(a.geom + tri);
        fresh = false;
      }
      stri[r][tid] = (outside | sliver) ? ~tri : tri;
      if (!LEAN) iters += it;   // (iters_tot: not counted by the lean instantiation)
      it = 0;
      ++r;
    }
    iters_sum = iters;
  }
  __syncthreads();

  evaluate_strip<R, FY, PEERS, ROWS, PLANE, LEAN>(a, stri, saxis, srow0, nrows, lane_f, active, cf);
  // one atomic per warp for the iteration counter (iters_tot, interpolation.f90:374)
  for (int off = 16; off > 0; off >>= 1) iters_sum += __shfl_down_sync(0xffffffffu, iters_sum, off);
  if (lane == 0 && iters_sum) atomicAdd((unsigned long long*)a.counters, (unsigned long long)iters_sum);
}

template <class R, bool FY, bool PEERS, int ROWS, int MINB, bool BP, bool CF, bool PLANE = false>
__global__ void __launch_bounds__(kWalkThreads, MINB) locate_eval_kernel(const WalkArgs<R> a) {
  __shared__ int stri[ROWS][kWalkThreads];
  __shared__ R saxis[ROWS];
  locate_eval_tile<R, FY, PEERS, ROWS, BP, CF, PLANE, (MINB > 8)>(a, blockIdx.x, blockIdx.y, stri, saxis);
}

// The same as a stand-by for the rasteriser (raster.cuh): a small grid that returns at once unless the device
// flag says the rasteriser could not do the window (an empty launch of the full grid costs 37 us at C4 size), and
// then loops over the strips.
template <class R, bool FY, bool PEERS, int ROWS>
__global__ void __launch_bounds__(kWalkThreads) locate_eval_standby_kernel(const WalkArgs<R> a, int nbx, int nby) {
  __shared__ int stri[ROWS][kWalkThreads];
  __shared__ R saxis[ROWS];
  if (!*a.only_if) return;
  for (int t = blockIdx.x; t < nbx * nby; t += gridDim.x) {
    locate_eval_tile<R, FY, PEERS, ROWS, true, true>(a, t % nbx, t / nbx, stri, saxis);
    __syncthreads();
  }
}


#include "raster.cuh"

// Result of a (re-)located target (f, s): value or mask bit, triangle id, work list of the hull fill.
template <class R>
__device__ __forceinline__ void publish_target(const WalkArgs<R>& a, int f, int s, int tri, bool inside, R tx, R ty) {
  const TargetWindow<R>& w = a.win;
  if (inside && (a.geom[tri].pad & kSliver)) inside = false;  // sliver_guard
  if (f >= a.bf0 && f < a.bf1 && s >= a.bs0 && s < a.bs1) {
    const int64_t o = (int64_t)(s - a.bs0) * (a.bf1 - a.bf0) + (f - a.bf0);
    if (inside) {
      TriGeom<R> g;
      load_eval(a.geom + tri, g);
      const int64_t oo = (int64_t)(s - a.bs0) * a.out_ld + (f - a.bf0);
      const R v = eval_any(g, tx, ty);
      a.out[oo] = v;
      for (int d = 0; d < a.n_peer; ++d) a.peer_out[d][oo] = v;
    }
    if (a.tri_out) a.tri_out[o] = (inside ? tri : ~tri) + 1;
  } else if (inside) {  // a halo target: the hull fill may take it as a source cell
    const int64_t hi = a.hfix.index(w.fast_is_y ? s : f, w.fast_is_y ? f : s);
    if (hi >= 0) {
      TriGeom<R> g;
      load_eval(a.geom + tri, g);
      a.hfix.vals[hi] = eval_any(g, tx, ty);
    }
  }
  const int lf = f - w.f0, wpr = (w.nf() + 31) >> 5;
  uint32_t* word = a.mask_words + (int64_t)(s - w.s0) * wpr + (lf >> 5);
  const uint32_t bit = 1u << (lf & 31);
  if (inside) {
    atomicAnd(word, ~bit);
  } else if (atomicOr(word, bit) == 0u) {  // the word was not on the fill's work list yet
    a.nz_words[atomicAdd((unsigned long long*)(a.counters + 4), 1ull)] = (int)((int64_t)(s - w.s0) * wpr + (lf >> 5));
  }
}

// ---------------------------------------------------------------------------------------------
// Ambiguity fix-up.  A target within 1e-10 (cross-product units) of a mesh edge passes the orientation test of
// both adjacent triangles (triangle_walk.f90:148-151); the reference returns whichever its own walk reaches
// first -- it walks along the grid row from the triangle found for the previous column, every row starting at
// ind_tri_start (interpolation.f90:342-355,365-373).  For random meshes this concerns about one target in 1e10,
// for meshes with grid-aligned edges (structured triangulations) it concerns many.  The flagged targets are
// re-located here by following the reference's path: back up along the row to the nearest column whose triangle
// is unambiguous (any walk finds that one), or to column 0 and the reference's start triangle, then walk forward
// column by column.
// ---------------------------------------------------------------------------------------------
template <class R> struct RawMesh {  // the caller's arrays: what a band-limited pack did not turn into records
  PointsView<R> pts;
  TopoView simp;
  int num_points;
  int use;  // 1: records may be missing, read the raw arrays
};
template <class R>
__device__ __forceinline__ unsigned long long start_tri_dist(const TriGeom<R>* geom, const RawMesh<R>& raw, int i, R opt_x,
                                                             R opt_y) {
  TriGeom<R> g;
  if (raw.use) {
    int a = raw.simp(i, 0), b = raw.simp(i, 1), c = raw.simp(i, 2);
    if ((unsigned)(a - 1) >= (unsigned)raw.num_points || (unsigned)(b - 1) >= (unsigned)raw.num_points ||
        (unsigned)(c - 1) >= (unsigned)raw.num_points)
      a = b = c = 1;  // as pack_mesh_kernel makes an invalid record harmless
    raw.pts.xy(a - 1, g.x1, g.y1);
    raw.pts.xy(b - 1, g.x2, g.y2);
    raw.pts.xy(c - 1, g.x3, g.y3);
  } else {
    load_walk(geom + i, g);
  }
  const R cx = ((g.x1 + g.x2) + g.x3) / R(3), cy = ((g.y1 + g.y2) + g.y3) / R(3);   // :348-349
  const double d = (double)((cx - opt_x) * (cx - opt_x) + (cy - opt_y) * (cy - opt_y));  // :350
  return (unsigned long long)__double_as_longlong(d);  // non-negative doubles order like their bit patterns
}

// The reference's ind_tri_start (interpolation.f90:342-355): arg-min of the centroid distance to
// (x_axis(1), SUM(y_axis)/len_y) over the first min(num_points, num_triangles) triangles, first minimum wins, i.e.
// the lexicographic minimum of (distance, index).  Needed only by fix-ups that reach column 0, so the kernel
// returns at once unless a target was flagged.  kStartBlocks CTAs each reduce a strided share, the last one to
// finish (ticket) reduces the partial results.  (The first version was one CTA: 6.4 ms for 5 M triangles
// whenever anything was flagged -- single-precision C4, structured meshes.)
constexpr int kStartBlocks = 296;
template <class R>
__global__ void __launch_bounds__(256)
start_tri_kernel(const TriGeom<R>* __restrict__ geom, const RawMesh<R> raw, int n_search, const R* __restrict__ x_axis,
                 const R* __restrict__ y_axis, int len_y, int64_t* counters, unsigned long long* part, int force) {
  if (counters[5] == 0 && !force) return;
  __shared__ R s_opt_y;
  __shared__ unsigned long long s_d[8], s_i[8];
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    R sum = R(0);
    for (int i = 0; i < len_y; ++i) sum = sum + y_axis[i];                // SUM(y_axis), :345 (serial order)
    s_opt_y = sum / R(len_y);
  }
  __syncthreads();
  const R ox = __ldg(x_axis), oy = s_opt_y;
  unsigned long long bd = ~0ull, bi = ~0ull;
  auto take = [&](unsigned long long d, unsigned long long i) {           // strict '<', first minimum wins (:351)
    if (d < bd || (d == bd && i < bi)) { bd = d; bi = i; }
  };
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_search; i += gridDim.x * blockDim.x)
    take(start_tri_dist(geom, raw, i, ox, oy), (unsigned long long)i);
  auto block_reduce = [&]() {
    for (int off = 16; off > 0; off >>= 1)
      take(__shfl_xor_sync(0xffffffffu, bd, off), __shfl_xor_sync(0xffffffffu, bi, off));
    if ((threadIdx.x & 31) == 0) { s_d[threadIdx.x >> 5] = bd; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
      bd = threadIdx.x < 8 ? s_d[threadIdx.x] : ~0ull;
      bi = threadIdx.x < 8 ? s_i[threadIdx.x] : ~0ull;
      for (int off = 4; off > 0; off >>= 1)
        take(__shfl_xor_sync(0xffffffffu, bd, off), __shfl_xor_sync(0xffffffffu, bi, off));
    }
  };
  block_reduce();
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = bd; part[2 * blockIdx.x + 1] = bi;
    __threadfence();
    s_last = atomicAdd(part + 2 * kStartBlocks, 1ull) == (unsigned long long)gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  bd = ~0ull; bi = ~0ull;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
    const volatile unsigned long long* vp = part;
    take(vp[2 * b], vp[2 * b + 1]);
  }
  __syncthreads();
  block_reduce();
  if (threadIdx.x == 0) {
    counters[7] = (int64_t)bi;
    part[2 * kStartBlocks] = 0ull;  // ticket ready for the next call on this scratch
  }
}

constexpr int kMaxBackup = 256;
// entries of the fix-up work list (gi_set_fixup_list_capacity shrinks it: tests of the overflow path)
inline int64_t amb_list_limit() { return fixup_list_capacity(); }  // gi_set_fixup_list_capacity (api.cu)

// Thread per flagged target for the usual case (the column before it is a certain start: one seed walk, one step);
// a target whose back-up has to go further is taken by its whole warp, 32 candidate columns at a time -- in single
// precision, or on a lattice, whole stretches of a row are ambiguous, and 256 candidates tested one after the other
// by one thread (each a seed walk of ~10 dependent record loads) made the kernel's tail (f32 C4: 1.2 ms).
template <class R>
__global__ void __launch_bounds__(128) fixup_kernel(const WalkArgs<R> a, int len_x, int len_y) {
  const TargetWindow<R>& w = a.win;
  const int64_t n = min(a.counters[5], (int64_t)a.amb_cap);
  const int nf = w.nf(), wpr = (nf + 31) >> 5;
  const R* x_axis = w.fast_is_y ? w.slow_axis : w.fast_axis;
  const R* y_axis = w.fast_is_y ? w.fast_axis : w.slow_axis;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // is column j of row iy a certain start?  (inside, and not within the margin band of an edge: any walk finds it)
  auto certain = [&](int j, int iy, R ty, int& tri) -> bool {
    // phase 1 has classified column j already: one it found outside the hull cannot be the certain INSIDE start
    // looked for here (passing over a candidate is always safe; a row outside the hull costs loads, not walks)
    const int jf = w.fast_is_y ? iy : j, js = w.fast_is_y ? j : iy;
    if (jf >= w.f0 && jf < w.f1 && js >= w.s0 && js < w.s1) {
      const uint32_t mw = a.mask_words[(int64_t)(js - w.s0) * wpr + ((jf - w.f0) >> 5)];
      if ((mw >> ((jf - w.f0) & 31)) & 1u) return false;
    }
    bool inside, amb;
    tri = locate_checked(a.geom, seed_lookup(a.seed, a.ncx, a.ncy, j, iy, a.num_tri), __ldg(x_axis + j), ty, a.iter_cap,
                         inside, amb, a.status);
    return inside && !amb;
  };
  // forward along the row from the state after column j0 exactly as the reference does (:367-373), then publish
  auto finish = [&](int tri, int j0, int ix, int iy, int f, int s, R ty) {
    bool inside = true, amb;
    TriGeom<R> g;
    int have = -1;
    for (int j = j0 + 1; j <= ix; ++j)
      tri = walk_checked_cached(a.geom, tri, g, have, __ldg(x_axis + j), ty, a.iter_cap, inside, amb, a.status);
    publish_target(a, f, s, tri, inside, __ldg(x_axis + ix), ty);
    atomicAdd((unsigned long long*)(a.counters + 3), 1ull);
  };
  for (int64_t base = warp * 32; base < n; base += nwarps * 32) {   // (warp-uniform trip count)
    const int64_t li = base + lane;
    bool far = false;
    int f = 0, s = 0, ix = 0, iy = 0;
    R ty = R(0);
    if (li < n) {
      const int lin = a.amb_list[li];
      s = w.s0 + lin / nf; f = w.f0 + lin % nf;
      ix = w.fast_is_y ? s : f; iy = w.fast_is_y ? f : s;
      ty = __ldg(y_axis + iy);
      int tri = -1;
      if (ix == 0) {                                                     // column 0 starts at ind_tri_start (:366)
        const int t0 = start_triangle(a);
        if (t0 >= 0) finish(t0, -1, ix, iy, f, s, ty);
      }
      else if (certain(ix - 1, iy, ty, tri)) finish(tri, ix - 1, ix, iy, f, s, ty);
      else far = true;
    }
    // back up to the nearest column of this row whose triangle does not depend on the path
    unsigned todo = __ballot_sync(0xffffffffu, far);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int t_ix = __shfl_sync(0xffffffffu, ix, src), t_iy = __shfl_sync(0xffffffffu, iy, src);
      const R t_ty = __ldg(y_axis + t_iy);
      const int j_stop = max(t_ix - kMaxBackup, 0);
      int tri = -1, j0 = -1;
      for (int jb = t_ix - 2; jb >= j_stop && tri < 0; jb -= 32) {
        const int j = jb - lane;
        int t = -1;
        const bool ok = j >= j_stop && certain(j, t_iy, t_ty, t);
        const unsigned hit = __ballot_sync(0xffffffffu, ok);
        if (hit) {
          const int l = __ffs(hit) - 1;   // the nearest one
          tri = __shfl_sync(0xffffffffu, t, l);
          j0 = jb - l;
        }
      }
      if (lane == src) {
        if (tri < 0 && j_stop > 0) {
          // no path-independent column within kMaxBackup (a grid row running along mesh edges): only column 0 (the
          // reference's ind_tri_start, :366) is another certain starting point -- leave the row to row_walk_kernel,
          // which redoes it along the reference's own path
          a.row_todo[iy] = 1;
          a.counters[6] = 1;
        } else {
          if (tri < 0) tri = start_triangle(a);
          if (tri >= 0) finish(tri, j0, ix, iy, f, s, ty);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The reference's own order: one thread per grid row starts at ind_tri_start and walks the row column by column,
// every target from the triangle of the previous one (interpolation.f90:362-404).  By construction this finds the
// reference's triangle for EVERY target, whatever the margin band does -- it is what the fix-up reproduces
// locally.  It runs (a) for the rows in which a fix-up found no certain column to start from within its back-up
// range (a grid row along mesh edges, or along the hull: every target of the row is flagged), (b) for the whole
// window when more targets were flagged than the fix-up list holds (a structured mesh whose edges run along the
// grid lines of a multi-million-node grid), so that nothing is silently left to phase 1's path, and (c) for
// all targets with GI_BARY_REFERENCE_ORDER (debug).  Rows are independent; within a row the chain of dependent walks
// is cut into speculative pieces (below) -- a fallback, not the production path.
// ---------------------------------------------------------------------------------------------
constexpr int kRowChunk = 1024;
constexpr int kRowSpecMaxCols = 50 * 1024;  // speculative form: one int of shared memory per column (200 KB)
constexpr int kRowThreads = 256;
// One CTA per grid row.
//   serial form       thread 0 walks a chunk of columns (the chain of dependent walks, nothing else in it), then all
//                     threads publish the chunk (values, mask bits, work list) in parallel.
//   speculative form  (rows of up to kRowSpecMaxCols columns) the chain is cut into one piece per thread: thread L
//                     walks piece L from a GUESSED state (the triangle a walk from the seed grid finds for the
//                     column before the piece) and records the state after every column.  The state after a
//                     column depends on nothing but the state before it and the column, so piece L is the true
//                     chain iff its guess equals the true state after the last column of piece L-1.  Thread 0 then
//                     goes through the pieces in order: a piece whose guess equals the (by then true) recorded
//                     state before it costs one comparison; otherwise the true chain is continued into the piece
//                     until its state equals the recorded one.  Exact -- and, since a guess is wrong only where the
//                     located triangle depends on the path, 64 + a few dependent walks per row at C4 size instead
//                     of 16 384 (round 2 had 32 pieces per row and one repair walk per piece: 544).
// GI_BARY_TRACE: rows walked, pieces whose guess was wrong, repair walks, their iterations, iterations of the pieces
__device__ unsigned long long g_row_walk_stats[10];  // + cycles: longest piece, repair, whole row; rounds; longest guess
template <class R>
__global__ void __launch_bounds__(kRowThreads) row_walk_kernel(const WalkArgs<R> a, int spec_cols) {
  __shared__ int s_tri[kRowChunk];      // found triangle, ~id = left through a hull edge
  __shared__ int s_guess[kRowThreads];  // speculative form: the state each piece started from
  extern __shared__ int s_state[];      // speculative form: the state after every column of the row
  const bool all_rows = a.reference_rows || a.counters[5] > (int64_t)a.amb_cap;
  if (!all_rows && a.counters[6] == 0) return;
  const TargetWindow<R>& w = a.win;
  const R* x_axis = w.fast_is_y ? w.slow_axis : w.fast_axis;
  const R* y_axis = w.fast_is_y ? w.fast_axis : w.slow_axis;
  const int ix0 = w.fast_is_y ? w.s0 : w.f0, ix1 = w.fast_is_y ? w.s1 : w.f1;
  const int iy0 = w.fast_is_y ? w.f0 : w.s0, iy1 = w.fast_is_y ? w.f1 : w.s1;
  const int tid = threadIdx.x;
  for (int iy = iy0 + blockIdx.x; iy < iy1; iy += gridDim.x) {
    if (!all_rows && !a.row_todo[iy]) continue;   // (the same for every thread of the CTA)
    __syncthreads();
    if (tid == 0) a.row_todo[iy] = 0;
    const R ty = __ldg(y_axis + iy);
    const int start = start_triangle(a);                                   // ind_tri = ind_tri_start   (:366)
    if (start < 0) continue;   // (no record for it: GI_ERR_PACK_WINDOW raised, the caller repeats with the whole mesh)
    bool inside, amb;
    const long long t_row0 = clock64();
    if (ix1 <= spec_cols && ix1 >= 64) {
      // pieces of at least 8 columns: the seed walk of a piece must not outweigh the piece
      const int piece = max((ix1 + kRowThreads - 1) / kRowThreads, 8), c_lo = tid * piece, c_hi = min(c_lo + piece, ix1);
      if (c_lo < ix1) {
        int tri = start;
        if (tid > 0) {  // any valid triangle may be guessed; a good guess makes the repair short
          tri = locate_checked(a.geom, seed_lookup(a.seed, a.ncx, a.ncy, c_lo - 1, iy, a.num_tri), __ldg(x_axis + c_lo - 1),
                               ty, a.iter_cap, inside, amb, (int*)nullptr);
          s_guess[tid] = inside ? tri : ~tri;
          atomicMax(&g_row_walk_stats[9], (unsigned long long)(clock64() - t_row0));
        }
        int its = 0, have = -1;
        TriGeom<R> g;
        R tx = __ldg(x_axis + c_lo);
        for (int ix = c_lo; ix < c_hi; ++ix) {
          const R tx_next = __ldg(x_axis + min(ix + 1, ix1 - 1));   // (off the dependent chain)
          tri = walk_checked_cached(a.geom, tri, g, have, tx, ty, a.iter_cap, inside, amb, tid ? (int*)nullptr : a.status,
                                    &its);
          s_state[ix] = inside ? tri : ~tri;
          tx = tx_next;
        }
        atomicAdd(&g_row_walk_stats[4], (unsigned long long)its);
        atomicMax(&g_row_walk_stats[5], (unsigned long long)(clock64() - t_row0));
      }
      // Repair in rounds, every piece at once: a piece whose start state differs from the state recorded after the
      // column before it is walked again from that state (until it meets its own recorded chain).  Piece 0 is true
      // from the start, so after round r the first r pieces are -- at most one round per piece, in practice as many
      // as a wrong state survives (a row outside the hull, where the reference keeps the hull triangle it left
      // through until that edge no longer sees the target: ~16 pieces at C4 size).
      const long long t_rep0 = clock64();
      int wrong = 0, repairs = 0, its2 = 0, rounds = 0;
      for (;;) {
        __syncthreads();
        int prev = 0;
        bool redo = false;
        if (tid > 0 && c_lo < ix1) {
          prev = s_state[c_lo - 1];
          redo = prev != s_guess[tid];
        }
        if (!__syncthreads_or(redo)) break;   // (also orders the reads above before the writes below)
        ++rounds;
        if (redo) {
          ++wrong;
          s_guess[tid] = prev;
          int tri = prev >= 0 ? prev : ~prev, have = -1;
          TriGeom<R> g;
          R tx = __ldg(x_axis + c_lo);
          for (int ix = c_lo; ix < c_hi; ++ix) {
            const R tx_next = __ldg(x_axis + min(ix + 1, ix1 - 1));
            const int recorded = s_state[ix];
            tri = walk_checked_cached(a.geom, tri, g, have, tx, ty, a.iter_cap, inside, amb, a.status, &its2);   // :372-373
            ++repairs;
            const int st = inside ? tri : ~tri;
            if (st == recorded) break;                                     // met the recorded chain
            s_state[ix] = st;
            tx = tx_next;
          }
        }
      }
      if (tid == 0) {
        atomicAdd(&g_row_walk_stats[0], 1ull);
        atomicMax(&g_row_walk_stats[6], (unsigned long long)(clock64() - t_rep0));
        atomicAdd(&g_row_walk_stats[8], (unsigned long long)rounds);
      }
      if (wrong) {
        atomicAdd(&g_row_walk_stats[1], (unsigned long long)wrong);
        atomicAdd(&g_row_walk_stats[2], (unsigned long long)repairs);
        atomicAdd(&g_row_walk_stats[3], (unsigned long long)its2);
      }
      __syncthreads();
      for (int ix = ix0 + tid; ix < ix1; ix += kRowThreads) {
        const int t = s_state[ix];
        publish_target(a, w.fast_is_y ? iy : ix, w.fast_is_y ? ix : iy, t >= 0 ? t : ~t, t >= 0, __ldg(x_axis + ix), ty);
      }
      atomicMax(&g_row_walk_stats[7], (unsigned long long)(clock64() - t_row0));
      continue;
    }
    int tri = start;
    for (int c0 = 0; c0 < ix1; c0 += kRowChunk) {                          // :367
      const int cn = min(kRowChunk, ix1 - c0);
      if (tid == 0) {
        TriGeom<R> g;
        int have = -1;
        for (int i = 0; i < cn; ++i) {
          tri = walk_checked_cached(a.geom, tri, g, have, __ldg(x_axis + c0 + i), ty, a.iter_cap, inside, amb, a.status);  // :372-373
          s_tri[i] = inside ? tri : ~tri;
        }
      }
      __syncthreads();
      for (int i = tid; i < cn; i += kRowThreads) {
        const int ix = c0 + i, t = s_tri[i];
        if (ix >= ix0)
          publish_target(a, w.fast_is_y ? iy : ix, w.fast_is_y ? ix : iy, t >= 0 ? t : ~t, t >= 0, __ldg(x_axis + ix), ty);
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// hull fill (triangle_walk.f90:90-132) on the bit mask of the walked window
// ---------------------------------------------------------------------------------------------
template <class R> struct FillArgs {
  TargetWindow<R> win;  // window E covered by the bit mask
  const uint32_t* mask_words;
  const int* nz_words;     // indices of the non-zero mask words, counters[4] of them
  int bf0, bf1, bs0, bs1;  // band B inside E: the cells to fill
  int vf0, vf1, vs0, vs1;  // V, B <= V: the cells whose final value can be read from `out` (out is V-shaped)
  int64_t nz_shift;        // added to the entries of nz_words (they may be relative to another window origin)
  R* out;                  // V-shaped, `out_ld` elements from one slow row to the next
  int64_t out_ld;
  R* peer_out[kMaxPeers];  // the same cells in the peer fields (filled values are stored there as well)
  int n_peer;
  int64_t* fill_src;  // optional, band-shaped
  int64_t* counters;  // [1] += masked cells, [2] = max level, [4] = number of nz_words
  int len_x, len_y;
  // to recompute the value of a source cell that lies in the halo (outside the band); geom == nullptr for
  // the standalone fill (no halo there)
  const TriGeom<R>* geom;
  const int* seed;
  int ncx, ncy, num_tri, iter_cap;
  HaloFix<R> hfix;    // values of halo cells that a fix-up re-located (see HaloFix)
  int* status;
};

template <class R>
__device__ __forceinline__ int fill_cell_state(const FillArgs<R>& a, int wpr, int iy, int ix) {
  // 0 = unmasked (valid source), 1 = masked, 2 = outside the walked window (unknown)
  const int f = a.win.fast_is_y ? iy : ix, s = a.win.fast_is_y ? ix : iy;
  if (f < a.win.f0 || f >= a.win.f1 || s < a.win.s0 || s >= a.win.s1) return 2;
  const int lf = f - a.win.f0;
  const uint32_t wd = __ldg(a.mask_words + (int64_t)(s - a.win.s0) * wpr + (lf >> 5));
  return (wd >> (lf & 31)) & 1u;
}

__device__ __forceinline__ int isqrt_floor(int v) {  // floor(sqrt(v)) for v >= 0
  int r = (int)sqrtf((float)v);
  while (r * r > v) --r;
  while ((r + 1) * (r + 1) <= v) ++r;
  return r;
}

template <class R>
__global__ void __launch_bounds__(128) fill_kernel(const FillArgs<R> a) {
  const int nf = a.win.nf(), wpr = (nf + 31) >> 5;
  const int64_t num_words = a.counters[4];
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t num_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int level_max = 0;
  long long masked = 0;
  const int level_cap = a.len_x + a.len_y;
  for (int64_t li = warp0; li < num_words; li += num_warps) {
    const int64_t wi = a.nz_words[li] + a.nz_shift;
    const uint32_t wd = __ldg(a.mask_words + wi);
    const int s = a.win.s0 + (int)(wi / wpr), f = a.win.f0 + (int)(wi % wpr) * 32 + lane;
    if (!((wd >> lane) & 1u)) continue;
    if (f < a.bf0 || f >= a.bf1 || s < a.bs0 || s >= a.bs1) continue;  // halo cell: not ours to fill
    ++masked;
    const int iy = a.win.fast_is_y ? f : s, ix = a.win.fast_is_y ? s : f;
    int best_i = -1, best_j = -1, level = 0;
    bool unknown = false;
    while (best_i < 0 && !unknown) {
      ++level;                                                     // :110
      if (level > level_cap) { raise_status(a.status, GI_ERR_ITER_CAP); break; }
      // annulus (L-0.5)^2 <= d2 < (L+0.5)^2 over integers: L*L-L < d2 <= L*L+L   (:111-112,117-118)
      const int lo = level * level - level, hi = level * level + level;
      int best_d2 = 0x7fffffff;
      for (int di = -level; di <= level; ++di) {                   // rows ascending (:113)
        const int i2 = iy + di;
        if (i2 < 0 || i2 >= a.len_y) continue;
        const int rem_hi = hi - di * di;
        if (rem_hi < 0) continue;
        const int dj_max = isqrt_floor(rem_hi);
        const int rem_lo = lo - di * di;                           // need dj*dj > rem_lo
        const int dj_min = rem_lo < 0 ? 0 : isqrt_floor(rem_lo) + 1;
        for (int pass = 0; pass < 2; ++pass) {                     // columns ascending (:114)
          for (int k = (pass == 0 ? dj_max : dj_min); pass == 0 ? k >= dj_min : k <= dj_max;
               k += (pass == 0 ? -1 : 1)) {
            if (pass == 1 && k == 0) continue;                     // dj = 0 handled in pass 0
            const int j2 = ix + (pass == 0 ? -k : k);
            if (j2 < 0 || j2 >= a.len_x) continue;
            const int st = fill_cell_state(a, wpr, i2, j2);
            if (st == 2) { unknown = true; continue; }
            if (st == 0) {
              const int d2 = di * di + k * k;
              if (d2 < best_d2) { best_d2 = d2; best_i = i2; best_j = j2; }  // strict '<' (:119)
            }
          }
        }
      }
      // a level is only trustworthy if none of its annulus cells was outside the walked window
      if (unknown) { raise_status(a.status, GI_ERR_HALO); best_i = -1; }
    }
    level_max = max(level_max, level);
    if (best_i < 0) continue;
    const int sf = a.win.fast_is_y ? best_i : best_j, ss = a.win.fast_is_y ? best_j : best_i;
    R v;
    if (sf >= a.vf0 && sf < a.vf1 && ss >= a.vs0 && ss < a.vs1) {
      v = a.out[(int64_t)(ss - a.vs0) * a.out_ld + (sf - a.vf0)];
    } else {  // source cell lies in the halo: recompute its (deterministic) value
      const int64_t hi = a.hfix.index(best_j, best_i);
      v = hi >= 0 ? a.hfix.vals[hi] : R(0);
      if (hi < 0 || is_no_fix(v)) {   // (a fix-up left no value: any walk finds this target's triangle)
        const R tx = __ldg((a.win.fast_is_y ? a.win.slow_axis : a.win.fast_axis) + best_j);
        const R ty = __ldg((a.win.fast_is_y ? a.win.fast_axis : a.win.slow_axis) + best_i);
        bool inside; int iters; TriGeom<R> g;
        const int tri = walk(a.geom, seed_lookup(a.seed, a.ncx, a.ncy, best_j, best_i, a.num_tri), tx, ty,
                             a.iter_cap, inside, iters, g, a.status);
        (void)tri;
        v = eval_any(g, tx, ty);
      }
    }
    const int64_t oo = (int64_t)(s - a.vs0) * a.out_ld + (f - a.vf0);
    a.out[oo] = v;  // :121
    for (int d = 0; d < a.n_peer; ++d) a.peer_out[d][oo] = v;
    if (a.fill_src)
      a.fill_src[(int64_t)(s - a.bs0) * (a.bf1 - a.bf0) + (f - a.bf0)] = (int64_t)best_i * a.len_x + best_j;
  }
  for (int off = 16; off > 0; off >>= 1) {
    masked += __shfl_down_sync(0xffffffffu, masked, off);
    level_max = max(level_max, __shfl_down_sync(0xffffffffu, level_max, off));
  }
  if (lane == 0) {
    if (masked) atomicAdd((unsigned long long*)(a.counters + 1), (unsigned long long)masked);
    if (level_max) atomicMax((unsigned long long*)(a.counters + 2), (unsigned long long)level_max);
  }
}

__global__ void expand_mask_kernel(const uint32_t* __restrict__ mask_words, int wpr, int e_f0, int e_s0,
                                   int bf0, int bf1, int bs0, int bs1, uint8_t* __restrict__ mask_out,
                                   int64_t* __restrict__ fill_src) {
  const int bnf = bf1 - bf0;
  const int64_t n = (int64_t)bnf * (bs1 - bs0);
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    const int s = bs0 + (int)(o / bnf), f = bf0 + (int)(o % bnf);
    const int lf = f - e_f0;
    const uint32_t wd = mask_words[(int64_t)(s - e_s0) * wpr + (lf >> 5)];
    if (mask_out) mask_out[o] = (wd >> (lf & 31)) & 1u;
    if (fill_src) fill_src[o] = -1;
  }
}

// standalone fill on a caller-supplied byte mask: pack it into the bit mask first
__global__ void bytes_to_bits_kernel(const uint8_t* __restrict__ mask, int nf, int wpr,
                                     uint32_t* __restrict__ words, int* __restrict__ nz_words,
                                     int64_t* __restrict__ counters) {
  const int lane_f = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  const bool m = lane_f < nf && mask[(int64_t)s * nf + lane_f] != 0;
  const unsigned b = __ballot_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && (lane_f >> 5) < wpr) {
    const int64_t wi = (int64_t)s * wpr + (lane_f >> 5);
    words[wi] = b;
    if (b) nz_words[atomicAdd((unsigned long long*)(counters + 4), 1ull)] = (int)wi;
  }
}

template <class R> struct PackedMesh {
  TriGeom<R>* geom;
  int* seed;
  int ncx, ncy, num_tri;
  RawMesh<R> raw;  // raw.use = 1: only the triangles near the row window have records
  const uint8_t* need;  // band-limited pack: which triangles have records (nullptr: all)
  TileBins bins;        // rasteriser: tile lists of the rows [bin_e0, bin_e1) (start == nullptr: none)
  int bin_e0, bin_e1;
  int plane;            // GI_BARY_PLANE_VALUES: records hold planes (kPlane) where the triangle's shape allows
};

// Rasteriser or walk?  Measured on C4 (profiles/r02_raster_*): the rasteriser needs 3.4 ms where the walk needs
// 3.1 -- testing every node of every bounding box (2.4 tests per target at 20 of 32 lanes) costs more instructions
// than the walk's 1.6 iterations per target -- so the walk stays the default and GI_BARY_RASTER=1 selects the
// rasteriser (the whole GPU test suite passes either way).
template <class R> bool use_raster(int num_tri, int64_t nodes) {
  static const int forced = getenv("GI_BARY_RASTER") ? atoi(getenv("GI_BARY_RASTER")) : 0;
  (void)num_tri; (void)nodes;
  return forced != 0;
}

// rows [w0, w1) +- this many rows are packed by a band-limited pack (GI_BARY_BAND_PACK)
constexpr int kPackMarginRows = 48;

// 1. pack the mesh (device pointers; records and seed grid live in `scr`)
template <class R>
int bary_pack(Scratch& scr, const R* points, int num_points, const R* data_in, const R* x_axis, int len_x,
              const R* y_axis, int len_y, const int32_t* simplices, const int32_t* neighbours, int num_tri, int flags,
              int* status, PackedMesh<R>* pm, int w0 = 0, int w1 = -1) {
  cudaStream_t st = scr.stream();
  GI_TRY(scr.alloc(&pm->geom, (size_t)num_tri));
  pm->ncx = (len_x + kSeedCell - 1) / kSeedCell;
  pm->ncy = (len_y + kSeedCell - 1) / kSeedCell;
  pm->num_tri = num_tri;
  GI_TRY(scr.alloc(&pm->seed, (size_t)pm->ncx * pm->ncy + 1));  // + the fallback seed
  GI_CUDA(cudaMemsetAsync(pm->seed, 0x7f, sizeof(int) * ((size_t)pm->ncx * pm->ncy + 1), st));
  const PointsView<R> pv = points_n2(points, num_points, flags & GI_POINTS_ROWMAJOR);
  const TopoView sv = topo_t3(simplices, num_tri, flags & GI_TOPO_ROWMAJOR);
  pm->raw = RawMesh<R>{pv, sv, num_points, 0};
  pm->plane = (flags & GI_BARY_PLANE_VALUES) ? 1 : 0;
  uint8_t* need = nullptr;
  // band-limited pack: rows [w0, w1) +- kPackMarginRows, when that is a real subset of the grid
  if (w1 < 0) w1 = len_y;
  const int p0 = std::max(0, w0 - kPackMarginRows), p1 = std::min(len_y, w1 + kPackMarginRows);
  if ((flags & GI_BARY_BAND_PACK) && (p0 > 0 || p1 < len_y) && p0 < p1) {
    // (need_kernel takes the row numbers and reads the axis itself: no host copy of the axis, no synchronisation)
    GI_TRY(scr.alloc(&need, (size_t)num_tri));
    need_kernel<R><<<div_up(num_tri, 256), 256, 0, st>>>(pv, sv, num_points, num_tri, y_axis, p0, p1 - 1, len_y, need);
    GI_CUDA(cudaGetLastError());
    pm->raw.use = 1;
  }
  pm->need = need;
  // rasteriser: the tile lists of the walked rows [w0, w1) are counted by the pack pass
  pm->bins = TileBins{};
  pm->bin_e0 = w0; pm->bin_e1 = w1;
  TargetWindow<R> bw{};
  int n_threads = num_tri;
  if (w0 < w1 && !pm->plane && use_raster<R>(num_tri, (int64_t)len_x * (w1 - w0))) {
    bw = (flags & GI_FIELD_ROWMAJOR) ? TargetWindow<R>{x_axis, y_axis, len_x, len_y, 0, len_x, w0, w1, 0}
                                     : TargetWindow<R>{y_axis, x_axis, len_y, len_x, w0, w1, 0, len_x, 1};
    GI_TRY(alloc_tile_bins<R>(scr, num_tri, bw, &pm->bins));
    n_threads = std::max(num_tri, std::max(len_x, len_y));
  }
  const TopoView nv = topo_t3(neighbours, num_tri, flags & GI_TOPO_ROWMAJOR);
  if (pm->bins.start)
    pack_mesh_kernel<R, true><<<div_up(n_threads, 256), 256, 0, st>>>(pv, data_in, sv, nv, num_points, num_tri, pm->geom,
                                                                     pm->seed, x_axis, len_x, y_axis, len_y, need, bw,
                                                                     pm->bins, (R)sliver_guard_quality(), pm->plane, status);
  else
    pack_mesh_kernel<R, false><<<div_up(n_threads, 256), 256, 0, st>>>(pv, data_in, sv, nv, num_points, num_tri, pm->geom,
                                                                      pm->seed, x_axis, len_x, y_axis, len_y, need, bw,
                                                                      pm->bins, (R)sliver_guard_quality(), pm->plane, status);
  GI_CUDA(cudaGetLastError());
  if (pm->bins.start) GI_TRY(finish_tile_bins(scr, num_tri, pm->bins));
  return GI_OK;
}

// 2. + 2b. locate / evaluate / fix-up over the window wa.win (mask words, work list and counters as set in wa)
template <class R>
int bary_walk(const WalkArgs<R>& wa_in, const PackedMesh<R>& pm, const R* x_axis, const R* y_axis, int len_x, int len_y,
              int num_points, cudaStream_t st, const TileBins* bins = nullptr, int tile_row0 = 0) {
  WalkArgs<R> wa = wa_in;
  wa.need = pm.need;
  const int nf = wa.win.nf(), ns = wa.win.ns();
  // 128 columns x 32 rows per CTA, 64 registers, >= 8 CTAs per SM: the best of the shapes measured in round 1
  // (profiles/r01_walk_variants.md)
  constexpr int kRows = 32;
  static_assert(kRows == kTileH, "the rasteriser's tiles are the walk kernel's strips");
  dim3 grid(div_up(nf, kWalkThreads), div_up(ns, kRows));
  GI_REQUIRE(grid.y <= 65535, "window too tall for one launch");
  const bool fy = wa.win.fast_is_y, peers = wa.n_peer > 0;
  wa.only_if = nullptr;
  if (bins) {
    if (!fy && !peers) raster_eval_kernel<R, false, false><<<grid, kWalkThreads, 0, st>>>(wa, *bins, tile_row0);
    else if (fy && !peers) raster_eval_kernel<R, true, false><<<grid, kWalkThreads, 0, st>>>(wa, *bins, tile_row0);
    else if (!fy) raster_eval_kernel<R, false, true><<<grid, kWalkThreads, 0, st>>>(wa, *bins, tile_row0);
    else raster_eval_kernel<R, true, true><<<grid, kWalkThreads, 0, st>>>(wa, *bins, tile_row0);
    GI_CUDA(cudaGetLastError());
    wa.only_if = bins->bad;  // descending axes / list overflow: the walk kernel does the window after all
    const int g = std::min<int64_t>((int64_t)grid.x * grid.y, kNumSMs * 8);
    if (!fy && !peers) locate_eval_standby_kernel<R, false, false, kRows><<<g, kWalkThreads, 0, st>>>(wa, grid.x, grid.y);
    else if (fy && !peers) locate_eval_standby_kernel<R, true, false, kRows><<<g, kWalkThreads, 0, st>>>(wa, grid.x, grid.y);
    else if (!fy) locate_eval_standby_kernel<R, false, true, kRows><<<g, kWalkThreads, 0, st>>>(wa, grid.x, grid.y);
    else locate_eval_standby_kernel<R, true, true, kRows><<<g, kWalkThreads, 0, st>>>(wa, grid.x, grid.y);
  } else {
    const bool bp = pm.need != nullptr;  // band-limited pack: links may lead to triangles without a record
    // careful exit-edge choice (CF): the sliver triangles along the hull cost the first / last strips of a window long
    // dependent walks without it (band of 2048 rows: 1.25 instead of 0.43 ms); in a window of many CTA waves those
    // strips are not the critical path and the test is 2 % of the kernel, so it is compiled out there
    const bool cf = (int64_t)grid.x * grid.y < (int64_t)kNumSMs * 8 * 24;
    // Small windows (the reference's own test size: 132 x 79 targets on 10 k triangles) are a handful of CTAs whose
    // threads each walk a dependent chain of 32 rows -- ~1.5 record fetches per row at L2 latency.  Shorter strips
    // give more, shorter chains (each pays its own seed walk): 132 x 79 targets 0.135 -> 0.082 ms for walk + fill +
    // copy-out with strips of 4 rows, 586 x 347 targets 0.19 -> 0.13 ms (profiles/r02_c1_rows_probe.txt).
    int small_rows = 0;
    if (!peers && !bp) {
      if ((int64_t)grid.x * div_up(ns, 8) <= kNumSMs * 4) small_rows = 4;
      else if ((int64_t)grid.x * div_up(ns, 16) <= kNumSMs * 4) small_rows = 8;
    }
#define GI_LAUNCH_SMALL(FY_, ROWS_)                                                                                  \
  do {                                                                                                               \
    dim3 g2(grid.x, div_up(ns, ROWS_));                                                                              \
    if (pm.plane) locate_eval_kernel<R, FY_, false, ROWS_, 4, false, true, true><<<g2, kWalkThreads, 0, st>>>(wa);  \
    else locate_eval_kernel<R, FY_, false, ROWS_, 4, false, true, false><<<g2, kWalkThreads, 0, st>>>(wa);          \
  } while (0)
    if (small_rows == 4 && !peers && !bp) {
      if (fy) GI_LAUNCH_SMALL(true, 4); else GI_LAUNCH_SMALL(false, 4);
    } else if (small_rows == 8 && !peers && !bp) {
      if (fy) GI_LAUNCH_SMALL(true, 8); else GI_LAUNCH_SMALL(false, 8);
    } else {
#define GI_LAUNCH_WALK3(FY_, PEERS_, BP_, CF_)                                                                    \
  do {                                                                                                            \
    if (pm.plane) {                                                                                               \
      locate_eval_kernel<R, FY_, PEERS_, kRows, 8, BP_, CF_, true><<<grid, kWalkThreads, 0, st>>>(wa);            \
    } else {                                                                                                      \
      bool launched = false;                                                                                      \
      if constexpr (walk_lean(FY_, PEERS_, BP_, CF_)) {   /* (not with debug outputs: they include iters_tot) */  \
        if (wa.tri_out == nullptr) {                                                                              \
          locate_eval_kernel<R, FY_, PEERS_, kRows, 9, BP_, CF_, false><<<grid, kWalkThreads, 0, st>>>(wa);       \
          launched = true;                                                                                        \
        }                                                                                                         \
      }                                                                                                           \
      if (!launched) locate_eval_kernel<R, FY_, PEERS_, kRows, 8, BP_, CF_, false><<<grid, kWalkThreads, 0, st>>>(wa); \
    }                                                                                                             \
  } while (0)
#define GI_LAUNCH_WALK2(FY_, PEERS_, BP_)           \
  do {                                              \
    if (cf) GI_LAUNCH_WALK3(FY_, PEERS_, BP_, true);  \
    else GI_LAUNCH_WALK3(FY_, PEERS_, BP_, false);    \
  } while (0)
#define GI_LAUNCH_WALK(FY_, PEERS_)                     \
  do {                                                  \
    if (bp) GI_LAUNCH_WALK2(FY_, PEERS_, true);         \
    else GI_LAUNCH_WALK2(FY_, PEERS_, false);           \
  } while (0)
    if (!fy && !peers) GI_LAUNCH_WALK(false, false);
    else if (fy && !peers) GI_LAUNCH_WALK(true, false);
    else if (!fy) GI_LAUNCH_WALK(false, true);
    else GI_LAUNCH_WALK(true, true);
    }
#undef GI_LAUNCH_SMALL
#undef GI_LAUNCH_WALK
#undef GI_LAUNCH_WALK2
#undef GI_LAUNCH_WALK3
  }
  GI_CUDA(cudaGetLastError());
  // targets inside the margin band of an edge: follow the reference's own path (normally none)
  const int n_search = std::min(num_points, pm.num_tri);  // the reference searches triangles 1..num_points (:347)
  start_tri_kernel<R><<<kStartBlocks, 256, 0, st>>>(pm.geom, pm.raw, n_search, x_axis, y_axis, len_y, wa.counters,
                                                    wa.start_part, wa.reference_rows);
  fixup_kernel<R><<<kNumSMs * 8, 128, 0, st>>>(wa, len_x, len_y);
  // more flagged targets than the list holds, or GI_BARY_REFERENCE_ORDER: the reference's row-sequential walk
  {
    const int cols = wa.win.fast_is_y ? wa.win.s1 : wa.win.f1;            // a row is walked from column 0
    const int spec_cols = cols <= kRowSpecMaxCols ? cols : 0;
    // the attribute belongs to the device's copy of the kernel: set once per device (bit mask, race-free; a second
    // thread that sets it again does no harm)
    static std::atomic<unsigned long long> configured{0ull};
    int dev = 0;
    GI_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (!bit || !(configured.load(std::memory_order_acquire) & bit)) {
      GI_CUDA(cudaFuncSetAttribute(row_walk_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kRowSpecMaxCols * (int)sizeof(int)));
      configured.fetch_or(bit, std::memory_order_release);
    }
    row_walk_kernel<R><<<std::min(wa.win.fast_is_y ? nf : ns, kNumSMs * 16), kRowThreads, (size_t)spec_cols * sizeof(int),
                         st>>>(wa, spec_cols);
  }
  GI_CUDA(cudaGetLastError());
  static const bool trace = getenv("GI_BARY_TRACE") != nullptr;
  if (trace) {  // debugging aid (synchronises): what the fix-up paths had to do in this window
    int64_t c[8];
    unsigned long long rs[10], zero[10] = {};
    GI_CUDA(cudaMemcpyAsync(c, wa.counters, sizeof(c), cudaMemcpyDeviceToHost, st));
    GI_CUDA(cudaMemcpyFromSymbolAsync(rs, g_row_walk_stats, sizeof(rs), 0, cudaMemcpyDeviceToHost, st));
    GI_CUDA(cudaMemcpyToSymbolAsync(g_row_walk_stats, zero, sizeof(zero), 0, cudaMemcpyHostToDevice, st));
    GI_CUDA(cudaStreamSynchronize(st));
    fprintf(stderr, "[gridinterp] walk window %d x %d: iterations %lld, flagged %lld (list %d), fixed up %lld, start "
                    "triangle %lld; row walk: %llu rows, %llu pieces with a wrong guess, %llu repair walks (%llu "
                    "iterations) in %llu rounds, %llu iterations in the pieces; cycles: longest guess %llu, longest piece "
                    "%llu, repair %llu, row %llu\n",
            nf, ns, (long long)c[0], (long long)c[5], wa.amb_cap, (long long)c[3], (long long)c[7], rs[0], rs[1], rs[2],
            rs[3], rs[8], rs[4], rs[9], rs[5], rs[6], rs[7]);
  }
  return GI_OK;
}

template <class R> void fill_from_walk(FillArgs<R>& fa, const WalkArgs<R>& wa, const PackedMesh<R>& pm, int len_x,
                                       int len_y) {
  fa.len_x = len_x; fa.len_y = len_y;
  fa.geom = pm.geom; fa.seed = pm.seed; fa.ncx = pm.ncx; fa.ncy = pm.ncy; fa.num_tri = pm.num_tri;
  fa.iter_cap = wa.iter_cap; fa.status = wa.status; fa.hfix = wa.hfix;
}

// the HaloFix store of a band [row_begin,row_end) walked as [e0,e1): allocated only when there are halo rows
template <class R>
int alloc_halo_fix(Scratch& scr, int e0, int e1, int row_begin, int row_end, int len_x, HaloFix<R>* h) {
  *h = HaloFix<R>{nullptr, e0, row_begin, row_end, len_x};
  const int64_t rows = (int64_t)(e1 - e0) - (row_end - row_begin);
  if (rows <= 0) return GI_OK;
  GI_TRY(scr.alloc(&h->vals, (size_t)rows * len_x));
  GI_CUDA(cudaMemsetAsync(h->vals, 0xff, sizeof(R) * (size_t)rows * len_x, scr.stream()));
  return GI_OK;
}

// 2. + 3. for the row band [row_begin,row_end) with `halo` extra rows walked on each side, on a packed mesh
template <class R>
int bary_run_window(Scratch& scr, const PackedMesh<R>& pm, const R* x_axis, int len_x, const R* y_axis, int len_y,
                    int num_points, int row_begin, int row_end, int halo, R* data_ip, int32_t* tri_out,
                    uint8_t* mask_out, int64_t* fill_src, int64_t* counters_out, int flags, const StatusSlot& slot,
                    cudaStream_t st, Phases* ph, const GatherDest<R>* dest = nullptr) {
  const int num_tri = pm.num_tri;
  int64_t* counters; int* nz_words;
  GI_TRY(scr.alloc(&counters, 8));
  GI_CUDA(cudaMemsetAsync(counters, 0, sizeof(int64_t) * 8, st));

  // ---- 2. walk + barycentric over the band plus halo rows -------------------------------------
  ph->mark("walk");
  const int e0 = max(0, row_begin - halo), e1 = min(len_y, row_end + halo);
  const bool rowmajor = flags & GI_FIELD_ROWMAJOR;
  WalkArgs<R> wa;
  if (rowmajor) {  // fast axis = x, slow = y
    wa.win = TargetWindow<R>{x_axis, y_axis, len_x, len_y, 0, len_x, e0, e1, 0};
    wa.bf0 = 0; wa.bf1 = len_x; wa.bs0 = row_begin; wa.bs1 = row_end;
  } else {         // column-major field: fast axis = y, slow = x
    wa.win = TargetWindow<R>{y_axis, x_axis, len_y, len_x, e0, e1, 0, len_x, 1};
    wa.bf0 = row_begin; wa.bf1 = row_end; wa.bs0 = 0; wa.bs1 = len_x;
  }
  const int nf = wa.win.nf(), ns = wa.win.ns(), wpr = (nf + 31) >> 5;
  uint32_t* mask_words;
  GI_REQUIRE((int64_t)wpr * ns < (int64_t)0x7fffffff, "window too large");
  GI_TRY(scr.alloc(&mask_words, (size_t)wpr * ns));
  GI_TRY(scr.alloc(&nz_words, (size_t)wpr * ns));
  wa.nz_words = nz_words;
  wa.amb_cap = (int)std::min<int64_t>((int64_t)nf * ns, amb_list_limit());
  wa.reference_rows = (flags & GI_BARY_REFERENCE_ORDER) ? 1 : 0;
  GI_TRY(scr.alloc(&wa.amb_list, (size_t)wa.amb_cap));
  GI_TRY(scr.alloc(&wa.start_part, (size_t)2 * kStartBlocks + 1));
  GI_CUDA(cudaMemsetAsync(wa.start_part, 0, sizeof(unsigned long long) * (2 * kStartBlocks + 1), st));
  GI_TRY(scr.alloc(&wa.row_todo, (size_t)len_y));
  GI_CUDA(cudaMemsetAsync(wa.row_todo, 0, sizeof(int) * (size_t)len_y, st));
  GI_TRY(alloc_halo_fix<R>(scr, e0, e1, row_begin, row_end, len_x, &wa.hfix));
  wa.geom = pm.geom; wa.seed = pm.seed; wa.ncx = pm.ncx; wa.ncy = pm.ncy; wa.num_tri = num_tri;
  wa.iter_cap = num_tri + 16;
  wa.out = data_ip; wa.tri_out = tri_out; wa.mask_words = mask_words; wa.counters = counters;
  wa.out_ld = (dest && dest->ld) ? dest->ld : wa.bf1 - wa.bf0;
  wa.n_peer = dest ? dest->n_peer : 0;
  for (int d = 0; d < wa.n_peer; ++d) wa.peer_out[d] = dest->peer[d];
  wa.status = slot.dev;
  const bool raster = pm.bins.start && pm.bin_e0 == e0 && pm.bin_e1 == e1;
  GI_TRY(bary_walk<R>(wa, pm, x_axis, y_axis, len_x, len_y, num_points, st, raster ? &pm.bins : nullptr, 0));
  if (mask_out || fill_src) {
    expand_mask_kernel<<<kNumSMs * 8, 256, 0, st>>>(mask_words, wpr, wa.win.f0, wa.win.s0, wa.bf0, wa.bf1, wa.bs0,
                                                   wa.bs1, mask_out, fill_src);
    GI_CUDA(cudaGetLastError());
  }

  // ---- 3. hull fill ---------------------------------------------------------------------------
  ph->mark("fill");
  FillArgs<R> fa;
  fa.win = wa.win; fa.mask_words = mask_words; fa.nz_words = nz_words; fa.nz_shift = 0;
  fa.bf0 = wa.bf0; fa.bf1 = wa.bf1; fa.bs0 = wa.bs0; fa.bs1 = wa.bs1;
  fa.vf0 = wa.bf0; fa.vf1 = wa.bf1; fa.vs0 = wa.bs0; fa.vs1 = wa.bs1;
  fa.out = data_ip; fa.fill_src = fill_src; fa.counters = counters;
  fa.out_ld = wa.out_ld; fa.n_peer = wa.n_peer;
  for (int d = 0; d < wa.n_peer; ++d) fa.peer_out[d] = wa.peer_out[d];
  fill_from_walk(fa, wa, pm, len_x, len_y);
  fill_kernel<R><<<kNumSMs * 8, 128, 0, st>>>(fa);
  GI_CUDA(cudaGetLastError());
  if (counters_out)
    GI_CUDA(cudaMemcpyAsync(counters_out, counters, sizeof(int64_t) * 4, cudaMemcpyDeviceToDevice, st));
  return GI_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host-side orchestration (device pointers in, work enqueued on `st`)
// ---------------------------------------------------------------------------------------------
namespace {
template <class R>
int barycentric_device_impl(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x,
                            const R* y_axis, int len_y, const int32_t* simplices, const int32_t* neighbours,
                            int num_tri, int row_begin, int row_end, int halo, R* data_ip, int32_t* tri_out,
                            uint8_t* mask_out, int64_t* fill_src, int64_t* counters_out, int flags,
                            cudaStream_t st, const GatherDest<R>* dest) {
  GI_REQUIRE(num_points >= 3 && num_tri >= 1, "need at least 3 points and 1 triangle");
  GI_REQUIRE(len_x >= 1 && len_y >= 1, "empty axis");
  GI_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= len_y, "row band out of range");
  GI_REQUIRE(halo >= 0, "negative halo");
  GI_TRY(ensure_device());
  if (row_begin == row_end) return GI_OK;
  StatusSlot slot;
  GI_TRY(get_status_slot(st, &slot));
  Scratch scr(st);
  Phases ph(st);

  // ---- 1. pack the mesh ---------------------------------------------------------------------
  ph.mark("pack_mesh");
  PackedMesh<R> pm;
  GI_TRY(bary_pack<R>(scr, points, num_points, data_in, x_axis, len_x, y_axis, len_y, simplices, neighbours, num_tri,
                      flags, slot.dev, &pm, std::max(0, row_begin - halo), std::min(len_y, row_end + halo)));
  GI_TRY(bary_run_window<R>(scr, pm, x_axis, len_x, y_axis, len_y, num_points, row_begin, row_end, halo, data_ip,
                            tri_out, mask_out, fill_src, counters_out, flags, slot, st, &ph, dest));
  ph.finish();
  GI_TRY(flush_status(st, slot));
  return GI_OK;
}
}  // namespace

template <class R>
int barycentric_device(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x,
                       const R* y_axis, int len_y, const int32_t* simplices, const int32_t* neighbours,
                       int num_tri, int row_begin, int row_end, int halo, R* data_ip, int32_t* tri_out,
                       uint8_t* mask_out, int64_t* fill_src, int64_t* counters_out, int flags,
                       cudaStream_t st) {
  return barycentric_device_impl<R>(points, num_points, data_in, x_axis, len_x, y_axis, len_y, simplices, neighbours,
                                    num_tri, row_begin, row_end, halo, data_ip, tri_out, mask_out, fill_src,
                                    counters_out, flags, st, nullptr);
}

// Row band [row_begin, row_end) computed here and stored at its place in `num_fields` full (len_y, len_x) fields:
// fields[0] is local (the hull fill reads source cells from it), the others are the same field on other GPUs
// (peer-mapped).  The stores of the walk / fix-up / fill kernels are the gather.
template <class R>
int barycentric_gather_device(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x,
                              const R* y_axis, int len_y, const int32_t* simplices, const int32_t* neighbours,
                              int num_tri, int row_begin, int row_end, int halo, R* const* fields, int num_fields,
                              int flags, cudaStream_t st) {
  GI_REQUIRE(fields && num_fields >= 1 && num_fields <= GI_MAX_GATHER_FIELDS, "1 .. GI_MAX_GATHER_FIELDS fields");
  for (int d = 0; d < num_fields; ++d) GI_REQUIRE(fields[d] != nullptr, "null field pointer");
  const bool rowmajor = flags & GI_FIELD_ROWMAJOR;
  // element (iy, ix) of a full field: iy*len_x + ix (row-major) or iy + ix*len_y (column-major)
  const int64_t off = rowmajor ? (int64_t)row_begin * len_x : (int64_t)row_begin;
  GatherDest<R> dest;
  dest.ld = rowmajor ? len_x : len_y;
  dest.n_peer = num_fields - 1;
  for (int d = 1; d < num_fields; ++d) dest.peer[d - 1] = fields[d] + off;
  return barycentric_device_impl<R>(points, num_points, data_in, x_axis, len_x, y_axis, len_y, simplices, neighbours,
                                    num_tri, row_begin, row_end, halo, fields[0] + off, nullptr, nullptr, nullptr,
                                    nullptr, flags, st, &dest);
}

namespace {
// Row band [row_begin,row_end) (+ `halo` rows walked on each side) on a packed mesh: the walked window is cut
// into sub-bands along the axis that is not contiguous in memory; sub-band b is located / evaluated, then
// sub-band b-1 is hull-filled (its fill may look into sub-bands b-2 .. b, all walked by then) and copied back on
// a second stream while sub-band b+1 is being walked.  Nothing is computed twice.  A fill that would have to
// look further than one sub-band away (GI_ERR_HALO) makes the run start over with a single band (where
// GI_ERR_HALO then means what it says: the caller's halo is too small).  dout is the device twin of data_ip
// (host), both (row_end-row_begin, len_x) in the layout `flags` selects.  Returns the status (synchronous).
template <class R>
int bary_run_banded(HostPipe* pipe, Scratch& scr, const PackedMesh<R>& pm, const R* dx, int len_x, const R* dy,
                    int len_y, int num_points, int row_begin, int row_end, int halo, R* dout, R* data_ip, int flags,
                    const StatusSlot& slot) {
  cudaStream_t st = pipe->compute;
  const int num_tri = pm.num_tri;
  const bool rowmajor = flags & GI_FIELD_ROWMAJOR;
  const int e0 = max(0, row_begin - halo), e1 = min(len_y, row_end + halo);
  // walked window E and written band B in (fast, slow) coordinates
  const int ef0 = rowmajor ? 0 : e0, ef1 = rowmajor ? len_x : e1, es0 = rowmajor ? e0 : 0, es1 = rowmajor ? e1 : len_x;
  const int bf0 = rowmajor ? 0 : row_begin, bf1 = rowmajor ? len_x : row_end;
  const int bs0 = rowmajor ? row_begin : 0, bs1 = rowmajor ? row_end : len_x;
  const int nfast = ef1 - ef0, nslow = es1 - es0, bnf = bf1 - bf0;
  const int wpr = (nfast + 31) >> 5;
  GI_REQUIRE((int64_t)wpr * nslow < (int64_t)0x7fffffff, "window too large");
  uint32_t* mask_words; int* nz_words; int64_t* counters; int* amb_list;
  GI_TRY(scr.alloc(&mask_words, (size_t)wpr * nslow));
  GI_TRY(scr.alloc(&nz_words, (size_t)wpr * nslow));
  GI_TRY(scr.alloc(&counters, (size_t)8 * kMaxBands));
  const int amb_cap = (int)std::min<int64_t>((int64_t)nfast * nslow, amb_list_limit());
  GI_TRY(scr.alloc(&amb_list, (size_t)amb_cap));  // shared by the bands: band b's fix-up runs before band b+1's walk
  unsigned long long* start_part;
  GI_TRY(scr.alloc(&start_part, (size_t)2 * kStartBlocks + 1));
  GI_CUDA(cudaMemsetAsync(start_part, 0, sizeof(unsigned long long) * (2 * kStartBlocks + 1), st));
  int* row_todo;
  GI_TRY(scr.alloc(&row_todo, (size_t)len_y));
  GI_CUDA(cudaMemsetAsync(row_todo, 0, sizeof(int) * (size_t)len_y, st));

  HaloFix<R> hfix;
  GI_TRY(alloc_halo_fix<R>(scr, e0, e1, row_begin, row_end, len_x, &hfix));

  // tile lists of the whole walked window (from the pack pass): the sub-bands below start at whole tile rows
  const bool raster = pm.bins.start && pm.bin_e0 == e0 && pm.bin_e1 == e1;
  const TileBins& bins = pm.bins;

  int nb = plan_bands(nslow, (int64_t)(row_end - row_begin) * len_x * sizeof(R));
  for (;;) {
    GI_CUDA(cudaMemsetAsync(counters, 0, sizeof(int64_t) * 8 * kMaxBands, st));
    int64_t start[kMaxBands + 1];  // sub-band b = slow indices [start[b], start[b+1])
    for (int b = 0; b < nb; ++b) {
      int64_t e;
      band_range(es0, es1, nb, b, 32, &start[b], &e);
      start[b + 1] = e;
    }
    auto window = [&](int64_t s0, int64_t s1) {
      return rowmajor ? TargetWindow<R>{dx, dy, len_x, len_y, ef0, ef1, (int)s0, (int)s1, 0}
                      : TargetWindow<R>{dy, dx, len_y, len_x, ef0, ef1, (int)s0, (int)s1, 1};
    };
    // the part of sub-band b that belongs to the band (empty for pure halo pieces)
    auto band_lo = [&](int b) { return (int)std::min<int64_t>(std::max<int64_t>(start[b], bs0), bs1); };
    auto band_hi = [&](int b) { return (int)std::max<int64_t>(std::min<int64_t>(start[b + 1], bs1), band_lo(b)); };
    WalkArgs<R> was[kMaxBands];
    auto fill_band = [&](int b) -> int {
      const int lo = band_lo(b), hi = band_hi(b);
      if (lo >= hi) return GI_OK;
      const int64_t f0 = start[b > 0 ? b - 1 : 0], f1 = start[b + 1 < nb ? b + 2 : nb];
      FillArgs<R> fa;
      fa.win = window(f0, f1);
      fa.mask_words = mask_words + (f0 - es0) * wpr;
      fa.nz_words = nz_words + (start[b] - es0) * wpr;
      fa.nz_shift = (start[b] - f0) * wpr;
      fa.bf0 = bf0; fa.bf1 = bf1; fa.bs0 = lo; fa.bs1 = hi;
      // every band cell walked so far holds its final value
      fa.vf0 = bf0; fa.vf1 = bf1; fa.vs0 = bs0; fa.vs1 = (int)std::max<int64_t>(std::min<int64_t>(f1, bs1), bs0);
      fa.out = dout; fa.fill_src = nullptr; fa.counters = counters + 8 * b;
      fa.out_ld = bnf; fa.n_peer = 0;
      fill_from_walk(fa, was[b], pm, len_x, len_y);
      fill_kernel<R><<<kNumSMs * 8, 128, 0, st>>>(fa);
      GI_CUDA(cudaGetLastError());
      if (nb > 1) {  // (a single band needs no second stream: small calls are latency-bound)
        GI_CUDA(cudaEventRecord(pipe->computed[b], st));
        GI_CUDA(cudaStreamWaitEvent(pipe->copy_out, pipe->computed[b], 0));
      }
      const int64_t o = (int64_t)(lo - bs0) * bnf, cnt = (int64_t)(hi - lo) * bnf;
      return to_host(nb > 1 ? pipe->copy_out : st, data_ip + o, dout + o, (size_t)cnt);
    };
    for (int b = 0; b < nb; ++b) {
      if (start[b] < start[b + 1]) {
        WalkArgs<R>& wa = was[b];
        wa.win = window(start[b], start[b + 1]);
        wa.bf0 = bf0; wa.bf1 = bf1; wa.bs0 = band_lo(b); wa.bs1 = band_hi(b);
        wa.geom = pm.geom; wa.seed = pm.seed; wa.ncx = pm.ncx; wa.ncy = pm.ncy; wa.num_tri = num_tri;
        wa.iter_cap = num_tri + 16;
        wa.out = dout + (int64_t)(wa.bs0 - bs0) * bnf; wa.tri_out = nullptr;
        wa.out_ld = bnf; wa.n_peer = 0;
        wa.mask_words = mask_words + (start[b] - es0) * wpr; wa.nz_words = nz_words + (start[b] - es0) * wpr;
        wa.counters = counters + 8 * b;
        wa.amb_list = amb_list; wa.amb_cap = amb_cap; wa.start_part = start_part;
        wa.reference_rows = (flags & GI_BARY_REFERENCE_ORDER) ? 1 : 0; wa.row_todo = row_todo;
        wa.hfix = hfix;
        wa.status = slot.dev;
        GI_TRY(bary_walk<R>(wa, pm, dx, dy, len_x, len_y, num_points, st, raster ? &bins : nullptr,
                            (int)((start[b] - es0) / kTileH)));
      }
      if (b > 0 && start[b - 1] < start[b]) GI_TRY(fill_band(b - 1));
    }
    if (start[nb - 1] < start[nb]) GI_TRY(fill_band(nb - 1));
    GI_TRY(flush_status(st, slot));
    if (nb > 1) GI_CUDA(cudaStreamSynchronize(pipe->copy_out));
    const int status = read_status(st, true);
    if (status == GI_ERR_HALO && nb > 1) {  // a hull fill reached beyond the neighbouring band: one band sees all
      nb = 1;
      continue;
    }
    return status;
  }
}
}  // namespace

// GI_HOST form without debug outputs: inputs in, mesh packed once, banded run.
template <class R>
int barycentric_host(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
                     int len_y, const int32_t* simplices, const int32_t* neighbours, int num_tri, int row_begin,
                     int row_end, int halo, R* data_ip, int flags) {
  GI_REQUIRE(num_points >= 3 && num_tri >= 1, "need at least 3 points and 1 triangle");
  GI_REQUIRE(len_x >= 1 && len_y >= 1, "empty axis");
  GI_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= len_y && halo >= 0, "bad row band");
  GI_TRY(ensure_device());
  if (row_begin == row_end) return GI_OK;
  HostPipe* pipe;
  GI_TRY(host_pipe(&pipe));
  cudaStream_t st = pipe->compute;
  StatusSlot slot;
  GI_TRY(get_status_slot(st, &slot));
  Scratch scr(st);
  Phases ph(st);
  ph.mark("h2d");
  R *dp, *dd, *dx, *dy, *dout;
  int32_t *ds, *dn;
  GI_TRY(to_device(scr, points, (size_t)2 * num_points, &dp));
  GI_TRY(to_device(scr, data_in, (size_t)num_points, &dd));
  GI_TRY(to_device(scr, x_axis, (size_t)len_x, &dx));
  GI_TRY(to_device(scr, y_axis, (size_t)len_y, &dy));
  GI_TRY(to_device(scr, simplices, (size_t)3 * num_tri, &ds));
  GI_TRY(to_device(scr, neighbours, (size_t)3 * num_tri, &dn));
  GI_TRY(scr.alloc(&dout, (size_t)len_x * (row_end - row_begin)));
  // a band that is a small part of the grid packs only the triangles near it; a walk that would leave them
  // (GI_ERR_PACK_WINDOW) makes the call start over with the whole mesh (this path synchronises anyway)
  const int e0 = std::max(0, row_begin - halo), e1 = std::min(len_y, row_end + halo);
  int pack_flags = flags & ~GI_BARY_BAND_PACK;
  if ((int64_t)(e1 - e0 + 2 * kPackMarginRows) * 2 <= len_y) pack_flags |= GI_BARY_BAND_PACK;
  for (;;) {
    ph.mark("pack_mesh");
    PackedMesh<R> pm;
    GI_TRY(bary_pack<R>(scr, dp, num_points, dd, dx, len_x, dy, len_y, ds, dn, num_tri, pack_flags, slot.dev, &pm, e0, e1));
    ph.mark("walk");
    const int status = bary_run_banded<R>(pipe, scr, pm, dx, len_x, dy, len_y, num_points, row_begin, row_end, halo,
                                          dout, data_ip, flags, slot);
    if (status == GI_ERR_PACK_WINDOW && (pack_flags & GI_BARY_BAND_PACK)) {
      pack_flags &= ~GI_BARY_BAND_PACK;
      continue;
    }
    ph.finish();
    return status;
  }
}

// ---------------------------------------------------------------------------------------------
// plan: packed mesh + target grid resident on the device, one execute() per field of nodal values
// ---------------------------------------------------------------------------------------------
namespace {
template <class R> struct BaryPlan : PlanBase {
  Scratch keep{nullptr, true};
  PackedMesh<R> pm;
  R *dx, *dy;
  int32_t* simp;  // device copy of simplices (for new nodal values)
  R* denom;       // the records' denominators (store_values)
  int num_points, len_x, len_y, row_begin, row_end, halo, flags;

  int execute(const void* data_in_v, void* data_ip_v, int mem, cudaStream_t st_in) override {
    const R* data_in = static_cast<const R*>(data_in_v);
    R* data_ip = static_cast<R*>(data_ip_v);
    if (row_begin == row_end) return GI_OK;
    HostPipe* pipe = nullptr;
    cudaStream_t st = st_in;
    if (mem == GI_HOST) { GI_TRY(host_pipe(&pipe)); st = pipe->compute; }
    StatusSlot slot;
    GI_TRY(get_status_slot(st, &slot));
    Scratch scr(st);
    Phases ph(st);
    const R* dd = data_in;
    R* dout = data_ip;
    const size_t band = (size_t)(row_end - row_begin) * len_x;
    if (mem == GI_HOST) {
      ph.mark("h2d");
      R* tmp;
      GI_TRY(to_device(scr, data_in, (size_t)num_points, &tmp));
      dd = tmp;
      GI_TRY(scr.alloc(&dout, band));
    }
    ph.mark("pack_values");
    pack_values_kernel<R><<<div_up(pm.num_tri, 256), 256, 0, st>>>(dd, topo_t3(simp, pm.num_tri, flags & GI_TOPO_ROWMAJOR),
                                                                  num_points, pm.num_tri, denom, pm.geom, pm.plane);
    GI_CUDA(cudaGetLastError());
    ph.mark("walk");
    if (mem == GI_HOST) {
      const int status = bary_run_banded<R>(pipe, scr, pm, dx, len_x, dy, len_y, num_points, row_begin, row_end, halo,
                                            dout, data_ip, flags, slot);
      ph.finish();
      return status;
    }
    GI_TRY(bary_run_window<R>(scr, pm, dx, len_x, dy, len_y, num_points, row_begin, row_end, halo, dout, nullptr,
                              nullptr, nullptr, nullptr, flags, slot, st, &ph));
    ph.finish();
    return flush_status(st, slot);
  }
};
}  // namespace

template <class R>
int barycentric_plan_create(const R* points, int num_points, const R* x_axis, int len_x, const R* y_axis, int len_y,
                            const int32_t* simplices, const int32_t* neighbours, int num_tri, int row_begin,
                            int row_end, int halo, int flags, int mem, cudaStream_t st_in, PlanBase** out) {
  GI_REQUIRE(num_points >= 3 && num_tri >= 1, "need at least 3 points and 1 triangle");
  GI_REQUIRE(len_x >= 1 && len_y >= 1, "empty axis");
  GI_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= len_y && halo >= 0, "bad row band");
  GI_TRY(ensure_device());
  cudaStream_t st = st_in;
  if (mem == GI_HOST) { HostPipe* pipe; GI_TRY(host_pipe(&pipe)); st = pipe->compute; }
  StatusSlot slot;
  GI_TRY(get_status_slot(st, &slot));
  std::unique_ptr<BaryPlan<R>> plan(new BaryPlan<R>());
  plan->real_bytes = sizeof(R);
  GI_CUDA(cudaGetDevice(&plan->device));
  plan->num_points = num_points; plan->len_x = len_x; plan->len_y = len_y;
  plan->row_begin = row_begin; plan->row_end = row_end; plan->halo = halo; plan->flags = flags;
  Scratch tmp(st);
  // bary_pack launches on its scratch's stream: the plan's creation stream, not the legacy default stream (the
  // pipeline streams are non-blocking, so work on stream 0 is not ordered with them)
  plan->keep.set_stream(st);
  const cudaMemcpyKind kind = mem == GI_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  GI_TRY(plan->keep.alloc(&plan->dx, (size_t)len_x));
  GI_TRY(plan->keep.alloc(&plan->dy, (size_t)len_y));
  GI_TRY(plan->keep.alloc(&plan->simp, (size_t)3 * num_tri));
  GI_CUDA(cudaMemcpyAsync(plan->dx, x_axis, sizeof(R) * (size_t)len_x, kind, st));
  GI_CUDA(cudaMemcpyAsync(plan->dy, y_axis, sizeof(R) * (size_t)len_y, kind, st));
  GI_CUDA(cudaMemcpyAsync(plan->simp, simplices, sizeof(int32_t) * (size_t)3 * num_tri, kind, st));
  const R* dp = points;
  const int32_t* dn = neighbours;
  R* zeros;  // nodal values are supplied by execute(): pack with zeros
  GI_TRY(tmp.alloc(&zeros, (size_t)num_points));
  GI_CUDA(cudaMemsetAsync(zeros, 0, sizeof(R) * (size_t)num_points, st));
  if (mem == GI_HOST) {
    R* t; int32_t* tn;
    GI_TRY(to_device(tmp, points, (size_t)2 * num_points, &t));
    GI_TRY(to_device(tmp, neighbours, (size_t)3 * num_tri, &tn));
    dp = t; dn = tn;
  }
  GI_TRY(bary_pack<R>(plan->keep, dp, num_points, zeros, plan->dx, len_x, plan->dy, len_y, plan->simp, dn, num_tri,
                      flags, slot.dev, &plan->pm, std::max(0, row_begin - halo), std::min(len_y, row_end + halo)));
  GI_TRY(plan->keep.alloc(&plan->denom, (size_t)num_tri));
  copy_denom_kernel<R><<<div_up(num_tri, 256), 256, 0, st>>>(plan->pm.geom, num_tri, plan->denom);
  GI_CUDA(cudaGetLastError());
  GI_TRY(flush_status(st, slot));
  const int status = read_status(st, true);  // invalid ids are reported at creation
  if (status != GI_OK) return status;
  *out = plan.release();
  return GI_OK;
}

template <class R>
int fill_device(const uint8_t* mask, int len_x, int len_y, R* data_ip, int64_t* fill_src, int flags,
                cudaStream_t st) {
  GI_REQUIRE(len_x >= 1 && len_y >= 1, "empty grid");
  GI_TRY(ensure_device());
  StatusSlot slot;
  GI_TRY(get_status_slot(st, &slot));
  Scratch scr(st);
  const bool rowmajor = flags & GI_FIELD_ROWMAJOR;
  FillArgs<R> fa;
  fa.win = rowmajor ? TargetWindow<R>{nullptr, nullptr, len_x, len_y, 0, len_x, 0, len_y, 0}
                    : TargetWindow<R>{nullptr, nullptr, len_y, len_x, 0, len_y, 0, len_x, 1};
  const int nf = fa.win.nf(), ns = fa.win.ns(), wpr = (nf + 31) >> 5;
  uint32_t* words; int64_t* counters; int* nz_words;
  GI_TRY(scr.alloc(&words, (size_t)wpr * ns));
  GI_TRY(scr.alloc(&nz_words, (size_t)wpr * ns));
  GI_TRY(scr.alloc(&counters, 8));
  GI_CUDA(cudaMemsetAsync(counters, 0, sizeof(int64_t) * 8, st));
  GI_REQUIRE(ns <= 65535, "grid too large for the standalone fill");
  bytes_to_bits_kernel<<<dim3(div_up(nf, 128), ns), 128, 0, st>>>(mask, nf, wpr, words, nz_words, counters);
  GI_CUDA(cudaGetLastError());
  if (fill_src) {
    expand_mask_kernel<<<kNumSMs * 8, 256, 0, st>>>(words, wpr, 0, 0, 0, nf, 0, ns, nullptr, fill_src);
    GI_CUDA(cudaGetLastError());
  }
  fa.mask_words = words; fa.nz_words = nz_words; fa.nz_shift = 0; fa.bf0 = 0; fa.bf1 = nf; fa.bs0 = 0; fa.bs1 = ns;
  fa.vf0 = 0; fa.vf1 = nf; fa.vs0 = 0; fa.vs1 = ns;
  fa.out = data_ip; fa.fill_src = fill_src; fa.counters = counters; fa.len_x = len_x; fa.len_y = len_y;
  fa.out_ld = nf; fa.n_peer = 0;
  fa.geom = nullptr; fa.seed = nullptr; fa.ncx = fa.ncy = fa.num_tri = fa.iter_cap = 0;
  fa.hfix = HaloFix<R>{nullptr, 0, 0, 0, 0};
  fa.status = slot.dev;
  fill_kernel<R><<<kNumSMs * 8, 128, 0, st>>>(fa);
  GI_CUDA(cudaGetLastError());
  GI_TRY(flush_status(st, slot));
  return GI_OK;
}

// ---------------------------------------------------------------------------------------------
// Scattered targets (the reference interpolates to the nodes of a regular grid only,
// interpolation.f90:365-371).  Same search structure -- packed records, a seed grid laid over the bounding box of the
// mesh -- same visibility walk (triangle_walk.f90:24-84) and the same weights (interpolation.f90:385-398) per target.
// A target outside the convex hull takes the value of the CLOSEST LOCATION ON THE MESH, which is what the reference's
// header comment promises (interpolation.f90:292-293) and its grid version approximates by the nearest node inside
// the hull: from the hull edge the walk left through, the hull is followed edge by edge (rotating about the shared
// hull vertex through the neighbour links) until the foot of the perpendicular falls on an edge or a vertex is
// closest; the weights are evaluated there.
// ---------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ unsigned long long ordered_key(double v) {  // monotone map double -> uint64
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double from_ordered_key(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double d;
#ifdef __CUDA_ARCH__
  d = __longlong_as_double((long long)u);
#else
  memcpy(&d, &u, sizeof d);
#endif
  return d;
}
// box[0..3] = keys of min x, min y, max x, max y (initialised to ~0, ~0, 0, 0)
template <class R>
__global__ void __launch_bounds__(256) points_box_kernel(PointsView<R> pts, int n, unsigned long long* box) {
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
    atomicMin(box + 0, ordered_key(x0)); atomicMin(box + 1, ordered_key(y0));
    atomicMax(box + 2, ordered_key(x1)); atomicMax(box + 3, ordered_key(y1));
  }
}
// virtual axes over the box: len nodes each (only the seed grid is built on them)
template <class R>
__global__ void virtual_axes_kernel(const unsigned long long* box, int len, R* xs, R* ys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  const double x0 = from_ordered_key(box[0]), y0 = from_ordered_key(box[1]);
  const double x1 = from_ordered_key(box[2]), y1 = from_ordered_key(box[3]);
  const double f = len > 1 ? (double)i / (double)(len - 1) : 0.0;
  xs[i] = (R)(x0 + (x1 - x0) * f);
  ys[i] = (R)(y0 + (y1 - y0) * f);
}

template <class R> struct TriView {  // vertices and neighbours of a record by index 0..2
  R x[3], y[3];
  int nb[3];  // neighbour opposite vertex i
  __device__ __forceinline__ void load(const TriGeom<R>* p) {
    TriGeom<R> g;
    load_walk(p, g);
    x[0] = g.x1; y[0] = g.y1; x[1] = g.x2; y[1] = g.y2; x[2] = g.x3; y[2] = g.y3;
    nb[0] = g.n1; nb[1] = g.n2; nb[2] = g.n3;
  }
};

// Closest point of the hull to (px, py), which left the mesh through hull edge e (vertices e, e+1) of triangle t.
template <class R>
__device__ void closest_on_hull(const TriGeom<R>* __restrict__ geom, int& t, int e, R px, R py, int cap, R& qx, R& qy,
                                int* status) {
  int dir = 0;  // +1: moving on past vertex b, -1: moving back past vertex a
  for (int it = 0;; ++it) {
    TriView<R> v;
    v.load(geom + t);
    const int ia = e, ib = (e + 1) % 3;
    const R abx = v.x[ib] - v.x[ia], aby = v.y[ib] - v.y[ia];
    const R len2 = abx * abx + aby * aby;
    const R tt = len2 > R(0) ? ((px - v.x[ia]) * abx + (py - v.y[ia]) * aby) / len2 : R(0);
    if (it >= cap) { raise_status(status, GI_ERR_ITER_CAP); qx = v.x[ia]; qy = v.y[ia]; return; }
    if (tt < R(0)) {
      if (dir > 0) { qx = v.x[ia]; qy = v.y[ia]; return; }  // the vertex between the two edges is closest
      dir = -1;
      const R ax = v.x[ia], ay = v.y[ia];
      int k = (e + 2) % 3;  // the edge entering a in t
      while (v.nb[(k + 2) % 3] >= 0) {
        t = v.nb[(k + 2) % 3];
        v.load(geom + t);
        const int j = (v.x[0] == ax && v.y[0] == ay) ? 0 : ((v.x[1] == ax && v.y[1] == ay) ? 1 : 2);
        k = (j + 2) % 3;
      }
      e = k;
    } else if (tt > R(1)) {
      if (dir < 0) { qx = v.x[ib]; qy = v.y[ib]; return; }
      dir = 1;
      const R bx = v.x[ib], by = v.y[ib];
      int k = (e + 1) % 3;  // the edge leaving b in t
      while (v.nb[(k + 2) % 3] >= 0) {
        t = v.nb[(k + 2) % 3];
        v.load(geom + t);
        const int j = (v.x[0] == bx && v.y[0] == by) ? 0 : ((v.x[1] == bx && v.y[1] == by) ? 1 : 2);
        k = j;
      }
      e = k;
    } else {
      qx = v.x[ia] + tt * abx; qy = v.y[ia] + tt * aby;
      return;
    }
  }
}

template <class R>
__global__ void __launch_bounds__(128)
bary_points_kernel(const TriGeom<R>* __restrict__ geom, const int* __restrict__ seed, int ncx, int ncy, int num_tri,
                   const R* __restrict__ xs, const R* __restrict__ ys, int vlen, PointsView<R> targets, int64_t m,
                   R* __restrict__ out, int32_t* __restrict__ tri_out, uint8_t* __restrict__ outside_out, int* status) {
  const R margin = Consts<R>::margin();
  const R x0 = __ldg(xs), y0 = __ldg(ys), ex = __ldg(xs + vlen - 1) - x0, ey = __ldg(ys + vlen - 1) - y0;
  const R sx = ex > R(0) ? R(vlen - 1) / ex : R(0), sy = ey > R(0) ? R(vlen - 1) / ey : R(0);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    R tx, ty;
    targets.xy(i, tx, ty);
    const int ix = (int)fmin(fmax((tx - x0) * sx, R(0)), R(vlen - 1));
    const int iy = (int)fmin(fmax((ty - y0) * sy, R(0)), R(vlen - 1));
    int tri = seed_lookup(seed, ncx, ncy, ix, iy, num_tri);
    int edge = -1;
    TriGeom<R> g;
    for (int it = 0;; ++it) {                                            // triangle_walk.f90:44-84
      load_walk(geom + tri, g);
      const R c1 = (g.x2 - g.x1) * (ty - g.y1) - (g.y2 - g.y1) * (tx - g.x1);
      const R c2 = (g.x3 - g.x2) * (ty - g.y2) - (g.y3 - g.y2) * (tx - g.x2);
      const R c3 = (g.x1 - g.x3) * (ty - g.y3) - (g.y1 - g.y3) * (tx - g.x3);
      int nxt, e;
      if (c1 < margin) { nxt = g.n3; e = 0; }
      else if (c2 < margin) { nxt = g.n1; e = 1; }
      else if (c3 < margin) { nxt = g.n2; e = 2; }
      else break;
      if (nxt < 0) { edge = e; break; }
      if (it >= num_tri + 16) { raise_status(status, GI_ERR_ITER_CAP); edge = e; break; }
      tri = nxt;
    }
    R qx = tx, qy = ty;
    if (edge >= 0) closest_on_hull(geom, tri, edge, tx, ty, num_tri + 16, qx, qy, status);
    load_eval(geom + tri, g);
    out[i] = eval_any(g, qx, qy);
    if (tri_out) tri_out[i] = tri + 1;
    if (outside_out) outside_out[i] = edge >= 0 ? 1 : 0;
  }
}
}  // namespace

template <class R>
int barycentric_points_device(const R* points, int num_points, const R* data_in, const int32_t* simplices,
                              const int32_t* neighbours, int num_tri, const R* targets, int64_t num_targets,
                              R* data_ip, int32_t* tri_out, uint8_t* outside_out, int flags, cudaStream_t st) {
  GI_REQUIRE(num_points >= 3 && num_tri >= 1, "need at least 3 points and 1 triangle");
  GI_REQUIRE(num_targets >= 0, "negative number of targets");
  GI_TRY(ensure_device());
  if (num_targets == 0) return GI_OK;
  StatusSlot slot;
  GI_TRY(get_status_slot(st, &slot));
  Scratch scr(st);
  Phases ph(st);
  ph.mark("pack_mesh");
  // seed grid: about one cell per four triangles over the bounding box of the mesh
  const int cells = std::max(1, std::min(4096, (int)std::sqrt((double)num_tri / 4.0)));
  const int vlen = cells * kSeedCell;
  unsigned long long* box;
  R *xs, *ys;
  GI_TRY(scr.alloc(&box, 4));
  GI_TRY(scr.alloc(&xs, (size_t)vlen));
  GI_TRY(scr.alloc(&ys, (size_t)vlen));
  GI_CUDA(cudaMemsetAsync(box, 0xff, sizeof(unsigned long long) * 2, st));
  GI_CUDA(cudaMemsetAsync(box + 2, 0, sizeof(unsigned long long) * 2, st));
  const PointsView<R> pv = points_n2(points, num_points, flags & GI_POINTS_ROWMAJOR);
  points_box_kernel<R><<<std::min(div_up(num_points, 256), kNumSMs * 8), 256, 0, st>>>(pv, num_points, box);
  virtual_axes_kernel<R><<<div_up(vlen, 256), 256, 0, st>>>(box, vlen, xs, ys);
  GI_CUDA(cudaGetLastError());
  PackedMesh<R> pm;
  GI_TRY(bary_pack<R>(scr, points, num_points, data_in, xs, vlen, ys, vlen, simplices, neighbours, num_tri,
                      flags & ~GI_BARY_BAND_PACK, slot.dev, &pm, 0, 0));
  ph.mark("walk");
  const PointsView<R> tv = points_n2(targets, num_targets, flags & GI_TARGETS_ROWMAJOR);
  const int blocks = (int)std::min<int64_t>(div_up(num_targets, 128), (int64_t)kNumSMs * 32);
  bary_points_kernel<R><<<blocks, 128, 0, st>>>(pm.geom, pm.seed, pm.ncx, pm.ncy, num_tri, xs, ys, vlen, tv, num_targets,
                                               data_ip, tri_out, outside_out, slot.dev);
  GI_CUDA(cudaGetLastError());
  ph.finish();
  GI_TRY(flush_status(st, slot));
  return GI_OK;
}

#define GI_INSTANTIATE(R)                                                                                  \
  template int barycentric_device<R>(const R*, int, const R*, const R*, int, const R*, int, const int32_t*, \
                                     const int32_t*, int, int, int, int, R*, int32_t*, uint8_t*, int64_t*,  \
                                     int64_t*, int, cudaStream_t);                                          \
  template int barycentric_gather_device<R>(const R*, int, const R*, const R*, int, const R*, int,         \
                                            const int32_t*, const int32_t*, int, int, int, int, R* const*, \
                                            int, int, cudaStream_t);                                        \
  template int fill_device<R>(const uint8_t*, int, int, R*, int64_t*, int, cudaStream_t);                   \
  template int barycentric_host<R>(const R*, int, const R*, const R*, int, const R*, int, const int32_t*,   \
                                   const int32_t*, int, int, int, int, R*, int);                            \
  template int barycentric_plan_create<R>(const R*, int, const R*, int, const R*, int, const int32_t*,      \
                                          const int32_t*, int, int, int, int, int, int, cudaStream_t,       \
                                          PlanBase**);                                                       \
  template int barycentric_points_device<R>(const R*, int, const R*, const int32_t*, const int32_t*, int,  \
                                            const R*, int64_t, R*, int32_t*, uint8_t*, int, cudaStream_t);
GI_INSTANTIATE(double)
GI_INSTANTIATE(float)

}  // namespace gi

// esrg.cu -- IDW from scattered points to an equally spaced regular grid with the cell-list ("esrg")
// k-nearest-neighbour search.
//
// Replaces (reference file:line):
//   interpolation.f90:201-288   idw_esrg_nearest (operator + IDW weights)
//   query_esrg.f90:23-91        assign_points_to_cells (cell binning, CSR by grid cell)
//   query_esrg.f90:97-249       nearest_neighbours_esrg (ring-expanding exact kNN from the cell centre)
//
// B200 design (not a port):
//   binning   each point's cell key (iy*len_x + ix, the reference's CSR order) is computed once and the points
//             are sorted by it with a stable LSD radix sort (radix.cuh; 4 passes for 2^26 cells) -- stable, so
//             the points of a cell stay in ascending id order exactly like the reference's fill
//             (query_esrg.f90:72-89).  ONE int32 table of len_y*len_x entries then receives the END offset of
//             every cell (start(c) = table[c-1]): a warp writes the contiguous cell range its 32 sorted positions
//             own, 32 consecutive cells at a time (esrg_table_kernel).  Coordinates, values and ids
//             are gathered into cell order so the query reads them contiguously.  (The first version counted
//             with 10 M atomics into the table, scanned all 67 M cells and scattered with atomics again:
//             1.75 ms at C3.)
//   query     three kernels:
//             esrg_tile_kernel   (production, k <= 16) a warp owns 8 x 4 targets, stages the points of the
//               tile's halo rectangle in shared memory and every lane selects its k+1 smallest 32-bit
//               monotone keys with comparator networks -- see the comment above the kernel.
//             esrg_fast_kernel   (k > 16, or more points per tile than can be staged) one thread per target.
//               The reference's result is "the k smallest d2 among the points of the (2L+1)^2 window, L =
//               first level at which k points lie within r_L = spac*(L+1/2)" (query_esrg.f90:141,159-169,
//               178-181); its candidate lists only decide the order among EQUAL d2.  So this
//               kernel keeps no list at all: it scans frame after frame, keeps the k smallest d2 seen in
//               registers and stops when the k-th is <= r_L^2.  A frame's top and bottom rows are contiguous
//               runs of the CSR (key = iy*len_x + ix).
//             Whenever equal d2 (the only case in which the processing order matters) or ambiguous keys turn
//             up, the target is handed to
//             esrg_query_kernel  which follows the reference's processing order exactly (carried candidates
//               first, then the frame in scan order, ties inserted before equal entries).  The reference's
//               two fixed 1000-entry candidate lists (query_esrg.f90:119-121) are one small in-place
//               deferred list per thread; when it overflows the thread switches to a list-free mode that
//               re-scans the inner frames with the predicate "was deferred at the previous level"
//               (d2 > r2_prev), which yields the same order for any point density.
#include <memory>

#include "kernels.cuh"
#include "radix.cuh"
#include "topk.cuh"

namespace gi {
namespace {

constexpr int kListCap = 40;  // deferred-candidate slots per thread before the list-free fallback

// query_esrg.f90:44-57: INT((p - axis(1) + spac/2)/spac) + 1 clamped to [1,len]  (0-based here)
template <class R> __device__ __forceinline__ int cell_index(R p, R axis0, R spac, int len) {
  const R f = trunc((p - axis0 + spac / R(2)) / spac);  // INT() truncates toward zero
  if (!(f > R(-1))) return 0;                           // also catches NaN
  if (f >= R(len)) return len - 1;
  return (int)f;
}

template <class R>
__global__ void __launch_bounds__(256)
esrg_key_kernel(PointsView<R> pts, int n, const R* __restrict__ x_axis, int len_x, const R* __restrict__ y_axis,
                int len_y, R spac, unsigned* __restrict__ keys, int* __restrict__ ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ix = cell_index(pts.x(i), __ldg(x_axis), spac, len_x);
  const int iy = cell_index(pts.y(i), __ldg(y_axis), spac, len_y);
  keys[i] = (unsigned)(iy * len_x + ix);  // CSR order: ind_y outer, ind_x inner (query_esrg.f90:63-68)
  ids[i] = i;
}

// table[c] = number of points with key <= c (the END offset of cell c), from the sorted keys.  The 32 positions
// p0..p0+31 of a warp own the contiguous cell range [key[p0], key[p0+32]) (from cell 0 for the first warp, up to
// `cells` for the last): the lanes write it 32 consecutive cells at a time, and the value of cell c is
// p0 + (number of the warp's keys <= c), found by bisection over the lanes' keys with shuffles.  Fully coalesced
// whatever the gaps between occupied cells (a thread-per-point fill wrote 27 B-strided words: 1 TB/s).
__global__ void __launch_bounds__(256)
esrg_table_kernel(const unsigned* __restrict__ keys, int n, long long cells, int* __restrict__ table) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int p0 = p - lane;
  if (p0 >= n) return;  // whole warp
  const long long key = p < n ? (long long)keys[p] : cells;  // non-decreasing over the lanes
  const long long begin = p0 == 0 ? 0 : __shfl_sync(0xffffffffu, key, 0);
  const long long end = p0 + 32 < n ? (long long)keys[p0 + 32] : cells;
  for (long long c = begin + lane; __any_sync(0xffffffffu, c < end); c += 32) {
    int lo = 0, hi = 32;  // first lane whose key is > c
#pragma unroll
    for (int step = 0; step < 6; ++step) {
      const int mid = (lo + hi) >> 1;
      const long long v = __shfl_sync(0xffffffffu, key, mid & 31);
      if (lo < hi) {
        if (v <= c) lo = mid + 1; else hi = mid;
      }
    }
    if (c < end) table[c] = p0 + lo;
  }
}

template <class R> struct alignas(16) XY { R x, y; };

// order[p] = 0-based id of the point at cell-order position p; sid_out (optional) receives the 1-based ids
template <class R>
__global__ void __launch_bounds__(256)
esrg_gather_kernel(PointsView<R> pts, const R* __restrict__ data, const int* __restrict__ order,
                   const unsigned* __restrict__ keys_sorted, int n, XY<R>* __restrict__ sxy, R* __restrict__ sval,
                   int* __restrict__ sid_out, unsigned* __restrict__ scell) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int i = order[p];
  if (scell) scell[p] = keys_sorted[p];
  if (sxy) { XY<R> v; pts.xy(i, v.x, v.y); sxy[p] = v; }
  if (sval) sval[p] = __ldg(data + i);
  if (sid_out) sid_out[p] = i + 1;  // 1-based ids like index_of_pts
}

// plans: new nodal values into cell order (sid holds the 1-based ids)
template <class R>
__global__ void __launch_bounds__(256)
esrg_values_kernel(const R* __restrict__ data, const int* __restrict__ sid, int n, R* __restrict__ sval) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) sval[p] = __ldg(data + sid[p] - 1);
}

__global__ void esrg_export_indptr_kernel(const int* __restrict__ table, int64_t cells, int* __restrict__ indptr) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c <= cells; c += (int64_t)gridDim.x * blockDim.x)
    indptr[c] = c == 0 ? 0 : table[c - 1];
}

template <class R> struct EsrgArgs {
  TargetWindow<R> win;
  const int* table;       // END offset per cell
  const XY<R>* sxy;       // coordinates in cell order
  const R* sval;          // values in cell order
  const int* sid;         // 1-based point ids in cell order
  const unsigned* scell;  // cell (iy*len_x + ix) of every point in cell order
  int len_x, len_y, k;
  R spac;
  R* out;                 // window-shaped
  int32_t* nn_index;      // optional (window, k)
  R* nn_dist2;            // optional
  int64_t* counters;      // [0] cells scanned, [1] max level, [2] list overflows, [3] targets sent to the exact kernel
  int* status;
  int* todo;              // window-linear ids of the targets the exact kernel must do (filled by the fast kernel)
  int64_t* todo_count;
  int all_exact;          // debug: the fast kernel hands every target over
  int stats;              // update counters[0..2] (debug: one atomic per target)
  R* tk_d;                // k > 32: neighbour lists in scratch memory (topk.cuh, KC = 0), k x tk_stride
  int* tk_id;
  int64_t tk_stride;
  PointsView<R> tgt;      // scattered targets (num_tgt > 0, esrg_query_kernel only): target i instead of a grid node
  int64_t num_tgt;
  const R* x_axis;        // (scattered targets: the cell of a target is computed like a point's)
  const R* y_axis;
};

// Smallest level L >= from whose radius is not exceeded by d2, i.e. the first level at which the reference accepts
// a deferred point with this d2: NOT (d2 > (spac*(L+0.5))**2), the radius evaluated as the reference does
// (query_esrg.f90:141,181,230).  Used once a target's frames have left the grid: no cell is added any more and
// nothing changes until the growing radius reaches the nearest deferred point (points clamped into the border
// cells from far outside the grid), so the levels in between are skipped instead of iterated.
template <class R> __device__ __forceinline__ int esrg_level_reaching(R spac, R d2, int from, R half = R(0.5)) {
  const double est = sqrt((double)d2) / (double)spac - (double)half;
  int L = est > 2.0e9 ? 2000000000 : (est > 0.0 ? (int)est : 0);
  L = max(L - 1, from);
  auto r2 = [&](int l) { const R rad = spac * (R(l) + half); return rad * rad; };
  while (L > from && !(d2 > r2(L - 1))) --L;
  while (d2 > r2(L) && L < 2000000000) ++L;
  return L;
}

template <class R, int KC>
__global__ void __launch_bounds__(128) esrg_query_kernel(const EsrgArgs<R> a) {
  const TargetWindow<R>& w = a.win;
  // Scattered targets (the reference's targets are the cell centres, query_esrg.f90:133-134): the ring
  // search runs around the target's cell, and a level's radius is spac * L instead of spac * (L + 1/2) -- a target
  // anywhere in its cell is at least L cells away from everything outside the (2L+1)^2 block.
  const bool scattered = a.num_tgt > 0;
  const R half = scattered ? R(0) : R(0.5);
  const int64_t n_todo = scattered ? a.num_tgt : *a.todo_count;
  for (int64_t li = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; li < n_todo; li += (int64_t)gridDim.x * blockDim.x) {
  const int len_x = a.len_x, len_y = a.len_y;
  int f = 0, s = 0, ix, iy;
  R cx, cy;
  if (scattered) {
    a.tgt.xy(li, cx, cy);
    ix = cell_index(cx, __ldg(a.x_axis), a.spac, len_x);
    iy = cell_index(cy, __ldg(a.y_axis), a.spac, len_y);
  } else {
    const int lin = a.todo[li];
    f = w.f0 + lin % w.nf();
    s = w.s0 + lin / w.nf();
    ix = w.fast_is_y ? s : f; iy = w.fast_is_y ? f : s;
    cx = __ldg((w.fast_is_y ? w.slow_axis : w.fast_axis) + ix);           // query_esrg.f90:133-134
    cy = __ldg((w.fast_is_y ? w.fast_axis : w.slow_axis) + iy);
  }

  TopK<R, KC> tk;
  if constexpr (KC == 0) tk.bind(a.tk_d, a.tk_id, a.tk_stride, blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
  tk.init(a.k, Consts<R>::huge(), -1);                                    // :129-130
  int found = 0;

  int list_pos[kListCap];
  R list_d2[kListCap];
  int n_list = 0;
  bool listfree = false;  // set once the deferred list overflowed
  int overflows = 0;
  long long cells_scanned = 0;

  R radius_sq = R(0), radius_sq_prev = R(0);
  R dmin_deferred = Consts<R>::huge();  // smallest d2 deferred at the current level
  // One candidate (query_esrg.f90:158-169 / :229-240).  `fresh` = first time this point is seen (frame ==
  // level); otherwise it is a re-scanned inner-frame point that counts only if it was deferred so far.
  auto consider = [&](int pos, R dist_sq, bool fresh) {
    if (!fresh && !(dist_sq > radius_sq_prev)) return;   // already consumed at an earlier level
    if (dist_sq > radius_sq) {                            // defer
      dmin_deferred = dist_sq < dmin_deferred ? dist_sq : dmin_deferred;
      if (!listfree) {
        if (n_list < kListCap) { list_pos[n_list] = pos; list_d2[n_list] = dist_sq; ++n_list; }
        else { listfree = true; ++overflows; }
      }
    } else if (dist_sq < tk.dlast) {
      tk.template insert<true>(dist_sq, pos);
      found = min(found + 1, a.k);
    }
  };
  auto scan_run = [&](int p0, int p1, bool fresh) {
    for (int p = p0; p < p1; ++p) {
      const XY<R> q = a.sxy[p];
      const R ddx = cx - q.x, ddy = cy - q.y;
      consider(p, ddx * ddx + ddy * ddy, fresh);                          // :147-148
    }
  };
  auto cell_start = [&](int c) { return c > 0 ? __ldg(a.table + c - 1) : 0; };
  // frame `lvl` around (ix,iy) in the reference's scan order (:186-223); lvl 0 = the centre cell
  auto scan_frame = [&](int lvl, bool fresh) {
    if (lvl == 0) {
      const int c = iy * len_x + ix;
      scan_run(cell_start(c), __ldg(a.table + c), fresh);
      ++cells_scanned;
      return;
    }
    const int jlo = max(ix - lvl, 0), jhi = min(ix + lvl, len_x - 1);
    for (int i = -lvl; i <= lvl; i += 2 * lvl) {                          // bottom / top rows, x ascending
      const int iy2 = iy + i;
      if (iy2 < 0 || iy2 >= len_y) continue;
      const int c0 = iy2 * len_x + jlo, c1 = iy2 * len_x + jhi;
      scan_run(cell_start(c0), __ldg(a.table + c1), fresh);
      cells_scanned += jhi - jlo + 1;
    }
    for (int j = -lvl; j <= lvl; j += 2 * lvl) {                          // left / right columns, y ascending
      const int ix2 = ix + j;
      if (ix2 < 0 || ix2 >= len_x) continue;
      for (int i = -lvl + 1; i <= lvl - 1; ++i) {
        const int iy2 = iy + i;
        if (iy2 < 0 || iy2 >= len_y) continue;
        const int c = iy2 * len_x + ix2;
        scan_run(cell_start(c), __ldg(a.table + c), fresh);
        ++cells_scanned;
      }
    }
  };

  int level = 0;
  // beyond this level the frame lies outside the grid: every cell has been scanned
  const int level_full = max(max(ix, len_x - 1 - ix), max(iy, len_y - 1 - iy));
  for (;;) {
    const R rad = a.spac * (R(level) + half);                             // :141,:181
    radius_sq_prev = radius_sq;
    radius_sq = rad * rad;
    dmin_deferred = Consts<R>::huge();
    if (listfree) {
      // exact list-free mode: inner frames first (the order the deferred list would have), then the new one
      for (int l = 0; l < min(level, level_full + 1); ++l) scan_frame(l, false);
    } else {
      // carried (deferred) candidates first, compacting in place (:184-185,:227-233)
      const int n_old = n_list;
      n_list = 0;
      for (int i = 0; i < n_old; ++i) {
        const int pos = list_pos[i];
        const R dist_sq = list_d2[i];
        if (dist_sq > radius_sq) {
          list_pos[n_list] = pos; list_d2[n_list] = dist_sq; ++n_list;
          dmin_deferred = dist_sq < dmin_deferred ? dist_sq : dmin_deferred;
        }
        else if (dist_sq < tk.dlast) { tk.template insert<true>(dist_sq, pos); found = min(found + 1, a.k); }
      }
    }
    if (level <= level_full) scan_frame(level, true);
    if (found == a.k) break;                                              // :178,:243
    ++level;
    if (level > level_full) {
      // The reference keeps incrementing the level (:178-181) until the radius reaches its deferred points; the
      // host checked num_points >= k, so that happens.  (No deferred point at all: NaN coordinates -- stop.)
      if (!(dmin_deferred < Consts<R>::huge())) { raise_status(a.status, GI_ERR_TOO_FEW_POINTS); break; }
      level = esrg_level_reaching(a.spac, dmin_deferred, level, half);
    }
  }

  // IDW (interpolation.f90:260-274): distances are square-rooted first (query_esrg.f90:247)
  const int64_t o = scattered ? li : (int64_t)(s - w.s0) * w.nf() + (f - w.f0);
  R numerator = R(0), denominator = R(0), first_val = R(0), first_dist = R(0);
  auto take = [&](int rank, int pos, R d2r) {
    const R dist = sqrt(d2r);
    const R v = pos >= 0 ? __ldg(a.sval + pos) : R(0);
    if (rank == 0) { first_val = v; first_dist = dist; }
    numerator = numerator + (v / dist);
    denominator = denominator + (R(1) / dist);
    if (a.nn_index) a.nn_index[o * a.k + rank] = pos >= 0 ? __ldg(a.sid + pos) : -1;
    if (a.nn_dist2) a.nn_dist2[o * a.k + rank] = d2r;
  };
  if constexpr (KC == 0) {
    for (int j = 0; j < a.k; ++j) take(j, tk.index(j), tk.dist2(j));
  } else {
#pragma unroll
    for (int j = 0; j < KC; ++j)
      if (j >= tk.base) take(j - tk.base, tk.id[j], tk.d[j]);  // rank j - tk.base of the list (topk.cuh)
  }
  a.out[o] = (first_dist <= Consts<R>::tiny_e4()) ? first_val : numerator / denominator;

  if (a.stats) {
    atomicAdd((unsigned long long*)a.counters, (unsigned long long)cells_scanned);
    atomicMax((unsigned long long*)(a.counters + 1), (unsigned long long)level);
    if (overflows) atomicAdd((unsigned long long*)(a.counters + 2), (unsigned long long)overflows);
  }
  }
}

// ---------------------------------------------------------------------------------------------
// fast kernel (see file header)
// ---------------------------------------------------------------------------------------------
template <class R, int KC>
__global__ void __launch_bounds__(128) esrg_fast_kernel(const EsrgArgs<R> a) {
  const TargetWindow<R>& w = a.win;
  const int f = w.f0 + blockIdx.x * 32 + (threadIdx.x & 31);
  const int s = w.s0 + blockIdx.y * 4 + (threadIdx.x >> 5);
  if (f >= w.f1 || s >= w.s1) return;
  const int lin = (s - w.s0) * w.nf() + (f - w.f0);
  if (a.all_exact) { a.todo[atomicAdd((unsigned long long*)a.todo_count, 1ull)] = lin; return; }
  const int ix = w.fast_is_y ? s : f, iy = w.fast_is_y ? f : s;
  const int len_x = a.len_x, len_y = a.len_y, k = a.k;
  const R cx = __ldg((w.fast_is_y ? w.slow_axis : w.fast_axis) + ix);     // query_esrg.f90:133-134
  const R cy = __ldg((w.fast_is_y ? w.fast_axis : w.slow_axis) + iy);

  // the k smallest d2 seen so far, unsorted; dmax = largest of them (HUGE until k points were seen)
  R d[KC];
  int id[KC];
#pragma unroll
  for (int j = 0; j < KC; ++j) { d[j] = Consts<R>::huge(); id[j] = -1; }
  R dmax = Consts<R>::huge();
  int jmax = 0;
  bool tie = false;

  auto scan_run = [&](int p0, int p1) {
    for (int p = p0; p < p1; ++p) {
      const XY<R> q = a.sxy[p];
      const R ddx = cx - q.x, ddy = cy - q.y;
      const R d2 = ddx * ddx + ddy * ddy;                                 // :147-148
      if (d2 < dmax) {
        R m = R(-1);
        int jm = 0;
#pragma unroll
        for (int j = 0; j < KC; ++j) {
          if (j < k) {
            if (j == jmax) { d[j] = d2; id[j] = p; }
            else tie |= (d[j] == d2);
            if (d[j] > m) { m = d[j]; jm = j; }
          }
        }
        dmax = m; jmax = jm;
      } else {
        tie |= (d2 == dmax);
      }
    }
  };
  auto cell_start = [&](int c) { return c > 0 ? __ldg(a.table + c - 1) : 0; };

  int level = 0;
  long long cells_scanned = 0;
  const int level_full = max(max(ix, len_x - 1 - ix), max(iy, len_y - 1 - iy));  // last frame that touches the grid
  for (;;) {
    if (level == 0) {
      const int c = iy * len_x + ix;
      scan_run(cell_start(c), __ldg(a.table + c));
      ++cells_scanned;
    } else {
      const int jlo = max(ix - level, 0), jhi = min(ix + level, len_x - 1);
      for (int i = -level; i <= level; ++i) {
        const int iy2 = iy + i;
        if (iy2 < 0 || iy2 >= len_y) continue;
        const int row = iy2 * len_x;
        if (i == -level || i == level) {                                  // a full row of the frame: one run
          scan_run(cell_start(row + jlo), __ldg(a.table + row + jhi));
          cells_scanned += jhi - jlo + 1;
        } else {                                                          // its two end cells
          if (ix - level >= 0) { const int c = row + ix - level; scan_run(cell_start(c), __ldg(a.table + c)); ++cells_scanned; }
          if (ix + level < len_x) { const int c = row + ix + level; scan_run(cell_start(c), __ldg(a.table + c)); ++cells_scanned; }
        }
      }
    }
    const R rad = a.spac * (R(level) + R(0.5));                             // :141,:181
    if (!(dmax > rad * rad)) break;                                       // k points within r_L: :178,:243
    ++level;
    if (level > level_full) {  // every cell scanned: the k smallest are final, the reference only grows its radius
      level = esrg_level_reaching(a.spac, dmax, level);
      break;
    }
  }
  if (tie) {  // equal distances: the reference's processing order decides -> exact kernel
    a.todo[atomicAdd((unsigned long long*)a.todo_count, 1ull)] = lin;
    return;
  }
  // ascending d2 (odd-even transposition network; unused slots hold HUGE and stay at the end)
#pragma unroll
  for (int round = 0; round < KC; ++round) {
#pragma unroll
    for (int j = (round & 1); j + 1 < KC; j += 2) {
      if (d[j + 1] < d[j]) {
        const R td = d[j]; d[j] = d[j + 1]; d[j + 1] = td;
        const int ti = id[j]; id[j] = id[j + 1]; id[j + 1] = ti;
      }
    }
  }
  // IDW (interpolation.f90:260-274): distances are square-rooted first (query_esrg.f90:247)
  const int64_t o = lin;
  R numerator = R(0), denominator = R(0), first_val = R(0), first_dist = R(0);
#pragma unroll
  for (int i = 0; i < KC; ++i) {
    if (i < k) {
      const int pos = id[i];
      const R dist = sqrt(d[i]);
      const R v = pos >= 0 ? __ldg(a.sval + pos) : R(0);
      if (i == 0) { first_val = v; first_dist = dist; }
      numerator = numerator + (v / dist);
      denominator = denominator + (R(1) / dist);
      if (a.nn_index) a.nn_index[o * k + i] = pos >= 0 ? __ldg(a.sid + pos) : -1;
      if (a.nn_dist2) a.nn_dist2[o * k + i] = d[i];
    }
  }
  a.out[o] = (first_dist <= Consts<R>::tiny_e4()) ? first_val : numerator / denominator;
  if (a.stats) {
    atomicAdd((unsigned long long*)a.counters, (unsigned long long)cells_scanned);
    atomicMax((unsigned long long*)(a.counters + 1), (unsigned long long)level);
  }
}

// ---------------------------------------------------------------------------------------------
// tile kernel: the production path for k <= 16
//
// ncu on the per-thread kernels above: 440 warp instructions per target at 7.6 of 32 lanes active -- every lane
// runs its own cell loops with its own trip counts.  Here a WARP owns a tile of 8 x 4 targets (8 along the
// contiguous output axis) and works in lock step:
//   1. stage   the points of the tile's halo rectangle (tile grown by H cells on every side; one CSR run per
//              grid row, one lane per run) are copied to shared memory (at most 223 of them);
//   2. scan    every lane computes its d2 to every staged point (shared-memory broadcast, no divergence) and
//              turns it into a 32-bit KEY = (fixed-point d2) << 8 | staged index, or 0xffffffff beyond r_H =
//              spac*(H+1/2), the radius the lane's own (2H+1)^2 window is guaranteed to cover.  The fixed-point
//              value is the low mantissa word of d2 + 2^(e+52) -- a MONOTONE function of d2 with 2^-23 r_H^2
//              resolution.  Keys carry no payload, so keeping the k+1 smallest is pure integer min/max in
//              registers: candidates are taken in batches of 8 (16 for k > 8), a batch is sorted by a
//              19-comparator network and merged into the kept keys by one bitonic merge (10 min/max per
//              candidate; an insertion ladder, the previous form, needs 17);
//   3. grow    if a lane found fewer than k points within r_H the warp repeats 1-2 with H+2 (rare: H starts
//              at the level where the mean density gives ~1.5k+6 points within r_H);
//   4. resolve monotonicity makes the k smallest keys the k nearest points in ascending order provided the
//              fixed-point values of the k+1 smallest keys are pairwise different; then the exact d2 of the k
//              winners is recomputed from the staged coordinates (same expression, same bits) and the IDW sum
//              is taken in that order.  Equal fixed-point values (exact ties, where the reference's candidate
//              order decides, or distances closer than the resolution: ~1e-5 of the targets) and tiles with
//              more than 223 staged points send the target to the exact kernel.
// The result is the exact k nearest neighbours, which is what the reference computes (see esrg_fast_kernel).
// The first tile version kept per-lane candidate lists in shared memory and selected with k+1 minimum-
// extraction passes: 112 warp instructions per target, 31 KB of shared memory per 2 warps (15.1 ms at C3).
// ---------------------------------------------------------------------------------------------
constexpr int kStageCap = 223;  // staged points per tile: the staged index is the low byte of a key; 224 x (16 + 4 + 2) B
                                // per warp keeps 10 CTAs of 4 warps on an SM
constexpr int kTileWarps = 4;

__device__ __forceinline__ unsigned lo_word(double x) { return (unsigned)__double2loint(x); }

// Key selection networks (registers only, fully unrolled; validated with the 0-1 principle).
#define GI_CE(x, y) { const unsigned lo__ = min(x, y), hi__ = max(x, y); x = lo__; y = hi__; }
// ascending sort of M keys: Batcher's 19-comparator odd-even merge sort for M = 8, a bitonic network otherwise
template <int M> __device__ __forceinline__ void sort_keys(unsigned (&b)[M]) {
  if (M == 8) {
    GI_CE(b[0], b[1]) GI_CE(b[2], b[3]) GI_CE(b[4], b[5]) GI_CE(b[6], b[7])
    GI_CE(b[0], b[2]) GI_CE(b[1], b[3]) GI_CE(b[4], b[6]) GI_CE(b[5], b[7])
    GI_CE(b[1], b[2]) GI_CE(b[5], b[6])
    GI_CE(b[0], b[4]) GI_CE(b[1], b[5]) GI_CE(b[2], b[6]) GI_CE(b[3], b[7])
    GI_CE(b[2], b[4]) GI_CE(b[3], b[5])
    GI_CE(b[1], b[2]) GI_CE(b[3], b[4]) GI_CE(b[5], b[6])
  } else {
#pragma unroll
    for (int k = 2; k <= M; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
        for (int i = 0; i < M; ++i) {
          const int l = i ^ j;
          if (l > i) {
            if ((i & k) == 0) { GI_CE(b[i], b[l]) } else { GI_CE(b[l], b[i]) }
          }
        }
      }
    }
  }
}
// a[] (ascending) <- the M smallest of a[] and b[] (both ascending), ascending; next <- min(next, the other M):
// min(a[i], b[M-1-i]) is the lower half of the union as a bitonic sequence, one bitonic merge sorts it
template <int M> __device__ __forceinline__ void merge_keys(unsigned (&a)[M], const unsigned (&b)[M], unsigned& next) {
  unsigned hi[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    const unsigned x = a[i], y = b[M - 1 - i];
    a[i] = min(x, y);
    hi[i] = max(x, y);
  }
#pragma unroll
  for (int i = 0; i < M; i += 2) next = __vimin3_u32(next, hi[i], hi[i + 1]);
#pragma unroll
  for (int j = M >> 1; j > 0; j >>= 1) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const int l = i ^ j;
      if (l > i) GI_CE(a[i], a[l])
    }
  }
}

// M = largest supported k.  The k live slots of key[] are the TOP ones (slots below hold 0 and stay 0), so the
// k-th smallest key sits in the compile-time slot M-1 for every run-time k; `next` is the (k+1)-th smallest.
template <class R, int M>
#ifndef GI_ESRG_TILE_MINB
#define GI_ESRG_TILE_MINB 10
#endif
__global__ void __launch_bounds__(32 * kTileWarps, GI_ESRG_TILE_MINB) esrg_tile_kernel(const EsrgArgs<R> a, int h0) {
  __shared__ XY<R> s_xy[kTileWarps][kStageCap + 1];
  __shared__ unsigned short s_cell[kTileWarps][kStageCap + 1];  // cell of a staged point inside the halo rectangle:
                                                                // column - rx0 (low byte), row - ry0 (high byte)
  __shared__ int s_pos[kTileWarps][kStageCap + 1];
  const TargetWindow<R>& w = a.win;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int len_x = a.len_x, len_y = a.len_y, k = a.k;
  // tile origin in (fast, slow); 8 lanes along fast, 4 along slow
  const int tiles_f = (w.nf() + 7) >> 3;
  const int tile = blockIdx.x * kTileWarps + wid;
  const int tf0 = w.f0 + (tile % tiles_f) * 8, ts0 = w.s0 + (tile / tiles_f) * 4;
  if (ts0 >= w.s1) return;  // whole warp
  const int f = tf0 + (lane & 7), s = ts0 + (lane >> 3);
  const bool active = f < w.f1 && s < w.s1;
  const int lin = (s - w.s0) * w.nf() + (f - w.f0);
  if (a.all_exact) {
    if (active) a.todo[atomicAdd((unsigned long long*)a.todo_count, 1ull)] = lin;
    return;
  }
  const int fa = min(f, w.f1 - 1), sa = min(s, w.s1 - 1);  // inactive lanes shadow a valid target
  const int ix = w.fast_is_y ? sa : fa, iy = w.fast_is_y ? fa : sa;
  const R cx = __ldg((w.fast_is_y ? w.slow_axis : w.fast_axis) + ix);     // query_esrg.f90:133-134
  const R cy = __ldg((w.fast_is_y ? w.fast_axis : w.slow_axis) + iy);
  // tile extent in grid cells
  const int f_hi = min(tf0 + 7, w.f1 - 1), s_hi = min(ts0 + 3, w.s1 - 1);
  const int tx0 = w.fast_is_y ? ts0 : tf0, tx1 = w.fast_is_y ? s_hi : f_hi;
  const int ty0 = w.fast_is_y ? tf0 : ts0, ty1 = w.fast_is_y ? f_hi : s_hi;

  const int base = M - k;  // first live slot
  unsigned key[M];
  unsigned next = 0xffffffffu;
  bool hand_over = false;
  int H = h0;
  unsigned tcell = 0;  // the target's own cell inside the halo rectangle, packed like s_cell
  for (;;) {
    const R rad = a.spac * (R(H) + R(0.5));                               // :141,:181
    const R r2 = rad * rad;
    // 2^(e+52) with r2 < 2^(e+23): d2 <= r2 maps to a fixed-point value below 2^23
    const int r2_exp = (__double2hiint((double)r2) >> 20) & 0x7ff;        // biased exponent of r2 (> 0)
    const double magic = __hiloint2double(min(r2_exp + 30, 2046) << 20, 0);
    // ---- 1. stage: one CSR run per grid row of the halo rectangle ----------------------------------
    const int rx0 = max(tx0 - H, 0), ry0 = max(ty0 - H, 0);
    tcell = (unsigned)(ix - rx0) | ((unsigned)(iy - ry0) << 8);
    const int rx1 = min(tx1 + H, len_x - 1), ry1 = min(ty1 + H, len_y - 1);
    const int nrows = ry1 - ry0 + 1, ncols = rx1 - rx0 + 1;
    const int row = (ry0 + min(lane, nrows - 1)) * len_x;
    int p0 = 0, n = 0;
    if (lane < nrows) {
      const int c0 = row + rx0;
      p0 = c0 > 0 ? __ldg(a.table + c0 - 1) : 0;
      n = __ldg(a.table + row + rx1) - p0;
    }
    int off = n;  // inclusive warp scan
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, off, d);
      if (lane >= d) off += t;
    }
    const int total = __shfl_sync(0xffffffffu, off, 31);
    off -= n;
    if (nrows > 32 || ncols > 255 || total > kStageCap) { hand_over = true; break; }  // dense cluster: exact kernel
    __syncwarp();
    for (int i = 0; i < n; ++i) {  // every staged point remembers its cell inside the rectangle (see 4.)
      s_xy[wid][off + i] = a.sxy[p0 + i];
      s_pos[wid][off + i] = p0 + i;
      s_cell[wid][off + i] = (unsigned short)((__ldg(a.scell + p0 + i) - (unsigned)(row + rx0)) | (lane << 8));
    }
    __syncwarp();
    // ---- 2. scan ------------------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < M; ++j) key[j] = j >= base ? 0xffffffffu : 0u;
    next = 0xffffffffu;
    auto key_of = [&](int i) -> unsigned {
      const XY<R> q = s_xy[wid][i];
      const R ddx = cx - q.x, ddy = cy - q.y;
      const R d2 = ddx * ddx + ddy * ddy;                                 // :147-148
      const unsigned fix = lo_word((double)d2 + magic);
      return d2 > r2 ? 0xffffffffu : (fix << 8) + (unsigned)i;
    };
    // batches of M candidates: sort the batch (network), merge it into the kept keys (one bitonic merge)
    int i0 = 0;
    for (; i0 + M <= total; i0 += M) {
      unsigned b[M];
#pragma unroll
      for (int u = 0; u < M; ++u) b[u] = key_of(i0 + u);
      sort_keys<M>(b);
      merge_keys<M>(key, b, next);
    }
    if (i0 < total) {  // ragged last batch (total is warp-uniform)
      unsigned b[M];
#pragma unroll
      for (int u = 0; u < M; ++u) b[u] = i0 + u < total ? key_of(i0 + u) : 0xffffffffu;
      sort_keys<M>(b);
      merge_keys<M>(key, b, next);
    }
    // ---- 3. grow? -----------------------------------------------------------------------------------
    if (__all_sync(0xffffffffu, key[M - 1] != 0xffffffffu)) break;        // k points within r_H: :178,:243
    // the rectangle already holds every point of the grid (neighbours clamped in from far outside): a larger H
    // adds nothing but radius -- the exact kernel jumps to the level the reference stops at
    if (rx0 == 0 && ry0 == 0 && rx1 == len_x - 1 && ry1 == len_y - 1) { hand_over = true; break; }
    H += 2;
  }
  if (hand_over) {
    if (active) a.todo[atomicAdd((unsigned long long*)a.todo_count, 1ull)] = lin;
    return;
  }
  // ---- 4. resolve -----------------------------------------------------------------------------------
  // The reference stops at the first level L whose (2L+1)^2 window holds k points within r_L and returns the k
  // nearest of THOSE (query_esrg.f90:178-243); the winners below come from the larger halo rectangle.  They are
  // the reference's result iff all of them lie inside the window of L* = the first level with r_L^2 >= the k-th
  // winner's d2 (no fewer than k points are within any smaller radius, and a subset's k smallest that are also the
  // superset's k smallest are the same points).  In exact arithmetic that always holds -- a point outside the
  // window is farther away than r_L -- but a point ON a cell border can be binned one cell further out and still
  // have a rounded distance <= r_L, and axes that are not exactly x(1) + i*spac shift the windows against the
  // distances; so it is checked, in cell indices, and the rare violations go to the exact kernel.
  bool ambiguous = false;
  unsigned cheb = 0;  // per byte: max |winner cell - target cell| (columns, rows)
  R dk2 = R(0);
  R numerator = R(0), denominator = R(0), first_val = R(0), first_dist = R(0);
#pragma unroll
  for (int j = 0; j < M; ++j) {
    if (j >= base) {
      ambiguous |= (key[j] >> 8) == ((j + 1 < M ? key[j + 1 < M ? j + 1 : j] : next) >> 8);
      const int e = key[j] & 0xff;
      const XY<R> q = s_xy[wid][e];
      const int pos = s_pos[wid][e];
      cheb = __vmaxu4(cheb, __vabsdiffu4(s_cell[wid][e], tcell));
      const R ddx = cx - q.x, ddy = cy - q.y;
      const R d2 = ddx * ddx + ddy * ddy;                                 // :147-148
      if (j == M - 1) dk2 = d2;
      // IDW (interpolation.f90:260-274): distances are square-rooted first (query_esrg.f90:247)
      const R dist = sqrt(d2);
      const R v = __ldg(a.sval + pos);
      if (j == base) { first_val = v; first_dist = dist; }
      numerator = numerator + (v / dist);
      denominator = denominator + (R(1) / dist);
      if (active) {  // (overwritten by the exact kernel if the target turns out to be ambiguous)
        if (a.nn_index) a.nn_index[(int64_t)lin * k + (j - base)] = __ldg(a.sid + pos);
        if (a.nn_dist2) a.nn_dist2[(int64_t)lin * k + (j - base)] = d2;
      }
    }
  }
  const int reach = (int)max(cheb & 0xffu, cheb >> 8);
  if (reach > 0) {  // all winners inside the window of L*  <=>  r_(reach-1)^2 < d2 of the k-th winner
    const R rad = a.spac * (R(reach - 1) + R(0.5));
    ambiguous |= !(dk2 > rad * rad);
  }
  if (!active) return;
  if (ambiguous) {  // equal (fixed-point) distances: the reference's candidate order may decide
    a.todo[atomicAdd((unsigned long long*)a.todo_count, 1ull)] = lin;
    return;
  }
  a.out[lin] = (first_dist <= Consts<R>::tiny_e4()) ? first_val : numerator / denominator;
}

template <class R> struct Binned {
  int* table;
  int* sid;
  unsigned* scell;
  XY<R>* sxy;
  R* sval;
};

template <class R>
int esrg_bin(Scratch& keep, Scratch& tmp, const R* points, int n, const R* data_in, const R* x_axis, int len_x,
             const R* y_axis, int len_y, R spac, int flags, bool want_values, Binned<R>* b) {
  cudaStream_t st = tmp.stream();
  const int64_t cells = (int64_t)len_x * len_y;
  GI_REQUIRE(cells < (int64_t)0x7fffffff, "grid has more than 2^31 cells");
  GI_TRY(keep.alloc(&b->table, (size_t)cells));
  GI_TRY(keep.alloc(&b->sid, (size_t)n));
  GI_TRY(keep.alloc(&b->scell, (size_t)n));
  GI_TRY(keep.alloc(&b->sxy, (size_t)n));
  b->sval = nullptr;
  if (want_values) GI_TRY(keep.alloc(&b->sval, (size_t)n));
  unsigned *keys, *keys_tmp, *keys_sorted;
  int *ids, *ids_tmp, *ids_sorted, *hist;
  GI_TRY(tmp.alloc(&keys, (size_t)n));
  GI_TRY(tmp.alloc(&keys_tmp, (size_t)n));
  GI_TRY(tmp.alloc(&ids, (size_t)n));
  GI_TRY(tmp.alloc(&ids_tmp, (size_t)n));
  GI_TRY(tmp.alloc(&hist, radix_hist_size<unsigned>(n)));
  const PointsView<R> pv = points_n2(points, n, flags & GI_POINTS_ROWMAJOR);
  const int blocks = div_up(n, 256);
  esrg_key_kernel<R><<<blocks, 256, 0, st>>>(pv, n, x_axis, len_x, y_axis, len_y, spac, keys, ids);
  GI_CUDA(cudaGetLastError());
  int passes = 1;
  while (passes < 4 && (cells - 1) >> (8 * passes)) ++passes;
  GI_TRY(radix_sort<unsigned>(keys, ids, keys_tmp, ids_tmp, hist, n, passes, st, &keys_sorted, &ids_sorted));
  esrg_table_kernel<<<blocks, 256, 0, st>>>(keys_sorted, n, (long long)cells, b->table);
  esrg_gather_kernel<R><<<blocks, 256, 0, st>>>(pv, data_in, ids_sorted, keys_sorted, n, b->sxy,
                                                data_in ? b->sval : nullptr, b->sid, b->scell);
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}
template <class R>
int esrg_bin(Scratch& scr, const R* points, int n, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
             int len_y, R spac, int flags, Binned<R>* b) {
  return esrg_bin<R>(scr, scr, points, n, data_in, x_axis, len_x, y_axis, len_y, spac, flags, data_in != nullptr, b);
}

// query + IDW for the targets rows [row0,row1) x columns [col0,col1); `out` (and nn_*) are window-shaped.
// counters: device int64[4], zeroed by the caller (slot 3 is the exact-order work counter of THIS call).
template <class R>
int esrg_query_window(Scratch& scr, const Binned<R>& b, int num_points, const R* x_axis, int len_x, const R* y_axis,
                      int len_y, R grid_spac, int k, int row0, int row1, int col0, int col1, R* out,
                      int32_t* nn_index, R* nn_dist2, int64_t* counters, bool stats, int flags, int* status,
                      Phases* ph) {
  if (row0 >= row1 || col0 >= col1) return GI_OK;
  cudaStream_t st = scr.stream();
  EsrgArgs<R> a;
  const bool rowmajor = flags & GI_FIELD_ROWMAJOR;
  a.win = rowmajor ? TargetWindow<R>{x_axis, y_axis, len_x, len_y, col0, col1, row0, row1, 0}
                   : TargetWindow<R>{y_axis, x_axis, len_y, len_x, row0, row1, col0, col1, 1};
  a.table = b.table; a.sxy = b.sxy; a.sval = b.sval; a.sid = b.sid; a.scell = b.scell;
  a.len_x = len_x; a.len_y = len_y; a.k = k; a.spac = grid_spac;
  a.out = out; a.nn_index = nn_index; a.nn_dist2 = nn_dist2;
  a.counters = counters; a.stats = stats ? 1 : 0; a.status = status;
  dim3 grid(div_up(a.win.nf(), 32), div_up(a.win.ns(), 4));
  GI_REQUIRE(grid.y <= 65535, "window too tall for one launch");
  const int64_t ntargets = (int64_t)a.win.nf() * a.win.ns();
  GI_REQUIRE(ntargets < (int64_t)0x7fffffff, "window has more than 2^31 targets");
  GI_TRY(scr.alloc(&a.todo, (size_t)ntargets));
  a.todo_count = counters + 3;
  a.all_exact = ((flags & GI_ESRG_REFERENCE_ORDER) || k > 32) ? 1 : 0;  // k > 32: only the exact kernel has the any-k list
  a.tk_d = nullptr; a.tk_id = nullptr; a.tk_stride = 0;
  a.tgt = PointsView<R>{nullptr, 0, 0}; a.num_tgt = 0; a.x_axis = x_axis; a.y_axis = y_axis;
  // first halo level: where the mean point density puts about 1.5k+6 points within r_H = spac*(H+1/2), so
  // that all 32 lanes of a tile find k of them in the first pass with high probability (Poisson tail)
  const double density = (double)num_points / ((double)len_x * (double)len_y);
  int h0 = 1;
  while (h0 < 12 && 3.141592653589793 * (h0 + 0.5) * (h0 + 0.5) * density < 1.5 * k + 6.0) ++h0;
  // the tile kernel stages at most kStageCap points per tile; very dense inputs go straight to the per-thread kernel
  const bool tiles_fit = (8.0 + 2 * h0) * (4.0 + 2 * h0) * density < 0.6 * kStageCap;
  if (k <= 16 && tiles_fit) {
    const int64_t tiles = (int64_t)div_up(a.win.nf(), 8) * div_up(a.win.ns(), 4);
    const unsigned tg = (unsigned)div_up(tiles, kTileWarps);
    if (k <= 8) esrg_tile_kernel<R, 8><<<tg, 32 * kTileWarps, 0, st>>>(a, h0);
    else esrg_tile_kernel<R, 16><<<tg, 32 * kTileWarps, 0, st>>>(a, h0);
  } else {
    esrg_fast_kernel<R, 32><<<grid, 128, 0, st>>>(a);
  }
  GI_CUDA(cudaGetLastError());
  if (ph) ph->mark("esrg_exact_ties");
  const int eg = kNumSMs * 8;
  if (k <= 4) esrg_query_kernel<R, 4><<<eg, 128, 0, st>>>(a);
  else if (k <= 8) esrg_query_kernel<R, 8><<<eg, 128, 0, st>>>(a);
  else if (k <= 16) esrg_query_kernel<R, 16><<<eg, 128, 0, st>>>(a);
  else if (k <= 32) esrg_query_kernel<R, 32><<<eg, 128, 0, st>>>(a);
  else {  // any k: the lists live in scratch memory (topk.cuh, KC = 0), one per thread of the grid-stride loop
    a.tk_stride = (int64_t)eg * 128;
    GI_TRY(scr.alloc(&a.tk_d, (size_t)a.tk_stride * k));
    GI_TRY(scr.alloc(&a.tk_id, (size_t)a.tk_stride * k));
    esrg_query_kernel<R, 0><<<eg, 128, 0, st>>>(a);
  }
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}

}  // namespace

// idw_esrg_nearest (interpolation.f90:201-288 + query_esrg.f90) at scattered targets (m, 2) instead of the nodes of
// the grid: the grid only defines the cell list; every target goes through the frame-by-frame kernel.
template <class R>
int esrg_points_device(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
                       int len_y, R grid_spac, const R* targets, int64_t num_targets, int k, R* data_ip,
                       int32_t* nn_index, R* nn_dist2, int flags, cudaStream_t st) {
  GI_REQUIRE(k >= 1, "num_neighbours must be at least 1");
  GI_REQUIRE(num_points >= k, "fewer points than num_neighbours");
  GI_REQUIRE(num_targets >= 0, "negative number of targets");
  GI_TRY(ensure_device());
  if (num_targets == 0) return GI_OK;
  StatusSlot slot;
  GI_TRY(get_status_slot(st, &slot));
  Scratch scr(st);
  Phases ph(st);
  ph.mark("esrg_bin");
  Binned<R> b;
  GI_TRY(esrg_bin<R>(scr, points, num_points, data_in, x_axis, len_x, y_axis, len_y, grid_spac, flags, &b));
  ph.mark("esrg_query");
  EsrgArgs<R> a;
  a.win = TargetWindow<R>{nullptr, nullptr, 0, 0, 0, 0, 0, 0, 0};
  a.table = b.table; a.sxy = b.sxy; a.sval = b.sval; a.sid = b.sid; a.scell = b.scell;
  a.len_x = len_x; a.len_y = len_y; a.k = k; a.spac = grid_spac;
  a.out = data_ip; a.nn_index = nn_index; a.nn_dist2 = nn_dist2;
  int64_t* counters;
  GI_TRY(scr.alloc(&counters, 4));
  GI_CUDA(cudaMemsetAsync(counters, 0, sizeof(int64_t) * 4, st));
  a.counters = counters; a.stats = 0; a.status = slot.dev;
  a.todo = nullptr; a.todo_count = counters + 3; a.all_exact = 1;
  a.tk_d = nullptr; a.tk_id = nullptr; a.tk_stride = 0;
  a.tgt = points_n2(targets, num_targets, flags & GI_TARGETS_ROWMAJOR); a.num_tgt = num_targets;
  a.x_axis = x_axis; a.y_axis = y_axis;
  const int eg = kNumSMs * 8;
  if (k <= 4) esrg_query_kernel<R, 4><<<eg, 128, 0, st>>>(a);
  else if (k <= 8) esrg_query_kernel<R, 8><<<eg, 128, 0, st>>>(a);
  else if (k <= 16) esrg_query_kernel<R, 16><<<eg, 128, 0, st>>>(a);
  else if (k <= 32) esrg_query_kernel<R, 32><<<eg, 128, 0, st>>>(a);
  else {
    a.tk_stride = (int64_t)eg * 128;
    GI_TRY(scr.alloc(&a.tk_d, (size_t)a.tk_stride * k));
    GI_TRY(scr.alloc(&a.tk_id, (size_t)a.tk_stride * k));
    esrg_query_kernel<R, 0><<<eg, 128, 0, st>>>(a);
  }
  GI_CUDA(cudaGetLastError());
  ph.finish();
  GI_TRY(flush_status(st, slot));
  return GI_OK;
}

template <class R>
int esrg_device(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
                int len_y, R grid_spac, int k, int row_begin, int row_end, R* data_ip, int32_t* nn_index,
                R* nn_dist2, int64_t* counters_out, int flags, cudaStream_t st) {
  GI_REQUIRE(k >= 1, "num_neighbours must be at least 1");
  GI_REQUIRE(num_points >= k, "fewer points than num_neighbours");
  GI_TRY(ensure_device());
  if (row_begin == row_end) return GI_OK;
  StatusSlot slot;
  GI_TRY(get_status_slot(st, &slot));
  Scratch scr(st);
  Phases ph(st);
  ph.mark("esrg_bin");
  Binned<R> b;
  GI_TRY(esrg_bin(scr, points, num_points, data_in, x_axis, len_x, y_axis, len_y, grid_spac, flags, &b));
  ph.mark("esrg_query");
  int64_t* counters;
  GI_TRY(scr.alloc(&counters, 4));
  GI_CUDA(cudaMemsetAsync(counters, 0, sizeof(int64_t) * 4, st));
  GI_TRY(esrg_query_window<R>(scr, b, num_points, x_axis, len_x, y_axis, len_y, grid_spac, k, row_begin, row_end, 0,
                              len_x, data_ip, nn_index, nn_dist2, counters, counters_out != nullptr, flags, slot.dev,
                              &ph));
  ph.finish();
  if (counters_out)
    GI_CUDA(cudaMemcpyAsync(counters_out, counters, sizeof(int64_t) * 4, cudaMemcpyDeviceToDevice, st));
  GI_TRY(flush_status(st, slot));
  return GI_OK;
}

namespace {
// output bands computed on the compute stream and copied back on the copy stream (common.cuh)
template <class R>
int esrg_run_banded(HostPipe* pipe, Scratch& scr, const Binned<R>& bins, int num_points, const R* dx, int len_x,
                    const R* dy, int len_y, R grid_spac, int k, int row_begin, int row_end, R* dout, R* data_ip,
                    int flags, const StatusSlot& slot) {
  cudaStream_t st = pipe->compute;
  const int nrows = row_end - row_begin;
  const bool rowmajor = flags & GI_FIELD_ROWMAJOR;
  const int64_t lo = rowmajor ? row_begin : 0, hi = rowmajor ? row_end : len_x;
  const int64_t line = rowmajor ? len_x : nrows;
  const int nb = plan_bands(hi - lo, (int64_t)nrows * len_x * sizeof(R));
  int64_t* counters;
  GI_TRY(scr.alloc(&counters, (size_t)4 * nb));
  GI_CUDA(cudaMemsetAsync(counters, 0, sizeof(int64_t) * 4 * nb, st));
  for (int b = 0; b < nb; ++b) {
    int64_t b0, b1;
    band_range(lo, hi, nb, b, 4, &b0, &b1);
    if (b0 >= b1) continue;
    R* band_out = dout + (b0 - lo) * line;
    if (rowmajor)
      GI_TRY(esrg_query_window<R>(scr, bins, num_points, dx, len_x, dy, len_y, grid_spac, k, (int)b0, (int)b1, 0,
                                  len_x, band_out, nullptr, nullptr, counters + 4 * b, false, flags, slot.dev,
                                  nullptr));
    else
      GI_TRY(esrg_query_window<R>(scr, bins, num_points, dx, len_x, dy, len_y, grid_spac, k, row_begin, row_end,
                                  (int)b0, (int)b1, band_out, nullptr, nullptr, counters + 4 * b, false, flags,
                                  slot.dev, nullptr));
    if (nb > 1) {  // (a single band needs no second stream: small calls are latency-bound)
      GI_CUDA(cudaEventRecord(pipe->computed[b], st));
      GI_CUDA(cudaStreamWaitEvent(pipe->copy_out, pipe->computed[b], 0));
    }
    GI_TRY(to_host(nb > 1 ? pipe->copy_out : st, data_ip + (b0 - lo) * line, band_out, (size_t)((b1 - b0) * line)));
  }
  GI_TRY(flush_status(st, slot));
  if (nb > 1) GI_CUDA(cudaStreamSynchronize(pipe->copy_out));
  return GI_OK;
}
}  // namespace

// GI_HOST form without debug outputs: inputs in, points binned once, output bands computed and copied back in a
// pipeline (common.cuh).
template <class R>
int esrg_host(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
              int len_y, R grid_spac, int k, int row_begin, int row_end, R* data_ip, int flags) {
  GI_REQUIRE(k >= 1, "num_neighbours must be at least 1");
  GI_REQUIRE(num_points >= k, "fewer points than num_neighbours");
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
  GI_TRY(to_device(scr, points, (size_t)2 * num_points, &dp));
  GI_TRY(to_device(scr, data_in, (size_t)num_points, &dd));
  GI_TRY(to_device(scr, x_axis, (size_t)len_x, &dx));
  GI_TRY(to_device(scr, y_axis, (size_t)len_y, &dy));
  const int nrows = row_end - row_begin;
  GI_TRY(scr.alloc(&dout, (size_t)nrows * len_x));
  ph.mark("esrg_bin");
  Binned<R> bins;
  GI_TRY(esrg_bin(scr, dp, num_points, dd, dx, len_x, dy, len_y, grid_spac, flags, &bins));
  ph.mark("esrg_query");
  GI_TRY(esrg_run_banded<R>(pipe, scr, bins, num_points, dx, len_x, dy, len_y, grid_spac, k, row_begin, row_end, dout,
                            data_ip, flags, slot));
  ph.finish();
  return read_status(st, true);
}

// ---------------------------------------------------------------------------------------------
// plan: binned points + target grid resident on the device, one execute() per field of nodal values
// ---------------------------------------------------------------------------------------------
namespace {
template <class R> struct EsrgPlan : PlanBase {
  Scratch keep{nullptr, true};
  Binned<R> bins;
  R *dx, *dy;
  R spac;
  int num_points, len_x, len_y, k, row_begin, row_end, flags;

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
    if (mem == GI_HOST) {
      ph.mark("h2d");
      R* tmp;
      GI_TRY(to_device(scr, data_in, (size_t)num_points, &tmp));
      dd = tmp;
      GI_TRY(scr.alloc(&dout, (size_t)(row_end - row_begin) * len_x));
    }
    ph.mark("esrg_values");
    esrg_values_kernel<R><<<div_up(num_points, 256), 256, 0, st>>>(dd, bins.sid, num_points, bins.sval);
    GI_CUDA(cudaGetLastError());
    ph.mark("esrg_query");
    if (mem == GI_HOST) {
      GI_TRY(esrg_run_banded<R>(pipe, scr, bins, num_points, dx, len_x, dy, len_y, spac, k, row_begin, row_end, dout,
                                data_ip, flags, slot));
      ph.finish();
      return read_status(st, true);
    }
    int64_t* counters;
    GI_TRY(scr.alloc(&counters, 4));
    GI_CUDA(cudaMemsetAsync(counters, 0, sizeof(int64_t) * 4, st));
    GI_TRY(esrg_query_window<R>(scr, bins, num_points, dx, len_x, dy, len_y, spac, k, row_begin, row_end, 0, len_x,
                                dout, nullptr, nullptr, counters, false, flags, slot.dev, nullptr));
    ph.finish();
    return flush_status(st, slot);
  }
};
}  // namespace

template <class R>
int esrg_plan_create(const R* points, int num_points, const R* x_axis, int len_x, const R* y_axis, int len_y,
                     R grid_spac, int k, int row_begin, int row_end, int flags, int mem, cudaStream_t st_in,
                     PlanBase** out) {
  GI_REQUIRE(k >= 1, "num_neighbours must be at least 1");
  GI_REQUIRE(num_points >= k, "fewer points than num_neighbours");
  GI_TRY(ensure_device());
  cudaStream_t st = st_in;
  if (mem == GI_HOST) { HostPipe* pipe; GI_TRY(host_pipe(&pipe)); st = pipe->compute; }
  std::unique_ptr<EsrgPlan<R>> plan(new EsrgPlan<R>());
  plan->real_bytes = sizeof(R);
  GI_CUDA(cudaGetDevice(&plan->device));
  plan->num_points = num_points; plan->len_x = len_x; plan->len_y = len_y; plan->k = k; plan->spac = grid_spac;
  plan->row_begin = row_begin; plan->row_end = row_end; plan->flags = flags;
  Scratch tmp(st);
  const R* dp = points;
  GI_TRY(plan->keep.alloc(&plan->dx, (size_t)len_x));
  GI_TRY(plan->keep.alloc(&plan->dy, (size_t)len_y));
  const cudaMemcpyKind kind = mem == GI_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  GI_CUDA(cudaMemcpyAsync(plan->dx, x_axis, sizeof(R) * (size_t)len_x, kind, st));
  GI_CUDA(cudaMemcpyAsync(plan->dy, y_axis, sizeof(R) * (size_t)len_y, kind, st));
  if (mem == GI_HOST) {
    R* t;
    GI_TRY(to_device(tmp, points, (size_t)2 * num_points, &t));
    dp = t;
  }
  GI_TRY(esrg_bin<R>(plan->keep, tmp, dp, num_points, (const R*)nullptr, plan->dx, len_x, plan->dy, len_y, grid_spac,
                     flags, true, &plan->bins));
  GI_CUDA(cudaStreamSynchronize(st));  // the plan may be used from any stream afterwards
  *out = plan.release();
  return GI_OK;
}

template <class R>
int esrg_bin_device(const R* points, int num_points, const R* x_axis, int len_x, const R* y_axis, int len_y,
                    R grid_spac, int32_t* index_of_pts, int32_t* indptr, int flags, cudaStream_t st) {
  GI_TRY(ensure_device());
  Scratch scr(st);
  Binned<R> b;
  GI_TRY(esrg_bin<R>(scr, points, num_points, nullptr, x_axis, len_x, y_axis, len_y, grid_spac, flags, &b));
  GI_CUDA(cudaMemcpyAsync(index_of_pts, b.sid, sizeof(int) * (size_t)num_points, cudaMemcpyDeviceToDevice, st));
  esrg_export_indptr_kernel<<<kNumSMs * 4, 256, 0, st>>>(b.table, (int64_t)len_x * len_y, indptr);
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}

#define GI_INSTANTIATE(R)                                                                                        \
  template int esrg_device<R>(const R*, int, const R*, const R*, int, const R*, int, R, int, int, int, R*, int32_t*, \
                              R*, int64_t*, int, cudaStream_t);                                                  \
  template int esrg_points_device<R>(const R*, int, const R*, const R*, int, const R*, int, R, const R*, int64_t,  \
                                     int, R*, int32_t*, R*, int, cudaStream_t);                                   \
  template int esrg_bin_device<R>(const R*, int, const R*, int, const R*, int, R, int32_t*, int32_t*, int,        \
                                  cudaStream_t);                                                                  \
  template int esrg_host<R>(const R*, int, const R*, const R*, int, const R*, int, R, int, int, int, R*, int);  \
  template int esrg_plan_create<R>(const R*, int, const R*, int, const R*, int, R, int, int, int, int, int,       \
                                   cudaStream_t, PlanBase**);
GI_INSTANTIATE(double)
GI_INSTANTIATE(float)

}  // namespace gi

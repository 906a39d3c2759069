// esrg.cu -- IDW from scattered points to an equally spaced regular grid with the cell-list ("esrg")
// k-nearest-neighbour search.
//
// Replaces (reference file:line):
//   interpolation.f90:201-288   idw_esrg_nearest (operator + IDW weights)
//   query_esrg.f90:23-91        assign_points_to_cells (cell binning, CSR by grid cell)
//   query_esrg.f90:97-249       nearest_neighbours_esrg (ring-expanding exact kNN from the cell centre)
//
// B200 design (not a port):
//   binning   one pass computes each point's cell key and a histogram (atomics into ONE int32 table of
//             len_y*len_x entries), an in-place exclusive scan turns it into offsets, a second pass scatters
//             the point ids with atomicAdd on the same table -- which leaves the table holding the END
//             offset of every cell (start(c) = table[c-1]), so no second 268 MB table is needed at 8192^2.
//             The reference's fill is stable (ascending point id inside a cell, query_esrg.f90:72-89); the
//             atomic scatter is not, so cells holding more than one point are insertion-sorted afterwards.
//             Coordinates, values and ids are then gathered into cell order so the query reads them
//             contiguously.
//   query     one thread per grid target, register-resident top-k (topk.cuh).  A frame's top and bottom
//             rows are contiguous in the CSR (key = iy*len_x + ix), so each costs two offset reads and one
//             contiguous run of points.  The reference's two fixed 1000-entry candidate lists
//             (query_esrg.f90:119-121) become one small in-place deferred list per thread; when it
//             overflows the thread switches to an exact list-free mode that re-scans the inner frames with
//             the predicate "was deferred at the previous level" (d2 > r2_prev), which reproduces the
//             reference's processing order and therefore its tie behaviour for any point density.
#include "kernels.cuh"
#include "scan.cuh"
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
esrg_key_count_kernel(PointsView<R> pts, int n, const R* __restrict__ x_axis, int len_x,
                      const R* __restrict__ y_axis, int len_y, R spac, int* __restrict__ keys,
                      int* __restrict__ table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ix = cell_index(pts.x(i), __ldg(x_axis), spac, len_x);
  const int iy = cell_index(pts.y(i), __ldg(y_axis), spac, len_y);
  const int key = iy * len_x + ix;  // CSR order: ind_y outer, ind_x inner (query_esrg.f90:63-68)
  keys[i] = key;
  atomicAdd(table + key, 1);
}

__global__ void __launch_bounds__(256)
esrg_scatter_kernel(const int* __restrict__ keys, int n, int* __restrict__ table, int* __restrict__ sid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int slot = atomicAdd(table + keys[i], 1);  // table[c] ends as the END offset of cell c
  sid[slot] = i + 1;                               // 1-based ids like index_of_pts
}

// restore the reference's stable order inside every cell that holds more than one point
__global__ void __launch_bounds__(256)
esrg_sort_cells_kernel(const int* __restrict__ keys, int n, const int* __restrict__ table, int* __restrict__ sid) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int key = keys[sid[p] - 1];
  const int start = key > 0 ? table[key - 1] : 0;
  if (start != p) return;
  const int end = table[key];
  for (int i = start + 1; i < end; ++i) {
    const int v = sid[i];
    int j = i - 1;
    while (j >= start && sid[j] > v) { sid[j + 1] = sid[j]; --j; }
    sid[j + 1] = v;
  }
}

template <class R> struct alignas(16) XY { R x, y; };

template <class R>
__global__ void __launch_bounds__(256)
esrg_gather_kernel(PointsView<R> pts, const R* __restrict__ data, const int* __restrict__ sid, int n,
                   XY<R>* __restrict__ sxy, R* __restrict__ sval) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int i = sid[p] - 1;
  XY<R> v; v.x = pts.x(i); v.y = pts.y(i);
  sxy[p] = v;
  if (sval) sval[p] = __ldg(data + i);
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
  int len_x, len_y, k;
  R spac;
  R* out;                 // window-shaped
  int32_t* nn_index;      // optional (window, k)
  R* nn_dist2;            // optional
  int64_t* counters;      // [0] cells scanned, [1] max level, [2] list overflows
  int* status;
};

template <class R, int KC>
__global__ void __launch_bounds__(128) esrg_query_kernel(const EsrgArgs<R> a) {
  const TargetWindow<R>& w = a.win;
  const int f = w.f0 + blockIdx.x * 32 + (threadIdx.x & 31);
  const int s = w.s0 + blockIdx.y * 4 + (threadIdx.x >> 5);
  if (f >= w.f1 || s >= w.s1) return;
  const int ix = w.fast_is_y ? s : f, iy = w.fast_is_y ? f : s;
  const int len_x = a.len_x, len_y = a.len_y;
  const R cx = __ldg((w.fast_is_y ? w.slow_axis : w.fast_axis) + ix);     // query_esrg.f90:133-134
  const R cy = __ldg((w.fast_is_y ? w.fast_axis : w.slow_axis) + iy);

  TopK<R, KC> tk;
  tk.init(a.k, Consts<R>::huge(), -1);                                    // :129-130
  int found = 0;

  int list_pos[kListCap];
  R list_d2[kListCap];
  int n_list = 0;
  bool listfree = false;  // set once the deferred list overflowed
  int overflows = 0;
  long long cells_scanned = 0;

  R radius_sq = R(0), radius_sq_prev = R(0);
  // One candidate (query_esrg.f90:158-169 / :229-240).  `fresh` = first time this point is seen (frame ==
  // level); otherwise it is a re-scanned inner-frame point that counts only if it was deferred so far.
  auto consider = [&](int pos, R dist_sq, bool fresh) {
    if (!fresh && !(dist_sq > radius_sq_prev)) return;   // already consumed at an earlier level
    if (dist_sq > radius_sq) {                            // defer
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
  const int level_cap = len_x + len_y;
  for (;;) {
    const R rad = a.spac * (R(level) + R(0.5));                           // :141,:181
    radius_sq_prev = radius_sq;
    radius_sq = rad * rad;
    if (listfree) {
      // exact list-free mode: inner frames first (the order the deferred list would have), then the new one
      for (int l = 0; l < level; ++l) scan_frame(l, false);
    } else {
      // carried (deferred) candidates first, compacting in place (:184-185,:227-233)
      const int n_old = n_list;
      n_list = 0;
      for (int i = 0; i < n_old; ++i) {
        const int pos = list_pos[i];
        const R dist_sq = list_d2[i];
        if (dist_sq > radius_sq) { list_pos[n_list] = pos; list_d2[n_list] = dist_sq; ++n_list; }
        else if (dist_sq < tk.dlast) { tk.template insert<true>(dist_sq, pos); found = min(found + 1, a.k); }
      }
    }
    scan_frame(level, true);
    if (found == a.k) break;                                              // :178,:243
    ++level;
    if (level > level_cap) { raise_status(a.status, GI_ERR_TOO_FEW_POINTS); break; }
  }

  // IDW (interpolation.f90:260-274): distances are square-rooted first (query_esrg.f90:247)
  const int64_t o = (int64_t)(s - w.s0) * w.nf() + (f - w.f0);
  R numerator = R(0), denominator = R(0), first_val = R(0), first_dist = R(0);
#pragma unroll
  for (int i = 0; i < KC; ++i) {
    if (i < a.k) {
      const int pos = tk.id[i];
      const R dist = sqrt(tk.d[i]);
      const R v = pos >= 0 ? __ldg(a.sval + pos) : R(0);
      if (i == 0) { first_val = v; first_dist = dist; }
      numerator = numerator + (v / dist);
      denominator = denominator + (R(1) / dist);
      if (a.nn_index) a.nn_index[o * a.k + i] = pos >= 0 ? __ldg(a.sid + pos) : -1;
      if (a.nn_dist2) a.nn_dist2[o * a.k + i] = tk.d[i];
    }
  }
  a.out[o] = (first_dist <= Consts<R>::tiny_e4()) ? first_val : numerator / denominator;

  if (a.counters) {
    atomicAdd((unsigned long long*)a.counters, (unsigned long long)cells_scanned);
    atomicMax((unsigned long long*)(a.counters + 1), (unsigned long long)level);
    if (overflows) atomicAdd((unsigned long long*)(a.counters + 2), (unsigned long long)overflows);
  }
}

template <class R> struct Binned {
  int* table;
  int* sid;
  int* keys;
  XY<R>* sxy;
  R* sval;
};

template <class R>
int esrg_bin(Scratch& scr, const R* points, int n, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
             int len_y, R spac, int flags, Binned<R>* b) {
  cudaStream_t st = scr.stream();
  const int64_t cells = (int64_t)len_x * len_y;
  GI_REQUIRE(cells < (int64_t)0x7fffffff, "grid has more than 2^31 cells");
  GI_TRY(scr.alloc(&b->table, (size_t)cells));
  GI_TRY(scr.alloc(&b->keys, (size_t)n));
  GI_TRY(scr.alloc(&b->sid, (size_t)n));
  GI_TRY(scr.alloc(&b->sxy, (size_t)n));
  b->sval = nullptr;
  if (data_in) GI_TRY(scr.alloc(&b->sval, (size_t)n));
  GI_CUDA(cudaMemsetAsync(b->table, 0, sizeof(int) * (size_t)cells, st));
  const PointsView<R> pv = points_n2(points, n, flags & GI_POINTS_ROWMAJOR);
  const int blocks = div_up(n, 256);
  esrg_key_count_kernel<R><<<blocks, 256, 0, st>>>(pv, n, x_axis, len_x, y_axis, len_y, spac, b->keys, b->table);
  GI_CUDA(cudaGetLastError());
  GI_TRY(exclusive_scan_i32(scr, b->table, b->table, cells));
  esrg_scatter_kernel<<<blocks, 256, 0, st>>>(b->keys, n, b->table, b->sid);
  esrg_sort_cells_kernel<<<blocks, 256, 0, st>>>(b->keys, n, b->table, b->sid);
  esrg_gather_kernel<R><<<blocks, 256, 0, st>>>(pv, data_in, b->sid, n, b->sxy, b->sval);
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}

}  // namespace

template <class R>
int esrg_device(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
                int len_y, R grid_spac, int k, int row_begin, int row_end, R* data_ip, int32_t* nn_index,
                R* nn_dist2, int64_t* counters_out, int flags, cudaStream_t st) {
  GI_REQUIRE(k >= 1 && k <= 32, "num_neighbours must be in 1..32");
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
  EsrgArgs<R> a;
  const bool rowmajor = flags & GI_FIELD_ROWMAJOR;
  a.win = rowmajor ? TargetWindow<R>{x_axis, y_axis, len_x, len_y, 0, len_x, row_begin, row_end, 0}
                   : TargetWindow<R>{y_axis, x_axis, len_y, len_x, row_begin, row_end, 0, len_x, 1};
  a.table = b.table; a.sxy = b.sxy; a.sval = b.sval; a.sid = b.sid;
  a.len_x = len_x; a.len_y = len_y; a.k = k; a.spac = grid_spac;
  a.out = data_ip; a.nn_index = nn_index; a.nn_dist2 = nn_dist2;
  a.counters = counters_out ? counters : nullptr; a.status = slot.dev;
  dim3 grid(div_up(a.win.nf(), 32), div_up(a.win.ns(), 4));
  GI_REQUIRE(grid.y <= 65535, "window too tall for one launch");
  if (k <= 4) esrg_query_kernel<R, 4><<<grid, 128, 0, st>>>(a);
  else if (k <= 8) esrg_query_kernel<R, 8><<<grid, 128, 0, st>>>(a);
  else if (k <= 16) esrg_query_kernel<R, 16><<<grid, 128, 0, st>>>(a);
  else esrg_query_kernel<R, 32><<<grid, 128, 0, st>>>(a);
  GI_CUDA(cudaGetLastError());
  ph.finish();
  if (counters_out)
    GI_CUDA(cudaMemcpyAsync(counters_out, counters, sizeof(int64_t) * 4, cudaMemcpyDeviceToDevice, st));
  GI_TRY(flush_status(st, slot));
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
  template int esrg_bin_device<R>(const R*, int, const R*, int, const R*, int, R, int32_t*, int32_t*, int,        \
                                  cudaStream_t);
GI_INSTANTIATE(double)
GI_INSTANTIATE(float)

}  // namespace gi

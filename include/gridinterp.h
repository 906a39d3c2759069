/* gridinterp.h -- C ABI of libgridinterp.so (CUDA, sm_100a)
 *
 * Drop-in boundary for the hot path of ChristianSteger/grid_interpolation: point
 * location + interpolation for every target point.  Each entry point below
 * replaces one Fortran module procedure that the reference's f2py binding calls;
 * the reference interface is cited as file:line (paths relative to the reference
 * checkout).  Argument meaning follows the Fortran dummies; the dimensions f2py
 * hides (intent(hide)) are explicit here.  No CUDA / torch types appear in any
 * signature: plain pointers, sizes, ints.
 *
 * Memory space (`mem`)
 *   GI_HOST    every pointer is host memory; the call copies in, computes on the
 *              current CUDA device, copies out and returns when the result is in
 *              the caller's buffer (pinned buffers -- gi_host_alloc -- make the
 *              copies run at full PCIe speed).  `stream` is ignored.
 *   GI_DEVICE  every pointer is device memory of the current device (e.g. the
 *              data_ptr() of a torch CUDA tensor); the work is enqueued on
 *              `stream` (a cudaStream_t passed as void*, NULL = default stream)
 *              and the call returns without waiting.  Errors detected on the
 *              device (iteration caps, invalid ids) are collected per stream:
 *              gi_stream_status() synchronises and returns them.
 *
 * Layouts (`flags`)
 *   0 means "exactly the reference's memory layout": every 2-D array is Fortran
 *   column-major as the f2py wrapper hands it to the Fortran --
 *       points(num_points,2) = x[0..n) then y[0..n)      (bilinear, esrg, barycentric)
 *       points(2,num_points) = x0,y0,x1,y1,...           (idw_kdtree, interpolation.f90:111)
 *       simplices/neighbours(num_triangles,3) = 3 columns of num_triangles
 *       data_in/data_ip(len_y,len_x): element (iy,ix) at iy + ix*len_y
 *   The GI_*_ROWMAJOR bits select the C-order (numpy default) alternative for one
 *   group of arguments so that no host-side transposition is ever needed.  All
 *   kernels are stride-generic: both layouts run at the same speed.
 *
 * Index conventions are the reference's: simplices / neighbours are 1-based,
 * neighbour 0 = no neighbour (interpolation.f90:338, test_interpolation.py:199-201).
 * Debug outputs return 1-based ids as the Fortran variables hold them; 0 = none.
 *
 * Every function returns a gi_status; gi_last_error() gives the message of the
 * last failure on the calling thread.
 */
#ifndef GRIDINTERP_H
#define GRIDINTERP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GI_VERSION 200

typedef enum {
  GI_OK = 0,
  GI_ERR_BAD_ARG = 1,        /* shape / pointer / flag error (f2py raises ValueError before the call) */
  GI_ERR_NO_DEVICE = 2,      /* no usable CUDA device: there is NO CPU fallback */
  GI_ERR_OOM = 3,
  GI_ERR_CUDA = 4,           /* CUDA runtime error, see gi_last_error() */
  GI_ERR_ITER_CAP = 5,       /* walk / ring search exceeded its cap (reference: infinite loop) */
  GI_ERR_TOO_FEW_POINTS = 6, /* fewer than num_neighbours points (reference: infinite loop, query_esrg.f90:178) */
  GI_ERR_BAD_INDEX = 7,      /* simplices / neighbours id out of range (reference: out-of-bounds read) */
  GI_ERR_HALO = 8,           /* row-band call: hull fill needs rows beyond the supplied halo */
  GI_ERR_PACK_WINDOW = 9,    /* GI_BARY_BAND_PACK: a walk left the triangles packed for the band; retry without the flag */
  GI_ERR_TIMEOUT = 10        /* gi_peer_wait: a peer did not signal within the timeout */
} gi_status;

#define GI_HOST 0
#define GI_DEVICE 1

#define GI_POINTS_ROWMAJOR 1 /* points array is C-order instead of Fortran order */
#define GI_TOPO_ROWMAJOR 2   /* simplices / neighbours are C-order (num_triangles,3) */
#define GI_FIELD_ROWMAJOR 4  /* data_in (bilinear) / data_ip (grid outputs): element (iy,ix) at iy*len_x + ix */
#define GI_KD_EXACT_PRUNE 32 /* idw_kdtree: prune the far subtree iff diff*diff >= max d2 (the geometrically correct
                               rule) instead of the reference's abs(diff) >= max d2 -> exact k nearest neighbours
                               even where the k-th d2 is < 1.  An extension (the reference README's to-do list);
                               results then differ from the reference's wherever its rule is too eager. */
#define GI_BILINEAR_SEARCH_AXES 64 /* bilinear: find the cell by bisection on the axis values instead of the
                                      reference's mean-spacing formula (interpolation.f90:49-58,67,73), which is
                                      only right for equally spaced axes.  An extension: results differ
                                      from the reference's wherever its formula picks the wrong cell. */
#define GI_KD_FAITHFUL_BUILD 128 /* idw_kdtree / gi_kd_build: build the tree by evaluating the reference's quickselect
                                    passes (kd_tree.f90:97-155) instead of by ranks.  The two trees are identical
                                    unless coordinates tie; the library detects ties and switches by itself, the
                                    flag forces the slower build (debug / tests). */
#define GI_BARY_REFERENCE_ORDER 256 /* barycentric_interpolation: locate EVERY target by the reference's own
                                       row-sequential walk from ind_tri_start (interpolation.f90:362-404; debug:
                                       one thread per grid row).  Default: only the targets inside the margin
                                       band of a mesh edge are re-located along that path -- or all of them
                                       when there are more than the fix-up list holds. */
#define GI_BARY_BAND_PACK 512 /* barycentric row-band calls with GI_DEVICE buffers (multi-GPU sharding): turn only the
                                 triangles whose y-range meets the walked rows (band + halo + 48 rows) into records
                                 instead of the whole mesh -- 1/8 of the packing work for one of 8 bands.  Links to
                                 skipped triangles are marked, and a walk that wants to cross one is reported as
                                 GI_ERR_PACK_WINDOW by gi_stream_status(): the caller then repeats the call without
                                 the flag (GI_HOST calls do both by themselves).  Results are identical. */
#define GI_BARY_PLANE_VALUES 2048 /* barycentric entry points and plans (opt-in): evaluate a target's value as the
                                     plane through its triangle's three nodal values, d3 + gx*(x-x3) + gy*(y-y3) with
                                     the gradient precomputed per triangle, instead of through the reference's weight
                                     expressions (interpolation.f90:385-398).  Triangle ids, the outside-hull mask and
                                     fill sources stay bit-identical; values agree with the reference's to
                                     1e-12 * max|data_in| in double precision (needle triangles, max|edge|^2 >
                                     128 |denom|, keep the exact expression).  The default is the bit-identical form. */
#define GI_TARGETS_ROWMAJOR 1024 /* *_points entry points: the (num_targets, 2) target array is C-order */
#define GI_ESRG_REFERENCE_ORDER 16 /* idw_esrg_nearest: process every target in the reference's candidate order
                                      (debug).  Default: only targets with exactly equal neighbour distances,
                                      the only case in which that order changes the result. */
#define GI_KD_REFERENCE_VISITS 8 /* idw_kdtree: visit exactly the nodes the reference visits (debug: makes the
                                    visit counter comparable).  Default: additionally skip far subtrees with
                                    diff*diff >= max d2, which provably cannot change the neighbour list. */

/* ---- library / device management ------------------------------------------------------------ */
int gi_version(void);
const char* gi_last_error(void);
int gi_device_count(void);
int gi_set_device(int device);
/* Pinned host memory for GI_HOST calls (page-locked: H2D / D2H at full PCIe rate). */
void* gi_host_alloc(uint64_t bytes);
void gi_host_free(void* p);
/* cudaStreamSynchronize(stream) + the first device-side error recorded on it since the last call. */
int gi_stream_status(void* stream);
/* Environment variables read by the library (debugging aids, none is needed in production):
 *   GI_KD_TRACE=1            print the number of rounds of the quickselect-faithful k-d tree build to stderr
 *   GI_BARY_TRACE=1          print, per walked window, what the barycentric fix-up paths had to do (flagged targets,
 *                            fix-ups, rows redone along the reference's path, repair rounds); synchronises
 *   GI_BARY_RASTER=1         the tile rasteriser instead of the walk for locate / evaluate (same results)
 *   GI_KD_NO_SMEM_FINISH=1, GI_KD_NO_SERIAL_FINISH=1   k-d builds by launches / rounds down to the leaves (A/B) */
/* Phase profiling: gi_set_profiling(1) starts a fresh recording on the calling thread; from then on every
 * call made by that thread records CUDA events around its phases (h2d, pack_mesh, walk, fill, d2h, ...) on
 * the stream it runs on, without synchronising.  gi_last_timings() waits for the last recorded event and
 * returns the number of entries since the recording started; ms[i] is the device time of entry i and
 * gi_last_timing_name(i) its phase name ("" marks the end of one call, 0 ms). */
void gi_set_profiling(int on);
int gi_last_timings(double* ms, int capacity);
const char* gi_last_timing_name(int i);
/* GI_HOST calls overlap their PCIe copies with the kernels: the grid operators compute and copy back their
 * output in bands along the axis that is not contiguous in memory, bilinear streams its points through in
 * chunks (results are identical to a one-shot call).  A band / chunk holds about `bytes` of output (default
 * 32 MB, at most 16 bands per call); 0 restores the default.  Process-wide tuning knob. */
void gi_set_pipeline_band_bytes(uint64_t bytes);
/* Return cached scratch memory of the current device to the driver: the stream-ordered pool and the 8 MB arenas
 * (one per stream a call ran on) that serve the small allocations of every call. */
void gi_release_workspace(void);
/* Capacity (targets per walked window) of the work list of the barycentric ambiguity fix-up, default 4 194 304;
 * when more targets lie within the margin band of a mesh edge the window is redone by the row-sequential walk.
 * A small value forces that path (tests).  <= 0 restores the default.  Process-wide. */
void gi_set_fixup_list_capacity(int64_t targets);
/* Sliver guard for barycentric_interpolation -- the reference's open to-do (README.md:41-43: "prohibiting barycentric
 * interpolation for sliver triangles close to the edge (and filling the subsequent missing values from the nearest
 * neighbours)").  A triangle with a hull edge whose quality |2 area| / max |edge|^2 (equilateral: 0.87) is below
 * min_quality is not interpolated in: its targets are masked like targets outside the hull and filled by
 * fill_nearest_neighbour (triangle_walk.f90:90-132).  0 (default) = off = the reference's behaviour.  Read when the
 * mesh is packed (one-shot calls: every call; plans: at creation).  Process-wide. */
void gi_set_sliver_guard(double min_quality);

/* ---- multi-GPU: peer memory and the fused gather (one process per GPU) --------------------------------- */
/* The reference is a single shared-memory process (4 OpenMP loops, interpolation.f90:62,152,249,362).  Here the
 * grid targets of interpolation.f90:362-404 are sharded in row bands over the GPUs of one node, and the final
 * gather of the field is not a separate collective: the kernels store every result into the field of every
 * destination GPU over NVLink (gi_barycentric_interpolation_gather_*).  The plumbing is CUDA IPC:
 *   gi_peer_alloc    device memory of the current device that other processes may map (zero-filled)
 *   gi_peer_export   GI_PEER_HANDLE_BYTES opaque bytes to send to the other ranks (any transport)
 *   gi_peer_open     map another rank's allocation into this process -> device pointer usable in kernels
 *   gi_peer_close / gi_peer_free
 *   gi_peer_signal   after all work enqueued so far on `stream`: flag_arrays[d][slot] = value for d < num
 *                    (system-scope release store; flag arrays are uint32 peer allocations, slot = this rank)
 *   gi_peer_wait     enqueue a wait on `stream` until flags[i] >= value for all i < num_slots (this rank's own
 *                    flag array); bounded by timeout_s (<= 30 s) -> GI_ERR_TIMEOUT via gi_stream_status()
 * so "all N bands have arrived in my copy of the field" is a stream-ordered event with no host round trip. */
#define GI_MAX_GATHER_FIELDS 8
#define GI_PEER_HANDLE_BYTES 64
void* gi_peer_alloc(uint64_t bytes);
void gi_peer_free(void* p);
int gi_peer_export(void* p, void* handle64);
int gi_peer_open(const void* handle64, void** p);
int gi_peer_close(void* p);
int gi_peer_signal(uint32_t* const* flag_arrays, int num, int slot, uint32_t value, void* stream);
int gi_peer_wait(const uint32_t* flags, int num_slots, uint32_t value, double timeout_s, void* stream);

/* ---- the four drop-in entry points ---------------------------------------------------------- */

/* interpolation.f90:20-100  SUBROUTINE bilinear(x_axis, len_x, y_axis, len_y, data_in, points,
 * num_points, data_ip): regular grid -> scattered points.  Out-of-grid points get -9999.0
 * (interpolation.f90:44,68-78; the reference then reads out of bounds, here the point is skipped). */
int gi_bilinear_f64(const double* x_axis, int len_x, const double* y_axis, int len_y,
                    const double* data_in, const double* points, int64_t num_points, double* data_ip,
                    int flags, int mem, void* stream);
int gi_bilinear_f32(const float* x_axis, int len_x, const float* y_axis, int len_y, const float* data_in,
                    const float* points, int64_t num_points, float* data_ip, int flags, int mem,
                    void* stream);

/* interpolation.f90:107-194  SUBROUTINE idw_kdtree(points, num_points, data_in, x_axis, len_x, y_axis,
 * len_y, num_neighbours, data_ip) with kd_tree.f90:30-240 (build + kNN query, including the reference's
 * prune rule |diff| < max d2, kd_tree.f90:191,198, and the 1.0e6 sentinel, interpolation.f90:159). */
int gi_idw_kdtree_f64(const double* points, int num_points, const double* data_in, const double* x_axis,
                      int len_x, const double* y_axis, int len_y, int num_neighbours, double* data_ip,
                      int flags, int mem, void* stream);

/* interpolation.f90:201-288  SUBROUTINE idw_esrg_nearest(points, num_points, data_in, x_axis, len_x,
 * y_axis, len_y, grid_spac, num_neighbours, data_ip) with query_esrg.f90:23-249 (cell binning + ring
 * search; no fixed 1000-entry candidate buffers). */
int gi_idw_esrg_nearest_f64(const double* points, int num_points, const double* data_in,
                            const double* x_axis, int len_x, const double* y_axis, int len_y,
                            double grid_spac, int num_neighbours, double* data_ip, int flags, int mem,
                            void* stream);

/* interpolation.f90:298-416  SUBROUTINE barycentric_interpolation(points, num_points, data_in, x_axis,
 * len_x, y_axis, len_y, simplices, neighbours, num_triangles, data_ip) with triangle_walk.f90:24-153
 * (visibility walk, orientation predicate with the -1e-10 margin, nearest-inside-cell hull fill). */
int gi_barycentric_interpolation_f64(const double* points, int num_points, const double* data_in,
                                     const double* x_axis, int len_x, const double* y_axis, int len_y,
                                     const int32_t* simplices, const int32_t* neighbours,
                                     int num_triangles, double* data_ip, int flags, int mem, void* stream);

/* ---- extended forms: row bands (multi-GPU sharding) + the indices the reference keeps internal -- */

/* Rows [row_begin, row_end) of the (len_y, len_x) output only; data_ip then holds just that band,
 * shape (row_end-row_begin, len_x) in the layout `flags` selects.  Optional outputs (NULL to skip), same
 * band shape and layout as data_ip unless stated:
 *   tri_out   triangle the walk ended in (find_triangle's ind_tri, triangle_walk.f90:32), 1-based
 *   mask_out  1 where the target is outside the convex hull (mask_outside, interpolation.f90:334,400)
 *   fill_src  for masked cells the 0-based row-major id iy*len_x+ix of the source cell chosen by
 *             fill_nearest_neighbour (triangle_walk.f90:115-123); -1 elsewhere
 *   counters  int64[4]: walk iterations total (iters_tot, interpolation.f90:374), masked cells,
 *             max fill level, targets resolved by the ambiguity fix-up pass
 * `halo` = extra rows walked on each side of the band so that the hull fill sees neighbouring bands'
 * cells; GI_ERR_HALO if a fill needed more. */
int gi_barycentric_interpolation_ex_f64(const double* points, int num_points, const double* data_in,
                                        const double* x_axis, int len_x, const double* y_axis, int len_y,
                                        const int32_t* simplices, const int32_t* neighbours,
                                        int num_triangles, int row_begin, int row_end, int halo,
                                        double* data_ip, int32_t* tri_out, uint8_t* mask_out,
                                        int64_t* fill_src, int64_t* counters, int flags, int mem,
                                        void* stream);

/* Morton order of bilinear's points (north_star: Morton-ordered targets / Morton ranges per GPU): order[j] = 0-based
 * index of the j-th point along the Z curve of the field's 16 x 16-node blocks (cell of a point as
 * interpolation.f90:67,73; points outside the grid last; stable).  bilinear's field reads are random gathers
 * (interpolation.f90:62-96): points taken in this order read each part of the field once, and a contiguous range
 * of the order touches a compact part of the field -- the shard for one of several GPUs.  Only the order is
 * computed; moving the points (and the results back) is left to the caller. */
int gi_bilinear_morton_order_f64(const double* x_axis, int len_x, const double* y_axis, int len_y,
                                 const double* points, int64_t num_points, int32_t* order, int flags, int mem,
                                 void* stream);
int gi_bilinear_morton_order_f32(const float* x_axis, int len_x, const float* y_axis, int len_y, const float* points,
                                 int64_t num_points, int32_t* order, int flags, int mem, void* stream);

/* ---- scattered targets: the reference interpolates to the nodes of a regular grid only
 * (interpolation.f90:154-158,365-371); these take any (num_targets, 2) array of target coordinates instead -------- */

/* barycentric_interpolation (interpolation.f90:298-416) at scattered targets: the same visibility walk
 * (triangle_walk.f90:24-84) from a seed grid laid over the mesh, the same weights (:385-398).  A target outside the
 * convex hull takes the value at the CLOSEST LOCATION ON THE MESH (the behaviour the reference's header comment
 * describes, interpolation.f90:292-293; its grid version takes the nearest node inside the hull instead): the hull is
 * followed from the edge the walk left through until the foot of the perpendicular lies on an edge or a vertex is
 * closest.  Optional outputs (NULL to skip): tri_out (num_targets) 1-based triangle the value was evaluated in,
 * outside_out (num_targets) 1 = outside the hull.  flags: GI_POINTS_ROWMAJOR, GI_TOPO_ROWMAJOR, GI_TARGETS_ROWMAJOR. */
int gi_barycentric_points_f64(const double* points, int num_points, const double* data_in, const int32_t* simplices,
                              const int32_t* neighbours, int num_triangles, const double* targets,
                              int64_t num_targets, double* data_ip, int32_t* tri_out, uint8_t* outside_out, int flags,
                              int mem, void* stream);
int gi_barycentric_points_f32(const float* points, int num_points, const float* data_in, const int32_t* simplices,
                              const int32_t* neighbours, int num_triangles, const float* targets, int64_t num_targets,
                              float* data_ip, int32_t* tri_out, uint8_t* outside_out, int flags, int mem, void* stream);
/* idw_kdtree (interpolation.f90:107-194 + kd_tree.f90) at scattered targets: same tree, same traversal and prune
 * rule, same sums.  points (2, num_points) as in gi_idw_kdtree; optional nn_index / nn_dist2 (num_targets, k). */
int gi_idw_kdtree_points_f64(const double* points, int num_points, const double* data_in, const double* targets,
                             int64_t num_targets, int num_neighbours, double* data_ip, int32_t* nn_index,
                             double* nn_dist2, int flags, int mem, void* stream);
int gi_idw_kdtree_points_f32(const float* points, int num_points, const float* data_in, const float* targets,
                             int64_t num_targets, int num_neighbours, float* data_ip, int32_t* nn_index, float* nn_dist2,
                             int flags, int mem, void* stream);
/* idw_esrg_nearest (interpolation.f90:201-288 + query_esrg.f90) at scattered targets: the grid (x_axis, y_axis,
 * grid_spac) only defines the cell list the points are binned into; the ring search runs around the target's own
 * cell and a level's radius is grid_spac * L instead of grid_spac * (L + 1/2) (the reference's targets sit at the
 * cell centres, query_esrg.f90:133-134,141; a target anywhere in its cell is at least L cells from everything
 * outside the (2L+1)^2 block) -> the exact k nearest points in ascending order, ties in the reference's order. */
int gi_idw_esrg_nearest_points_f64(const double* points, int num_points, const double* data_in, const double* x_axis,
                                   int len_x, const double* y_axis, int len_y, double grid_spac,
                                   const double* targets, int64_t num_targets, int num_neighbours, double* data_ip,
                                   int32_t* nn_index, double* nn_dist2, int flags, int mem, void* stream);
int gi_idw_esrg_nearest_points_f32(const float* points, int num_points, const float* data_in, const float* x_axis,
                                   int len_x, const float* y_axis, int len_y, float grid_spac, const float* targets,
                                   int64_t num_targets, int num_neighbours, float* data_ip, int32_t* nn_index,
                                   float* nn_dist2, int flags, int mem, void* stream);

/* Mesh ingest ("GPU Delaunay / mesh ingest"; no counterpart in the reference, which takes simplices and
 * neighbours in Qhull's facet order and the points in the caller's order, test_interpolation.py:64-67): the same
 * mesh renumbered along the Z curve of position, so that every kernel that follows a link reads nearby memory.
 *   triangle_order[j] = 1-based input id of the triangle that becomes triangle j+1: triangles sorted (stable) by the
 *     Z value of their centroid ((x1+x2+x3)/3, (y1+y2+y3)/3); Z value of a position = 16 bits per axis,
 *     q = trunc((v - v_min) * (65535 / (v_max - v_min))) over the points' bounding box, x in the even bits;
 *   simplices_out / neighbours_out: row j = the input row of that triangle (vertex order and neighbour columns
 *     kept, so every interpolated value is computed from the same operands), triangle ids renumbered, 0 = none;
 *   points_out / point_order (both or neither): the points sorted (stable) by their own Z value, point_order[j] =
 *     1-based input id of the point that becomes point j+1, simplices_out then holds the new vertex ids and the
 *     caller permutes its nodal values the same way (data_in[point_order - 1]).
 * Layouts as `flags` says for the inputs (GI_POINTS_ROWMAJOR, GI_TOPO_ROWMAJOR); outputs in the same layouts.
 * An id out of range raises GI_ERR_BAD_INDEX. */
int gi_mesh_reorder_f64(const double* points, int num_points, const int32_t* simplices, const int32_t* neighbours,
                        int num_triangles, double* points_out, int32_t* point_order, int32_t* simplices_out,
                        int32_t* neighbours_out, int32_t* triangle_order, int flags, int mem, void* stream);
int gi_mesh_reorder_f32(const float* points, int num_points, const int32_t* simplices, const int32_t* neighbours,
                        int num_triangles, float* points_out, int32_t* point_order, int32_t* simplices_out,
                        int32_t* neighbours_out, int32_t* triangle_order, int flags, int mem, void* stream);

/* Row band [row_begin, row_end) of barycentric_interpolation computed on the current device and stored at its
 * place in num_fields (<= GI_MAX_GATHER_FIELDS) FULL (len_y, len_x) fields in the layout `flags` selects:
 * fields[0] must be memory of the current device (the hull fill reads its source cells there), the others are the
 * same field on other GPUs (gi_peer_open).  All pointers are device pointers, the work is enqueued on `stream`
 * (GI_DEVICE semantics).  Replaces the OpenMP row loop interpolation.f90:362-404 + the fill :412 for one band. */
int gi_barycentric_interpolation_gather_f64(const double* points, int num_points, const double* data_in,
                                            const double* x_axis, int len_x, const double* y_axis, int len_y,
                                            const int32_t* simplices, const int32_t* neighbours,
                                            int num_triangles, int row_begin, int row_end, int halo,
                                            double* const* fields, int num_fields, int flags, void* stream);

/* nn_index (band rows, len_x, k): 1-based ids in ascending-distance order as neighbours_index /
 * index_nn hold them (kd_tree.f90:236, query_esrg.f90:168), k fastest; nn_dist2 squared distances
 * (kd: neighbours(3,:), esrg: dist_sq_nn).  counters int64[4]: node visits (kd) or cells scanned (esrg),
 * max ring level, candidate-list overflows, targets resolved in the reference's candidate order (ties). */
int gi_idw_kdtree_ex_f64(const double* points, int num_points, const double* data_in,
                         const double* x_axis, int len_x, const double* y_axis, int len_y,
                         int num_neighbours, int row_begin, int row_end, double* data_ip,
                         int32_t* nn_index, double* nn_dist2, int64_t* counters, int flags, int mem,
                         void* stream);
int gi_idw_esrg_nearest_ex_f64(const double* points, int num_points, const double* data_in,
                               const double* x_axis, int len_x, const double* y_axis, int len_y,
                               double grid_spac, int num_neighbours, int row_begin, int row_end,
                               double* data_ip, int32_t* nn_index, double* nn_dist2, int64_t* counters,
                               int flags, int mem, void* stream);

/* ---- plans: geometry prepared once, many fields ------------------------------------------------- */

/* The reference rebuilds its search structure in every call (interpolation.f90:139,235-239; the start-triangle
 * search :342-355).  A plan keeps what depends on the geometry only -- the packed triangle records + seed grid,
 * the k-d tree, or the cell list -- together with the target grid (axes, row band, layout flags) on the device;
 * gi_plan_execute_* then interpolates one field of nodal values `data_in` (num_points) into `data_ip`
 * (row_end-row_begin, len_x), with results identical to the one-shot call on the same inputs.
 * `mem` / `stream` of create describe points / axes / topology, those of execute describe data_in / data_ip
 * (a GI_HOST execute pipelines its copies like the one-shot calls).  A plan belongs to the device that was
 * current at creation and serves one execute at a time.  Extension: no counterpart in the
 * reference's f2py module. */
typedef struct gi_plan gi_plan;
int gi_barycentric_plan_create_f64(const double* points, int num_points, const double* x_axis, int len_x,
                                   const double* y_axis, int len_y, const int32_t* simplices,
                                   const int32_t* neighbours, int num_triangles, int row_begin, int row_end,
                                   int halo, int flags, int mem, void* stream, gi_plan** plan);
int gi_idw_kdtree_plan_create_f64(const double* points, int num_points, const double* x_axis, int len_x,
                                  const double* y_axis, int len_y, int num_neighbours, int row_begin,
                                  int row_end, int flags, int mem, void* stream, gi_plan** plan);
int gi_idw_esrg_plan_create_f64(const double* points, int num_points, const double* x_axis, int len_x,
                                const double* y_axis, int len_y, double grid_spac, int num_neighbours,
                                int row_begin, int row_end, int flags, int mem, void* stream, gi_plan** plan);
int gi_plan_execute_f64(gi_plan* plan, const double* data_in, double* data_ip, int mem, void* stream);
void gi_plan_destroy(gi_plan* plan);

/* ---- parity / debug exports for the search structures ----------------------------------------- */

/* kd_tree.f90:30-95: the canonical balanced tree (node = rank (n+1)/2 element along axis depth mod 2)
 * as an implicit array: order[] (num_points, 1-based point ids) lists the points so that the node of any
 * segment [lo,hi) is element lo+(hi-lo+1)/2-1, left subtree [lo,m), right subtree (m,hi). */
int gi_kd_build_f64(const double* points, int num_points, int32_t* order, int flags, int mem, void* stream);

/* query_esrg.f90:23-91 assign_points_to_cells: index_of_pts (num_points, 1-based, stable inside a cell),
 * indptr (len_y*len_x + 1, 0-based offsets, CSR order ind_y outer / ind_x inner as query_esrg.f90:63-68). */
int gi_esrg_assign_points_to_cells_f64(const double* points, int num_points, const double* x_axis, int len_x,
                                       const double* y_axis, int len_y, double grid_spac,
                                       int32_t* index_of_pts, int32_t* indptr, int flags, int mem,
                                       void* stream);

/* triangle_walk.f90:90-132 fill_nearest_neighbour on a caller-supplied mask (uint8, 1 = masked) and
 * field, both (len_y, len_x) in the layout `flags` selects; data is filled in place. */
int gi_fill_nearest_neighbour_f64(const uint8_t* mask_outside, int len_x, int len_y, double* data_ip,
                                  int64_t* fill_src, int flags, int mem, void* stream);

/* ---- single precision (the reference as shipped: every REAL is default kind) ------------- */
/* Same entry points, argument meaning and flags with float instead of double (grid_spac included). */
int gi_idw_kdtree_f32(const float* points, int num_points, const float* data_in, const float* x_axis, int len_x,
                      const float* y_axis, int len_y, int num_neighbours, float* data_ip, int flags, int mem,
                      void* stream);
int gi_idw_esrg_nearest_f32(const float* points, int num_points, const float* data_in, const float* x_axis,
                            int len_x, const float* y_axis, int len_y, float grid_spac, int num_neighbours,
                            float* data_ip, int flags, int mem, void* stream);
int gi_barycentric_interpolation_f32(const float* points, int num_points, const float* data_in,
                                     const float* x_axis, int len_x, const float* y_axis, int len_y,
                                     const int32_t* simplices, const int32_t* neighbours, int num_triangles,
                                     float* data_ip, int flags, int mem, void* stream);
int gi_barycentric_interpolation_ex_f32(const float* points, int num_points, const float* data_in,
                                        const float* x_axis, int len_x, const float* y_axis, int len_y,
                                        const int32_t* simplices, const int32_t* neighbours, int num_triangles,
                                        int row_begin, int row_end, int halo, float* data_ip, int32_t* tri_out,
                                        uint8_t* mask_out, int64_t* fill_src, int64_t* counters, int flags,
                                        int mem, void* stream);
int gi_barycentric_interpolation_gather_f32(const float* points, int num_points, const float* data_in,
                                            const float* x_axis, int len_x, const float* y_axis, int len_y,
                                            const int32_t* simplices, const int32_t* neighbours, int num_triangles,
                                            int row_begin, int row_end, int halo, float* const* fields,
                                            int num_fields, int flags, void* stream);
int gi_idw_kdtree_ex_f32(const float* points, int num_points, const float* data_in, const float* x_axis,
                         int len_x, const float* y_axis, int len_y, int num_neighbours, int row_begin,
                         int row_end, float* data_ip, int32_t* nn_index, float* nn_dist2, int64_t* counters,
                         int flags, int mem, void* stream);
int gi_idw_esrg_nearest_ex_f32(const float* points, int num_points, const float* data_in, const float* x_axis,
                               int len_x, const float* y_axis, int len_y, float grid_spac, int num_neighbours,
                               int row_begin, int row_end, float* data_ip, int32_t* nn_index, float* nn_dist2,
                               int64_t* counters, int flags, int mem, void* stream);
int gi_barycentric_plan_create_f32(const float* points, int num_points, const float* x_axis, int len_x,
                                   const float* y_axis, int len_y, const int32_t* simplices,
                                   const int32_t* neighbours, int num_triangles, int row_begin, int row_end,
                                   int halo, int flags, int mem, void* stream, gi_plan** plan);
int gi_idw_kdtree_plan_create_f32(const float* points, int num_points, const float* x_axis, int len_x,
                                  const float* y_axis, int len_y, int num_neighbours, int row_begin, int row_end,
                                  int flags, int mem, void* stream, gi_plan** plan);
int gi_idw_esrg_plan_create_f32(const float* points, int num_points, const float* x_axis, int len_x,
                                const float* y_axis, int len_y, float grid_spac, int num_neighbours, int row_begin,
                                int row_end, int flags, int mem, void* stream, gi_plan** plan);
int gi_plan_execute_f32(gi_plan* plan, const float* data_in, float* data_ip, int mem, void* stream);
int gi_kd_build_f32(const float* points, int num_points, int32_t* order, int flags, int mem, void* stream);
int gi_esrg_assign_points_to_cells_f32(const float* points, int num_points, const float* x_axis, int len_x,
                                       const float* y_axis, int len_y, float grid_spac, int32_t* index_of_pts,
                                       int32_t* indptr, int flags, int mem, void* stream);
int gi_fill_nearest_neighbour_f32(const uint8_t* mask_outside, int len_x, int len_y, float* data_ip,
                                  int64_t* fill_src, int flags, int mem, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRIDINTERP_H */

// kernels.cuh -- device-level entry points (device pointers in, work enqueued on a stream).
// api.cu wraps these into the C ABI of include/gridinterp.h (host / device memory variants).
#pragma once
#include "common.cuh"

namespace gi {

// bilinear.cu -- interpolation.f90:20-100
template <class R>
int bilinear_device(const R* x_axis, int len_x, const R* y_axis, int len_y, const R* data_in, const R* points,
                    int64_t num_points, R* data_ip, int flags, cudaStream_t st);

// order[j] = 0-based index of the j-th point along the Z curve of the field's 16 x 16-node blocks (device pointers)
template <class R>
int bilinear_morton_order_device(const R* x_axis, int len_x, const R* y_axis, int len_y, const R* points,
                                 int64_t num_points, int32_t* order, int flags, cudaStream_t st);

// barycentric.cu -- interpolation.f90:298-416 + triangle_walk.f90
// counters (device, optional): int64[4] = walk iterations, masked cells, max fill level, fix-ups
template <class R>
int barycentric_device(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x,
                       const R* y_axis, int len_y, const int32_t* simplices, const int32_t* neighbours,
                       int num_tri, int row_begin, int row_end, int halo, R* data_ip, int32_t* tri_out,
                       uint8_t* mask_out, int64_t* fill_src, int64_t* counters, int flags, cudaStream_t st);
// multi-GPU form: the band is stored at its place in num_fields full (len_y, len_x) fields (fields[0] local, the
// others peer-mapped), see barycentric.cu GatherDest
template <class R>
int barycentric_gather_device(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x,
                              const R* y_axis, int len_y, const int32_t* simplices, const int32_t* neighbours,
                              int num_tri, int row_begin, int row_end, int halo, R* const* fields, int num_fields,
                              int flags, cudaStream_t st);
template <class R>
int fill_device(const uint8_t* mask, int len_x, int len_y, R* data_ip, int64_t* fill_src, int flags,
                cudaStream_t st);

// esrg.cu -- interpolation.f90:201-288 + query_esrg.f90
template <class R>
int esrg_device(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
                int len_y, R grid_spac, int k, int row_begin, int row_end, R* data_ip, int32_t* nn_index,
                R* nn_dist2, int64_t* counters, int flags, cudaStream_t st);
template <class R>
int esrg_bin_device(const R* points, int num_points, const R* x_axis, int len_x, const R* y_axis, int len_y,
                    R grid_spac, int32_t* index_of_pts, int32_t* indptr, int flags, cudaStream_t st);

// kdtree.cu -- interpolation.f90:107-194 + kd_tree.f90
template <class R>
int kdtree_device(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x,
                  const R* y_axis, int len_y, int k, int row_begin, int row_end, R* data_ip,
                  int32_t* nn_index, R* nn_dist2, int64_t* counters, int flags, cudaStream_t st);
template <class R>
int kd_build_device(const R* points, int num_points, int32_t* order, int flags, cudaStream_t st);

// scattered targets (m, 2) instead of grid nodes; tri_out / outside_out / nn_* optional
template <class R>
int barycentric_points_device(const R* points, int num_points, const R* data_in, const int32_t* simplices,
                              const int32_t* neighbours, int num_tri, const R* targets, int64_t num_targets,
                              R* data_ip, int32_t* tri_out, uint8_t* outside_out, int flags, cudaStream_t st);
template <class R>
int kdtree_points_device(const R* points, int num_points, const R* data_in, const R* targets, int64_t num_targets, int k,
                         R* data_ip, int32_t* nn_index, R* nn_dist2, int flags, cudaStream_t st);

template <class R>
int esrg_points_device(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
                       int len_y, R grid_spac, const R* targets, int64_t num_targets, int k, R* data_ip,
                       int32_t* nn_index, R* nn_dist2, int flags, cudaStream_t st);

// mesh.cu -- mesh ingest: triangles (and optionally points) renumbered along the Z curve of their position
template <class R>
int mesh_reorder_device(const R* points, int num_points, const int32_t* simplices, const int32_t* neighbours,
                        int num_tri, R* points_out, int32_t* point_order, int32_t* simplices_out,
                        int32_t* neighbours_out, int32_t* triangle_order, int flags, cudaStream_t st);

// GI_HOST forms of the four operators (host pointers in, synchronous): copies and compute are pipelined over
// output bands / point chunks (common.cuh, HostPipe).  No debug outputs; api.cu falls back to "copy in, *_device,
// copy out" when those are requested.
template <class R>
int kdtree_host(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
                int len_y, int k, int row_begin, int row_end, R* data_ip, int flags);
template <class R>
int bilinear_host(const R* x_axis, int len_x, const R* y_axis, int len_y, const R* data_in, const R* points,
                  int64_t num_points, R* data_ip, int flags);
template <class R>
int barycentric_host(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
                     int len_y, const int32_t* simplices, const int32_t* neighbours, int num_tri, int row_begin,
                     int row_end, int halo, R* data_ip, int flags);
template <class R>
int esrg_host(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
              int len_y, R grid_spac, int k, int row_begin, int row_end, R* data_ip, int flags);

// plans (gi_plan): geometry + target grid prepared once, execute() per field of nodal values
template <class R>
int barycentric_plan_create(const R* points, int num_points, const R* x_axis, int len_x, const R* y_axis, int len_y,
                            const int32_t* simplices, const int32_t* neighbours, int num_tri, int row_begin,
                            int row_end, int halo, int flags, int mem, cudaStream_t st, PlanBase** out);
template <class R>
int kdtree_plan_create(const R* points, int num_points, const R* x_axis, int len_x, const R* y_axis, int len_y, int k,
                       int row_begin, int row_end, int flags, int mem, cudaStream_t st, PlanBase** out);
template <class R>
int esrg_plan_create(const R* points, int num_points, const R* x_axis, int len_x, const R* y_axis, int len_y,
                     R grid_spac, int k, int row_begin, int row_end, int flags, int mem, cudaStream_t st,
                     PlanBase** out);

}  // namespace gi

"""The reference's Python surface, re-hosted on libgridinterp.so.

Mirrors the f2py module that the reference builds from interpolation.f90 (``from interpolation import
interpolation as ip_fortran``, test_interpolation.py:27): same four callables, same positional order and
keyword names (the Fortran dummy names with the ``intent(hide)`` dimensions removed):

    bilinear(x_axis, y_axis, data_in, points)                                        interpolation.f90:20-38
    idw_kdtree(points, data_in, x_axis, y_axis, num_neighbours)                      interpolation.f90:107-127
    idw_esrg_nearest(points, data_in, x_axis, y_axis, grid_spac, num_neighbours)     interpolation.f90:201-223
    barycentric_interpolation(points, data_in, x_axis, y_axis, simplices, neighbours) interpolation.f90:298-322

Like f2py the functions accept any array-like in any memory order and allocate the output.  Differences:
arithmetic is float64 (the shipped reference is float32); 2-D outputs are returned with the same
shape ``(len_y, len_x)`` and, by default, the same Fortran memory order as f2py (``order="C"`` selects C order;
both are produced directly by the kernels, nothing is transposed); shape errors raise ValueError as f2py's do,
device-side failures (which hang or corrupt memory in the reference) raise RuntimeError.

Inputs may also be torch CUDA tensors: the call then runs on the current torch stream on device buffers and
returns a torch tensor (PyTorch is only a device-buffer provider here).

The ``*_ex`` variants additionally take a row band (multi-GPU sharding) and return the indices the reference
keeps internal (triangle ids, neighbour ids) for parity testing.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from types import SimpleNamespace

import numpy as np

from . import _lib
from ._lib import (GI_DEVICE, GI_FIELD_ROWMAJOR, GI_HOST, GI_KD_REFERENCE_VISITS, GI_POINTS_ROWMAJOR,
                   GI_TOPO_ROWMAJOR, check, lib)

__all__ = ["bilinear", "idw_kdtree", "idw_esrg_nearest", "barycentric_interpolation",
           "barycentric_interpolation_ex", "idw_kdtree_ex", "idw_esrg_nearest_ex", "kd_build",
           "assign_points_to_cells", "fill_nearest_neighbour", "pinned_empty", "pinned_copy",
           "set_profiling", "set_pipeline_band_bytes", "set_fixup_list_capacity", "set_sliver_guard", "barycentric_points", "idw_kdtree_points", "idw_esrg_nearest_points", "bilinear_morton_order", "mesh_reorder", "release_workspace", "pinned_trim",
           "last_timings", "BarycentricPlan", "KdTreePlan", "EsrgPlan"]

_F64 = np.dtype(np.float64)


def _real(dtype):
    """(numpy dtype, C-symbol suffix) of the arithmetic kind: float64 (graded) or float32 (as shipped)."""
    dt = np.dtype(dtype)
    if dt == np.float64:
        return dt, "f64"
    if dt == np.float32:
        return dt, "f32"
    raise ValueError("dtype must be float64 or float32")


# -------------------------------------------------------------------------------------------------
# pinned host buffers (page-locked: H2D / D2H at PCIe rate), cached because cudaHostAlloc is slow
# -------------------------------------------------------------------------------------------------
class _PinnedPool:
    """Cache of page-locked buffers in exact-size buckets.  The cap (default 4 GiB, $GI_PINNED_CACHE_BYTES) bounds
    what stays cached after the arrays are collected; pinned_trim() releases it."""

    def __init__(self, cap_bytes=None):
        if cap_bytes is None:
            cap_bytes = int(os.environ.get("GI_PINNED_CACHE_BYTES", 4 << 30))
        self.free = {}
        self.cached = 0
        self.cap = cap_bytes
        self.lock = threading.Lock()

    def take(self, nbytes):
        with self.lock:
            lst = self.free.get(nbytes)
            if lst:
                self.cached -= nbytes
                return lst.pop()
        p = lib().gi_host_alloc(nbytes)
        if not p:
            raise MemoryError(f"gi_host_alloc({nbytes}) failed: {_lib.last_error()}")
        return p

    def give(self, ptr, nbytes):
        with self.lock:
            if self.cached + nbytes <= self.cap:
                self.free.setdefault(nbytes, []).append(ptr)
                self.cached += nbytes
                return
        lib().gi_host_free(ptr)

    def trim(self, keep_bytes=0):
        with self.lock:
            for nbytes in sorted(self.free, reverse=True):
                lst = self.free[nbytes]
                while lst and self.cached > keep_bytes:
                    lib().gi_host_free(lst.pop())
                    self.cached -= nbytes


_pool = _PinnedPool()


class _PinnedOwner:
    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes

    def __del__(self):
        try:
            _pool.give(self.ptr, self.nbytes)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64, order="C"):
    """numpy array backed by page-locked memory from libgridinterp (returned to a pool when collected)."""
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(int(s) for s in shape)
    nbytes = max(int(np.prod(shape, dtype=np.int64)) * dtype.itemsize, 1)
    nbytes = (nbytes + 4095) // 4096 * 4096
    ptr = _pool.take(nbytes)
    buf = (C.c_char * nbytes).from_address(ptr)
    buf._gi_owner = _PinnedOwner(ptr, nbytes)
    flat = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64)))
    return flat.reshape(shape, order=order)


def pinned_copy(a, dtype=None, order="K"):
    """Copy an array into pinned memory keeping its memory order (``order="K"``) or forcing C / F."""
    a = np.asarray(a)
    dtype = np.dtype(dtype or a.dtype)
    if order == "K":
        order = "F" if (a.ndim > 1 and a.flags.f_contiguous and not a.flags.c_contiguous) else "C"
    out = pinned_empty(a.shape, dtype, order)
    out[...] = a
    return out


# -------------------------------------------------------------------------------------------------
# argument marshalling
# -------------------------------------------------------------------------------------------------
def _is_torch(x):
    return type(x).__module__.split(".")[0] == "torch" and hasattr(x, "data_ptr")


class _Arg:
    """A marshalled array argument: raw pointer + whether the 2-D layout is C (row-major) order."""
    __slots__ = ("ptr", "rowmajor", "keep", "shape")


def _np_arg(a, dtype, ndim, name):
    a = np.asarray(a)
    if a.ndim != ndim:
        raise ValueError(f"{name} must be {ndim}-dimensional, got shape {a.shape}")
    if a.dtype != dtype:
        a = a.astype(dtype, order="K")  # f2py casts silently as well
    arg = _Arg()
    if ndim == 1:
        a = np.ascontiguousarray(a)
        arg.rowmajor = True
    elif a.flags.f_contiguous and not a.flags.c_contiguous:
        arg.rowmajor = False
    else:
        a = np.ascontiguousarray(a)
        arg.rowmajor = True
    arg.ptr, arg.keep, arg.shape = a.ctypes.data, a, a.shape
    return arg


def _torch_arg(t, dtype, ndim, name):
    import torch
    tdtype = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
              np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
    if not t.is_cuda:
        raise ValueError(f"{name}: torch tensors must live on a CUDA device (got {t.device})")
    if t.dim() != ndim:
        raise ValueError(f"{name} must be {ndim}-dimensional, got shape {tuple(t.shape)}")
    if t.dtype != tdtype:
        t = t.to(tdtype)
    arg = _Arg()
    if ndim == 2 and not t.is_contiguous() and t.t().is_contiguous():
        arg.rowmajor = False
    else:
        t = t.contiguous()
        arg.rowmajor = True
    arg.ptr, arg.keep, arg.shape = t.data_ptr(), t, tuple(t.shape)
    return arg


_have_device = None


def _require_device():
    """There is no CPU fallback: fail loudly when no CUDA device is visible."""
    global _have_device
    if _have_device is None:
        _have_device = lib().gi_device_count() > 0
    if not _have_device:
        raise RuntimeError("no CUDA device available: grid_interpolation_b200 has no CPU fallback")


class _Call:
    """Decides host vs device path from the first array argument and marshals the rest consistently."""

    def __init__(self, first):
        self.torch = _is_torch(first)
        self.mem = GI_DEVICE if self.torch else GI_HOST
        self.stream = None
        self.device = None
        if self.torch:
            import torch
            self.device = first.device
            self.stream = torch.cuda.current_stream(self.device).cuda_stream

    def arg(self, a, dtype, ndim, name):
        if self.torch:
            if not _is_torch(a):
                import torch
                a = torch.as_tensor(np.asarray(a), device=self.device)
            return _torch_arg(a, dtype, ndim, name)
        if _is_torch(a):
            raise ValueError(f"{name}: mixing torch CUDA tensors and host arrays is not supported")
        return _np_arg(a, dtype, ndim, name)

    def out(self, shape, dtype, order):
        """(array, pointer, rowmajor) for an output of the given shape."""
        _require_device()
        if self.torch:
            import torch
            tdtype = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                      np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8,
                      np.dtype(np.int64): torch.int64}[np.dtype(dtype)]
            if order == "F" and len(shape) > 1:
                t = torch.empty(tuple(reversed(shape)), dtype=tdtype, device=self.device)
                view = t.permute(*reversed(range(len(shape))))
                return view, t.data_ptr(), False
            t = torch.empty(tuple(shape), dtype=tdtype, device=self.device)
            return t, t.data_ptr(), True
        a = pinned_empty(shape, dtype, order)
        return a, a.ctypes.data, order != "F" or len(shape) == 1

    def finish(self, status, what, sync=True):
        check(status, what)
        if self.torch and sync:
            import torch
            with torch.cuda.device(self.device):
                check(lib().gi_stream_status(self.stream), what)

    def __enter__(self):
        if self.torch:
            import torch
            self._guard = torch.cuda.device(self.device)
            self._guard.__enter__()
        return self

    def __exit__(self, *exc):
        if self.torch:
            self._guard.__exit__(*exc)
        return False


def _use_out(call, out, shape, dtype, order):
    """Honour a caller-supplied output array (numpy, C- or F-contiguous) instead of allocating one."""
    _require_device()  # after the shape checks, like f2py raising ValueError before the call
    if out is None:
        return call.out(shape, dtype, order)
    if call.torch:
        import torch
        tdtype = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
        # the kernels write sizeof(dtype)-byte elements through out.data_ptr(): anything else is a device fault
        if not _is_torch(out) or not out.is_cuda or out.device != call.device:
            raise ValueError(f"out must be a torch CUDA tensor on {call.device} (the device of the inputs)")
        if out.dtype != tdtype:
            raise ValueError(f"out must have dtype {tdtype}, got {out.dtype}")
        if tuple(out.shape) != tuple(shape):
            raise ValueError(f"out must have shape {shape}")
        if out.is_contiguous():
            return out, out.data_ptr(), True
        if out.dim() == 2 and out.t().is_contiguous():
            return out, out.data_ptr(), False
        raise ValueError("out must be contiguous")
    if not isinstance(out, np.ndarray):
        raise ValueError("out must be a numpy array when the inputs are host arrays")
    if out.shape != tuple(shape) or out.dtype != np.dtype(dtype):
        raise ValueError(f"out must have shape {shape} and dtype {np.dtype(dtype)}")
    if out.flags.c_contiguous:
        return out, out.ctypes.data, True
    if out.flags.f_contiguous:
        return out, out.ctypes.data, False
    raise ValueError("out must be C- or F-contiguous")


# -------------------------------------------------------------------------------------------------
# the four drop-in entry points
# -------------------------------------------------------------------------------------------------
def bilinear(x_axis, y_axis, data_in, points, *, search_axes=False, dtype=np.float64, out=None, sync=True):
    """Regular grid -> scattered points (interpolation.f90:20-100).

    data_in has shape (len_y, len_x), points (num_points, 2); returns data_ip (num_points,).  Points outside
    [x_axis[0], x_axis[-1]) x [y_axis[0], y_axis[-1]) get -9999.0 (interpolation.f90:44,68-78).
    ``search_axes=True`` (extension) finds the cell by bisection on the axis values, which is also right for axes
    that are not equally spaced (the reference assumes they are, interpolation.f90:49-58).
    """
    rt, suf = _real(dtype)
    with _Call(data_in) as c:
        x = c.arg(x_axis, rt, 1, "x_axis"); y = c.arg(y_axis, rt, 1, "y_axis")
        d = c.arg(data_in, rt, 2, "data_in"); p = c.arg(points, rt, 2, "points")
        len_x, len_y = x.shape[0], y.shape[0]
        if d.shape != (len_y, len_x):
            raise ValueError(f"data_in must have shape (len_y, len_x) = {(len_y, len_x)}, got {d.shape}")
        if p.shape[1] != 2:
            raise ValueError(f"points must have shape (num_points, 2), got {p.shape}")
        n = p.shape[0]
        res, res_ptr, _ = _use_out(c, out, (n,), rt, "C")
        flags = (GI_FIELD_ROWMAJOR if d.rowmajor else 0) | (GI_POINTS_ROWMAJOR if p.rowmajor else 0)
        flags |= _lib.GI_BILINEAR_SEARCH_AXES if search_axes else 0
        st = getattr(lib(), "gi_bilinear_" + suf)(x.ptr, len_x, y.ptr, len_y, d.ptr, p.ptr, n, res_ptr, flags, c.mem, c.stream)
        c.finish(st, "bilinear", sync)
        return res


def bilinear_morton_order(x_axis, y_axis, points, *, dtype=np.float64, sync=True):
    """Order of bilinear's ``points`` (num_points, 2) along the Z curve of the field's 16 x 16-node blocks (0-based
    int32 permutation, points outside the grid last).  ``bilinear(x, y, field, points[order])`` reads every part of
    the field once instead of gathering at random, and ``order[a:b]`` is a spatially compact shard (one GPU's share);
    un-permute with ``out[order] = result``.  See gridinterp.h."""
    rt, suf = _real(dtype)
    with _Call(points) as c:
        x = c.arg(x_axis, rt, 1, "x_axis"); y = c.arg(y_axis, rt, 1, "y_axis")
        p = c.arg(points, rt, 2, "points")
        if p.shape[1] != 2:
            raise ValueError(f"points must have shape (num_points, 2), got {p.shape}")
        n = p.shape[0]
        res, res_ptr, _ = c.out((n,), np.int32, "C")
        flags = GI_POINTS_ROWMAJOR if p.rowmajor else 0
        st = getattr(lib(), "gi_bilinear_morton_order_" + suf)(x.ptr, x.shape[0], y.ptr, y.shape[0], p.ptr, n, res_ptr, flags,
                                                              c.mem, c.stream)
        c.finish(st, "bilinear_morton_order", sync)
        return res


def mesh_reorder(points, simplices, neighbours, *, reorder_points=True, dtype=np.float64, sync=True):
    """Mesh ingest (gi_mesh_reorder, see gridinterp.h): the same triangle mesh renumbered along the Z curve of position.

    points (num_points, 2); simplices / neighbours (num_triangles, 3), 1-based, 0 = no neighbour (scipy's arrays + 1, as
    for barycentric_interpolation).  Returns a namespace with

    * ``simplices``, ``neighbours``: triangles sorted by the Z value of their centroid, ids renumbered (vertex order
      and neighbour columns kept: every interpolated value is computed from the same operands as on the input mesh);
    * ``triangle_order``: 1-based input id of each new triangle (``triangle_order[tri - 1]`` maps a triangle id found
      on the new mesh back);
    * with ``reorder_points`` (default): ``points`` sorted by their own Z value and ``point_order`` (1-based input
      ids) -- pass ``data_in[point_order - 1]`` as nodal values; otherwise ``points`` is the input and ``point_order``
      None.

    The reference takes the mesh in Qhull's facet order (test_interpolation.py:64-67), which is unrelated to position:
    on the 10 M-triangle headline mesh the pack pass of every stateless barycentric call drops from 0.50 to 0.33 ms
    on an ingested mesh."""
    rt, suf = _real(dtype)
    with _Call(points) as c:
        p = c.arg(points, rt, 2, "points")
        if p.shape[1] != 2:
            raise ValueError(f"points must have shape (num_points, 2), got {p.shape}")
        s = c.arg(simplices, np.dtype(np.int32), 2, "simplices")
        nb = c.arg(neighbours, np.dtype(np.int32), 2, "neighbours")
        t = s.shape[0]
        if s.shape != (t, 3):
            raise ValueError(f"simplices must have shape (num_triangles, 3), got {s.shape}")
        if nb.shape != (t, 3):
            raise ValueError(f"neighbours must have shape (num_triangles, 3) = {(t, 3)}, got {nb.shape}")
        if s.rowmajor != nb.rowmajor:
            nb = c.arg(_reorder(nb.keep, s.rowmajor), np.dtype(np.int32), 2, "neighbours")
        n = p.shape[0]
        if n < 3 or t < 1:
            raise ValueError("need at least 3 points and 1 triangle")
        so, so_ptr, _ = c.out((t, 3), np.int32, "C" if s.rowmajor else "F")
        no, no_ptr, _ = c.out((t, 3), np.int32, "C" if s.rowmajor else "F")
        to, to_ptr, _ = c.out((t,), np.int32, "C")
        po = porder = po_ptr = porder_ptr = None
        if reorder_points:
            po, po_ptr, _ = c.out((n, 2), rt, "C" if p.rowmajor else "F")
            porder, porder_ptr, _ = c.out((n,), np.int32, "C")
        flags = (GI_POINTS_ROWMAJOR if p.rowmajor else 0) | (GI_TOPO_ROWMAJOR if s.rowmajor else 0)
        st = getattr(lib(), "gi_mesh_reorder_" + suf)(p.ptr, n, s.ptr, nb.ptr, t, po_ptr, porder_ptr, so_ptr, no_ptr, to_ptr,
                                                     flags, c.mem, c.stream)
        c.finish(st, "mesh_reorder", sync)
        return SimpleNamespace(points=po if reorder_points else points, point_order=porder, simplices=so, neighbours=no,
                               triangle_order=to)


def _grid_common(c, points, data_in, x_axis, y_axis, points_shape, rt=_F64):
    x = c.arg(x_axis, rt, 1, "x_axis"); y = c.arg(y_axis, rt, 1, "y_axis")
    p = c.arg(points, rt, 2, "points"); d = c.arg(data_in, rt, 1, "data_in")
    if points_shape == "n2":
        if p.shape[1] != 2:
            raise ValueError(f"points must have shape (num_points, 2), got {p.shape}")
        n = p.shape[0]
    else:
        if p.shape[0] != 2:
            raise ValueError(f"points must have shape (2, num_points), got {p.shape}")
        n = p.shape[1]
    if d.shape != (n,):
        raise ValueError(f"data_in must have shape (num_points,) = ({n},), got {d.shape}")
    return x, y, p, d, n


def _band(rows, len_y):
    r0, r1 = (0, len_y) if rows is None else (int(rows[0]), int(rows[1]))
    if not (0 <= r0 <= r1 <= len_y):
        raise ValueError(f"rows must satisfy 0 <= begin <= end <= len_y, got {(r0, r1)}")
    return r0, r1


def idw_kdtree_ex(points, data_in, x_axis, y_axis, num_neighbours, *, rows=None, order="F", debug=False,
                  reference_visits=False, exact_prune=False, dtype=np.float64, out=None, sync=True):
    """idw_kdtree on the row band ``rows=(begin, end)``; debug=True also returns neighbour ids / d2 / visits.
    ``exact_prune=True`` (extension) replaces the reference's prune rule by the geometrically correct one."""
    rt, suf = _real(dtype)
    with _Call(points) as c:
        x, y, p, d, n = _grid_common(c, points, data_in, x_axis, y_axis, "2n", rt)
        len_x, len_y = x.shape[0], y.shape[0]
        r0, r1 = _band(rows, len_y)
        k = int(num_neighbours)
        res, res_ptr, rowmajor = _use_out(c, out, (r1 - r0, len_x), rt, order)
        flags = (GI_FIELD_ROWMAJOR if rowmajor else 0) | (GI_POINTS_ROWMAJOR if p.rowmajor else 0)
        flags |= GI_KD_REFERENCE_VISITS if reference_visits else 0
        flags |= _lib.GI_KD_EXACT_PRUNE if exact_prune else 0
        nn = d2 = cnt = None
        nn_ptr = d2_ptr = cnt_ptr = None
        if debug:
            shp = (r1 - r0, len_x, k) if rowmajor else (k, r1 - r0, len_x)
            nn, nn_ptr, _ = c.out(shp, np.int32, "C" if rowmajor else "F")
            d2, d2_ptr, _ = c.out(shp, rt, "C" if rowmajor else "F")
            cnt, cnt_ptr, _ = c.out((4,), np.int64, "C")
        st = getattr(lib(), "gi_idw_kdtree_ex_" + suf)(p.ptr, n, d.ptr, x.ptr, len_x, y.ptr, len_y, k, r0, r1, res_ptr, nn_ptr,
                                        d2_ptr, cnt_ptr, flags, c.mem, c.stream)
        c.finish(st, "idw_kdtree", sync)
        if not debug:
            return res
        if not rowmajor:  # present (rows, len_x, k) regardless of memory order
            nn, d2 = _moveaxis(nn, 0, -1), _moveaxis(d2, 0, -1)
        return res, SimpleNamespace(nn_index=nn, nn_dist2=d2, visits=int(cnt[0]))


def idw_kdtree(points, data_in, x_axis, y_axis, num_neighbours, **kw):
    """Scattered points -> regular grid, IDW over the k nearest neighbours found with a k-d tree
    (interpolation.f90:107-194).  points has shape (2, num_points); returns data_ip (len_y, len_x)."""
    return idw_kdtree_ex(points, data_in, x_axis, y_axis, num_neighbours, **kw)


def idw_esrg_nearest_ex(points, data_in, x_axis, y_axis, grid_spac, num_neighbours, *, rows=None, order="F",
                        debug=False, reference_order=False, dtype=np.float64, out=None, sync=True):
    """idw_esrg_nearest on a row band; debug=True also returns neighbour ids / d2 / counters."""
    rt, suf = _real(dtype)
    with _Call(points) as c:
        x, y, p, d, n = _grid_common(c, points, data_in, x_axis, y_axis, "n2", rt)
        len_x, len_y = x.shape[0], y.shape[0]
        r0, r1 = _band(rows, len_y)
        k = int(num_neighbours)
        res, res_ptr, rowmajor = _use_out(c, out, (r1 - r0, len_x), rt, order)
        flags = (GI_FIELD_ROWMAJOR if rowmajor else 0) | (GI_POINTS_ROWMAJOR if p.rowmajor else 0)
        flags |= _lib.GI_ESRG_REFERENCE_ORDER if reference_order else 0
        nn = d2 = cnt = None
        nn_ptr = d2_ptr = cnt_ptr = None
        if debug:
            shp = (r1 - r0, len_x, k) if rowmajor else (k, r1 - r0, len_x)
            nn, nn_ptr, _ = c.out(shp, np.int32, "C" if rowmajor else "F")
            d2, d2_ptr, _ = c.out(shp, rt, "C" if rowmajor else "F")
            cnt, cnt_ptr, _ = c.out((4,), np.int64, "C")
        st = getattr(lib(), "gi_idw_esrg_nearest_ex_" + suf)(p.ptr, n, d.ptr, x.ptr, len_x, y.ptr, len_y, float(grid_spac), k,
                                              r0, r1, res_ptr, nn_ptr, d2_ptr, cnt_ptr, flags, c.mem, c.stream)
        c.finish(st, "idw_esrg_nearest", sync)
        if not debug:
            return res
        if not rowmajor:
            nn, d2 = _moveaxis(nn, 0, -1), _moveaxis(d2, 0, -1)
        return res, SimpleNamespace(nn_index=nn, nn_dist2=d2, cells_scanned=int(cnt[0]), level_max=int(cnt[1]),
                                    list_overflows=int(cnt[2]), exact_order_targets=int(cnt[3]))


def idw_esrg_nearest(points, data_in, x_axis, y_axis, grid_spac, num_neighbours, **kw):
    """Scattered points -> equally spaced regular grid, IDW over the k nearest neighbours found with the
    cell-list ring search (interpolation.f90:201-288).  points (num_points, 2) -> data_ip (len_y, len_x)."""
    return idw_esrg_nearest_ex(points, data_in, x_axis, y_axis, grid_spac, num_neighbours, **kw)


def barycentric_interpolation_ex(points, data_in, x_axis, y_axis, simplices, neighbours, *, rows=None, halo=0,
                                 order="F", debug=False, reference_order=False, band_pack=False, plane_values=False,
                                 dtype=np.float64, out=None, sync=True):
    """barycentric_interpolation on a row band (+ ``halo`` rows for the hull fill); debug=True also returns
    triangle ids (1-based), the outside-hull mask, fill sources and counters.  reference_order=True locates every
    target by the reference's own row-sequential walk (debug: slow; the default reproduces its result).
    band_pack=True (device tensors; host calls decide by themselves) packs only the triangles near the band
    (GI_BARY_BAND_PACK): a GiError with status GI_ERR_PACK_WINDOW then asks for a retry without it.
    plane_values=True (GI_BARY_PLANE_VALUES, opt-in) evaluates each value as the plane through the triangle's nodal
    values with a per-triangle gradient: same triangles, mask and fill sources, values within 1e-12 * max|data_in|
    of the reference's (double precision) instead of bit-identical, the locate / evaluate kernel ~25 % faster."""
    rt, suf = _real(dtype)
    with _Call(points) as c:
        x, y, p, d, n = _grid_common(c, points, data_in, x_axis, y_axis, "n2", rt)
        s = c.arg(simplices, np.dtype(np.int32), 2, "simplices")
        nb = c.arg(neighbours, np.dtype(np.int32), 2, "neighbours")
        t = s.shape[0]
        if s.shape != (t, 3):
            raise ValueError(f"simplices must have shape (num_triangles, 3), got {s.shape}")
        if nb.shape != (t, 3):
            raise ValueError(f"neighbours must have shape (num_triangles, 3) = {(t, 3)}, got {nb.shape}")
        if s.rowmajor != nb.rowmajor:  # one flag covers both: bring neighbours to simplices' order
            nb = c.arg(_reorder(nb.keep, s.rowmajor), np.dtype(np.int32), 2, "neighbours")
        len_x, len_y = x.shape[0], y.shape[0]
        r0, r1 = _band(rows, len_y)
        res, res_ptr, rowmajor = _use_out(c, out, (r1 - r0, len_x), rt, order)
        flags = ((GI_FIELD_ROWMAJOR if rowmajor else 0) | (GI_POINTS_ROWMAJOR if p.rowmajor else 0)
                 | (GI_TOPO_ROWMAJOR if s.rowmajor else 0) | (_lib.GI_BARY_REFERENCE_ORDER if reference_order else 0)
                 | (_lib.GI_BARY_BAND_PACK if band_pack else 0)
                 | (_lib.GI_BARY_PLANE_VALUES if plane_values else 0))
        tri = mask = src = cnt = None
        tri_ptr = mask_ptr = src_ptr = cnt_ptr = None
        if debug:
            o = "C" if rowmajor else "F"
            tri, tri_ptr, _ = c.out((r1 - r0, len_x), np.int32, o)
            mask, mask_ptr, _ = c.out((r1 - r0, len_x), np.uint8, o)
            src, src_ptr, _ = c.out((r1 - r0, len_x), np.int64, o)
            cnt, cnt_ptr, _ = c.out((4,), np.int64, "C")
        st = getattr(lib(), "gi_barycentric_interpolation_ex_" + suf)(p.ptr, n, d.ptr, x.ptr, len_x, y.ptr, len_y, s.ptr, nb.ptr,
                                                       t, r0, r1, int(halo), res_ptr, tri_ptr, mask_ptr, src_ptr,
                                                       cnt_ptr, flags, c.mem, c.stream)
        c.finish(st, "barycentric_interpolation", sync)
        if not debug:
            return res
        return res, SimpleNamespace(tri=tri, mask=mask != 0, fill_src=src, iters_tot=int(cnt[0]),
                                    masked=int(cnt[1]), fill_level_max=int(cnt[2]), fixups=int(cnt[3]))


def barycentric_interpolation(points, data_in, x_axis, y_axis, simplices, neighbours, **kw):
    """Triangle mesh -> regular grid: visibility walk + barycentric weights + hull fill
    (interpolation.f90:298-416).  simplices / neighbours (num_triangles, 3) are 1-based, neighbour 0 = none
    (test_interpolation.py:199-201).  Returns data_ip (len_y, len_x)."""
    return barycentric_interpolation_ex(points, data_in, x_axis, y_axis, simplices, neighbours, **kw)


# -------------------------------------------------------------------------------------------------
# scattered targets: the reference interpolates to grid nodes only
# -------------------------------------------------------------------------------------------------
def barycentric_points(points, data_in, simplices, neighbours, targets, *, debug=False, plane_values=False,
                       dtype=np.float64, sync=True):
    """barycentric_interpolation (interpolation.f90:298-416) at scattered ``targets`` (num_targets, 2) instead of
    grid nodes: same visibility walk, same weights.  A target outside the convex hull takes the value at the closest
    location on the mesh (interpolation.f90:292-293), found by following the hull.  Returns data_ip (num_targets,);
    debug=True also returns the 1-based triangle the value was evaluated in and the outside-hull flags."""
    rt, suf = _real(dtype)
    with _Call(points) as c:
        p = c.arg(points, rt, 2, "points"); d = c.arg(data_in, rt, 1, "data_in")
        s = c.arg(simplices, np.dtype(np.int32), 2, "simplices")
        nb = c.arg(neighbours, np.dtype(np.int32), 2, "neighbours")
        tg = c.arg(targets, rt, 2, "targets")
        if p.shape[1] != 2:
            raise ValueError(f"points must have shape (num_points, 2), got {p.shape}")
        n, t = p.shape[0], s.shape[0]
        if d.shape != (n,):
            raise ValueError(f"data_in must have shape (num_points,) = {(n,)}, got {d.shape}")
        if s.shape != (t, 3) or nb.shape != (t, 3):
            raise ValueError("simplices / neighbours must have shape (num_triangles, 3)")
        if tg.shape[1] != 2:
            raise ValueError(f"targets must have shape (num_targets, 2), got {tg.shape}")
        if s.rowmajor != nb.rowmajor:
            nb = c.arg(_reorder(nb.keep, s.rowmajor), np.dtype(np.int32), 2, "neighbours")
        m = tg.shape[0]
        res, res_ptr, _ = c.out((m,), rt, "C")
        tri = outside = None
        tri_ptr = out_ptr = None
        if debug:
            tri, tri_ptr, _ = c.out((m,), np.int32, "C")
            outside, out_ptr, _ = c.out((m,), np.uint8, "C")
        flags = ((GI_POINTS_ROWMAJOR if p.rowmajor else 0) | (GI_TOPO_ROWMAJOR if s.rowmajor else 0)
                 | (_lib.GI_TARGETS_ROWMAJOR if tg.rowmajor else 0) | (_lib.GI_BARY_PLANE_VALUES if plane_values else 0))
        st = getattr(lib(), "gi_barycentric_points_" + suf)(p.ptr, n, d.ptr, s.ptr, nb.ptr, t, tg.ptr, m, res_ptr, tri_ptr,
                                                           out_ptr, flags, c.mem, c.stream)
        c.finish(st, "barycentric_points", sync)
        if not debug:
            return res
        return res, SimpleNamespace(tri=tri, outside=outside != 0)


def idw_kdtree_points(points, data_in, targets, num_neighbours, *, debug=False, exact_prune=False, dtype=np.float64,
                      sync=True):
    """idw_kdtree (interpolation.f90:107-194) at scattered ``targets`` (num_targets, 2) instead of grid nodes: same
    tree, same traversal and prune rule, same sums.  points (2, num_points) as in idw_kdtree.  debug=True also
    returns neighbour ids (1-based, (num_targets, k)) and squared distances."""
    rt, suf = _real(dtype)
    with _Call(points) as c:
        p = c.arg(points, rt, 2, "points"); d = c.arg(data_in, rt, 1, "data_in")
        tg = c.arg(targets, rt, 2, "targets")
        if p.shape[0] != 2:
            raise ValueError(f"points must have shape (2, num_points), got {p.shape}")
        n = p.shape[1]
        if d.shape != (n,):
            raise ValueError(f"data_in must have shape (num_points,) = {(n,)}, got {d.shape}")
        if tg.shape[1] != 2:
            raise ValueError(f"targets must have shape (num_targets, 2), got {tg.shape}")
        k = int(num_neighbours)
        if k < 1:
            raise ValueError("num_neighbours must be at least 1")
        m = tg.shape[0]
        res, res_ptr, _ = c.out((m,), rt, "C")
        nn = d2 = None
        nn_ptr = d2_ptr = None
        if debug:
            nn, nn_ptr, _ = c.out((m, k), np.int32, "C")
            d2, d2_ptr, _ = c.out((m, k), rt, "C")
        flags = ((GI_POINTS_ROWMAJOR if p.rowmajor else 0) | (_lib.GI_TARGETS_ROWMAJOR if tg.rowmajor else 0)
                 | (_lib.GI_KD_EXACT_PRUNE if exact_prune else 0))
        st = getattr(lib(), "gi_idw_kdtree_points_" + suf)(p.ptr, n, d.ptr, tg.ptr, m, k, res_ptr, nn_ptr, d2_ptr, flags,
                                                          c.mem, c.stream)
        c.finish(st, "idw_kdtree_points", sync)
        if not debug:
            return res
        return res, SimpleNamespace(nn_index=nn, nn_dist2=d2)


def idw_esrg_nearest_points(points, data_in, x_axis, y_axis, grid_spac, targets, num_neighbours, *, debug=False,
                            dtype=np.float64, sync=True):
    """idw_esrg_nearest (interpolation.f90:201-288) at scattered ``targets`` (num_targets, 2): the grid only defines
    the cell list; the ring search runs around each target's own cell with radius grid_spac * L per level (the exact
    k nearest points, ascending).  debug=True also returns neighbour ids (1-based) and squared distances."""
    rt, suf = _real(dtype)
    with _Call(points) as c:
        x, y, p, d, n = _grid_common(c, points, data_in, x_axis, y_axis, "n2", rt)
        tg = c.arg(targets, rt, 2, "targets")
        if tg.shape[1] != 2:
            raise ValueError(f"targets must have shape (num_targets, 2), got {tg.shape}")
        k = int(num_neighbours)
        if k < 1:
            raise ValueError("num_neighbours must be at least 1")
        m = tg.shape[0]
        res, res_ptr, _ = c.out((m,), rt, "C")
        nn = d2 = None
        nn_ptr = d2_ptr = None
        if debug:
            nn, nn_ptr, _ = c.out((m, k), np.int32, "C")
            d2, d2_ptr, _ = c.out((m, k), rt, "C")
        flags = (GI_POINTS_ROWMAJOR if p.rowmajor else 0) | (_lib.GI_TARGETS_ROWMAJOR if tg.rowmajor else 0)
        st = getattr(lib(), "gi_idw_esrg_nearest_points_" + suf)(p.ptr, n, d.ptr, x.ptr, x.shape[0], y.ptr, y.shape[0],
                                                                float(grid_spac), tg.ptr, m, k, res_ptr, nn_ptr, d2_ptr,
                                                                flags, c.mem, c.stream)
        c.finish(st, "idw_esrg_nearest_points", sync)
        if not debug:
            return res
        return res, SimpleNamespace(nn_index=nn, nn_dist2=d2)


# -------------------------------------------------------------------------------------------------
# plans: geometry + target grid prepared once on the device, one call per field of nodal values
# -------------------------------------------------------------------------------------------------
class _Plan:
    """Base of the three plan classes (gi_plan of gridinterp.h).  ``plan(data_in)`` returns data_ip with the
    shape / memory order fixed at creation; numpy in -> numpy out (pinned), torch CUDA tensor in -> torch out on
    the current stream.  Results are identical to the one-shot call on the same inputs."""

    def __init__(self):
        self._handle = C.c_void_p()
        self._suf = "f64"
        self._rt = _F64
        self._n = 0
        self._shape = (0, 0)
        self._order = "F"

    def _create(self, name, call, args):
        _require_device()
        st = getattr(lib(), name + "_" + self._suf)(*args, call.mem, call.stream, C.byref(self._handle))
        check(st, name)

    def __call__(self, data_in, *, out=None, sync=True):
        if not self._handle:
            raise RuntimeError("plan has been closed")
        with _Call(data_in) as c:
            d = c.arg(data_in, self._rt, 1, "data_in")
            if d.shape != (self._n,):
                raise ValueError(f"data_in must have shape (num_points,) = ({self._n},), got {d.shape}")
            res, res_ptr, rowmajor = _use_out(c, out, self._shape, self._rt, self._order)
            if rowmajor != (self._order == "C") and min(self._shape) > 1:
                raise ValueError(f"out must be {self._order}-contiguous (the plan's layout)")
            st = getattr(lib(), "gi_plan_execute_" + self._suf)(self._handle, d.ptr, res_ptr, c.mem, c.stream)
            c.finish(st, type(self).__name__, sync)
            return res

    def close(self):
        if self._handle:
            lib().gi_plan_destroy(self._handle)
            self._handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _plan_geometry(plan, c, points, x_axis, y_axis, points_shape, rows, order, dtype):
    plan._rt, plan._suf = _real(dtype)
    x = c.arg(x_axis, plan._rt, 1, "x_axis"); y = c.arg(y_axis, plan._rt, 1, "y_axis")
    p = c.arg(points, plan._rt, 2, "points")
    if points_shape == "n2":
        if p.shape[1] != 2:
            raise ValueError(f"points must have shape (num_points, 2), got {p.shape}")
        plan._n = p.shape[0]
    else:
        if p.shape[0] != 2:
            raise ValueError(f"points must have shape (2, num_points), got {p.shape}")
        plan._n = p.shape[1]
    if order not in ("C", "F"):
        raise ValueError("order must be 'C' or 'F'")
    r0, r1 = _band(rows, y.shape[0])
    plan._shape, plan._order = (r1 - r0, x.shape[0]), order
    flags = (GI_FIELD_ROWMAJOR if order == "C" else 0) | (GI_POINTS_ROWMAJOR if p.rowmajor else 0)
    return x, y, p, r0, r1, flags


class BarycentricPlan(_Plan):
    """barycentric_interpolation with the mesh packed once: ``plan = BarycentricPlan(points, x_axis, y_axis,
    simplices, neighbours); field = plan(data_in)``."""

    def __init__(self, points, x_axis, y_axis, simplices, neighbours, *, rows=None, halo=0, order="F",
                 plane_values=False, dtype=np.float64):
        super().__init__()
        with _Call(points) as c:
            x, y, p, r0, r1, flags = _plan_geometry(self, c, points, x_axis, y_axis, "n2", rows, order, dtype)
            s = c.arg(simplices, np.dtype(np.int32), 2, "simplices")
            nb = c.arg(neighbours, np.dtype(np.int32), 2, "neighbours")
            t = s.shape[0]
            if s.shape != (t, 3) or nb.shape != (t, 3):
                raise ValueError("simplices and neighbours must both have shape (num_triangles, 3)")
            if s.rowmajor != nb.rowmajor:
                nb = c.arg(_reorder(nb.keep, s.rowmajor), np.dtype(np.int32), 2, "neighbours")
            flags |= GI_TOPO_ROWMAJOR if s.rowmajor else 0
            flags |= _lib.GI_BARY_PLANE_VALUES if plane_values else 0
            self._create("gi_barycentric_plan_create", c,
                         (p.ptr, self._n, x.ptr, x.shape[0], y.ptr, y.shape[0], s.ptr, nb.ptr, t, r0, r1, int(halo), flags))


class KdTreePlan(_Plan):
    """idw_kdtree with the tree built once: ``plan = KdTreePlan(points, x_axis, y_axis, num_neighbours)``
    (points (2, num_points) as in idw_kdtree)."""

    def __init__(self, points, x_axis, y_axis, num_neighbours, *, rows=None, order="F", exact_prune=False,
                 dtype=np.float64):
        super().__init__()
        with _Call(points) as c:
            x, y, p, r0, r1, flags = _plan_geometry(self, c, points, x_axis, y_axis, "2n", rows, order, dtype)
            flags |= _lib.GI_KD_EXACT_PRUNE if exact_prune else 0
            self._create("gi_idw_kdtree_plan_create", c,
                         (p.ptr, self._n, x.ptr, x.shape[0], y.ptr, y.shape[0], int(num_neighbours), r0, r1, flags))


class EsrgPlan(_Plan):
    """idw_esrg_nearest with the points binned once: ``plan = EsrgPlan(points, x_axis, y_axis, grid_spac,
    num_neighbours)``."""

    def __init__(self, points, x_axis, y_axis, grid_spac, num_neighbours, *, rows=None, order="F", dtype=np.float64):
        super().__init__()
        with _Call(points) as c:
            x, y, p, r0, r1, flags = _plan_geometry(self, c, points, x_axis, y_axis, "n2", rows, order, dtype)
            self._create("gi_idw_esrg_plan_create", c,
                         (p.ptr, self._n, x.ptr, x.shape[0], y.ptr, y.shape[0], float(grid_spac), int(num_neighbours),
                          r0, r1, flags))


# -------------------------------------------------------------------------------------------------
# parity / debug exports of the search structures
# -------------------------------------------------------------------------------------------------
def kd_build(points, faithful=False):
    """Implicit k-d tree (kd_tree.f90:30-95): 1-based point ids in tree order (see gridinterp.h).  faithful=True
    forces the build that evaluates the reference's quickselect passes (chosen automatically when coordinates tie)."""
    with _Call(points) as c:
        p = c.arg(points, _F64, 2, "points")
        if p.shape[0] != 2:
            raise ValueError("points must have shape (2, num_points)")
        n = p.shape[1]
        order, ptr, _ = c.out((n,), np.int32, "C")
        flags = (GI_POINTS_ROWMAJOR if p.rowmajor else 0) | (_lib.GI_KD_FAITHFUL_BUILD if faithful else 0)
        st = lib().gi_kd_build_f64(p.ptr, n, ptr, flags, c.mem, c.stream)
        c.finish(st, "kd_build")
        return order


def assign_points_to_cells(points, x_axis, y_axis, grid_spac):
    """query_esrg.f90:23-91 -> (index_of_pts 1-based, indptr 0-based offsets of len_y*len_x+1 cells)."""
    with _Call(points) as c:
        p = c.arg(points, _F64, 2, "points")
        x = c.arg(x_axis, _F64, 1, "x_axis"); y = c.arg(y_axis, _F64, 1, "y_axis")
        n = p.shape[0]
        idx, idx_ptr, _ = c.out((n,), np.int32, "C")
        ptr, ptr_ptr, _ = c.out((x.shape[0] * y.shape[0] + 1,), np.int32, "C")
        st = lib().gi_esrg_assign_points_to_cells_f64(p.ptr, n, x.ptr, x.shape[0], y.ptr, y.shape[0],
                                                      float(grid_spac), idx_ptr, ptr_ptr,
                                                      GI_POINTS_ROWMAJOR if p.rowmajor else 0, c.mem, c.stream)
        c.finish(st, "assign_points_to_cells")
        return idx, ptr


def fill_nearest_neighbour(mask_outside, data_ip):
    """triangle_walk.f90:90-132 on copies: returns (filled field, fill_src)."""
    mask = np.ascontiguousarray(np.asarray(mask_outside) != 0, dtype=np.uint8)
    data = np.array(data_ip, dtype=np.float64, order="C", copy=True)
    if mask.shape != data.shape or mask.ndim != 2:
        raise ValueError("mask_outside and data_ip must be 2-D arrays of the same shape")
    ly, lx = mask.shape
    src = np.empty((ly, lx), np.int64)
    st = lib().gi_fill_nearest_neighbour_f64(mask.ctypes.data, lx, ly, data.ctypes.data, src.ctypes.data,
                                             GI_FIELD_ROWMAJOR, GI_HOST, None)
    check(st, "fill_nearest_neighbour")
    return data, src


def set_profiling(on: bool) -> None:
    lib().gi_set_profiling(1 if on else 0)


def set_fixup_list_capacity(targets: int) -> None:
    """Capacity of the barycentric ambiguity fix-up work list (0 = default 4 Mi targets); see gridinterp.h."""
    lib().gi_set_fixup_list_capacity(int(targets))


def set_sliver_guard(min_quality: float) -> None:
    """Sliver guard of barycentric_interpolation (the reference's README to-do): triangles with a hull edge whose
    quality |2 area| / max |edge|^2 is below ``min_quality`` are not interpolated in; their targets are masked and
    filled from the nearest neighbours.  0 = off (the reference's behaviour); see gridinterp.h."""
    lib().gi_set_sliver_guard(float(min_quality))


def release_workspace() -> None:
    """Return the library's cached device scratch memory (stream-ordered pool) to the driver."""
    lib().gi_release_workspace()


def pinned_trim(keep_bytes: int = 0) -> None:
    """Free cached page-locked host buffers until at most ``keep_bytes`` stay cached."""
    _pool.trim(keep_bytes)


def set_pipeline_band_bytes(nbytes: int) -> None:
    """Granularity of the copy / compute pipeline of host (numpy) calls: bytes of output per band (0 = default
    32 MB).  Results do not depend on it."""
    lib().gi_set_pipeline_band_bytes(int(nbytes))


def last_timings(capacity=8192):
    """[(phase name, device milliseconds)] of every call made by this thread since set_profiling(True)."""
    buf = (C.c_double * capacity)()
    n = lib().gi_last_timings(buf, capacity)
    res = [(lib().gi_last_timing_name(i).decode(), buf[i]) for i in range(min(n, capacity))]
    return [(name, ms) for name, ms in res if name]


def _moveaxis(a, src, dst):
    if _is_torch(a):
        return a.movedim(src, dst)
    return np.moveaxis(a, src, dst)


def _reorder(a, rowmajor):
    if _is_torch(a):
        return a.contiguous() if rowmajor else a.t().contiguous().t()
    return np.ascontiguousarray(a) if rowmajor else np.asfortranarray(a)

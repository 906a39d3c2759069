"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).  TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``grid_interpolation_b200``) never does.

The four public callables mirror the f2py module of the reference
(``from interpolation import interpolation as ip_fortran``; names, positional
order and keyword names from interpolation.f90:20-21, :107-108, :201-202,
:298-300 with the ``intent(hide)`` dimensions removed).  Like f2py they accept
any array-like, convert to the build precision / Fortran order, and return
Fortran-ordered outputs.  ``debug=True`` additionally returns the internal
indices that the reference never exposes (neighbour ids, triangle ids, ...).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ORC_OK, ORC_BAD_ARG, ORC_ITER_CAP, ORC_TOO_FEW_POINTS = 0, 1, 5, 6


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ only, seconds)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def set_sliver_guard(min_quality: float) -> None:
    """The library's opt-in sliver guard (NOT in the reference: README.md:41-43 lists it as a to-do), see oracle.cpp."""
    lib().orc_set_sliver_guard(C.c_double(float(min_quality)))


def get_max_threads() -> int:
    return int(lib().orc_get_max_threads())


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64", C.c_double
    if dtype == np.float32:
        return "f32", C.c_float
    raise ValueError("dtype must be float64 or float32")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a, dtype):
    """Fortran-contiguous array of the given dtype (what f2py would pass)."""
    return np.asfortranarray(np.asarray(a), dtype=dtype)


def _check(status, what):
    if status == ORC_OK:
        return
    msg = {ORC_BAD_ARG: "bad argument", ORC_ITER_CAP: "iteration cap hit (reference would hang)",
           ORC_TOO_FEW_POINTS: "fewer points than num_neighbours (reference would hang)"}.get(status, "error")
    raise RuntimeError(f"oracle.{what}: {msg} (status {status})")


# ---------------------------------------------------------------------------
def bilinear(x_axis, y_axis, data_in, points, dtype=np.float64):
    """interpolation.f90:20-100.  data_in (len_y, len_x), points (N, 2) -> data_ip (N,)."""
    suf, _ = _suffix(dtype)
    x_axis = _f(x_axis, dtype); y_axis = _f(y_axis, dtype)
    data_in = _f(data_in, dtype); points = _f(points, dtype)
    if data_in.shape != (y_axis.size, x_axis.size):
        raise ValueError("data_in must have shape (len_y, len_x)")
    if points.ndim != 2 or points.shape[1] != 2:
        raise ValueError("points must have shape (num_points, 2)")
    n = points.shape[0]
    out = np.empty(n, dtype=dtype)
    fn = getattr(lib(), "orc_bilinear_" + suf)
    st = fn(_p(x_axis), C.c_int(x_axis.size), _p(y_axis), C.c_int(y_axis.size), _p(data_in), _p(points),
            C.c_int64(n), _p(out))
    _check(st, "bilinear")
    return out


def idw_kdtree(points, data_in, x_axis, y_axis, num_neighbours, dtype=np.float64, debug=False):
    """interpolation.f90:107-194.  points (2, N) -> data_ip (len_y, len_x) [Fortran order]."""
    suf, creal = _suffix(dtype)
    points = _f(points, dtype); data_in = _f(data_in, dtype)
    x_axis = _f(x_axis, dtype); y_axis = _f(y_axis, dtype)
    if points.ndim != 2 or points.shape[0] != 2:
        raise ValueError("points must have shape (2, num_points)")
    n = points.shape[1]
    if data_in.shape != (n,):
        raise ValueError("data_in must have shape (num_points,)")
    k = int(num_neighbours)
    ly, lx = y_axis.size, x_axis.size
    out = np.empty((ly, lx), dtype=dtype, order="F")
    nn_index = np.empty((k, ly, lx), dtype=np.int32, order="F") if debug else None
    nn_d2 = np.empty((k, ly, lx), dtype=dtype, order="F") if debug else None
    visits = C.c_int64(0); tb = C.c_double(0); tq = C.c_double(0)
    fn = getattr(lib(), "orc_idw_kdtree_" + suf)
    st = fn(_p(points), C.c_int(n), _p(data_in), _p(x_axis), C.c_int(lx), _p(y_axis), C.c_int(ly), C.c_int(k),
            _p(out), _p(nn_index), _p(nn_d2), C.byref(visits) if debug else None, C.byref(tb), C.byref(tq))
    _check(st, "idw_kdtree")
    if debug:
        return out, SimpleNamespace(nn_index=nn_index, nn_dist2=nn_d2, visits=visits.value,
                                    t_build=tb.value, t_query=tq.value)
    return out


def kd_build(points, dtype=np.float64):
    """kd_tree.f90:30-155: returns the reference's pointer tree as arrays (pre-order node ids)."""
    suf, _ = _suffix(dtype)
    points = _f(points, dtype)
    n = points.shape[1]
    xy = np.empty((n, 2), dtype=dtype)
    idx = np.empty(n, np.int32); left = np.empty(n, np.int32); right = np.empty(n, np.int32)
    root = getattr(lib(), "orc_kd_build_" + suf)(_p(points), C.c_int(n), _p(xy), _p(idx), _p(left), _p(right))
    return SimpleNamespace(root=root, xy=xy, index=idx, left=left, right=right)


def assign_points_to_cells(points, x_axis, y_axis, grid_spac, dtype=np.float64):
    """query_esrg.f90:23-91 -> (index_of_pts[N] 1-based, indptr (len_y,len_x) 1-based, num_ppgc)."""
    suf, creal = _suffix(dtype)
    points = _f(points, dtype); x_axis = _f(x_axis, dtype); y_axis = _f(y_axis, dtype)
    n = points.shape[0]
    ly, lx = y_axis.size, x_axis.size
    index_of_pts = np.empty(n, np.int32)
    indptr = np.empty((ly, lx), np.int32, order="F"); num_ppgc = np.empty((ly, lx), np.int32, order="F")
    getattr(lib(), "orc_assign_points_to_cells_" + suf)(
        _p(points), C.c_int(n), _p(x_axis), C.c_int(lx), _p(y_axis), C.c_int(ly), creal(grid_spac),
        _p(index_of_pts), _p(indptr), _p(num_ppgc))
    return index_of_pts, indptr, num_ppgc


def idw_esrg_nearest(points, data_in, x_axis, y_axis, grid_spac, num_neighbours, dtype=np.float64, debug=False):
    """interpolation.f90:201-288.  points (N, 2) -> data_ip (len_y, len_x) [Fortran order]."""
    suf, creal = _suffix(dtype)
    points = _f(points, dtype); data_in = _f(data_in, dtype)
    x_axis = _f(x_axis, dtype); y_axis = _f(y_axis, dtype)
    if points.ndim != 2 or points.shape[1] != 2:
        raise ValueError("points must have shape (num_points, 2)")
    n = points.shape[0]
    if data_in.shape != (n,):
        raise ValueError("data_in must have shape (num_points,)")
    k = int(num_neighbours)
    ly, lx = y_axis.size, x_axis.size
    out = np.empty((ly, lx), dtype=dtype, order="F")
    nn_index = np.empty((k, ly, lx), dtype=np.int32, order="F") if debug else None
    nn_dist = np.empty((k, ly, lx), dtype=dtype, order="F") if debug else None
    lvl = C.c_int(0); lst = C.c_int(0); tb = C.c_double(0); tq = C.c_double(0)
    fn = getattr(lib(), "orc_idw_esrg_nearest_" + suf)
    st = fn(_p(points), C.c_int(n), _p(data_in), _p(x_axis), C.c_int(lx), _p(y_axis), C.c_int(ly),
            creal(grid_spac), C.c_int(k), _p(out), _p(nn_index), _p(nn_dist), C.byref(lvl), C.byref(lst),
            C.byref(tb), C.byref(tq))
    _check(st, "idw_esrg_nearest")
    if debug:
        return out, SimpleNamespace(nn_index=nn_index, nn_dist=nn_dist, level_max=lvl.value,
                                    list_max=lst.value, t_build=tb.value, t_query=tq.value)
    return out


def barycentric_interpolation(points, data_in, x_axis, y_axis, simplices, neighbours, dtype=np.float64,
                              debug=False):
    """interpolation.f90:298-416.  simplices / neighbours (T, 3) 1-based, 0 = no neighbour."""
    suf, _ = _suffix(dtype)
    points = _f(points, dtype); data_in = _f(data_in, dtype)
    x_axis = _f(x_axis, dtype); y_axis = _f(y_axis, dtype)
    simplices = _f(simplices, np.int32); neighbours = _f(neighbours, np.int32)
    if points.ndim != 2 or points.shape[1] != 2:
        raise ValueError("points must have shape (num_points, 2)")
    n = points.shape[0]
    if data_in.shape != (n,):
        raise ValueError("data_in must have shape (num_points,)")
    t = simplices.shape[0]
    if simplices.shape != (t, 3) or neighbours.shape != (t, 3):
        raise ValueError("simplices / neighbours must have shape (num_triangles, 3)")
    ly, lx = y_axis.size, x_axis.size
    out = np.empty((ly, lx), dtype=dtype, order="F")
    tri = np.empty((ly, lx), np.int32, order="F") if debug else None
    mask = np.empty((ly, lx), np.uint8, order="F") if debug else None
    src = np.empty((ly, lx), np.int64, order="F") if debug else None
    tri_start = C.c_int(0); iters = C.c_int64(0); lvl = C.c_int(0)
    ts = C.c_double(0); tw = C.c_double(0); tf = C.c_double(0)
    fn = getattr(lib(), "orc_barycentric_interpolation_" + suf)
    st = fn(_p(points), C.c_int(n), _p(data_in), _p(x_axis), C.c_int(lx), _p(y_axis), C.c_int(ly),
            _p(simplices), _p(neighbours), C.c_int(t), _p(out), _p(tri), _p(mask), _p(src),
            C.byref(tri_start), C.byref(iters), C.byref(lvl), C.byref(ts), C.byref(tw), C.byref(tf))
    _check(st, "barycentric_interpolation")
    if debug:
        return out, SimpleNamespace(tri=tri, mask=mask.astype(bool), fill_src=src, tri_start=tri_start.value,
                                    iters_tot=iters.value, fill_level_max=lvl.value, t_start=ts.value,
                                    t_walk=tw.value, t_fill=tf.value)
    return out


def barycentric_points(points, data_in, simplices, neighbours, targets, dtype=np.float64, debug=False):
    """Extension (not in the reference, which is grid-only): barycentric interpolation at scattered targets (m, 2);
    outside the hull the closest location on the mesh.  See oracle.cpp."""
    suf, _ = _suffix(dtype)
    points = _f(points, dtype); data_in = _f(data_in, dtype); targets = _f(targets, dtype)
    simplices = _f(simplices, np.int32); neighbours = _f(neighbours, np.int32)
    n, t, m = points.shape[0], simplices.shape[0], targets.shape[0]
    out = np.empty(m, dtype=dtype)
    tri = np.empty(m, np.int32) if debug else None
    outside = np.empty(m, np.uint8) if debug else None
    st = getattr(lib(), "orc_barycentric_points_" + suf)(_p(points), C.c_int(n), _p(data_in), _p(simplices), _p(neighbours),
                                                        C.c_int(t), _p(targets), C.c_int64(m), _p(out), _p(tri), _p(outside))
    _check(st, "barycentric_points")
    return (out, SimpleNamespace(tri=tri, outside=outside.astype(bool))) if debug else out


def idw_kdtree_points(points, data_in, targets, num_neighbours, dtype=np.float64, debug=False):
    """Extension (not in the reference): idw_kdtree at scattered targets (m, 2); points (2, n).  See oracle.cpp."""
    suf, _ = _suffix(dtype)
    points = _f(points, dtype); data_in = _f(data_in, dtype); targets = _f(targets, dtype)
    n, m, k = points.shape[1], targets.shape[0], int(num_neighbours)
    out = np.empty(m, dtype=dtype)
    nn = np.empty((m, k), np.int32) if debug else None
    d2 = np.empty((m, k), dtype) if debug else None
    st = getattr(lib(), "orc_idw_kdtree_points_" + suf)(_p(points), C.c_int(n), _p(data_in), _p(targets), C.c_int64(m),
                                                       C.c_int(k), _p(out), _p(nn), _p(d2))
    _check(st, "idw_kdtree_points")
    return (out, SimpleNamespace(nn_index=nn, nn_dist2=d2)) if debug else out


def find_triangle(points, simplices, neighbours, target, ind_tri, dtype=np.float64):
    """triangle_walk.f90:24-84 for a single target; ind_tri 1-based seed -> (ind_tri, inside, iters)."""
    suf, creal = _suffix(dtype)
    points = _f(points, dtype)
    simplices = _f(simplices, np.int32); neighbours = _f(neighbours, np.int32)
    tri = C.c_int(int(ind_tri)); inside = C.c_int(0); iters = C.c_int(0)
    st = getattr(lib(), "orc_find_triangle_" + suf)(
        _p(points), C.c_int(points.shape[0]), _p(simplices), _p(neighbours), C.c_int(simplices.shape[0]),
        creal(target[0]), creal(target[1]), C.byref(tri), C.byref(inside), C.byref(iters))
    _check(st, "find_triangle")
    return tri.value, bool(inside.value), iters.value


def fill_nearest_neighbour(mask_outside, data_ip, dtype=np.float64):
    """triangle_walk.f90:90-132 on copies -> (filled data, fill_src, level_max)."""
    suf, _ = _suffix(dtype)
    mask = _f(np.asarray(mask_outside).astype(np.uint8), np.uint8)
    data = np.array(data_ip, dtype=dtype, order="F", copy=True)
    ly, lx = mask.shape
    src = np.full((ly, lx), -1, np.int64, order="F")
    lvl = C.c_int(0)
    st = getattr(lib(), "orc_fill_nearest_neighbour_" + suf)(_p(mask), C.c_int(lx), C.c_int(ly), _p(data),
                                                             _p(src), C.byref(lvl))
    _check(st, "fill_nearest_neighbour")
    return data, src, lvl.value

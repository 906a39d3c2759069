"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN FORTRAN SOURCE.

There is no Fortran compiler in the image, so the four reference files cannot be built; oracle/f90interp.py is a
small interpreter for the Fortran 90 subset they are written in (it implements language semantics, nothing about
interpolation).  This script loads the unmodified text of

    /root/reference/kd_tree.f90  query_esrg.f90  triangle_walk.f90  interpolation.f90

and calls the module procedures the f2py binding exports -- bilinear, idw_kdtree, idw_esrg_nearest,
barycentric_interpolation -- plus the search primitives behind them (build_kd_tree + nearest_neighbours,
assign_points_to_cells + nearest_neighbours_esrg) on small seeded inputs, once with the default REAL kind set to
double (the graded build, -fdefault-real-8) and once with single precision (the reference as shipped).  The
outputs are what the reference computes for these inputs; tests/test_oracle_golden.py pins the C++ oracle against
them bit for bit and tests/test_gpu_parity.py the CUDA library.

Run in the build container only (needs /root/reference); takes about two minutes:

    python tests/golden/make_golden_f90.py        # rewrites tests/golden/f90_f64.npz, f90_f32.npz
    python tests/golden/make_golden_f90.py --c1   # the reference's own 5 000-point case (fields only, ~10 min)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import workloads  # noqa: E402  (host-side input generation only)
from oracle.f90interp import Cell, Interpreter  # noqa: E402

REF = os.environ.get("GI_REFERENCE", "/root/reference")
FILES = ("kd_tree.f90", "query_esrg.f90", "triangle_walk.f90", "interpolation.f90")


def farr(a, dtype):
    return np.asfortranarray(np.asarray(a, dtype=dtype))


def kd_case(itp, R, pts_2n, data, x, y, k):
    """idw_kdtree (interpolation.f90:107-194) + the neighbour lists behind it (kd_tree.f90:30-240)."""
    n = pts_2n.shape[1]
    field = np.zeros((y.size, x.size), dtype=R, order="F")
    itp.call("idw_kdtree", pts_2n, n, data, x, x.size, y, y.size, k, field)
    # the same query per target, through the module procedures, to expose the neighbour ids / squared distances
    tree = Cell(itp._new_obj("kdtree_type", {}))
    index = np.arange(1, n + 1, dtype=np.int64)
    itp.call("build_kd_tree", tree, pts_2n, n, 0, index)
    nn_index = np.zeros((y.size, x.size, k), np.int64)
    nn_dist2 = np.zeros((y.size, x.size, k), R)
    root = Cell(tree.get().f["root"])
    for iy in range(y.size):
        for ix in range(x.size):
            nb = np.full((3, k), R(1.0e6), dtype=R, order="F")       # interpolation.f90:159
            nbi = np.zeros(k, np.int64)                               # (never-filled slots stay 0 here)
            itp.call("nearest_neighbours", root, np.array([x[ix], y[iy]], dtype=R), 0, nb, k, nbi)
            nn_index[iy, ix] = nbi
            nn_dist2[iy, ix] = nb[2]
    return field, nn_index, nn_dist2


def esrg_case(itp, R, pts_n2, data, x, y, spac, k):
    """idw_esrg_nearest (interpolation.f90:201-288) + binning and per-target lists (query_esrg.f90:23-249)."""
    n = pts_n2.shape[0]
    field = np.zeros((y.size, x.size), dtype=R, order="F")
    itp.call("idw_esrg_nearest", pts_n2, n, data, x, x.size, y, y.size, R(spac), k, field)
    index_of_pts = np.zeros(n, np.int64)
    indptr = np.zeros((y.size, x.size), np.int64, order="F")
    num_ppgc = np.zeros((y.size, x.size), np.int64, order="F")
    itp.call("assign_points_to_cells", pts_n2, x, y, R(spac), index_of_pts, indptr, num_ppgc)
    nn_index = np.zeros((y.size, x.size, k), np.int64)
    nn_dist = np.zeros((y.size, x.size, k), R)
    for iy in range(y.size):
        for ix in range(x.size):
            inn = np.zeros(k, np.int64)
            dnn = np.zeros(k, R)
            itp.call("nearest_neighbours_esrg", pts_n2, index_of_pts, indptr, num_ppgc, k, x, y, R(spac), ix + 1, iy + 1,
                     inn, dnn)
            nn_index[iy, ix] = inn
            nn_dist[iy, ix] = dnn
    return field, index_of_pts, indptr, num_ppgc, nn_index, nn_dist


def main():
    src = [open(os.path.join(REF, f)).read() for f in FILES]
    for R, tag in ((np.float64, "f64"), (np.float32, "f32")):
        itp = Interpreter(src, R)
        out = {}
        w = workloads.c1(num_points=300)                     # test_interpolation.py:54-93 geometry, 300 points
        pts = farr(w.points, R); dat = w.data_pts.astype(R)
        x = w.x_axis.astype(R); y = w.y_axis.astype(R)
        simp = farr(w.simplices, np.int64); nbr = farr(w.neighbours, np.int64)
        out.update(points=pts, data=dat, x_axis=x, y_axis=y, simplices=w.simplices, neighbours=w.neighbours,
                   grid_spac=R(w.grid_spac))

        # ---- barycentric_interpolation (interpolation.f90:298-416 + triangle_walk.f90)
        bary = np.zeros((y.size, x.size), dtype=R, order="F")
        itp.call("barycentric_interpolation", pts, pts.shape[0], dat, x, x.size, y, y.size, simp, nbr, simp.shape[0], bary)
        out["bary_field"] = bary

        # ---- idw_esrg_nearest, k = 6
        f, iop, ptr, ppgc, nni, nnd = esrg_case(itp, R, pts, dat, x, y, w.grid_spac, 6)
        out.update(esrg_field=f, esrg_index_of_pts=iop, esrg_indptr=ptr, esrg_num_ppgc=ppgc, esrg_nn_index=nni,
                   esrg_nn_dist=nnd)

        # ---- idw_kdtree, k = 6: (a) as is, (b) coordinates scaled by 0.2 -> every squared distance < 1, the regime in
        # which the reference's prune rule abs(diff) < max d2 (kd_tree.f90:191,198) returns approximate sets
        pts_2n = farr(w.points.T, R)
        f, nni, nnd = kd_case(itp, R, pts_2n, dat, x, y, 6)
        out.update(kd_field=f, kd_nn_index=nni, kd_nn_dist2=nnd)
        s = R(0.2)
        pts_s = farr(w.points.T * 0.2, R); xs = (w.x_axis * 0.2).astype(R); ys = (w.y_axis * 0.2).astype(R)
        f, nni, nnd = kd_case(itp, R, pts_s, dat, xs, ys, 6)
        out.update(kds_points=pts_s, kds_x_axis=xs, kds_y_axis=ys, kds_field=f, kds_nn_index=nni, kds_nn_dist2=nnd)

        # ---- bilinear (interpolation.f90:20-100): the esrg field back to the points, and non-uniform axes
        inside = ((w.points[:, 0] >= w.x_axis[0]) & (w.points[:, 0] < w.x_axis[-1]) &
                  (w.points[:, 1] >= w.y_axis[0]) & (w.points[:, 1] < w.y_axis[-1]))
        bp = farr(w.points[inside], R)
        res = np.zeros(bp.shape[0], R)
        itp.call("bilinear", x, x.size, y, y.size, out["esrg_field"], bp, bp.shape[0], res)
        out.update(bil_points=bp, bil_values=res)
        rng = np.random.default_rng(12)
        xn = np.cumsum(rng.uniform(0.97, 1.03, 23)).astype(R); yn = np.cumsum(rng.uniform(0.97, 1.03, 17)).astype(R)
        fld = farr(rng.normal(size=(17, 23)), R)
        # (stay one cell away from the borders: the reference reads out of bounds there, interpolation.f90:68-78)
        bq = farr(np.column_stack([rng.uniform(xn[1], xn[-2], 200), rng.uniform(yn[1], yn[-2], 200)]), R)
        res2 = np.zeros(200, R)
        itp.call("bilinear", xn, xn.size, yn, yn.size, fld, bq, 200, res2)
        out.update(biln_x_axis=xn, biln_y_axis=yn, biln_field=fld, biln_points=bq, biln_values=res2)

        # ---- tie-heavy lattice: points on a spacing-2 lattice (shuffled), targets on the spacing-1 grid over the same
        # extent + one ring -> targets ON vertices and edges (walk margin / path dependence, triangle_walk.f90:148-151),
        # exact distance ties (kd_tree.f90:222-231 stable, query_esrg.f90:234-240 new-before-equal), zero distances
        # (nearest-copy branches interpolation.f90:164-169,260-265), equal split keys (kd_tree.f90:129); then the
        # same lattice with 20 exact duplicates for the two IDW paths
        from scipy.spatial import Delaunay
        gx, gy = np.meshgrid(np.arange(0, 21, 2.0), np.arange(0, 15, 2.0))
        lp = np.column_stack([gx.ravel(), gy.ravel()])
        rng = np.random.default_rng(7)
        lp = lp[rng.permutation(len(lp))]
        ld = rng.normal(size=len(lp)).astype(R)
        lx = np.arange(-1.0, 22.0, 1.0).astype(R); ly = np.arange(-1.0, 16.0, 1.0).astype(R)
        tri = Delaunay(lp)
        ls = tri.simplices.astype(np.int64) + 1; ln = tri.neighbors.astype(np.int64) + 1
        f = np.zeros((ly.size, lx.size), dtype=R, order="F")
        itp.call("barycentric_interpolation", farr(lp, R), len(lp), ld, lx, lx.size, ly, ly.size, farr(ls, np.int64),
                 farr(ln, np.int64), len(ls), f)
        out.update(lat_points=farr(lp, R), lat_data=ld, lat_x_axis=lx, lat_y_axis=ly, lat_simplices=ls.astype(np.int32),
                   lat_neighbours=ln.astype(np.int32), lat_bary_field=f)
        dp = np.vstack([lp, lp[rng.integers(0, len(lp), 20)]])
        dd = rng.normal(size=len(dp)).astype(R)
        out.update(dup_points=farr(dp, R), dup_data=dd)
        for name, P, D in (("lat", lp, ld), ("dup", dp, dd)):
            for k in (1, 4, 6):
                f = np.zeros((ly.size, lx.size), dtype=R, order="F")
                itp.call("idw_kdtree", farr(P.T, R), len(P), D, lx, lx.size, ly, ly.size, k, f)
                out[f"{name}_kd_field_k{k}"] = f
                f = np.zeros((ly.size, lx.size), dtype=R, order="F")
                itp.call("idw_esrg_nearest", farr(P, R), len(P), D, lx, lx.size, ly, ly.size, R(1.0), k, f)
                out[f"{name}_esrg_field_k{k}"] = f

        # ---- a small grid among far-away points (sparse stations around a local grid): most points are clamped into
        # the border cells (query_esrg.f90:44-57) and the ring search keeps growing its radius after its frames have
        # left the grid (:178-181) until it reaches them
        rng = np.random.default_rng(48)
        fx = (-4.75 + np.arange(5)).astype(R); fy = (2.5 + np.arange(7)).astype(R)
        fp = np.column_stack([rng.uniform(-17.0, 11.0, 14), rng.uniform(-15.0, 27.0, 14)]).astype(R)
        fd = rng.normal(size=14).astype(R)
        out.update(far_points=farr(fp, R), far_data=fd, far_x_axis=fx, far_y_axis=fy)
        for k in (1, 7, 13):
            f = np.zeros((fy.size, fx.size), dtype=R, order="F")
            itp.call("idw_esrg_nearest", farr(fp, R), 14, fd, fx, fx.size, fy, fy.size, R(1.0), k, f)
            out[f"far_esrg_field_k{k}"] = f
            f = np.zeros((fy.size, fx.size), dtype=R, order="F")
            itp.call("idw_kdtree", farr(fp.T, R), 14, fd, fx, fx.size, fy, fy.size, k, f)
            out[f"far_kd_field_k{k}"] = f

        path = os.path.join(HERE, f"f90_{tag}.npz")
        np.savez_compressed(path, **out)
        print(f"{path}: {len(out)} arrays, {itp.calls} procedure calls interpreted")


def c1_full():
    """The reference's own test case at its own size (test_interpolation.py:54-93 with 5 000 points -> 132 x 79 grid,
    num_nn = 6): the four fields only, double precision.  About ten minutes of interpretation."""
    src = [open(os.path.join(REF, f)).read() for f in FILES]
    R = np.float64
    itp = Interpreter(src, R)
    w = workloads.c1()
    pts = farr(w.points, R); dat = w.data_pts.astype(R); x = w.x_axis.astype(R); y = w.y_axis.astype(R)
    simp = farr(w.simplices, np.int64); nbr = farr(w.neighbours, np.int64)
    out = {}
    f = np.zeros((y.size, x.size), dtype=R, order="F")
    itp.call("barycentric_interpolation", pts, pts.shape[0], dat, x, x.size, y, y.size, simp, nbr, simp.shape[0], f)
    out["bary_field"] = f.copy()
    print("barycentric done", itp.calls, flush=True)
    f = np.zeros((y.size, x.size), dtype=R, order="F")
    itp.call("idw_esrg_nearest", pts, pts.shape[0], dat, x, x.size, y, y.size, R(w.grid_spac), 6, f)
    out["esrg_field"] = f.copy()
    print("esrg done", itp.calls, flush=True)
    f = np.zeros((y.size, x.size), dtype=R, order="F")
    itp.call("idw_kdtree", farr(w.points.T, R), pts.shape[0], dat, x, x.size, y, y.size, 6, f)
    out["kd_field"] = f.copy()
    print("kd done", itp.calls, flush=True)
    inside = ((w.points[:, 0] >= w.x_axis[0]) & (w.points[:, 0] < w.x_axis[-1]) &
              (w.points[:, 1] >= w.y_axis[0]) & (w.points[:, 1] < w.y_axis[-1]))
    bp = farr(w.points[inside], R)
    res = np.zeros(bp.shape[0], R)
    itp.call("bilinear", x, x.size, y, y.size, out["bary_field"], bp, bp.shape[0], res)
    out.update(bil_inside=inside, bil_values=res, inputs_sha256=np.array(workloads.digest(w)))
    path = os.path.join(HERE, "f90_c1_f64.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {itp.calls} procedure calls interpreted")


if __name__ == "__main__":
    if "--c1" in sys.argv:
        c1_full()
    else:
        main()

"""Generate golden vectors from the REFERENCE'S OWN executable Python twins.

Run in the build container only (needs /root/reference); the GPU box never runs
this.  It does not import the reference dev scripts (they import matplotlib and
execute plotting code at module level); instead it extracts, with ``ast``, the
*unmodified source text* of the function definitions named below from

    development/idw_interp_esrg_dev.py   assign_points_to_cells (:121-168),
                                         nearest_neighbours_reg (:174-274)
    development/triangle_walk_dev.py     Point/LineSegment (:19-29), triangle_point_outside (:37-45),
                                         triangle_walk (:82-165), barycentric_interpolation (:173-183),
                                         fill_nearest_neighbour (:189-221)

and executes exactly that text on small seeded inputs.  The outputs are what the
reference itself computes for these inputs; tests/test_oracle_golden.py pins the
C++ oracle (and through it the CUDA path) against them.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import ast
import contextlib
import io
import os
import sys

import numpy as np
from scipy.spatial import Delaunay

REF = os.environ.get("GI_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def extract(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np}
    chunks = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            chunks.append(ast.get_source_segment(src, node))
    assert len(chunks) == len(names), (path, names, len(chunks))
    exec("\n\n".join(chunks), ns)
    return ns


def sin_mountains(x, y, omega=0.6, amplitude=25.0):  # test_interpolation.py:38-40
    return amplitude * np.sin(omega * x + y) + amplitude


def make_esrg():
    ns = extract(os.path.join(REF, "development", "idw_interp_esrg_dev.py"),
                 ["assign_points_to_cells", "nearest_neighbours_reg"])
    # same construction as idw_interp_esrg_dev.py:41-46,66-70 at the "very small" size (:27-30) ...
    cases = {}
    for name, (x_size, y_size, n, spac, k, seed) in {
        "a": (10, 8, 50, 10.0, 6, 33),       # the script's "very small grid/mesh"
        "b": (24, 17, 400, 0.37, 8, 7),      # denser, k = 8, non-integer spacing
        "c": (16, 12, 40, 1.0, 4, 11),       # sparse: many empty cells, deep ring levels
    }.items():
        x_axis = np.arange(-50.0, -50.0 + spac * x_size, spac)[:x_size]
        y_axis = np.arange(-30.0, -30.0 + spac * y_size, spac)[:y_size]
        rs = np.random.RandomState(seed)
        points = np.empty((n, 2))
        points[:, 0] = rs.uniform(x_axis[0], x_axis[-1], n)
        points[:, 1] = rs.uniform(y_axis[0], y_axis[-1], n)
        if name == "c":  # some points outside the grid -> exercises the clamping branches (:125-138)
            points[:5] += np.array([-3.0 * spac, 2.0 * spac])
            points[5:10] += np.array([4.0 * spac, -5.0 * spac])
        index_of_pts, indptr, num_ppgc = ns["assign_points_to_cells"](points, x_axis, y_axis, spac)
        nn_index = np.empty((y_axis.size, x_axis.size, k), np.int32)
        nn_dist = np.empty((y_axis.size, x_axis.size, k))
        for iy in range(y_axis.size):
            for ix in range(x_axis.size):
                a, b = ns["nearest_neighbours_reg"](points, index_of_pts, indptr, num_ppgc, k, x_axis, y_axis,
                                                    spac, ix, iy)
                nn_index[iy, ix], nn_dist[iy, ix] = a, b
        cases[name] = dict(points=points, x_axis=x_axis, y_axis=y_axis, grid_spac=np.float64(spac),
                           k=np.int32(k), index_of_pts=index_of_pts, indptr=indptr, num_ppgc=num_ppgc,
                           nn_index=nn_index, nn_dist=nn_dist)
    flat = {f"{c}_{k}": v for c, d in cases.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(HERE, "esrg_twin.npz"), **flat)
    print("esrg_twin.npz:", {c: d["nn_index"].shape for c, d in cases.items()})


def make_walk():
    ns = extract(os.path.join(REF, "development", "triangle_walk_dev.py"),
                 ["Point", "LineSegment", "triangle_point_outside", "triangle_walk",
                  "barycentric_interpolation", "fill_nearest_neighbour"])
    Point = ns["Point"]
    cases = {}
    for name, (n, lx, ly, seed) in {"a": (50, 14, 9, 33), "b": (600, 41, 29, 5)}.items():
        rs = np.random.RandomState(seed)
        points = np.empty((n, 2))
        points[:, 0] = rs.uniform(-0.33, 20.57, n)   # triangle_walk_dev.py:252-253
        points[:, 1] = rs.uniform(-2.47, 9.89, n)
        tri = Delaunay(points)
        data_pts = sin_mountains(points[:, 0], points[:, 1])
        # grid slightly larger than the hull so that outside cells exist
        x_axis = np.linspace(points[:, 0].min() - 0.4, points[:, 0].max() + 0.4, lx)
        y_axis = np.linspace(points[:, 1].min() - 0.4, points[:, 1].max() + 0.4, ly)
        simplices, neighbours = tri.simplices, tri.neighbors
        # row-sequential scan, every row restarting from the same triangle, as interpolation.f90:365-367
        # (the dev script's own driver :376-423 is a serpentine variant of the same calls)
        ind_tri_start = 0
        tri_out = np.empty((ly, lx), np.int32)
        inside = np.empty((ly, lx), bool)
        iters_out = np.empty((ly, lx), np.int32)
        data_ip = np.full((ly, lx), np.nan)
        for iy in range(ly):
            ind_tri = ind_tri_start
            for ix in range(lx):
                pt = Point(x_axis[ix], y_axis[iy])
                ind_tri, ins, iters = ns["triangle_walk"](points, simplices, neighbours, -1, pt, ind_tri,
                                                          plot=False, print_stat=False)
                tri_out[iy, ix], inside[iy, ix], iters_out[iy, ix] = ind_tri, ins, iters
                if ins:
                    ind = simplices[ind_tri, :]
                    w1, w2, w3 = ns["barycentric_interpolation"](Point(*points[ind[0]]), Point(*points[ind[1]]),
                                                                 Point(*points[ind[2]]), pt)
                    data_ip[iy, ix] = data_pts[ind[0]] * w1 + data_pts[ind[1]] * w2 + data_pts[ind[2]] * w3
        mask = ~inside
        filled = data_ip.copy()
        with contextlib.redirect_stdout(io.StringIO()):  # the twin prints per cell
            ns["fill_nearest_neighbour"](mask, x_axis, y_axis, filled)
        cases[name] = dict(points=points, data_pts=data_pts, x_axis=x_axis, y_axis=y_axis,
                           simplices=simplices.astype(np.int32), neighbours=neighbours.astype(np.int32),
                           tri_start=np.int32(ind_tri_start), tri=tri_out, inside=inside, iters=iters_out,
                           data_walk=data_ip, data_filled=filled)
    # the predicate sanity case of triangle_walk_dev.py:48-59 (P = (3.0, 2.0 - 1e-15) is accepted by the margin)
    verts = (Point(3.0, 2.0), Point(9.0, 3.0), Point(5.0, 7.0))
    pts = [(5.0, 5.0), (7.0, 7.0), (3.0, 2.0 - 1e-15), (4.079, 4.7), (3.0, 2.0 - 1e-9)]
    res = []
    for p in pts:
        P = Point(*p)
        res.append([bool(ns["triangle_point_outside"](ns["LineSegment"](verts[i], verts[(i + 1) % 3]), P))
                    for i in range(3)])
    flat = {f"{c}_{k}": v for c, d in cases.items() for k, v in d.items()}
    flat["pred_points"] = np.array(pts)
    flat["pred_outside"] = np.array(res)
    np.savez_compressed(os.path.join(HERE, "walk_twin.npz"), **flat)
    print("walk_twin.npz:", {c: (d["tri"].shape, int((~d["inside"]).sum())) for c, d in cases.items()},
          flat["pred_outside"].tolist())


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not present: golden vectors can only be regenerated in the build container")
    make_esrg()
    make_walk()

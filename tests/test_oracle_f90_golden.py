"""Pin the C++ oracle against the reference's OWN FORTRAN SOURCE, executed statement by statement by
oracle/f90interp.py (tests/golden/make_golden_f90.py wrote tests/golden/f90_f64.npz / f90_f32.npz in the build
container).  Every index and every floating-point value must agree bit for bit, in double (the graded kind) and in
single precision (the reference as shipped).  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle as orc


@pytest.fixture(scope="module", params=["f64", "f32"])
def g(request, golden_dir):
    z = np.load(os.path.join(golden_dir, f"f90_{request.param}.npz"))
    return z, (np.float64 if request.param == "f64" else np.float32)


def same(a, b):
    assert np.array_equal(np.asarray(a), np.asarray(b))


def test_bilinear_matches_fortran_source(g):
    z, R = g
    same(orc.bilinear(z["x_axis"], z["y_axis"], z["esrg_field"], z["bil_points"], dtype=R), z["bil_values"])
    same(orc.bilinear(z["biln_x_axis"], z["biln_y_axis"], z["biln_field"], z["biln_points"], dtype=R), z["biln_values"])


@pytest.mark.parametrize("case", ["kd", "kds"])
def test_idw_kdtree_matches_fortran_source(g, case):
    """kds = coordinates scaled so that every squared distance is < 1: the regime in which the reference's prune
    rule abs(diff) < max d2 (kd_tree.f90:191,198) drops true neighbours -- the quirk must be reproduced."""
    z, R = g
    pts = z["points"].T if case == "kd" else z["kds_points"]
    x = z["x_axis"] if case == "kd" else z["kds_x_axis"]
    y = z["y_axis"] if case == "kd" else z["kds_y_axis"]
    field, w = orc.idw_kdtree(pts, z["data"], x, y, 6, dtype=R, debug=True)
    same(np.moveaxis(w.nn_index, 0, -1), z[case + "_nn_index"])
    same(np.moveaxis(w.nn_dist2, 0, -1), z[case + "_nn_dist2"])
    same(field, z[case + "_field"])
    if case == "kds" and R is np.float64:  # the vectors really exercise the approximate regime
        from scipy.spatial import cKDTree
        xx, yy = np.meshgrid(x, y)
        _, exact = cKDTree(pts.T).query(np.stack([xx.ravel(), yy.ravel()], 1), k=6)
        assert (np.sort(exact + 1, 1) != np.sort(z["kds_nn_index"].reshape(-1, 6), 1)).any(1).mean() > 0.05


def test_idw_esrg_matches_fortran_source(g):
    z, R = g
    iop, ptr, ppgc = orc.assign_points_to_cells(z["points"], z["x_axis"], z["y_axis"], float(z["grid_spac"]), dtype=R)
    same(iop, z["esrg_index_of_pts"]); same(ptr, z["esrg_indptr"]); same(ppgc, z["esrg_num_ppgc"])
    field, w = orc.idw_esrg_nearest(z["points"], z["data"], z["x_axis"], z["y_axis"], float(z["grid_spac"]), 6,
                                    dtype=R, debug=True)
    same(np.moveaxis(w.nn_index, 0, -1), z["esrg_nn_index"])
    same(np.moveaxis(w.nn_dist, 0, -1), z["esrg_nn_dist"])
    same(field, z["esrg_field"])


def test_barycentric_matches_fortran_source(g):
    z, R = g
    same(orc.barycentric_interpolation(z["points"], z["data"], z["x_axis"], z["y_axis"], z["simplices"],
                                       z["neighbours"], dtype=R), z["bary_field"])


def test_reference_test_case_at_full_size(golden_dir):
    """The reference's own test case at its own size (test_interpolation.py:54-93: 5 000 points -> 132 x 79 grid,
    num_nn = 6), Fortran source interpreted end to end (make_golden_f90.py --c1): the four fields bit for bit."""
    import workloads
    z = np.load(os.path.join(golden_dir, "f90_c1_f64.npz"))
    w = workloads.c1()
    assert workloads.digest(w) == str(z["inputs_sha256"]), "seeded inputs differ from the ones the vectors were made from"
    same(orc.barycentric_interpolation(w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours),
         z["bary_field"])
    same(orc.idw_esrg_nearest(w.points, w.data_pts, w.x_axis, w.y_axis, w.grid_spac, 6), z["esrg_field"])
    same(orc.idw_kdtree(w.points.T, w.data_pts, w.x_axis, w.y_axis, 6), z["kd_field"])
    same(orc.bilinear(w.x_axis, w.y_axis, z["bary_field"], w.points[z["bil_inside"]]), z["bil_values"])


def test_lattice_ties_match_fortran_source(g):
    """Tie-heavy lattice (targets on vertices and edges, exact distance ties, zero distances, equal split keys) and
    the same lattice with exact duplicate points: every field bit for bit."""
    z, R = g
    same(orc.barycentric_interpolation(z["lat_points"], z["lat_data"], z["lat_x_axis"], z["lat_y_axis"],
                                       z["lat_simplices"], z["lat_neighbours"], dtype=R), z["lat_bary_field"])
    for name in ("lat", "dup"):
        P, D = z[name + "_points"], z[name + "_data"]
        for k in (1, 4, 6):
            same(orc.idw_kdtree(P.T, D, z["lat_x_axis"], z["lat_y_axis"], k, dtype=R), z[f"{name}_kd_field_k{k}"])
            same(orc.idw_esrg_nearest(P, D, z["lat_x_axis"], z["lat_y_axis"], 1.0, k, dtype=R),
                 z[f"{name}_esrg_field_k{k}"])


def test_far_away_points_match_fortran_source(g):
    """A small grid among far-away points: they are clamped into the border cells and the ring search grows its radius
    long after its frames have left the grid (query_esrg.f90:44-57,178-181)."""
    z, R = g
    for k in (1, 7, 13):
        same(orc.idw_esrg_nearest(z["far_points"], z["far_data"], z["far_x_axis"], z["far_y_axis"], 1.0, k, dtype=R),
             z[f"far_esrg_field_k{k}"])
        same(orc.idw_kdtree(z["far_points"].T, z["far_data"], z["far_x_axis"], z["far_y_axis"], k, dtype=R),
             z[f"far_kd_field_k{k}"])

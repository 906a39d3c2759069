"""The CUDA library against the reference's OWN FORTRAN SOURCE (executed by oracle/f90interp.py in the build
container, vectors in tests/golden/f90_f64.npz / f90_f32.npz): every neighbour id and every value bit for bit, in
double (graded) and single precision (as shipped).  Runs on the GPU box; reads only the committed vectors."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from grid_interpolation_b200 import interpolation as ip  # noqa: E402


@pytest.fixture(scope="module", params=["f64", "f32"])
def g(request, golden_dir):
    z = np.load(os.path.join(golden_dir, f"f90_{request.param}.npz"))
    return z, (np.float64 if request.param == "f64" else np.float32)


def same(a, b):
    assert np.array_equal(np.asarray(a), np.asarray(b))


def test_bilinear(g):
    z, R = g
    same(ip.bilinear(z["x_axis"], z["y_axis"], z["esrg_field"], z["bil_points"], dtype=R), z["bil_values"])
    same(ip.bilinear(z["biln_x_axis"], z["biln_y_axis"], z["biln_field"], z["biln_points"], dtype=R), z["biln_values"])


@pytest.mark.parametrize("case", ["kd", "kds"])
def test_idw_kdtree(g, case):
    z, R = g
    pts = z["points"].T if case == "kd" else z["kds_points"]
    x = z["x_axis"] if case == "kd" else z["kds_x_axis"]
    y = z["y_axis"] if case == "kd" else z["kds_y_axis"]
    field, d = ip.idw_kdtree_ex(pts, z["data"], x, y, 6, debug=True, dtype=R)
    same(d.nn_index, z[case + "_nn_index"])
    same(d.nn_dist2, z[case + "_nn_dist2"])
    same(field, z[case + "_field"])
    same(ip.idw_kdtree_ex(pts, z["data"], x, y, 6, reference_visits=True, dtype=R), z[case + "_field"])


def test_idw_esrg(g):
    z, R = g
    if R is np.float64:
        iop, ptr = ip.assign_points_to_cells(z["points"], z["x_axis"], z["y_axis"], float(z["grid_spac"]))
        same(iop, z["esrg_index_of_pts"])
        same(ptr[:-1].reshape(z["y_axis"].size, z["x_axis"].size) + 1, z["esrg_indptr"])
    field, d = ip.idw_esrg_nearest_ex(z["points"], z["data"], z["x_axis"], z["y_axis"], float(z["grid_spac"]), 6,
                                      debug=True, dtype=R)
    same(d.nn_index, z["esrg_nn_index"])
    same(np.sqrt(d.nn_dist2), z["esrg_nn_dist"])
    same(field, z["esrg_field"])
    same(ip.idw_esrg_nearest_ex(z["points"], z["data"], z["x_axis"], z["y_axis"], float(z["grid_spac"]), 6,
                                reference_order=True, dtype=R), z["esrg_field"])


@pytest.mark.parametrize("order", ["F", "C"])
def test_barycentric(g, order):
    z, R = g
    same(ip.barycentric_interpolation(z["points"], z["data"], z["x_axis"], z["y_axis"], z["simplices"],
                                      z["neighbours"], order=order, dtype=R), z["bary_field"])
    with ip.BarycentricPlan(z["points"], z["x_axis"], z["y_axis"], z["simplices"], z["neighbours"], order=order,
                            dtype=R) as plan:
        same(plan(z["data"]), z["bary_field"])


def test_reference_test_case_at_full_size(golden_dir):
    """The reference's own test case at its own size (test_interpolation.py:54-93: 5 000 points -> 132 x 79 grid,
    num_nn = 6) against the Fortran source interpreted end to end (make_golden_f90.py --c1)."""
    import workloads
    z = np.load(os.path.join(golden_dir, "f90_c1_f64.npz"))
    w = workloads.c1()
    assert workloads.digest(w) == str(z["inputs_sha256"]), "seeded inputs differ from the ones the vectors were made from"
    same(ip.barycentric_interpolation(w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours),
         z["bary_field"])
    same(ip.idw_esrg_nearest(w.points, w.data_pts, w.x_axis, w.y_axis, w.grid_spac, 6), z["esrg_field"])
    same(ip.idw_kdtree(w.points.T, w.data_pts, w.x_axis, w.y_axis, 6), z["kd_field"])
    same(ip.bilinear(w.x_axis, w.y_axis, z["bary_field"], w.points[z["bil_inside"]]), z["bil_values"])


def test_lattice_ties(g):
    """Tie-heavy lattice (targets on vertices and edges: the walk's margin band and path dependence; exact distance
    ties, zero distances, equal split keys) and the same lattice with exact duplicate points."""
    z, R = g
    for order in ("F", "C"):
        same(ip.barycentric_interpolation(z["lat_points"], z["lat_data"], z["lat_x_axis"], z["lat_y_axis"],
                                          z["lat_simplices"], z["lat_neighbours"], order=order, dtype=R),
             z["lat_bary_field"])
    for name in ("lat", "dup"):
        P, D = z[name + "_points"], z[name + "_data"]
        for k in (1, 4, 6):
            same(ip.idw_kdtree(P.T, D, z["lat_x_axis"], z["lat_y_axis"], k, dtype=R), z[f"{name}_kd_field_k{k}"])
            same(ip.idw_esrg_nearest(P, D, z["lat_x_axis"], z["lat_y_axis"], 1.0, k, dtype=R),
                 z[f"{name}_esrg_field_k{k}"])


def test_far_away_points(g):
    """A small grid among far-away points: they are clamped into the border cells and the ring search grows its radius
    long after its frames have left the grid (query_esrg.f90:44-57,178-181)."""
    z, R = g
    for k in (1, 7, 13):
        same(ip.idw_esrg_nearest(z["far_points"], z["far_data"], z["far_x_axis"], z["far_y_axis"], 1.0, k, dtype=R),
             z[f"far_esrg_field_k{k}"])
        same(ip.idw_kdtree(z["far_points"].T, z["far_data"], z["far_x_axis"], z["far_y_axis"], k, dtype=R),
             z[f"far_kd_field_k{k}"])

"""Headless port of the reference's only end-to-end flow, test_interpolation.py, through the reference's own
import line.  The script only *prints* deviation statistics against scipy (test_interpolation.py:166-205,
254-268); here they are assertions.  Plotting and the shapely "inner domain" polygon are visualisation only and
are left out (components 6)."""
import numpy as np
import pytest
from scipy import interpolate
from scipy.spatial import KDTree

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("num_points", [5_000, 100_000, 500_000])   # 500 000 = the script's default (:54)
def test_test_interpolation_flow(num_points):
    from interpolation import interpolation as ip_fortran          # test_interpolation.py:27
    import workloads
    w = workloads.c1(num_points=num_points)                         # :54-93 (seed 33, sine mountains, grid)
    points, data_pts, x_axis, y_axis, grid_spac = w.points, w.data_pts, w.x_axis, w.y_axis, w.grid_spac

    # scipy reference solutions (:142-159)
    num_nn = 6
    kd_tree = KDTree(points)
    points_ip = np.vstack([i.ravel() for i in np.meshgrid(x_axis, y_axis)]).T
    dist, indices = kd_tree.query(points_ip, k=num_nn, workers=-1)
    data_reg_iwd_ref = ((data_pts[indices] / dist).sum(axis=1) / (1.0 / dist).sum(axis=1)).reshape(
        y_axis.size, x_axis.size)
    data_reg_bi_ref = interpolate.griddata(points, data_pts, points_ip, method="linear").reshape(
        y_axis.size, x_axis.size)

    # "Fortran" solutions, called exactly as the script calls them (:185-204)
    points_ft_trans = np.asfortranarray(points.transpose())
    data_reg_iwd_kdtree_ft = ip_fortran.idw_kdtree(points_ft_trans, data_pts, x_axis, y_axis, num_nn)
    points_ft = np.asfortranarray(points)
    data_reg_iwd_esrg_ft = ip_fortran.idw_esrg_nearest(points_ft, data_pts, x_axis, y_axis, grid_spac, num_nn)
    simplices_ft = np.asfortranarray(w.simplices)                   # already + 1
    neighbours_ft = np.asfortranarray(w.neighbours)
    data_reg_bi_ft = ip_fortran.barycentric_interpolation(points_ft, data_pts, x_axis, y_axis, simplices_ft,
                                                          neighbours_ft)
    for a in (data_reg_iwd_kdtree_ft, data_reg_iwd_esrg_ft, data_reg_bi_ft):
        assert a.shape == (y_axis.size, x_axis.size) and a.flags.f_contiguous    # what f2py returns

    # esrg is an exact kNN: agrees with scipy to rounding
    assert np.abs(data_reg_iwd_esrg_ft - data_reg_iwd_ref).max() < 1e-10
    # the k-d tree uses the reference's (too eager) prune rule: approximate neighbours, bounded deviation
    dev = np.abs(data_reg_iwd_kdtree_ft - data_reg_iwd_ref)
    assert np.isfinite(dev).all() and dev.mean() < 0.5
    # barycentric: equal to griddata inside the hull (NaN outside, where the reference fills)
    inside = ~np.isnan(data_reg_bi_ref)
    assert np.abs(data_reg_bi_ft[inside] - data_reg_bi_ref[inside]).max() < 1e-11
    assert np.isfinite(data_reg_bi_ft).all()

    # regular -> unstructured (:243-259)
    f_ip = interpolate.RegularGridInterpolator((x_axis, y_axis), data_reg_iwd_esrg_ft.transpose(),
                                               method="linear", bounds_error=False)
    data_pts_rec = f_ip(points)
    data_pts_rec_ft = ip_fortran.bilinear(x_axis, y_axis, np.asfortranarray(data_reg_iwd_esrg_ft),
                                          np.asfortranarray(points))
    assert not np.any(data_pts_rec_ft == -9999.0)                   # :257-258
    assert np.abs(data_pts_rec_ft - data_pts_rec).max() < 1e-10
    # re-interpolation error vs the original data (:262-268): small on average
    assert np.abs(data_pts_rec_ft - data_pts).mean() < 0.5

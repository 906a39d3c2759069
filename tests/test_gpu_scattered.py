"""Scattered targets (the reference interpolates to grid nodes only, interpolation.f90:154-158,365-371):
the CUDA entry points against the oracle's restatement of the same definition (GPU)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from grid_interpolation_b200 import interpolation as ip  # noqa: E402
import workloads  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def _targets(points, m, seed, margin=0.15):
    rng = np.random.default_rng(seed)
    lo, hi = points.min(0), points.max(0)
    return rng.uniform(lo - margin * (hi - lo), hi + margin * (hi - lo), (m, 2))


def _np(a):
    return a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)


def _check_bary(got, g, want, w):
    got, tri, outside = _np(got), _np(g.tri), _np(g.outside)
    same_tri = tri == w.tri
    ins = ~w.outside
    # inside the hull the triangle is unique -- except within 1e-10 of an edge, which both neighbours accept
    # (triangle_walk.f90:148): the two walks start from different triangles, so such a target may end in either
    assert same_tri[ins].mean() > 0.9995
    assert (outside != w.outside).sum() <= (~same_tri[ins]).sum()
    # outside the hull the closest location is unique, the triangle it is evaluated in is not when that location is
    # a hull vertex (the walks may leave the hull through different edges)
    assert same_tri[~ins].mean() > 0.5
    assert np.array_equal(got[same_tri], want[same_tri]), "same triangle, different bits"
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("order", ["C", "F"])
def test_barycentric_points_c1(c1, order):
    targets = np.asarray(_targets(c1.points, 20_000, 1), order=order)
    want, w = orc.barycentric_points(c1.points, c1.data_pts, c1.simplices, c1.neighbours, targets, debug=True)
    assert 2000 < w.outside.sum() < 12_000
    got, g = ip.barycentric_points(c1.points, c1.data_pts, c1.simplices, c1.neighbours, targets, debug=True)
    _check_bary(got, g, want, w)
    # the grid version's in-hull values are the same expression at the same targets
    xx, yy = np.meshgrid(c1.x_axis, c1.y_axis)
    nodes = np.stack([xx.ravel(), yy.ravel()], 1)
    field, f = ip.barycentric_interpolation_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                               c1.neighbours, order="C", debug=True)
    at_nodes, gn = ip.barycentric_points(c1.points, c1.data_pts, c1.simplices, c1.neighbours, nodes, debug=True)
    ins = ~f.mask.ravel()
    assert np.array_equal(gn.outside, ~ins)
    same = gn.tri[ins] == f.tri.ravel()[ins]
    assert same.mean() > 0.9995 and np.array_equal(at_nodes[ins][same], field.ravel()[ins][same])


def test_barycentric_points_device_tensors_medium_mesh_and_f32():
    import torch
    rng = np.random.default_rng(3)
    pts = rng.uniform(0.0, 1000.0, (200_000, 2))
    simp, nbr = workloads.delaunay(pts)
    data = np.sin(pts[:, 0] * 0.01) + 0.003 * pts[:, 1]
    targets = _targets(pts, 1_000_000, 4, margin=0.05)
    want, w = orc.barycentric_points(pts, data, simp, nbr, targets, debug=True)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
    got, g = ip.barycentric_points(dev(pts), dev(data), dev(simp), dev(nbr), dev(targets), debug=True)
    _check_bary(got, g, want, w)
    assert ip.barycentric_points(pts, data, simp, nbr, np.zeros((0, 2))).shape == (0,)
    # single precision: same walk in float; the oracle's f32 build is the checker
    p32, d32, t32 = pts.astype(np.float32), data.astype(np.float32), targets[:100_000].astype(np.float32)
    want32, w32 = orc.barycentric_points(p32, d32, simp, nbr, t32, dtype=np.float32, debug=True)
    got32, g32 = ip.barycentric_points(p32, d32, simp, nbr, t32, dtype=np.float32, debug=True)
    same = g32.tri == w32.tri
    assert same.mean() > 0.99 and np.array_equal(got32[same], want32[same])


@pytest.mark.parametrize("k", [6, 40])
def test_idw_kdtree_points(c1, k):
    targets = _targets(c1.points, 30_000, 5)
    pts = np.asfortranarray(c1.points.T)
    want, w = orc.idw_kdtree_points(pts, c1.data_pts, targets, k, debug=True)
    got, g = ip.idw_kdtree_points(pts, c1.data_pts, targets, k, debug=True)
    assert np.array_equal(g.nn_index, w.nn_index) and np.array_equal(g.nn_dist2, w.nn_dist2)
    assert np.array_equal(got, want)
    # grid nodes as scattered targets: the grid entry point's bits
    xx, yy = np.meshgrid(c1.x_axis, c1.y_axis)
    nodes = np.stack([xx.ravel(), yy.ravel()], 1)
    field = ip.idw_kdtree(pts, c1.data_pts, c1.x_axis, c1.y_axis, k, order="C")
    assert np.array_equal(ip.idw_kdtree_points(pts, c1.data_pts, nodes, k), field.ravel())
    import torch
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
    got_d = ip.idw_kdtree_points(dev(pts), dev(c1.data_pts), dev(targets), k)
    assert np.array_equal(got_d.cpu().numpy(), want)


def _exact_knn_idw(points, data, targets, k, R=np.float64):
    """The exact k nearest points of every target (scipy cKDTree for the ids; the k-d ORACLE cannot serve here: like
    the reference it prunes with |diff| < d^2, kd_tree.f90:182, and misses neighbours when distances are below 1 or
    above the 1e6 sentinel), with the reference's arithmetic for everything that is compared bit for bit: squared
    distances as dx*dx + dy*dy, IDW sums in ascending neighbour order (interpolation.f90:260-274)."""
    from scipy.spatial import cKDTree
    p = points.astype(R); t = targets.astype(R); d = data.astype(R)
    _, idx = cKDTree(p.astype(np.float64)).query(t.astype(np.float64), k=k)
    idx = idx.reshape(len(t), k)
    dx = p[idx, 0] - t[:, None, 0]; dy = p[idx, 1] - t[:, None, 1]
    d2 = dx * dx + dy * dy
    order = np.argsort(d2, axis=1, kind="stable")            # the distances as the kernel rounds them decide the order
    idx = np.take_along_axis(idx, order, 1); d2 = np.take_along_axis(d2, order, 1)
    dist = np.sqrt(d2)
    num = np.zeros(len(t), R); den = np.zeros(len(t), R)
    for j in range(k):
        with np.errstate(divide="ignore", invalid="ignore"):
            num = num + d[idx[:, j]] / dist[:, j]
            den = den + R(1) / dist[:, j]
    with np.errstate(divide="ignore", invalid="ignore"):
        out = num / den
    return out.astype(R), (idx + 1).astype(np.int32), d2


@pytest.mark.parametrize("k", [1, 6, 40])
def test_idw_esrg_nearest_points(c1, k):
    """Scattered targets through the cell list: the exact k nearest points in ascending order and the reference's IDW
    sums over them (continuous random coordinates: no distance ties, no zero distances)."""
    targets = _targets(c1.points, 30_000, 7)
    spac = float(c1.x_axis[1] - c1.x_axis[0])
    want, want_nn, want_d2 = _exact_knn_idw(c1.points, c1.data_pts, targets, k)
    got, g = ip.idw_esrg_nearest_points(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, spac, targets, k, debug=True)
    assert np.array_equal(g.nn_index, want_nn) and np.array_equal(g.nn_dist2, want_d2)
    assert np.array_equal(got, want)
    # grid nodes as scattered targets: the grid entry point's bits (and so the oracle's)
    xx, yy = np.meshgrid(c1.x_axis, c1.y_axis)
    nodes = np.stack([xx.ravel(), yy.ravel()], 1)
    field = ip.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, spac, k, order="C")
    assert np.array_equal(field, orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, spac, k))
    assert np.array_equal(ip.idw_esrg_nearest_points(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, spac, nodes, k),
                          field.ravel())
    # device tensors, Fortran-ordered targets, single precision
    import torch
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
    got_d = ip.idw_esrg_nearest_points(dev(c1.points), dev(c1.data_pts), dev(c1.x_axis), dev(c1.y_axis), spac,
                                       dev(targets), k)
    assert np.array_equal(got_d.cpu().numpy(), want)
    assert np.array_equal(ip.idw_esrg_nearest_points(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, spac,
                                                     np.asfortranarray(targets), k), want)
    assert ip.idw_esrg_nearest_points(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, spac, np.zeros((0, 2)), k).shape == (0,)
    f32 = np.float32
    p32, d32, t32 = c1.points.astype(f32), c1.data_pts.astype(f32), targets[:5000].astype(f32)
    want32, nn32, d232 = _exact_knn_idw(p32, d32, t32, k, f32)
    got32, g32 = ip.idw_esrg_nearest_points(p32, d32, c1.x_axis.astype(f32), c1.y_axis.astype(f32), f32(spac), t32, k,
                                            dtype=f32, debug=True)
    assert np.array_equal(g32.nn_dist2, d232)                  # (ids may differ where two f32 distances tie)
    same = (g32.nn_index == nn32).all(1)
    assert same.mean() > 0.99 and np.array_equal(got32[same], want32[same])


def test_idw_esrg_nearest_points_targets_far_outside_the_grid(c1):
    rng = np.random.default_rng(11)
    lo, hi = c1.points.min(0), c1.points.max(0)
    targets = np.concatenate([rng.uniform(lo - 5 * (hi - lo), hi + 5 * (hi - lo), (5000, 2)),
                              rng.uniform(lo - 1e4 * (hi - lo), hi + 1e4 * (hi - lo), (500, 2))])
    spac = float(c1.x_axis[1] - c1.x_axis[0])
    want, want_nn, want_d2 = _exact_knn_idw(c1.points, c1.data_pts, targets, 6)
    got, g = ip.idw_esrg_nearest_points(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, spac, targets, 6, debug=True)
    assert np.array_equal(g.nn_dist2, want_d2)
    same = (g.nn_index == want_nn).all(1)                     # far targets: neighbour distances may tie after rounding
    assert same.mean() > 0.99 and np.array_equal(got[same], want[same])

"""Independent second opinion: oracle vs scipy on the reference's only fixture (C1) -- the comparisons
test_interpolation.py:142-205,243-259 prints, turned into assertions, plus the known-answer numbers of
SURVEY.md section 4.  k-d tree and bilinear have no reference-side golden vector ("parity unpinned"): these
cross-checks and line fidelity are what pins them."""
import numpy as np
import pytest
from scipy import interpolate
from scipy.spatial import Delaunay, KDTree

from oracle import oracle as orc


@pytest.fixture(scope="module")
def ref(c1):
    kd = KDTree(c1.points)
    points_ip = np.vstack([i.ravel() for i in np.meshgrid(c1.x_axis, c1.y_axis)]).T
    dist, indices = kd.query(points_ip, k=c1.num_nn)
    shape = (c1.y_axis.size, c1.x_axis.size)
    idw = ((c1.data_pts[indices] / dist).sum(axis=1) / (1.0 / dist).sum(axis=1)).reshape(shape)
    bi = interpolate.griddata(c1.points, c1.data_pts, points_ip, method="linear").reshape(shape)
    return dict(points_ip=points_ip, dist=dist, indices=indices, idw=idw, bi=bi, shape=shape)


def test_c1_known_answers(c1):
    assert c1.simplices.shape == (9976, 3)
    assert c1.x_axis.size == 132 and c1.y_axis.size == 79
    assert c1.grid_spac == pytest.approx(0.16059899634317984, rel=1e-14)
    assert c1.area_min > 0  # CCW


def test_esrg_equals_scipy_kdtree(c1, ref):
    out, dbg = orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, c1.num_nn,
                                    debug=True)
    idx = np.moveaxis(dbg.nn_index, 0, -1).reshape(-1, c1.num_nn) - 1
    dist = np.moveaxis(dbg.nn_dist, 0, -1).reshape(-1, c1.num_nn)
    assert np.array_equal(idx, ref["indices"])                     # idw_interp_esrg_dev.py:308-309
    assert np.abs(dist - ref["dist"]).max() <= 1e-10               # :310-311
    np.testing.assert_allclose(out, ref["idw"], rtol=1e-12)
    assert dbg.level_max == 5 and dbg.list_max == 30


def test_kdtree_quirk_and_scaled_exactness(c1, ref):
    n = ref["indices"].shape[0]
    out, dbg = orc.idw_kdtree(c1.points.T, c1.data_pts, c1.x_axis, c1.y_axis, c1.num_nn, debug=True)
    idx = np.moveaxis(dbg.nn_index, 0, -1).reshape(-1, c1.num_nn) - 1
    mism = (np.sort(idx, 1) != np.sort(ref["indices"], 1)).any(1).mean()
    assert 0.45 < mism < 0.60            # kd_tree.f90:191,198 prunes too eagerly when d2 < 1
    assert dbg.visits / n == pytest.approx(21.1, abs=0.2)
    # coordinates x10: k-th d2 > 1 -> the rule is lax -> exact kNN
    out10, dbg10 = orc.idw_kdtree(c1.points.T * 10, c1.data_pts, c1.x_axis * 10, c1.y_axis * 10, c1.num_nn,
                                  debug=True)
    idx10 = np.moveaxis(dbg10.nn_index, 0, -1).reshape(-1, c1.num_nn) - 1
    assert np.array_equal(idx10, ref["indices"])
    np.testing.assert_allclose(out10, ref["idw"], rtol=1e-12)


def test_kd_tree_is_canonical(c1):
    """kd_tree.f90:45,54-68: node = rank (n+1)/2 element along axis depth%2 (distinct coordinates)."""
    t = orc.kd_build(c1.points.T)
    pts = c1.points

    def check(node, ids, depth):
        if node < 0:
            assert len(ids) == 0
            return
        axis = depth % 2
        order = ids[np.argsort(pts[ids, axis], kind="stable")]
        m = (len(ids) + 1) // 2
        assert t.index[node] - 1 == order[m - 1]
        check(t.left[node], order[:m - 1], depth + 1)
        check(t.right[node], order[m:], depth + 1)

    check(t.root, np.arange(pts.shape[0]), 0)


def test_barycentric_equals_griddata_and_find_simplex(c1, ref):
    out, dbg = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                             c1.neighbours, debug=True)
    assert dbg.tri_start - 1 == 5 and dbg.iters_tot == 28318 and dbg.fill_level_max == 3
    assert dbg.mask.sum() == 433 and np.array_equal(dbg.mask, np.isnan(ref["bi"]))
    ins = ~dbg.mask
    assert np.abs(out[ins] - ref["bi"][ins]).max() < 1e-13
    fs = Delaunay(c1.points).find_simplex(ref["points_ip"]).reshape(ref["shape"])
    assert np.array_equal(dbg.tri[ins] - 1, fs[ins])
    # fill == brute-force nearest unmasked cell, ties -> smallest row then column
    ii, jj = np.nonzero(ins)
    for (i, j) in zip(*np.nonzero(dbg.mask)):
        d2 = (ii - i) ** 2 + (jj - j) ** 2
        sel = np.flatnonzero(d2 == d2.min())[0]    # nonzero() is row-major -> (i, j) lexicographic
        assert dbg.fill_src[i, j] == ii[sel] * c1.x_axis.size + jj[sel]
        assert out[i, j] == out[ii[sel], jj[sel]]


def test_bilinear_equals_regular_grid_interpolator(c1):
    field = orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, c1.num_nn)
    f_ip = interpolate.RegularGridInterpolator((c1.x_axis, c1.y_axis), field.T, method="linear",
                                               bounds_error=False)
    got = orc.bilinear(c1.x_axis, c1.y_axis, field, c1.points)
    np.testing.assert_allclose(got, f_ip(c1.points), rtol=1e-12)
    assert not (got == -9999.0).any()
    # defined out-of-bounds behaviour (reference: UB after writing -9999.0)
    oob = np.array([[c1.x_axis[0] - 1.0, 0.0], [0.0, c1.y_axis[-1] + 0.5], [c1.x_axis[-1] + 1.0, 0.0], [np.nan, 0.0]])
    assert (orc.bilinear(c1.x_axis, c1.y_axis, field, oob) == -9999.0).all()


def test_f32_build_tracks_f64(c1):
    """The shipped reference is f32: same algorithm, float rounding."""
    a = orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 6)
    b = orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 6, dtype=np.float32)
    assert b.dtype == np.float32 and np.abs(a - b).max() < 1e-3


def test_sliver_guard_extension_masks_hull_slivers_only(c1):
    """The library's opt-in sliver guard (the reference's README to-do, README.md:41-43) as the oracle states it:
    off by default; on, it only adds masked targets, all of them in triangles with a hull edge."""
    plain, p = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                             c1.neighbours, debug=True)
    orc.set_sliver_guard(0.3)
    try:
        got, g = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                               c1.neighbours, debug=True)
    finally:
        orc.set_sliver_guard(0.0)
    added = g.mask & ~p.mask
    assert added.any() and (g.mask | ~p.mask).all()
    assert np.array_equal(got[~g.mask], plain[~g.mask])
    hull = (c1.neighbours == 0).any(axis=1)
    assert hull[g.tri[added] - 1].all()
    again = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours)
    assert np.array_equal(again, plain)


# ------------------------------------------------------------------------------------------------------------
# extensions: scattered targets (the reference is grid-only) -- the oracle's restatement against scipy
# ------------------------------------------------------------------------------------------------------------
def _closest_on_hull_brute(points, hull_edges, p):
    best, q = np.inf, None
    for a, b in hull_edges:
        ab = points[b] - points[a]
        t = np.clip(np.dot(p - points[a], ab) / np.dot(ab, ab), 0.0, 1.0)
        c = points[a] + t * ab
        d = np.dot(p - c, p - c)
        if d < best:
            best, q = d, c
    return q


def test_scattered_barycentric_extension_matches_scipy(c1):
    from scipy.interpolate import LinearNDInterpolator
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(12)
    lo, hi = c1.points.min(0), c1.points.max(0)
    targets = rng.uniform(lo - 0.15 * (hi - lo), hi + 0.15 * (hi - lo), (4000, 2))
    got, g = orc.barycentric_points(c1.points, c1.data_pts, c1.simplices, c1.neighbours, targets, debug=True)
    lin = LinearNDInterpolator(c1.points, c1.data_pts)
    want = lin(targets)
    inside = ~np.isnan(want)
    assert inside.sum() > 1000 and (~inside).sum() > 500
    assert np.array_equal(g.outside, ~inside) or (g.outside != ~inside).sum() <= 2      # margin band only
    ok = inside & ~g.outside
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-9, atol=1e-9)
    # outside the hull: the value at the closest point of the hull
    hull = ConvexHull(c1.points)
    edges = [(s[0], s[1]) for s in hull.simplices]
    for i in np.flatnonzero(g.outside & ~inside)[:150]:
        q = _closest_on_hull_brute(c1.points, edges, targets[i])
        np.testing.assert_allclose(got[i], lin(q[None] * (1 - 1e-13) + c1.points.mean(0) * 1e-13)[0], rtol=1e-6, atol=1e-6)


def test_scattered_kdtree_extension_matches_grid_version_and_scipy(c1):
    from scipy.spatial import cKDTree
    k = 6
    scale = 10.0                                         # the regime in which the reference's prune rule is exact
    pts = np.asfortranarray(c1.points.T * scale)
    xa, ya = c1.x_axis[::5] * scale, c1.y_axis[::5] * scale
    grid, w = orc.idw_kdtree(pts, c1.data_pts, xa, ya, k, debug=True)
    xx, yy = np.meshgrid(xa, ya)
    targets = np.stack([xx.ravel(), yy.ravel()], 1)
    got, g = orc.idw_kdtree_points(pts, c1.data_pts, targets, k, debug=True)
    assert np.array_equal(got.reshape(grid.shape), grid)                       # same targets, same bits
    d, idx = cKDTree(pts.T).query(targets, k=k)
    assert np.array_equal(g.nn_index - 1, idx)

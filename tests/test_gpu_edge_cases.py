"""Edge cases of the four entry points against the oracle (GPU): tiny / degenerate sizes, k at its limits,
empty bands, points outside the grid, mostly-masked fields."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from grid_interpolation_b200 import interpolation as ip  # noqa: E402
import workloads  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def same(a, b):
    assert np.array_equal(np.asarray(a), np.asarray(b))


def test_single_triangle_mostly_outside():
    pts = np.array([[2.0, 1.0], [9.0, 2.5], [4.0, 8.0]])
    simp = np.array([[1, 2, 3]], np.int32); nbr = np.zeros((1, 3), np.int32)
    data = np.array([1.0, -2.0, 5.0])
    x = np.linspace(0.0, 12.0, 37); y = np.linspace(-3.0, 11.0, 29)
    want, w = orc.barycentric_interpolation(pts, data, x, y, simp, nbr, debug=True)
    got, g = ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, debug=True)
    assert w.mask.mean() > 0.7 and w.fill_level_max >= 8
    same(g.mask, w.mask); same(g.fill_src, w.fill_src); same(got, want)


def test_grid_entirely_outside_the_mesh_reports_instead_of_hanging():
    pts = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    simp = np.array([[1, 2, 3]], np.int32); nbr = np.zeros((1, 3), np.int32)
    x = np.linspace(5.0, 6.0, 4); y = np.linspace(5.0, 6.0, 3)
    with pytest.raises(RuntimeError, match="cap"):       # reference: fill_nearest_neighbour loops forever
        ip.barycentric_interpolation(pts, np.ones(3), x, y, simp, nbr)


def test_one_by_one_grid_and_empty_band(c1):
    x = c1.x_axis[40:41]; y = c1.y_axis[30:31]
    same(ip.barycentric_interpolation(c1.points, c1.data_pts, x, y, c1.simplices, c1.neighbours),
         orc.barycentric_interpolation(c1.points, c1.data_pts, x, y, c1.simplices, c1.neighbours))
    same(ip.idw_esrg_nearest(c1.points, c1.data_pts, x, y, c1.grid_spac, 6),
         orc.idw_esrg_nearest(c1.points, c1.data_pts, x, y, c1.grid_spac, 6))
    same(ip.idw_kdtree(c1.points.T, c1.data_pts, x, y, 6), orc.idw_kdtree(c1.points.T, c1.data_pts, x, y, 6))
    empty = ip.idw_esrg_nearest_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 6, rows=(7, 7))
    assert empty.shape == (0, c1.x_axis.size)


@pytest.mark.parametrize("k", [1, 2, 16, 17, 32])
def test_k_limits(c1, k):
    sub = slice(0, 600)
    pts, dat = c1.points[sub], c1.data_pts[sub]
    x, y = c1.x_axis[::3], c1.y_axis[::3]
    spac = c1.grid_spac * 3
    want, w = orc.idw_esrg_nearest(pts, dat, x, y, spac, k, debug=True)
    got, g = ip.idw_esrg_nearest_ex(pts, dat, x, y, spac, k, debug=True)
    same(g.nn_index, np.moveaxis(w.nn_index, 0, -1)); same(got, want)
    want, w = orc.idw_kdtree(pts.T, dat, x, y, k, debug=True)
    got, g = ip.idw_kdtree_ex(pts.T, dat, x, y, k, debug=True)
    same(g.nn_index, np.moveaxis(w.nn_index, 0, -1)); same(got, want)


@pytest.mark.parametrize("k", [33, 50, 200])
def test_k_above_32_uses_the_any_k_lists(c1, k):
    """The reference takes any num_neighbours (interpolation.f90:107-108,201-202); above 32 the neighbour lists
    live in scratch memory instead of registers (topk.cuh, KC = 0).  Same ids, same order, same bits."""
    sub = slice(0, 900)
    pts, dat = c1.points[sub], c1.data_pts[sub]
    x, y = c1.x_axis[::4], c1.y_axis[::4]
    spac = c1.grid_spac * 4
    want, w = orc.idw_esrg_nearest(pts, dat, x, y, spac, k, debug=True)
    for order in ("F", "C"):
        got, g = ip.idw_esrg_nearest_ex(pts, dat, x, y, spac, k, debug=True, order=order)
        same(g.nn_index, np.moveaxis(w.nn_index, 0, -1)); same(got, want)
    got = ip.idw_esrg_nearest(pts, dat, x, y, spac, k)     # the pipelined host path
    same(got, want)
    want, w = orc.idw_kdtree(pts.T, dat, x, y, k, debug=True)
    got, g = ip.idw_kdtree_ex(pts.T, dat, x, y, k, debug=True)
    same(g.nn_index, np.moveaxis(w.nn_index, 0, -1)); same(got, want)
    got = ip.idw_kdtree(pts.T, dat, x, y, k)
    same(got, want)
    with ip.KdTreePlan(pts.T, x, y, k) as plan:
        same(plan(dat), want)


def test_esrg_exactly_k_points_and_points_outside_grid():
    rng = np.random.default_rng(8)
    pts = rng.uniform(-4.0, 14.0, (8, 2))           # several outside the 10 x 7 grid -> clamped cells
    dat = rng.normal(size=8)
    x = np.arange(10.0); y = np.arange(7.0)
    want, w = orc.idw_esrg_nearest(pts, dat, x, y, 1.0, 8, debug=True)
    got, g = ip.idw_esrg_nearest_ex(pts, dat, x, y, 1.0, 8, debug=True)
    same(g.nn_index, np.moveaxis(w.nn_index, 0, -1)); same(got, want)
    idx, ptr = ip.assign_points_to_cells(pts, x, y, 1.0)
    oi, op, _ = orc.assign_points_to_cells(pts, x, y, 1.0)
    same(idx, oi); same(ptr[:-1].reshape(7, 10) + 1, op)


def _far_case(k, spread):
    rng = np.random.default_rng(int(spread) + k)
    x = 100.0 + 0.5 * np.arange(40.0); y = -7.0 + 0.5 * np.arange(30.0)
    pts = np.column_stack([x.mean() + rng.uniform(-1, 1, 300) * 20.0 * spread,
                           y.mean() + rng.uniform(-1, 1, 300) * 15.0 * spread])
    return x, y, pts, rng.normal(size=300)


@pytest.mark.parametrize("k", [3, 8, 14, 20, 32])
@pytest.mark.parametrize("spread", [3.0, 40.0])
def test_esrg_points_far_outside_the_grid(k, spread):
    """Sparse points around a small grid: most are clamped into the border cells and are accepted only when the
    search radius has grown far beyond the grid (query_esrg.f90:178-181 keeps incrementing the level).  Tile kernel
    (k <= 16), per-thread kernel (k > 16) and the reference-order kernel; neighbour ids and max level included."""
    x, y, pts, dat = _far_case(k, spread)
    want, w = orc.idw_esrg_nearest(pts, dat, x, y, 0.5, k, debug=True)
    for ref_order in (False, True):
        got, g = ip.idw_esrg_nearest_ex(pts, dat, x, y, 0.5, k, debug=True, reference_order=ref_order)
        same(g.nn_index, np.moveaxis(w.nn_index, 0, -1)); same(got, want)
        if ref_order:
            assert g.level_max == w.level_max


@pytest.mark.parametrize("k", [8, 20])
def test_esrg_points_extremely_far_outside_the_grid(k):
    """~1e6 ring levels between the grid and the points: the levels are skipped, not iterated.  (The line-faithful
    oracle would iterate them, so the check is a brute-force exact kNN with the reference's IDW expression.)"""
    x, y, pts, dat = _far_case(k, 3.0e4)
    xx, yy = np.meshgrid(x, y)
    d2 = (xx[..., None] - pts[:, 0]) ** 2 + (yy[..., None] - pts[:, 1]) ** 2       # query_esrg.f90:147-148
    nn = np.argsort(d2, axis=-1, kind="stable")[..., :k]
    dist = np.sqrt(np.take_along_axis(d2, nn, -1))
    num = np.zeros_like(xx); den = np.zeros_like(xx)
    for i in range(k):                                                              # interpolation.f90:267-273
        num = num + dat[nn[..., i]] / dist[..., i]
        den = den + 1.0 / dist[..., i]
    for ref_order in (False, True):
        got, g = ip.idw_esrg_nearest_ex(pts, dat, x, y, 0.5, k, debug=True, reference_order=ref_order)
        same(g.nn_index, nn + 1); same(got, num / den)
        if ref_order:
            assert g.level_max > 100_000


def test_kdtree_fewer_points_than_k():
    """Slots that never fill keep the sentinel distance and an undefined id in the reference
    (interpolation.f90:151,159); both sides define id 0 / value 0 for them."""
    rng = np.random.default_rng(9)
    pts = rng.uniform(0.0, 5.0, (2, 3)); dat = rng.normal(size=3)
    x = np.linspace(0, 5, 6); y = np.linspace(0, 5, 5)
    want, w = orc.idw_kdtree(pts, dat, x, y, 5, debug=True)
    got, g = ip.idw_kdtree_ex(pts, dat, x, y, 5, debug=True)
    assert (w.nn_index == 0).any()
    same(g.nn_index, np.moveaxis(w.nn_index, 0, -1)); same(got, want)


def test_bilinear_two_node_axes_and_boundaries():
    x = np.array([1.0, 3.0]); y = np.array([-1.0, 0.5])
    f = np.array([[1.0, 2.0], [3.0, 5.0]])
    pts = np.array([[1.0, -1.0], [2.9999999, 0.4999999], [3.0, 0.0], [2.0, 0.5], [0.999, 0.0], [2.0, -0.25]])
    same(ip.bilinear(x, y, f, pts), orc.bilinear(x, y, f, pts))


def test_non_contiguous_and_integer_inputs_are_converted_like_f2py(c1):
    pts = np.asfortranarray(c1.points)[::2]                       # strided view
    dat = c1.data_pts[::2]
    simp, nbr = workloads.delaunay(np.ascontiguousarray(pts))
    want = orc.barycentric_interpolation(pts, dat, c1.x_axis, c1.y_axis, simp, nbr)
    got = ip.barycentric_interpolation(pts, dat.astype(np.float32).astype(np.float64), list(c1.x_axis), c1.y_axis,
                                       simp.astype(np.int64), nbr.astype(np.int64))
    want2 = orc.barycentric_interpolation(pts, dat.astype(np.float32).astype(np.float64), c1.x_axis, c1.y_axis,
                                          simp, nbr)
    same(got, want2)
    assert want.shape == got.shape


@pytest.mark.parametrize("order", ["C", "F"])
def test_bilinear_search_axes_extension(order):
    """search_axes=True (extension): bisection on the axis values -> correct on non-uniform axes, where
    the reference's mean-spacing formula picks wrong cells; checked against scipy."""
    from scipy.interpolate import RegularGridInterpolator
    rng = np.random.default_rng(21)
    x = np.cumsum(rng.uniform(0.2, 3.0, 60)); y = np.cumsum(rng.uniform(0.2, 3.0, 44))     # strongly non-uniform
    field = np.asarray(rng.normal(size=(44, 60)), order=order)
    n = 20_000
    pts = np.column_stack([rng.uniform(x[0] - 1, x[-1] + 1, n), rng.uniform(y[0] - 1, y[-1] + 1, n)])
    got = ip.bilinear(x, y, field, pts, search_axes=True)
    inside = (pts[:, 0] >= x[0]) & (pts[:, 0] < x[-1]) & (pts[:, 1] >= y[0]) & (pts[:, 1] < y[-1])
    assert np.array_equal(got[~inside], np.full((~inside).sum(), -9999.0))
    want = RegularGridInterpolator((y, x), field)(pts[inside][:, ::-1])
    # the reference divides by the MEAN spacings (interpolation.f90:93): rescale to the true cell area
    ix = np.searchsorted(x, pts[inside, 0], side="right") - 1; iy = np.searchsorted(y, pts[inside, 1], side="right") - 1
    mean_area = ((x[-1] - x[0]) / (x.size - 1)) * ((y[-1] - y[0]) / (y.size - 1))
    scale = mean_area / ((x[ix + 1] - x[ix]) * (y[iy + 1] - y[iy]))
    np.testing.assert_allclose(got[inside] * scale, want, rtol=1e-9, atol=1e-11)
    # on equally spaced axes both cell rules agree (up to points that sit on a node to within rounding)
    xu = np.linspace(0.0, 59.0, 60); yu = np.linspace(0.0, 43.0, 44)
    p2 = np.column_stack([rng.uniform(0, 59, n), rng.uniform(0, 43, n)])
    np.testing.assert_allclose(ip.bilinear(xu, yu, field, p2, search_axes=True), ip.bilinear(xu, yu, field, p2),
                               rtol=1e-12, atol=1e-12)


def test_out_tensor_is_validated_before_the_kernel_writes_through_it(c1):
    """A caller-supplied torch `out` of the wrong dtype / device would make the kernel write 8-byte elements through
    a foreign pointer: it must be a ValueError, like the numpy branch."""
    import torch
    d = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
    args = [d(a) for a in (c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours)]
    shape = (c1.y_axis.size, c1.x_axis.size)
    good = torch.empty(shape, dtype=torch.float64, device="cuda:0")
    ip.barycentric_interpolation(*args, out=good)
    same(good.cpu().numpy(), orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis,
                                                          c1.simplices, c1.neighbours))
    for bad in (torch.empty(shape, dtype=torch.float32, device="cuda:0"),      # 4-byte elements on the f64 path
                torch.empty(shape, dtype=torch.int64, device="cuda:0"),
                torch.empty(shape, dtype=torch.float64),                        # host tensor
                np.empty(shape)):                                               # numpy out on the device path
        with pytest.raises(ValueError):
            ip.barycentric_interpolation(*args, out=bad)
    with pytest.raises(ValueError):                                             # and through a plan
        with ip.BarycentricPlan(args[0], args[2], args[3], args[4], args[5]) as plan:
            plan(args[1], out=torch.empty(shape, dtype=torch.float32, device="cuda:0"))
    with pytest.raises(ValueError):                                             # torch out on the host path
        ip.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours,
                                     out=good)

"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Everything goes through the C ABI via the
reference-facing Python layer and is compared with the CPU oracle on the same inputs:
indices bit-exact, float64 values bit-exact (both sides evaluate the reference's expressions without FMA
contraction), plus the committed golden vectors of the reference's own Python twins."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from grid_interpolation_b200 import interpolation as ip  # noqa: E402
import workloads  # noqa: E402
from oracle import oracle as orc  # noqa: E402

RTOL = 1e-12  # north_star: FP64 values within 1e-12 relative (we additionally assert bit equality)


def assert_values(got, want):
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    assert np.array_equal(np.asarray(got), np.asarray(want)), "values are within 1e-12 but not bit-identical"


# ------------------------------------------------------------------------------------------ bilinear
@pytest.mark.parametrize("order", ["C", "F"])
def test_bilinear_c1(c1, order):
    field = orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, c1.num_nn)
    want = orc.bilinear(c1.x_axis, c1.y_axis, field, c1.points)
    data = np.ascontiguousarray(field) if order == "C" else np.asfortranarray(field)
    pts = np.ascontiguousarray(c1.points) if order == "C" else np.asfortranarray(c1.points)
    got = ip.bilinear(c1.x_axis, c1.y_axis, data, pts)
    assert got.shape == (c1.points.shape[0],)
    assert_values(got, want)


def test_bilinear_random_nonuniform_and_oob():
    rng = np.random.default_rng(0)
    x = np.cumsum(rng.uniform(0.5, 1.5, 301)); y = np.cumsum(rng.uniform(0.2, 0.4, 157)) - 7.0
    field = rng.normal(size=(157, 301))
    pts = np.stack([rng.uniform(x[0] - 5, x[-1] + 5, 20011), rng.uniform(y[0] - 2, y[-1] + 2, 20011)], 1)
    pts[:3] = [[np.nan, 0.0], [x[0], y[0]], [x[-1], y[-1]]]
    want = orc.bilinear(x, y, field, pts)
    got = ip.bilinear(x, y, field, pts)
    assert (want == -9999.0).sum() > 100
    assert_values(got, want)


def test_bilinear_empty_and_tiny():
    x, y = np.array([0.0, 1.0]), np.array([0.0, 2.0])
    f = np.array([[1.0, 2.0], [3.0, 4.0]])
    assert ip.bilinear(x, y, f, np.zeros((0, 2))).shape == (0,)
    assert_values(ip.bilinear(x, y, f, np.array([[0.25, 0.5]])), orc.bilinear(x, y, f, np.array([[0.25, 0.5]])))


# --------------------------------------------------------------------------------------- barycentric
@pytest.mark.parametrize("order", ["F", "C"])
def test_barycentric_c1(c1, order):
    want, w = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                            c1.neighbours, debug=True)
    got, g = ip.barycentric_interpolation_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                             c1.neighbours, order=order, debug=True)
    assert got.shape == want.shape == (79, 132)
    assert got.flags.f_contiguous if order == "F" else got.flags.c_contiguous
    assert np.array_equal(g.mask, w.mask) and g.masked == 433 and g.fixups == 0
    ins = ~w.mask
    assert np.array_equal(g.tri[ins], w.tri[ins])              # found triangle ids, bit-exact
    assert np.array_equal(g.fill_src, w.fill_src)              # hull-fill sources incl. tie-breaks
    assert g.fill_level_max == w.fill_level_max == 3
    assert_values(got, want)
    # plain entry point, Fortran-ordered inputs exactly as test_interpolation.py:192-204 passes them
    plain = ip.barycentric_interpolation(np.asfortranarray(c1.points), c1.data_pts, c1.x_axis, c1.y_axis,
                                         np.asfortranarray(c1.simplices), np.asfortranarray(c1.neighbours))
    assert_values(plain, want)


@pytest.mark.parametrize("case", ["a", "b"])
def test_barycentric_golden_twin(golden_dir, case):
    z = np.load(os.path.join(golden_dir, "walk_twin.npz"))
    g = lambda k: z[f"{case}_{k}"]
    got, dbg = ip.barycentric_interpolation_ex(g("points"), g("data_pts"), g("x_axis"), g("y_axis"),
                                               g("simplices") + 1, g("neighbours") + 1, debug=True)
    inside = g("inside")
    assert np.array_equal(~dbg.mask, inside)
    assert np.array_equal(dbg.tri[inside] - 1, g("tri")[inside])
    assert np.array_equal(got, g("data_filled"))


def test_barycentric_row_bands_with_halo(c1):
    want = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                         c1.neighbours)
    ly = c1.y_axis.size
    for order in ("C", "F"):
        parts = []
        for r0, r1 in [(0, 20), (20, 41), (41, 42), (42, ly)]:
            parts.append(ip.barycentric_interpolation_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis,
                                                         c1.simplices, c1.neighbours, rows=(r0, r1), halo=4,
                                                         order=order))
        assert_values(np.concatenate(parts, 0), want)
    # a band whose fill needs rows outside the walked window must say so instead of returning wrong cells
    with pytest.raises(RuntimeError, match="halo"):
        ip.barycentric_interpolation_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                        c1.neighbours, rows=(0, 2), halo=0)


def test_barycentric_targets_on_mesh_edges():
    """Structured mesh, grid nodes exactly on mesh vertices and edges: such targets pass the orientation test of
    two (or more) triangles and the reference returns the one its row-wise walk reaches first
    (triangle_walk.f90:148-151, interpolation.f90:365-373).  The fix-up pass must reproduce that choice."""
    gx, gy = np.meshgrid(np.arange(0.0, 12.0), np.arange(0.0, 9.0))
    pts = np.stack([gx.ravel(), gy.ravel()], 1)
    simp, nbr = workloads.delaunay(pts)
    a = pts[simp - 1]
    area = 0.5 * ((a[:, 1, 0] - a[:, 0, 0]) * (a[:, 2, 1] - a[:, 0, 1]) - (a[:, 1, 1] - a[:, 0, 1]) * (a[:, 2, 0] - a[:, 0, 0]))
    keep = area > 1e-9                          # Qhull may emit flat triangles for co-circular lattice points
    assert keep.all(), "degenerate triangles in the lattice triangulation"
    data = np.sin(pts[:, 0]) + 0.3 * pts[:, 1] ** 2
    x = np.arange(-0.5, 12.0, 0.5); y = np.arange(-0.5, 9.0, 0.25)
    want, w = orc.barycentric_interpolation(pts, data, x, y, simp, nbr, debug=True)
    for order in ("C", "F"):
        got, g = ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, debug=True, order=order)
        assert g.fixups > 20                   # many targets sit on edges / vertices
        assert np.array_equal(g.mask, w.mask)
        ins = ~w.mask
        assert np.array_equal(g.tri[ins], w.tri[ins])
        assert np.array_equal(g.fill_src, w.fill_src)
        assert_values(got, want)


@pytest.mark.parametrize("cols,halo", [((-4, 70), 17), ((13, 51), 4)])
def test_band_fill_reads_fixed_up_halo_sources(cols, halo):
    """Row bands whose hull fill takes its source cells from HALO rows that lie exactly on mesh edges (a sheared,
    scaled lattice: coordinates are not exactly representable, so the two triangles sharing an edge give values that
    differ in the last bit).  The band must read the value of the triangle the reference's row-wise walk chooses --
    the fix-up paths leave it in the HaloFix store -- i.e. equal the one-shot field and the oracle bit for bit.
    Columns (-4, 70) cross the sheared hull on both sides (the nearest source of a corner target is ~16 cells away,
    hence the wide halo); columns (13, 51) stay inside the x-range every lattice row covers (sources 3 rows away)."""
    import torch
    gx, gy = np.meshgrid(np.arange(0.0, 14.0), np.arange(0.0, 11.0))
    pts = np.stack([0.37 * gx.ravel() + 0.11 * gy.ravel() + 0.1, 0.29 * gy.ravel() + 0.2], 1)
    simp, nbr = workloads.delaunay(pts)
    data = np.sin(3.0 * pts[:, 0]) + 0.3 * pts[:, 1] ** 2 + 7.0
    # target rows ON the lattice rows y = 0.29 j + 0.2 (and between them)
    x = 0.1 + 0.0925 * np.arange(*cols)
    y = 0.2 + 0.145 * np.arange(-3, 24)
    want, w = orc.barycentric_interpolation(pts, data, x, y, simp, nbr, debug=True)
    assert w.mask[:3].all() and not w.mask[3].all()          # rows 0..2 lie below the hull: filled from row 3 upwards
    ly = y.size
    for order in ("C", "F"):
        full, g = ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, debug=True, order=order)
        assert g.fixups > 50
        assert_values(full, want)
        for bands in ([(0, 3), (3, 4), (4, 21), (21, ly)], [(0, 2), (2, 23), (23, ly)]):
            parts = [ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, rows=b, halo=halo, order=order)
                     for b in bands]
            assert_values(np.concatenate(parts, 0), want)
            dev = [torch.as_tensor(np.ascontiguousarray(a), device="cuda:0") for a in (pts, data, x, y, simp, nbr)]
            parts = [ip.barycentric_interpolation_ex(*dev, rows=b, halo=halo, order=order).cpu().numpy() for b in bands]
            assert_values(np.concatenate(parts, 0), want)
            parts = [ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, rows=b, halo=halo, order=order,
                                                     plane_values=True) for b in bands]
            assert np.array_equal(np.concatenate(parts, 0),
                                  ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, order=order, plane_values=True))


def _lattice_case():
    gx, gy = np.meshgrid(np.arange(0.0, 12.0), np.arange(0.0, 9.0))
    pts = np.stack([gx.ravel(), gy.ravel()], 1)
    simp, nbr = workloads.delaunay(pts)
    data = np.sin(pts[:, 0]) + 0.3 * pts[:, 1] ** 2
    return pts, data, np.arange(-0.5, 12.0, 0.5), np.arange(-0.5, 9.0, 0.25), simp, nbr


@pytest.mark.parametrize("order", ["C", "F"])
def test_barycentric_reference_order_walk(c1, order):
    """GI_BARY_REFERENCE_ORDER: one thread per grid row walks the row from ind_tri_start exactly like the reference
    (interpolation.f90:362-404) -- same triangles (also for targets on edges), mask, fill sources and values, on the
    reference's own test case, on the lattice mesh, and on a row band with halo."""
    for pts, data, x, y, simp, nbr in ((c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours),
                                       _lattice_case()):
        want, w = orc.barycentric_interpolation(pts, data, x, y, simp, nbr, debug=True)
        got, g = ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, debug=True, order=order,
                                                 reference_order=True)
        assert np.array_equal(g.mask, w.mask)
        assert np.array_equal(g.tri[~w.mask], w.tri[~w.mask])
        assert np.array_equal(g.fill_src, w.fill_src)
        assert_values(got, want)
        r0, r1 = 5, y.size - 7
        band = ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, rows=(r0, r1), halo=6, order=order,
                                               reference_order=True)
        assert_values(band, want[r0:r1])


@pytest.fixture
def tiny_fixup_list():
    ip.set_fixup_list_capacity(64)
    yield
    ip.set_fixup_list_capacity(0)


def test_barycentric_more_flagged_targets_than_the_fixup_list_holds(tiny_fixup_list):
    """Every lattice target on an edge or vertex is flagged; with a 64-entry fix-up list (normally 4 M) the list
    overflows and the row-sequential walk takes over -- nothing may be left to phase 1's own path."""
    pts, data, x, y, simp, nbr = _lattice_case()
    want, w = orc.barycentric_interpolation(pts, data, x, y, simp, nbr, debug=True)
    for order in ("C", "F"):
        got, g = ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, debug=True, order=order)
        assert np.array_equal(g.mask, w.mask)
        assert np.array_equal(g.tri[~w.mask], w.tri[~w.mask])
        assert np.array_equal(g.fill_src, w.fill_src)
        assert_values(got, want)
        host = ip.barycentric_interpolation(pts, data, x, y, simp, nbr, order=order)      # pipelined host path
        assert_values(host, want)


def test_barycentric_grid_rows_along_mesh_edges_for_hundreds_of_columns():
    """A lattice mesh 330 columns wide whose horizontal edges coincide with grid rows: along such a row EVERY target is
    inside the margin band, so a fix-up finds no path-independent column within its 256-column back-up range; the
    row-sequential walk then redoes the window along the reference's own path (interpolation.f90:362-404)."""
    gx, gy = np.meshgrid(np.arange(0.0, 330.0), np.arange(0.0, 6.0))
    pts = np.stack([gx.ravel(), gy.ravel()], 1)
    simp, nbr = workloads.delaunay(pts)
    data = np.cos(0.1 * pts[:, 0]) + 0.2 * pts[:, 1] ** 2
    x = np.arange(0.0, 329.01, 0.5); y = np.arange(0.0, 5.01, 0.5)
    want, w = orc.barycentric_interpolation(pts, data, x, y, simp, nbr, debug=True)
    for order in ("C", "F"):
        got, g = ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, debug=True, order=order)
        assert np.array_equal(g.mask, w.mask)
        assert np.array_equal(g.tri[~w.mask], w.tri[~w.mask])
        assert_values(got, want)
        with ip.BarycentricPlan(pts, x, y, simp, nbr, order=order) as plan:
            assert_values(plan(data), want)


@pytest.mark.parametrize("q", [0.05, 0.3])
def test_barycentric_sliver_guard(c1, q):
    """The reference's README to-do (README.md:41-43), opt-in: hull slivers below the quality threshold are not
    interpolated in, their targets are masked and filled from the nearest neighbours -- same switch, same
    expression, same bits in the oracle and the library; every API door."""
    plain, p = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                             c1.neighbours, debug=True)
    orc.set_sliver_guard(q); ip.set_sliver_guard(q)
    try:
        want, w = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                                c1.neighbours, debug=True)
        assert w.mask.sum() > p.mask.sum() and (w.mask | ~p.mask).all()      # the guard only adds masked targets
        assert np.array_equal(want[~w.mask], plain[~w.mask])
        for order in ("F", "C"):
            got, g = ip.barycentric_interpolation_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                                     c1.neighbours, order=order, debug=True)
            assert np.array_equal(g.mask, w.mask) and np.array_equal(g.fill_src, w.fill_src)
            assert_values(got, want)
        assert_values(ip.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                                   c1.neighbours), want)
        with ip.BarycentricPlan(c1.points, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours) as plan:
            assert_values(plan(c1.data_pts), want)
        rows, halo = (0, 40), 12
        band = ip.barycentric_interpolation_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                               c1.neighbours, rows=rows, halo=halo, order="C")
        assert_values(band, want[rows[0]:rows[1]])
    finally:
        orc.set_sliver_guard(0.0); ip.set_sliver_guard(0.0)


def test_barycentric_bad_topology_is_reported(c1):
    bad = c1.simplices.copy(); bad[7, 1] = c1.points.shape[0] + 5
    with pytest.raises(RuntimeError, match="out of range"):
        ip.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, bad, c1.neighbours)


def test_fill_standalone_matches_oracle():
    rng = np.random.default_rng(5)
    mask = rng.random((61, 83)) < 0.55
    mask[20:45, 30:70] = True          # a big hole: deep levels
    mask[0, 0] = False
    data = rng.normal(size=mask.shape)
    want, wsrc, lvl = orc.fill_nearest_neighbour(mask, data)
    got, gsrc = ip.fill_nearest_neighbour(mask, data)
    assert lvl >= 10
    assert np.array_equal(gsrc, wsrc)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------- esrg
@pytest.mark.parametrize("order", ["F", "C"])
@pytest.mark.parametrize("k", [6, 8, 1, 13])
def test_esrg_c1(c1, order, k):
    want, w = orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, k, debug=True)
    got, g = ip.idw_esrg_nearest_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, k, order=order,
                                    debug=True)
    assert np.array_equal(g.nn_index, np.moveaxis(w.nn_index, 0, -1))     # neighbour ids in order, bit-exact
    assert np.array_equal(np.sqrt(g.nn_dist2), np.moveaxis(w.nn_dist, 0, -1))
    assert_values(got, want)
    # random FP64 inputs have no equal distances: only a handful of targets with an overfull candidate list
    # (local density far above the mean) are handed to the exact-order kernel
    assert g.exact_order_targets <= got.size // 50
    # the kernel that follows the reference's candidate order, forced on every target
    got2, g2 = ip.idw_esrg_nearest_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, k, order=order,
                                      debug=True, reference_order=True)
    assert g2.exact_order_targets == got.size and g2.level_max == w.level_max
    assert np.array_equal(g2.nn_index, np.moveaxis(w.nn_index, 0, -1))
    assert_values(got2, want)


def test_esrg_golden_twin(golden_dir):
    z = np.load(os.path.join(golden_dir, "esrg_twin.npz"))
    for case in "abc":
        g = lambda k: z[f"{case}_{k}"]
        pts = g("points")
        data = np.cos(pts[:, 0]) + pts[:, 1]
        k = int(g("k"))
        _, dbg = ip.idw_esrg_nearest_ex(pts, data, g("x_axis"), g("y_axis"), float(g("grid_spac")), k, debug=True)
        assert np.array_equal(dbg.nn_index - 1, g("nn_index"))
        idx, ptr = ip.assign_points_to_cells(pts, g("x_axis"), g("y_axis"), float(g("grid_spac")))
        assert np.array_equal(idx - 1, g("index_of_pts"))
        assert np.array_equal(ptr[:-1].reshape(g("indptr").shape), g("indptr"))


def test_esrg_clustered_overflows_candidate_list():
    """Dense clusters: thousands of points per cell -> the reference's 1000-entry lists would overflow
    (query_esrg.f90:119-121); the list-free fallback must still give the oracle's (growable-list) answer."""
    rng = np.random.default_rng(11)
    x = np.arange(40.0); y = np.arange(30.0)
    pts = np.concatenate([rng.normal([10.2, 11.7], 0.6, (4000, 2)), rng.uniform([0, 0], [39, 29], (300, 2))])
    data = np.sin(pts[:, 0]) * pts[:, 1]
    want, w = orc.idw_esrg_nearest(pts, data, x, y, 1.0, 8, debug=True)
    got, g = ip.idw_esrg_nearest_ex(pts, data, x, y, 1.0, 8, debug=True)
    assert w.list_max > 1000
    assert np.array_equal(g.nn_index, np.moveaxis(w.nn_index, 0, -1))
    assert_values(got, want)
    got2, g2 = ip.idw_esrg_nearest_ex(pts, data, x, y, 1.0, 8, debug=True, reference_order=True)
    assert g2.list_overflows > 0                # the deferred list overflowed -> list-free re-scan mode
    assert np.array_equal(g2.nn_index, np.moveaxis(w.nn_index, 0, -1))
    assert_values(got2, want)


def test_esrg_ties_on_lattice():
    """Points on a lattice: many exactly equal distances -> exercises the tie order (new point goes before
    equal ones, boundary tie keeps the earlier one; query_esrg.f90:163-168) and the stable cell order."""
    gx, gy = np.meshgrid(np.arange(0.0, 12.0, 0.5), np.arange(0.0, 9.0, 0.5))
    pts = np.stack([gx.ravel(), gy.ravel()], 1)
    rng = np.random.default_rng(3)
    pts = pts[rng.permutation(pts.shape[0])]
    data = rng.normal(size=pts.shape[0])
    x = np.arange(12.0); y = np.arange(9.0)
    want, w = orc.idw_esrg_nearest(pts, data, x, y, 1.0, 7, debug=True)
    got, g = ip.idw_esrg_nearest_ex(pts, data, x, y, 1.0, 7, debug=True)
    assert g.exact_order_targets > got.size // 2   # ties everywhere: resolved in the reference's candidate order
    assert np.array_equal(g.nn_index, np.moveaxis(w.nn_index, 0, -1))
    assert_values(got, want)     # includes the "extremely close neighbour" branch (distance 0)


def test_esrg_too_few_points():
    with pytest.raises(RuntimeError, match="fewer"):
        ip.idw_esrg_nearest(np.random.rand(3, 2), np.zeros(3), np.arange(4.0), np.arange(4.0), 1.0, 5)


# ------------------------------------------------------------------------------------------- k-d tree
def test_kd_build_is_canonical(c1):
    order = ip.kd_build(np.ascontiguousarray(c1.points.T)) - 1
    ref = orc.kd_build(c1.points.T)

    def walk(node, lo, hi):     # oracle pointer tree vs implicit array
        if node < 0:
            assert lo == hi
            return
        m = lo + (hi - lo + 1) // 2 - 1
        assert order[m] == ref.index[node] - 1
        walk(ref.left[node], lo, m)
        walk(ref.right[node], m + 1, hi)

    walk(ref.root, 0, order.size)


@pytest.mark.parametrize("scale", [1.0, 10.0, 0.3])
@pytest.mark.parametrize("k", [6, 8])
def test_kdtree_c1(c1, scale, k):
    """scale 1.0 / 0.3: k-th d2 < 1 -> the reference's prune rule is too eager and the sets are NOT the exact
    kNN; they must still equal the oracle's bit for bit.  scale 10: exact regime."""
    pts = np.asfortranarray(c1.points.T * scale)
    xa, ya = c1.x_axis * scale, c1.y_axis * scale
    want, w = orc.idw_kdtree(pts, c1.data_pts, xa, ya, k, debug=True)
    got, g = ip.idw_kdtree_ex(pts, c1.data_pts, xa, ya, k, debug=True)
    assert np.array_equal(g.nn_index, np.moveaxis(w.nn_index, 0, -1))
    assert np.array_equal(g.nn_dist2, np.moveaxis(w.nn_dist2, 0, -1))
    assert_values(got, want)
    assert g.visits <= w.visits
    # with the extra (state-preserving) prune switched off the traversal visits exactly the reference's nodes
    got2, g2 = ip.idw_kdtree_ex(pts, c1.data_pts, xa, ya, k, debug=True, reference_visits=True, order="C")
    assert g2.visits == w.visits
    assert np.array_equal(g2.nn_index, np.moveaxis(w.nn_index, 0, -1))
    assert_values(got2, want)


def test_kdtree_sentinel_far_points():
    """Neighbours with d2 >= 1e6 are never found (interpolation.f90:159, kd_tree.f90:222-223)."""
    rng = np.random.default_rng(2)
    pts = rng.uniform(0, 5000.0, (2, 300))
    data = rng.normal(size=300)
    x = np.linspace(0, 5000.0, 23); y = np.linspace(0, 5000.0, 17)
    want, w = orc.idw_kdtree(pts, data, x, y, 8, debug=True)
    got, g = ip.idw_kdtree_ex(pts, data, x, y, 8, debug=True)
    assert (w.nn_index == 0).any()
    assert np.array_equal(g.nn_index, np.moveaxis(w.nn_index, 0, -1))
    assert_values(got, want)


# ------------------------------------------------------------------------------- torch device buffers
def test_torch_device_path(c1):
    import torch
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    want = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                         c1.neighbours)
    got = ip.barycentric_interpolation(t(c1.points), t(c1.data_pts), t(c1.x_axis), t(c1.y_axis),
                                       t(c1.simplices), t(c1.neighbours), order="C")
    assert got.is_cuda and tuple(got.shape) == want.shape
    assert_values(got.cpu().numpy(), want)
    back = ip.bilinear(t(c1.x_axis), t(c1.y_axis), got, t(c1.points))
    assert_values(back.cpu().numpy(), orc.bilinear(c1.x_axis, c1.y_axis, want, c1.points))
    e = ip.idw_esrg_nearest(t(c1.points), t(c1.data_pts), t(c1.x_axis), t(c1.y_axis), c1.grid_spac, 6)
    assert_values(e.cpu().numpy(), orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis,
                                                        c1.grid_spac, 6))
    kd = ip.idw_kdtree(t(c1.points).t(), t(c1.data_pts), t(c1.x_axis), t(c1.y_axis), 6)
    assert_values(kd.cpu().numpy(), orc.idw_kdtree(c1.points.T, c1.data_pts, c1.x_axis, c1.y_axis, 6))


# --------------------------------------------------------------- larger, size-independent properties
def test_medium_mesh_matches_oracle_and_properties():
    """200k-point mesh on a 1024^2 grid (the oracle still finishes in seconds)."""
    rng = np.random.default_rng(4)
    n, grid = 200_000, 1024
    pts = rng.uniform(0.0, grid - 1.0, (n, 2))
    simp, nbr = workloads.delaunay(pts)
    data = workloads.sin_mountains(pts[:, 0] * 0.02, pts[:, 1] * 0.02)
    axis = np.arange(float(grid))
    want, w = orc.barycentric_interpolation(pts, data, axis, axis, simp, nbr, debug=True)
    got, g = ip.barycentric_interpolation_ex(pts, data, axis, axis, simp, nbr, debug=True, order="C")
    ins = ~w.mask
    assert np.array_equal(g.mask, w.mask)
    assert np.array_equal(g.tri[ins], w.tri[ins])
    assert_values(got, want)
    # linearity in the data: interp(a*d1 + d2) is affine in the nodal values -> a linear field is reproduced
    lin = 3.0 * pts[:, 0] - 2.0 * pts[:, 1] + 5.0
    out = ip.barycentric_interpolation(pts, lin, axis, axis, simp, nbr, order="C")
    xx, yy = np.meshgrid(axis, axis)
    np.testing.assert_allclose(out[ins], (3.0 * xx - 2.0 * yy + 5.0)[ins], rtol=0, atol=1e-7)
    # bilinear back at the grid nodes themselves is the identity
    nodes = np.stack([xx[:-1, :-1].ravel(), yy[:-1, :-1].ravel()], 1)
    back = ip.bilinear(axis, axis, got, nodes)
    assert np.array_equal(back, got[:-1, :-1].ravel())


# ------------------------------------------------------------- single precision (the shipped numerics)
def test_f32_entry_points_match_f32_oracle(c1):
    """The reference as shipped is single precision: the _f32 entry points must reproduce the f32
    oracle bit for bit (same operations, IEEE single division / sqrt, no FMA contraction)."""
    f = np.float32
    pts, dat, xa, ya = c1.points.astype(f), c1.data_pts.astype(f), c1.x_axis.astype(f), c1.y_axis.astype(f)
    # barycentric: the f32 margin (-1e-10f) and f32 cross products decide the triangles
    want, w = orc.barycentric_interpolation(pts, dat, xa, ya, c1.simplices, c1.neighbours, dtype=f, debug=True)
    got, g = ip.barycentric_interpolation_ex(pts, dat, xa, ya, c1.simplices, c1.neighbours, dtype=f, debug=True)
    assert got.dtype == f and np.array_equal(g.mask, w.mask)
    ins = ~w.mask
    assert np.array_equal(g.tri[ins], w.tri[ins])
    assert np.array_equal(got, want)
    # esrg / k-d tree
    spac = f(c1.grid_spac)
    want_e, we = orc.idw_esrg_nearest(pts, dat, xa, ya, spac, 6, dtype=f, debug=True)
    got_e, ge = ip.idw_esrg_nearest_ex(pts, dat, xa, ya, spac, 6, dtype=f, debug=True)
    assert np.array_equal(ge.nn_index, np.moveaxis(we.nn_index, 0, -1)) and np.array_equal(got_e, want_e)
    want_k, wk = orc.idw_kdtree(pts.T, dat, xa, ya, 6, dtype=f, debug=True)
    got_k, gk = ip.idw_kdtree_ex(pts.T, dat, xa, ya, 6, dtype=f, debug=True)
    assert np.array_equal(gk.nn_index, np.moveaxis(wk.nn_index, 0, -1)) and np.array_equal(got_k, want_k)
    # bilinear
    want_b = orc.bilinear(xa, ya, want_e, pts, dtype=f)
    got_b = ip.bilinear(xa, ya, got_e, pts, dtype=f)
    assert got_b.dtype == f and np.array_equal(got_b, want_b)


def test_f32_barycentric_hull_rows_and_fixups_match_f32_oracle():
    """Single precision on a mesh large enough for what f32 exercises at C4 size: long hull triangles raise the
    rounding-aware ambiguity threshold, hundreds of targets are flagged, the first and last grid rows lie wholly
    outside the hull (no certain column to back up to: the row walk redoes them along the reference's path from
    ind_tri_start, pieces repaired in rounds), flagged targets near column 0 start at ind_tri_start as well.
    Values, mask and triangle ids must equal the f32 oracle bit for bit, in both layouts and in row bands."""
    f = np.float32
    rng = np.random.default_rng(77)
    n_grid = 1200
    pts = rng.uniform(0.0, n_grid - 1.0, (150_000, 2))
    simp, nbr = workloads.delaunay(pts)
    data = workloads.sin_mountains(pts[:, 0] * 0.01, pts[:, 1] * 0.01)
    p32, d32 = pts.astype(f), data.astype(f)
    x32 = np.arange(n_grid, dtype=f); y32 = np.arange(n_grid, dtype=f)
    want, w = orc.barycentric_interpolation(p32, d32, x32, y32, simp, nbr, dtype=f, debug=True)
    assert w.mask[0].all() and w.mask[-1].all()            # rows outside the hull
    ins = ~w.mask
    for order in ("F", "C"):
        got, g = ip.barycentric_interpolation_ex(p32, d32, x32, y32, simp, nbr, dtype=f, debug=True, order=order)
        assert g.fixups > 20
        assert np.array_equal(g.mask, w.mask)
        assert np.array_equal(g.tri[ins], w.tri[ins])
        assert np.array_equal(got, want)
        assert np.array_equal(g.fill_src[w.mask], w.fill_src[w.mask])
    parts = [ip.barycentric_interpolation_ex(p32, d32, x32, y32, simp, nbr, dtype=f, rows=b, halo=24)
             for b in ((0, 400), (400, 1000), (1000, n_grid))]
    assert np.array_equal(np.concatenate(parts, 0), want)


def test_kdtree_exact_prune_extension(c1):
    """exact_prune=True (README to-do of the reference): the exact k nearest neighbours even where the k-th d2 < 1
    makes the reference's rule too eager -- checked against scipy."""
    from scipy.spatial import KDTree
    k = 6
    pts_ip = np.vstack([i.ravel() for i in np.meshgrid(c1.x_axis, c1.y_axis)]).T
    dist, idx = KDTree(c1.points).query(pts_ip, k=k)
    got, g = ip.idw_kdtree_ex(c1.points.T, c1.data_pts, c1.x_axis, c1.y_axis, k, debug=True, exact_prune=True,
                              order="C")
    assert np.array_equal(g.nn_index.reshape(-1, k) - 1, idx)
    ref = ((c1.data_pts[idx] / dist).sum(1) / (1.0 / dist).sum(1)).reshape(got.shape)
    np.testing.assert_allclose(got, ref, rtol=1e-12)
    # ... and it differs from the reference rule on this geometry (~53 % of the targets)
    _, gr = ip.idw_kdtree_ex(c1.points.T, c1.data_pts, c1.x_axis, c1.y_axis, k, debug=True, order="C")
    assert (np.sort(gr.nn_index, -1) != np.sort(g.nn_index, -1)).any(-1).mean() > 0.4


@pytest.mark.parametrize("order", ["C", "F"])
def test_bilinear_sector_load_path(order):
    """Axes whose lengths are multiples of 4 take the 256-bit sector-load kernel (bilinear.cu); it must agree with
    the scalar kernel's oracle bit for bit, including points in every position modulo 4 and out-of-grid points."""
    rng = np.random.default_rng(12)
    x = np.cumsum(rng.uniform(0.5, 1.5, 64)); y = np.cumsum(rng.uniform(0.2, 0.4, 48)) - 3.0
    field = rng.normal(size=(48, 64))
    pts = np.stack([rng.uniform(x[0] - 1, x[-1] + 1, 50021), rng.uniform(y[0] - 1, y[-1] + 1, 50021)], 1)
    pts[:4] = [[x[0], y[0]], [x[-1], y[-1]], [np.nextafter(x[-1], 0), np.nextafter(y[-1], 0)], [np.nan, 1.0]]
    want = orc.bilinear(x, y, field, pts)
    f = np.ascontiguousarray(field) if order == "C" else np.asfortranarray(field)
    p = np.ascontiguousarray(pts) if order == "C" else np.asfortranarray(pts)
    assert_values(ip.bilinear(x, y, f, p), want)


def test_bilinear_morton_order_is_a_spatial_permutation():
    """gi_bilinear_morton_order: a permutation that sorts the points by the Z value of their 16 x 16-node block
    (points outside the grid last); bilinear on the ordered points, un-permuted, equals the direct call."""
    rng = np.random.default_rng(21)
    lx, ly, n = 1024, 768, 300_000
    x = np.arange(float(lx)) * 0.5 + 3.0; y = np.arange(float(ly)) * 0.25 - 10.0
    field = rng.normal(size=(ly, lx))
    pts = np.column_stack([rng.uniform(x[0] - 5, x[-1] + 5, n), rng.uniform(y[0] - 5, y[-1] + 5, n)])
    order = ip.bilinear_morton_order(x, y, pts)
    assert order.dtype == np.int32 and np.array_equal(np.sort(order), np.arange(n))
    ix = np.floor((pts[:, 0] - x[0]) / 0.5); iy = np.floor((pts[:, 1] - y[0]) / 0.25)
    inside = (ix >= 0) & (ix < lx - 1) & (iy >= 0) & (iy < ly - 1)
    def spread(v):
        v = v.astype(np.uint32)
        v = (v | (v << 8)) & 0x00ff00ff; v = (v | (v << 4)) & 0x0f0f0f0f
        v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555
        return v
    key = np.where(inside, spread(np.clip(ix, 0, lx).astype(np.int64) >> 4) | (spread(np.clip(iy, 0, ly).astype(np.int64) >> 4) << 1),
                   np.uint32(1 << 12))
    ks = key[order]
    assert (np.diff(ks.astype(np.int64)) >= 0).all()
    assert (np.diff(order)[np.diff(ks.astype(np.int64)) == 0] > 0).all()         # stable
    direct = ip.bilinear(x, y, field, pts)
    back = np.empty(n)
    back[order] = ip.bilinear(x, y, field, pts[order])
    assert np.array_equal(back, direct)
    # a contiguous eighth of the order is spatially compact: the shard of one of 8 GPUs touches ~1/8 of the field
    part = pts[order[:n // 8]]
    part = part[inside[order[:n // 8]]]
    bbox = np.ptp(part[:, 0]) * np.ptp(part[:, 1])
    assert bbox < 0.3 * np.ptp(x) * np.ptp(y)
    import torch
    d = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
    assert np.array_equal(ip.bilinear_morton_order(d(x), d(y), d(pts)).cpu().numpy(), order)

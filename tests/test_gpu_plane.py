"""GI_BARY_PLANE_VALUES (opt-in): values as the plane through the triangle's nodal values (GPU).

What must hold against the oracle (the restatement of interpolation.f90:298-416):
  * triangle ids, outside-hull mask and fill sources: bit-identical (the locate phase is the same code);
  * values: |got - want| <= TOL * max|data_in| with TOL = 1e-12 in double precision (north_star's tolerance), and
    2e-4 in single precision (the same bound, 16 * 128 * eps, with float's eps);
  * needle triangles (max|edge|^2 > 128 |denom|) keep the reference's expression: bit-identical values there;
  * every path that produces a value (one-shot, bands + halo, fix-ups, hull fill, plans, scattered targets, both
    field layouts, host and device buffers) produces the same bits.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from grid_interpolation_b200 import interpolation as ip  # noqa: E402
import workloads  # noqa: E402
from oracle import oracle as orc  # noqa: E402

TOL64 = 1e-12
TOL32 = 2e-4


def _cond(points, simplices):
    """max|edge|^2 / |denom| per triangle, the quantity plane_of() thresholds at 128 (barycentric.cu)."""
    a = points[simplices - 1]
    x1, y1, x2, y2, x3, y3 = a[:, 0, 0], a[:, 0, 1], a[:, 1, 0], a[:, 1, 1], a[:, 2, 0], a[:, 2, 1]
    denom = (y2 - y3) * (x1 - x3) + (x3 - x2) * (y1 - y3)
    l1 = (x2 - x1) ** 2 + (y2 - y1) ** 2
    l2 = (x3 - x2) ** 2 + (y3 - y2) ** 2
    l3 = (x1 - x3) ** 2 + (y1 - y3) ** 2
    with np.errstate(divide="ignore"):
        return np.maximum(np.maximum(l1, l2), l3) / np.abs(denom)


def _check(got, g, want, w, data, tol, points=None, simplices=None):
    assert np.array_equal(g.mask, w.mask)
    ins = ~w.mask
    assert np.array_equal(g.tri[ins], w.tri[ins])
    assert np.array_equal(g.fill_src, w.fill_src)
    scale = np.abs(data).max()
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)).max() / scale
    assert err <= tol, f"max |diff| / max|data| = {err:.3e} > {tol:.1e}"
    if points is not None:  # needles keep the reference's expression
        needle = _cond(points, simplices) > 130.0
        in_needle = ins & needle[np.where(ins, w.tri, 1) - 1]
        assert np.array_equal(got[in_needle], want[in_needle])
        return int(in_needle.sum()), err
    return 0, err


@pytest.mark.parametrize("order", ["C", "F"])
def test_plane_values_c1(c1, order):
    args = (c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours)
    want, w = orc.barycentric_interpolation(*args, debug=True)
    got, g = ip.barycentric_interpolation_ex(*args, debug=True, order=order, plane_values=True)
    _check(got, g, want, w, c1.data_pts, TOL64, c1.points, c1.simplices)
    exact = ip.barycentric_interpolation_ex(*args, order=order)
    assert np.array_equal(exact, want)                       # the default stays bit-identical
    assert not np.array_equal(got, want)                     # ... and the flag really selects another expression
    # both layouts, bands + halo, device tensors, plan: the same bits
    other = ip.barycentric_interpolation_ex(*args, order="F" if order == "C" else "C", plane_values=True)
    assert np.array_equal(other, got)
    ly = c1.y_axis.size
    parts = [ip.barycentric_interpolation_ex(*args, rows=(r0, r1), halo=4, order=order, plane_values=True)
             for r0, r1 in [(0, 20), (20, 41), (41, 42), (42, ly)]]
    assert np.array_equal(np.concatenate(parts, 0), got)
    import torch
    dev = [torch.as_tensor(np.ascontiguousarray(a), device="cuda:0") for a in args]
    assert np.array_equal(ip.barycentric_interpolation_ex(*dev, order=order, plane_values=True).cpu().numpy(), got)
    plan = ip.BarycentricPlan(c1.points, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours, order=order,
                              plane_values=True)
    assert np.array_equal(plan(c1.data_pts), got)
    other_data = np.cos(c1.points[:, 0]) * 3.0 - c1.points[:, 1]
    assert np.array_equal(plan(other_data), ip.barycentric_interpolation_ex(
        c1.points, other_data, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours, order=order, plane_values=True))
    plan.close() if hasattr(plan, "close") else None


def test_plane_values_targets_on_mesh_edges():
    """Structured lattice: hundreds of targets on edges and vertices go through the fix-up path (publish_target)."""
    gx, gy = np.meshgrid(np.arange(0.0, 12.0), np.arange(0.0, 9.0))
    pts = np.stack([gx.ravel(), gy.ravel()], 1)
    simp, nbr = workloads.delaunay(pts)
    data = np.sin(pts[:, 0]) + 0.3 * pts[:, 1] ** 2
    x = np.arange(-0.5, 12.0, 0.5); y = np.arange(-0.5, 9.0, 0.25)
    want, w = orc.barycentric_interpolation(pts, data, x, y, simp, nbr, debug=True)
    for order in ("C", "F"):
        got, g = ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, debug=True, order=order, plane_values=True)
        assert g.fixups > 100
        _check(got, g, want, w, data, TOL64)
        ref_rows, _ = ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, debug=True, order=order,
                                                      plane_values=True, reference_order=True)
        assert np.array_equal(ref_rows, got)


def test_plane_values_needles_and_large_coordinates():
    """A cloud with nearly collinear triples (needle triangles inside the hull and along it) at coordinates around
    1e6: the needles keep the exact expression, everything else stays inside the tolerance."""
    rng = np.random.default_rng(12)
    base = rng.uniform(0.0, 400.0, (60_000, 2))
    k = 3000
    i = rng.integers(0, base.shape[0], k); j = rng.integers(0, base.shape[0], k)
    tt = rng.uniform(0.2, 0.8, k)[:, None]
    mid = base[i] * tt + base[j] * (1 - tt) + rng.normal(0.0, 1e-4, (k, 2))
    near = base[:k] + rng.normal(0.0, 2e-3, (k, 2))                     # very short edges
    pts = np.concatenate([base, mid, near]) + np.array([1.0e6, -3.0e6])
    simp, nbr = workloads.delaunay(pts)
    data = 40.0 * np.sin(0.05 * pts[:, 0]) + 0.01 * (pts[:, 1] + 3.0e6) + 500.0
    x = np.linspace(pts[:, 0].min() - 3.0, pts[:, 0].max() + 3.0, 1500)
    y = np.linspace(pts[:, 1].min() - 3.0, pts[:, 1].max() + 3.0, 1100)
    want, w = orc.barycentric_interpolation(pts, data, x, y, simp, nbr, debug=True)
    got, g = ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, debug=True, order="C", plane_values=True)
    n_needle, err = _check(got, g, want, w, data, TOL64, pts, simp)
    assert (_cond(pts, simp) > 130.0).sum() > 50
    assert n_needle > 0


def test_plane_values_medium_mesh_f64_f32_and_scattered_targets():
    rng = np.random.default_rng(21)
    pts = rng.uniform(0.0, 2047.0, (400_000, 2))
    simp, nbr = workloads.delaunay(pts)
    data = workloads.sin_mountains(pts[:, 0] * 0.01, pts[:, 1] * 0.01)
    x = np.arange(2048.0); y = np.arange(1024.0) * 2.0
    want, w = orc.barycentric_interpolation(pts, data, x, y, simp, nbr, debug=True)
    got, g = ip.barycentric_interpolation_ex(pts, data, x, y, simp, nbr, debug=True, order="C", plane_values=True)
    _check(got, g, want, w, data, TOL64, pts, simp)
    # single precision: same bound in float's eps, checked against the f32 oracle
    p32, d32, x32, y32 = (a.astype(np.float32) for a in (pts, data, x, y))
    want32, w32 = orc.barycentric_interpolation(p32, d32, x32, y32, simp, nbr, dtype=np.float32, debug=True)
    got32, g32 = ip.barycentric_interpolation_ex(p32, d32, x32, y32, simp, nbr, debug=True, order="C",
                                                 dtype=np.float32, plane_values=True)
    _check(got32, g32, want32, w32, d32, TOL32)
    # scattered targets: the same plane records
    tg = rng.uniform(-20.0, 2070.0, (300_000, 2))
    wp, wd = orc.barycentric_points(pts, data, simp, nbr, tg, debug=True)
    gp, gd = ip.barycentric_points(pts, data, simp, nbr, tg, debug=True, plane_values=True)
    same = gd.tri == wd.tri
    assert same.mean() > 0.999
    assert np.abs(gp[same] - wp[same]).max() <= TOL64 * np.abs(data).max()
    # grid nodes as scattered targets: the grid call's bits inside the hull
    xx, yy = np.meshgrid(x[:256], y[:64])
    nodes = np.stack([xx.ravel(), yy.ravel()], 1)
    at_nodes, gn = ip.barycentric_points(pts, data, simp, nbr, nodes, debug=True, plane_values=True)
    sub = got[:64, :256].ravel(); ins = ~w.mask[:64, :256].ravel()
    same_tri = gn.tri == w.tri[:64, :256].ravel()
    assert np.array_equal(at_nodes[ins & same_tri], sub[ins & same_tri])

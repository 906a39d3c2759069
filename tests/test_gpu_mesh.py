"""Mesh ingest (gi_mesh_reorder): the library's definition restated in numpy and compared bit for
bit, and the property that makes it an ingest step -- the reordered mesh gives the same field."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grid_interpolation_b200 import interpolation as ip  # noqa: E402
from grid_interpolation_b200._lib import GiError  # noqa: E402
import workloads  # noqa: E402
from oracle import oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


def _spread16(v):
    v = v.astype(np.uint32) & np.uint32(0xffff)
    v = (v | (v << np.uint32(8))) & np.uint32(0x00ff00ff)
    v = (v | (v << np.uint32(4))) & np.uint32(0x0f0f0f0f)
    v = (v | (v << np.uint32(2))) & np.uint32(0x33333333)
    v = (v | (v << np.uint32(1))) & np.uint32(0x55555555)
    return v


def _z_value(x, y, box, R):
    """gridinterp.h: q = trunc((v - v_min) * (65535 / (v_max - v_min))) clamped to 0..65535, evaluated in R."""
    x0, y0, x1, y1 = box
    sx = R(65535) / (x1 - x0) if x1 > x0 else R(0)
    sy = R(65535) / (y1 - y0) if y1 > y0 else R(0)
    qx = np.clip(((x - x0) * sx), 0, 65535).astype(np.int64)
    qy = np.clip(((y - y0) * sy), 0, 65535).astype(np.int64)
    return _spread16(qx) | (_spread16(qy) << np.uint32(1))


def mesh_reorder_np(points, simplices, neighbours, reorder_points=True, dtype=np.float64):
    R = np.dtype(dtype).type
    p = np.asarray(points, dtype=dtype)
    x, y = p[:, 0], p[:, 1]
    box = (x.min(), y.min(), x.max(), y.max())
    s = np.asarray(simplices, dtype=np.int64); nb = np.asarray(neighbours, dtype=np.int64)
    xs, ys = x[s - 1], y[s - 1]
    cx = (xs[:, 0] + xs[:, 1] + xs[:, 2]) / R(3)
    cy = (ys[:, 0] + ys[:, 1] + ys[:, 2]) / R(3)
    t_order = np.argsort(_z_value(cx, cy, box, R), kind="stable")
    t_inv = np.empty_like(t_order); t_inv[t_order] = np.arange(t_order.size)
    s_new = s[t_order]
    nb_old = nb[t_order]
    nb_new = np.where(nb_old >= 1, t_inv[np.maximum(nb_old, 1) - 1] + 1, 0)
    p_new, p_order = p, None
    if reorder_points:
        po = np.argsort(_z_value(x, y, box, R), kind="stable")
        p_inv = np.empty_like(po); p_inv[po] = np.arange(po.size)
        s_new = p_inv[s_new - 1] + 1
        p_new, p_order = p[po], (po + 1).astype(np.int32)
    return p_new, p_order, s_new.astype(np.int32), nb_new.astype(np.int32), (t_order + 1).astype(np.int32)


@pytest.fixture(scope="module")
def c1():
    return workloads.c1()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("reorder_points", [True, False])
def test_mesh_reorder_matches_its_definition(c1, dtype, reorder_points):
    pts = c1.points.astype(dtype)
    want = mesh_reorder_np(pts, c1.simplices, c1.neighbours, reorder_points, dtype)
    for lay in ("C", "F"):
        conv = np.ascontiguousarray if lay == "C" else np.asfortranarray
        m = ip.mesh_reorder(conv(pts), conv(c1.simplices), conv(c1.neighbours), reorder_points=reorder_points, dtype=dtype)
        assert np.array_equal(m.simplices, want[2]) and np.array_equal(m.neighbours, want[3])
        assert np.array_equal(m.triangle_order, want[4])
        if reorder_points:
            assert np.array_equal(m.points, want[0]) and np.array_equal(m.point_order, want[1])
            assert m.points.dtype == dtype
        else:
            assert m.point_order is None
    import torch
    d = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
    m = ip.mesh_reorder(d(pts), d(c1.simplices), d(c1.neighbours), reorder_points=reorder_points, dtype=dtype)
    assert np.array_equal(m.simplices.cpu().numpy(), want[2]) and np.array_equal(m.neighbours.cpu().numpy(), want[3])
    assert np.array_equal(m.triangle_order.cpu().numpy(), want[4])
    if reorder_points:
        assert np.array_equal(m.points.cpu().numpy(), want[0]) and np.array_equal(m.point_order.cpu().numpy(), want[1])


def test_reordered_mesh_is_the_same_mesh(c1):
    """Permutations, consistent adjacency, and the same field: values bit for bit, the same outside-hull mask, triangle
    ids that map back through triangle_order; the oracle agrees on the reordered mesh as well."""
    m = ip.mesh_reorder(c1.points, c1.simplices, c1.neighbours)
    nt, npnt = c1.simplices.shape[0], c1.points.shape[0]
    assert np.array_equal(np.sort(m.triangle_order), np.arange(1, nt + 1))
    assert np.array_equal(np.sort(m.point_order), np.arange(1, npnt + 1))
    assert np.array_equal(m.points, c1.points[m.point_order - 1])
    # the same triangles: vertex coordinates row by row
    assert np.array_equal(m.points[m.simplices - 1], c1.points[c1.simplices[m.triangle_order - 1] - 1])
    # adjacency is symmetric: my neighbour lists me
    t_idx, col = np.nonzero(m.neighbours > 0)
    back = m.neighbours[m.neighbours[t_idx, col] - 1]
    assert (back == (t_idx + 1)[:, None]).any(1).all()
    assert (m.neighbours == 0).sum() == (c1.neighbours == 0).sum()
    data = c1.data_pts[m.point_order - 1]
    got0, g0 = ip.barycentric_interpolation_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours,
                                               debug=True)
    got1, g1 = ip.barycentric_interpolation_ex(m.points, data, c1.x_axis, c1.y_axis, m.simplices, m.neighbours, debug=True)
    assert np.array_equal(g0.mask, g1.mask)
    assert np.array_equal(got0, got1)
    ins = ~g0.mask
    assert np.array_equal(m.triangle_order[g1.tri[ins] - 1], g0.tri[ins])
    want1, w1 = orc.barycentric_interpolation(m.points, data, c1.x_axis, c1.y_axis, m.simplices, m.neighbours, debug=True)
    assert np.array_equal(got1, want1) and np.array_equal(g1.tri[ins], w1.tri[ins])


def test_reordered_mesh_is_spatially_ordered():
    """What the ingest step is for: neighbours in space become neighbours in memory."""
    rng = np.random.default_rng(5)
    pts = rng.uniform(0.0, 1000.0, (200_000, 2))
    simp, nbr = workloads.delaunay(pts)
    m = ip.mesh_reorder(pts, simp, nbr)

    def link_distance(nb):   # Qhull's order has short links mostly (median 4) and a heavy tail (mean 25 000)
        t_idx, col = np.nonzero(nb > 0)
        return np.abs(nb[t_idx, col] - 1 - t_idx).mean()

    def vertex_distance(s):  # the caller's point order knows nothing of position
        return np.median(np.abs(s[:, 0] - s[:, 1]))

    assert link_distance(m.neighbours) * 20 < link_distance(nbr)
    assert vertex_distance(m.simplices) * 1000 < vertex_distance(simp)
    # and it is the same mesh to the operator
    data = workloads.sin_mountains(pts[:, 0] * 0.01, pts[:, 1] * 0.01)
    x = np.linspace(-5.0, 1005.0, 700); y = np.linspace(-5.0, 1005.0, 500)
    a = ip.barycentric_interpolation(pts, data, x, y, simp, nbr)
    b = ip.barycentric_interpolation(m.points, data[m.point_order - 1], x, y, m.simplices, m.neighbours)
    assert np.array_equal(a, b)


def test_mesh_reorder_rejects_bad_input(c1):
    bad = c1.simplices.copy(); bad[7, 1] = c1.points.shape[0] + 5
    with pytest.raises(GiError):
        ip.mesh_reorder(c1.points, bad, c1.neighbours)
    with pytest.raises(ValueError):
        ip.mesh_reorder(c1.points, c1.simplices[:, :2], c1.neighbours)
    with pytest.raises(ValueError):
        ip.mesh_reorder(c1.points[:, :1], c1.simplices, c1.neighbours)

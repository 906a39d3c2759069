"""Plans (gi_plan): the search structure / packed mesh is prepared once, every field of nodal values gives the
same result as the reference-shaped one-shot call and as the oracle (GPU)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from grid_interpolation_b200 import interpolation as ip  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def same(a, b):
    assert np.array_equal(np.asarray(a), np.asarray(b))


def fields(c1, n=3):
    rng = np.random.default_rng(7)
    yield c1.data_pts
    for _ in range(n - 1):
        yield rng.normal(size=c1.data_pts.shape)


@pytest.mark.parametrize("order", ["F", "C"])
def test_barycentric_plan(c1, order):
    with ip.BarycentricPlan(c1.points, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours, order=order) as plan:
        for data in fields(c1):
            got = plan(data)
            same(got, orc.barycentric_interpolation(c1.points, data, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours))
            assert got.flags.c_contiguous if order == "C" else got.flags.f_contiguous


def test_barycentric_plan_row_band_with_halo(c1):
    rows, halo = (20, 61), 12
    want = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours)
    with ip.BarycentricPlan(c1.points, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours, rows=rows, halo=halo,
                            order="C") as plan:
        same(plan(c1.data_pts), want[rows[0]:rows[1]])


@pytest.mark.parametrize("k", [6, 8])
def test_kdtree_and_esrg_plans(c1, k):
    with ip.KdTreePlan(c1.points.T, c1.x_axis, c1.y_axis, k) as kd, \
         ip.EsrgPlan(c1.points, c1.x_axis, c1.y_axis, c1.grid_spac, k, order="C") as es:
        for data in fields(c1):
            same(kd(data), orc.idw_kdtree(c1.points.T, data, c1.x_axis, c1.y_axis, k))
            same(es(data), orc.idw_esrg_nearest(c1.points, data, c1.x_axis, c1.y_axis, c1.grid_spac, k))


def test_plans_on_device_tensors_and_banded_host_path(c1):
    import torch
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    plan = ip.EsrgPlan(t(c1.points), t(c1.x_axis), t(c1.y_axis), c1.grid_spac, 6, order="C")
    want = orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 6)
    got = plan(t(c1.data_pts))                      # device in -> device out
    assert got.is_cuda
    same(got.cpu().numpy(), want)
    ip.set_pipeline_band_bytes(1 << 10)             # the same plan from host memory, 16 bands
    try:
        same(plan(c1.data_pts), want)
        with ip.BarycentricPlan(c1.points, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours, order="C") as bp:
            same(bp(c1.data_pts), orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis,
                                                                c1.simplices, c1.neighbours))
    finally:
        ip.set_pipeline_band_bytes(0)
    plan.close()
    with pytest.raises(RuntimeError):
        plan(c1.data_pts)


def test_plan_argument_errors(c1):
    with pytest.raises(ValueError):
        ip.KdTreePlan(c1.points, c1.x_axis, c1.y_axis, 6)          # points must be (2, n)
    with ip.EsrgPlan(c1.points, c1.x_axis, c1.y_axis, c1.grid_spac, 6) as plan:
        with pytest.raises(ValueError):
            plan(c1.data_pts[:-1])
    bad = c1.simplices.copy(); bad[3, 1] = c1.points.shape[0] + 5
    with pytest.raises(RuntimeError, match="out of range"):
        ip.BarycentricPlan(c1.points, c1.x_axis, c1.y_axis, bad, c1.neighbours)

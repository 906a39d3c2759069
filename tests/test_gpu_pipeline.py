"""The copy / compute pipeline of host (numpy) calls: many small bands must give exactly the one-shot result
(GPU).  The band size is a process-wide knob (gi_set_pipeline_band_bytes); every test restores the default."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from grid_interpolation_b200 import interpolation as ip  # noqa: E402
import workloads  # noqa: E402
from oracle import oracle as orc  # noqa: E402


@pytest.fixture
def tiny_bands():
    ip.set_pipeline_band_bytes(1 << 10)   # 1 KB of output per band -> the maximum of 16 bands on every case below
    yield
    ip.set_pipeline_band_bytes(0)


def same(a, b):
    assert np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.parametrize("order", ["C", "F"])
def test_grid_operators_banded_equal_oracle(c1, tiny_bands, order):
    """16 bands along rows (C order) or columns (F order): barycentric incl. hull fill across band seams."""
    want = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours)
    got = ip.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours,
                                       order=order)
    same(got, want)
    assert got.flags.c_contiguous if order == "C" else got.flags.f_contiguous
    same(ip.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 6, order=order),
         orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 6))
    same(ip.idw_kdtree(c1.points.T, c1.data_pts, c1.x_axis, c1.y_axis, 6, order=order),
         orc.idw_kdtree(c1.points.T, c1.data_pts, c1.x_axis, c1.y_axis, 6))


def test_row_band_calls_are_banded_too(c1, tiny_bands):
    rows = (17, 70)
    same(ip.idw_esrg_nearest_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 8, rows=rows, order="C"),
         orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 8)[rows[0]:rows[1]])
    same(ip.idw_kdtree_ex(c1.points.T, c1.data_pts, c1.x_axis, c1.y_axis, 8, rows=rows, order="F"),
         orc.idw_kdtree(c1.points.T, c1.data_pts, c1.x_axis, c1.y_axis, 8)[rows[0]:rows[1]])


@pytest.mark.parametrize("order", ["C", "F"])
def test_barycentric_row_band_with_halo_is_banded_too(c1, tiny_bands, order):
    """multi-GPU style call (row band + halo rows) through the host path: sub-bands include pure halo pieces"""
    want = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours)
    for rows, halo in (((0, 40), 8), ((33, 79), 8), ((20, 45), 40)):
        got = ip.barycentric_interpolation_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                              c1.neighbours, rows=rows, halo=halo, order=order)
        same(got, want[rows[0]:rows[1]])
    with pytest.raises(RuntimeError, match="halo"):      # a halo that is too small is still reported
        ip.barycentric_interpolation_ex(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours,
                                        rows=(0, 2), halo=0, order=order)


def test_fill_deeper_than_a_band_restarts_with_one_band(tiny_bands):
    """A hull fill that has to look further than the neighbouring band (GI_ERR_HALO inside) must not leak out:
    the call starts over with a single band and returns the oracle's field."""
    rng = np.random.default_rng(11)
    from scipy.spatial import Delaunay
    pts = rng.uniform(40.0, 60.0, (300, 2))          # a small mesh in the middle of a 200 x 180 grid
    tri = Delaunay(pts)
    simp = (tri.simplices + 1).astype(np.int32); nbr = (tri.neighbors + 1).astype(np.int32)
    data = np.sin(pts[:, 0] * 0.3) + pts[:, 1] * 0.01
    x = np.linspace(0.0, 100.0, 200); y = np.linspace(0.0, 100.0, 180)
    want, w = orc.barycentric_interpolation(pts, data, x, y, simp, nbr, debug=True)
    assert w.fill_level_max > 64                      # deeper than two 32-row bands
    for order in ("C", "F"):
        same(ip.barycentric_interpolation(pts, data, x, y, simp, nbr, order=order), want)


@pytest.mark.parametrize("order", ["C", "F"])
def test_bilinear_chunked_equals_oracle(tiny_bands, order):
    rng = np.random.default_rng(5)
    x = np.cumsum(rng.uniform(0.5, 1.5, 64)); y = np.cumsum(rng.uniform(0.5, 1.5, 48))
    field = rng.normal(size=(48, 64))
    n = 70_001                                        # 16 chunks of 4096-aligned size + a ragged tail
    pts = np.column_stack([rng.uniform(x[0] - 1, x[-1] + 1, n), rng.uniform(y[0] - 1, y[-1] + 1, n)])
    pts = np.asarray(pts, order=order)
    same(ip.bilinear(x, y, field, pts), orc.bilinear(x, y, field, pts))


def test_default_band_size_is_restored(c1):
    """(sanity) with the default 32 MB bands the small cases run as one band and still match."""
    same(ip.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 6),
         orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 6))


def test_host_calls_from_concurrent_threads(c1):
    """ctypes releases the GIL (like f2py's `threadsafe`, interpolation.f90:30,118,213,311): four host threads run
    the four operators at once, each on its own thread-local pipeline, and every result matches the oracle."""
    import threading
    want_b = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours)
    want_e = orc.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 6)
    want_k = orc.idw_kdtree(c1.points.T, c1.data_pts, c1.x_axis, c1.y_axis, 6)
    want_l = orc.bilinear(c1.x_axis, c1.y_axis, want_e, c1.points)
    jobs = {
        "bary": (lambda: ip.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices,
                                                      c1.neighbours), want_b),
        "esrg": (lambda: ip.idw_esrg_nearest(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.grid_spac, 6), want_e),
        "kd": (lambda: ip.idw_kdtree(c1.points.T, c1.data_pts, c1.x_axis, c1.y_axis, 6), want_k),
        "bil": (lambda: ip.bilinear(c1.x_axis, c1.y_axis, want_e, c1.points), want_l),
    }
    errors = []

    def worker(name):
        fn, want = jobs[name]
        try:
            for _ in range(25):
                if not np.array_equal(fn(), want):
                    errors.append(f"{name}: result differs")
                    return
        except Exception as e:  # noqa: BLE001
            errors.append(f"{name}: {e!r}")

    threads = [threading.Thread(target=worker, args=(n,)) for n in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors

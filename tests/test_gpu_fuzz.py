"""Differential fuzz of the CUDA library against the oracle (GPU): random small problems of every flavour the
interpreted Fortran source was fuzzed with against the oracle -- both precisions, odd spacings, clustered points,
points on grid nodes and cell borders (duplicates, exact ties), points far outside the grid, rounded coordinates,
k from 1 to num_points -- all four entry points, bit for bit."""
import numpy as np
import pytest
from scipy.spatial import Delaunay

pytestmark = pytest.mark.gpu

from grid_interpolation_b200 import interpolation as ip  # noqa: E402
from grid_interpolation_b200._lib import GiError, GI_ERR_HALO  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def same(a, b, what):
    assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True), what


def dev(a, R):
    import torch
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a).astype(R) if np.asarray(a).dtype.kind == "f" else a),
                           device="cuda:0")


def api_variant(vr, R, want, tag, call, plan, ly, halo=None):
    """The same problem through another door of the API, picked at random: C-ordered output, a plan, device tensors
    in and out, a row band (with a halo that covers the grid where the operator needs one)."""
    v = int(vr.integers(0, 4))
    if v == 0:
        same(call(order="C"), want, "order=C " + tag)
    elif v == 1:
        with plan() as pl:
            same(pl(pl._fuzz_data), want, "plan " + tag)
    elif v == 2:
        same(call(device=True).cpu().numpy(), want, "device tensors " + tag)
    else:
        r0 = int(vr.integers(0, ly)); r1 = int(vr.integers(r0, ly + 1))
        kw = {} if halo is None else {"halo": halo}
        same(call(rows=(r0, r1), **kw), want[r0:r1], f"rows ({r0},{r1}) " + tag)


def run_trials(seed, trials, n_max, lx_max, ly_max):
    rng = np.random.default_rng(1000 + seed)
    for trial in range(trials):
        R = np.float64 if rng.uniform() < 0.6 else np.float32
        n = int(rng.integers(3, n_max))
        mode = int(rng.integers(0, 6))
        lx = int(rng.integers(2, lx_max)); ly = int(rng.integers(2, ly_max))
        spac = float(rng.choice([1.0, 0.5, 2.0, 0.37]))
        x = (rng.uniform(-5, 5) + spac * np.arange(lx)).astype(R); y = (rng.uniform(-5, 5) + spac * np.arange(ly)).astype(R)
        if rng.uniform() < 0.3:   # axes that are not exactly x(1) + i*spac: the ring windows no longer match the radii
            x = (x + rng.uniform(-0.2, 0.2, lx).astype(R) * R(spac)).astype(R)
            y = (y + rng.uniform(-0.2, 0.2, ly).astype(R) * R(spac)).astype(R)
        ex = float(x[-1] - x[0]); ey = float(y[-1] - y[0])
        if mode == 0:
            pts = np.column_stack([rng.uniform(x[0] - spac, x[-1] + spac, n), rng.uniform(y[0] - spac, y[-1] + spac, n)])
        elif mode == 1:       # clustered: hull well inside the grid, deep hull fill
            pts = np.column_stack([rng.normal(x[0] + ex / 2, ex / 8, n), rng.normal(y[0] + ey / 2, ey / 8, n)])
        elif mode == 2:       # on grid nodes and cell borders: duplicates, exact ties, zero distances
            pts = np.column_stack([rng.choice(x, n) + rng.choice([0, 0, spac / 2], n),
                                   rng.choice(y, n) + rng.choice([0, 0, spac / 2], n)])
        elif mode == 3:       # mostly far outside the grid
            pts = np.column_stack([rng.uniform(x[0] - 3 * ex, x[-1] + 3 * ex, n), rng.uniform(y[0] - 3 * ey, y[-1] + 3 * ey, n)])
        elif mode == 5:       # unique lattice nodes / half steps: targets ON vertices and edges (margin band, path dependence)
            gx, gy = np.meshgrid(x[0] + 0.5 * spac * np.arange(2 * lx - 1), y[0] + 0.5 * spac * np.arange(2 * ly - 1))
            allp = np.column_stack([gx.ravel(), gy.ravel()])
            pts = allp[rng.permutation(len(allp))[:max(min(n, len(allp)), 4)]]; n = len(pts)
        else:                 # rounded coordinates: coordinate ties
            pts = np.round(np.column_stack([rng.uniform(x[0], x[-1], n), rng.uniform(y[0], y[-1], n)]), 1)
        if rng.uniform() < 0.3:   # another length scale / offset: the walk's margin 1e-10 and the k-d sentinel 1e6 are absolute
            sc = float(rng.choice([1.0e3, 1.0e-3, 37.0])); off = rng.uniform(-1.0, 1.0, 2) * 1.0e3 * sc
            x = (x.astype(np.float64) * sc + off[0]).astype(R); y = (y.astype(np.float64) * sc + off[1]).astype(R)
            pts = pts * sc + off; spac = float(R(spac * sc))
        pts = pts.astype(R).astype(np.float64)
        dat = rng.normal(size=n).astype(R)
        k = int(rng.integers(1, min(n, 9) + 1)) if rng.uniform() < 0.8 else min(n, 32)
        tag = f"seed {seed} trial {trial}: {R.__name__} mode {mode} n {n} k {k} grid {lx}x{ly} spac {spac}"

        want, w = orc.idw_kdtree(pts.T, dat, x, y, k, dtype=R, debug=True)
        got, g = ip.idw_kdtree_ex(pts.T, dat, x, y, k, debug=True, dtype=R)
        same(g.nn_index, np.moveaxis(w.nn_index, 0, -1), "kd ids " + tag); same(got, want, "kd " + tag)
        vr = np.random.default_rng(77_000 + 1000 * seed + trial)

        def kd_call(device=False, **kw):
            a = (dev(pts.T, R), dev(dat, R), dev(x, R), dev(y, R)) if device else (pts.T, dat, x, y)
            return ip.idw_kdtree_ex(*a, k, dtype=R, **kw)

        def kd_plan():
            pl = ip.KdTreePlan(pts.T, x, y, k, dtype=R); pl._fuzz_data = dat
            return pl
        api_variant(vr, R, want, "kd " + tag, kd_call, kd_plan, ly)

        want, w = orc.idw_esrg_nearest(pts, dat, x, y, spac, k, dtype=R, debug=True)
        got, g = ip.idw_esrg_nearest_ex(pts, dat, x, y, spac, k, debug=True, dtype=R)
        same(g.nn_index, np.moveaxis(w.nn_index, 0, -1), "esrg ids " + tag); same(got, want, "esrg " + tag)

        def es_call(device=False, **kw):
            a = (dev(pts, R), dev(dat, R), dev(x, R), dev(y, R)) if device else (pts, dat, x, y)
            return ip.idw_esrg_nearest_ex(*a, spac, k, dtype=R, **kw)

        def es_plan():
            pl = ip.EsrgPlan(pts, x, y, spac, k, dtype=R); pl._fuzz_data = dat
            return pl
        api_variant(vr, R, want, "esrg " + tag, es_call, es_plan, ly)

        if mode != 2:
            tri = Delaunay(pts)
            simp = (tri.simplices + 1).astype(np.int32); nbr = (tri.neighbors + 1).astype(np.int32)
            # (fewer triangles than points -- tiny or duplicated inputs: the reference's start-triangle loop reads out
            #  of bounds, interpolation.f90:347; oracle and library both stop at the last triangle)
            try:
                want = orc.barycentric_interpolation(pts, dat, x, y, simp, nbr, dtype=R)
            except RuntimeError:      # no grid node inside the hull: the reference's fill never ends -- both report
                with pytest.raises(RuntimeError):
                    ip.barycentric_interpolation(pts, dat, x, y, simp, nbr, dtype=R)
            else:
                same(ip.barycentric_interpolation(pts, dat, x, y, simp, nbr, dtype=R), want, "barycentric " + tag)

                def ba_call(device=False, **kw):
                    a = ((dev(pts, R), dev(dat, R), dev(x, R), dev(y, R), dev(simp, R), dev(nbr, R)) if device
                         else (pts, dat, x, y, simp, nbr))
                    return ip.barycentric_interpolation_ex(*a, dtype=R, **kw)

                def ba_plan():
                    pl = ip.BarycentricPlan(pts, x, y, simp, nbr, dtype=R); pl._fuzz_data = dat
                    return pl
                api_variant(vr, R, want, "barycentric " + tag, ba_call, ba_plan, ly, halo=ly)
                if ly >= 230:
                    # a narrow band with a tight halo: the host path packs only the triangles near the band
                    # (GI_BARY_BAND_PACK, retried with the whole mesh by itself when a walk -- or the reference's
                    # start triangle -- needs more); a halo that turns out too small is reported, not guessed
                    r0 = int(vr.integers(0, ly - 16)); r1 = r0 + int(vr.integers(1, 17))
                    for halo in (6, 40, ly):
                        try:
                            part = ba_call(rows=(r0, r1), halo=halo)
                        except GiError as e:
                            assert e.status == GI_ERR_HALO and halo < ly, f"band ({r0},{r1}) halo {halo}: {e} " + tag
                            continue
                        same(part, want[r0:r1], f"band ({r0},{r1}) halo {halo} " + tag)
                        break

        fld = np.asfortranarray(rng.normal(size=(ly, lx)).astype(R))
        q = np.column_stack([rng.uniform(x[0], x[-1] - 1e-3 * (x[-1] - x[0]), 60),
                             rng.uniform(y[0], y[-1] - 1e-3 * (y[-1] - y[0]), 60)])
        q[:10, 0] = rng.choice(x[:-1], 10); q[5:15, 1] = rng.choice(y[:-1], 10)
        q = q.astype(R)
        same(ip.bilinear(x, y, fld, q, dtype=R), orc.bilinear(x, y, fld, q, dtype=R), "bilinear " + tag)


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_small(seed):
    run_trials(seed, 12, 400, 40, 30)


@pytest.mark.parametrize("seed", range(100, 108))
def test_fuzz_medium(seed):
    """tens of thousands of points, grids of a few hundred nodes per side: several CTAs / tiles / sort tiles per call"""
    run_trials(seed, 3, 30_000, 300, 250)


@pytest.mark.parametrize("seed", range(300, 308))
def test_fuzz_tall_grids_narrow_bands(seed):
    """grids of several hundred rows: the narrow-band trial of run_trials (band-limited pack + tight halo) takes part"""
    run_trials(seed, 2, 20_000, 160, 640)

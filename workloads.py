"""Seeded synthetic workloads C1..C5 (BASELINE.md / SURVEY.md section 8d).

Host-side input generation only (numpy / scipy.spatial.Delaunay for the mesh --
mesh ingest is outside the hot path).  Used by the
tests and by bench.py; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np


def sin_mountains(x, y, omega=0.6, amplitude=25.0):
    """test_interpolation.py:38-40."""
    return amplitude * np.sin(omega * x + y) + amplitude


def delaunay(points):
    """scipy/Qhull Delaunay -> (simplices, neighbours) int32, 1-based, 0 = no neighbour, CCW.

    Mirrors test_interpolation.py:64-67,199-201 (``simplices + 1`` / ``neighbors + 1``).
    """
    from scipy.spatial import Delaunay
    tri = Delaunay(points)
    simplices = (tri.simplices + 1).astype(np.int32)
    neighbours = (tri.neighbors + 1).astype(np.int32)
    return simplices, neighbours


def c1(num_points=5000, seed=33):
    """test_interpolation.py:54-93 with the "small" 5 000-point mesh (README.md:32)."""
    points = np.empty((num_points, 2))
    np.random.seed(seed)
    points[:, 0] = np.random.uniform(-0.33, 20.57, num_points)
    points[:, 1] = np.random.uniform(-2.47, 9.89, num_points)
    simplices, neighbours = delaunay(points)
    data_pts = sin_mountains(points[:, 0], points[:, 1])
    vertices = points[simplices - 1]
    temp = np.concatenate((vertices, np.ones(vertices.shape[:2] + (1,))), axis=2)
    area_tri = 0.5 * np.linalg.det(temp)
    grid_spac = np.sqrt(area_tri.sum() / simplices.shape[0])
    safety_multi = 1.005
    x_ext = points[:, 0].max() - points[:, 0].min()
    y_ext = points[:, 1].max() - points[:, 1].min()
    len_x = int(np.ceil(x_ext * safety_multi / grid_spac))
    len_y = int(np.ceil(y_ext * safety_multi / grid_spac))
    x_add = ((len_x * grid_spac) - x_ext) / 2.0
    x_axis = np.linspace(points[:, 0].min() - x_add, points[:, 0].max() + x_add, len_x + 1)
    y_add = ((len_y * grid_spac) - y_ext) / 2.0
    y_axis = np.linspace(points[:, 1].min() - y_add, points[:, 1].max() + y_add, len_y + 1)
    return SimpleNamespace(name="C1", points=points, data_pts=data_pts, simplices=simplices,
                           neighbours=neighbours, x_axis=x_axis, y_axis=y_axis, grid_spac=float(grid_spac),
                           num_nn=6, area_min=float(area_tri.min()))


def digest(w):
    """SHA-256 of a C1-style workload's inputs.  Golden fixtures that store outputs only keep this beside them; the
    tests rebuild the seeded inputs and compare digests first, so an input mismatch cannot pass for a parity failure."""
    import hashlib
    h = hashlib.sha256()
    for a in (w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours):
        a = np.asarray(a)
        h.update(np.ascontiguousarray(a.astype(np.int64) if a.dtype.kind in "iu" else a).tobytes())
    return h.hexdigest()


def _z(x, y, extent):
    s = 20.9 / extent
    return sin_mountains(x * s, y * s)


def scattered(n, grid, seed, lo=0.0, hi=None, k=8):
    """C2 / C3 style: n uniform points over a ``grid`` x ``grid`` unit-spaced regular grid."""
    hi = float(grid - 1) if hi is None else hi
    rng = np.random.default_rng(seed)
    points = rng.uniform(lo, hi, size=(n, 2))
    axis = np.arange(float(grid))
    return SimpleNamespace(points=points, data_pts=_z(points[:, 0], points[:, 1], grid - 1.0),
                           x_axis=axis, y_axis=axis.copy(), grid_spac=1.0, num_nn=k)


def c2(n=1_000_000, grid=4096):
    w = scattered(n, grid, seed=2); w.name = "C2"; return w


def c3(n=10_000_000, grid=8192):
    w = scattered(n, grid, seed=3, lo=-0.5, hi=grid - 0.5); w.name = "C3"; return w


def _default_cache():
    """<repo>/.gi_cache when it exists (git-ignored; it travels with a gpurun snapshot, so the one-minute Qhull run
    is not repeated on every fresh box), else /tmp/gi_cache."""
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), ".gi_cache")
    return here if os.path.isdir(here) else "/tmp/gi_cache"


def c4(n=5_000_000, grid=16384, cache_dir=None, seed=4):
    """Barycentric headline config: n uniform points -> Delaunay (about 2n triangles) -> grid x grid.

    The Qhull triangulation of 5 M points takes about a minute and 4 GB; it is cached as .npy files under
    ``cache_dir`` (default $GI_CACHE or /tmp/gi_cache) so that the two bench arms and all ranks share it.
    """
    cache_dir = cache_dir or os.environ.get("GI_CACHE") or _default_cache()
    tag = f"c4_n{n}_g{grid}_s{seed}"
    fs = os.path.join(cache_dir, tag + "_simplices.npy")
    fn = os.path.join(cache_dir, tag + "_neighbours.npy")
    rng = np.random.default_rng(seed)
    points = rng.uniform(0.0, grid - 1.0, size=(n, 2))
    if os.path.exists(fs) and os.path.exists(fn):
        simplices, neighbours = np.load(fs), np.load(fn)
    else:
        simplices, neighbours = delaunay(points)
        try:
            os.makedirs(cache_dir, exist_ok=True)
            for f, a in ((fs, simplices), (fn, neighbours)):
                tmp = f + f".tmp{os.getpid()}.npy"
                np.save(tmp, a)
                os.replace(tmp, f)
        except OSError:
            pass
    axis = np.arange(float(grid))
    return SimpleNamespace(name="C4", points=points, data_pts=_z(points[:, 0], points[:, 1], grid - 1.0),
                           simplices=simplices, neighbours=neighbours, x_axis=axis, y_axis=axis.copy(),
                           grid_spac=1.0)


def c5(n=100_000_000, grid=16384, seed=5):
    """Bilinear: analytic field on the grid x grid regular grid -> n scattered points strictly inside."""
    rng = np.random.default_rng(seed)
    axis = np.arange(float(grid))
    px = rng.uniform(0.0, grid - 1.0, size=n)
    py = rng.uniform(0.0, grid - 1.0, size=n)
    np.minimum(px, np.nextafter(grid - 1.0, 0.0), out=px)
    np.minimum(py, np.nextafter(grid - 1.0, 0.0), out=py)
    s = 20.9 / (grid - 1.0)
    field = sin_mountains((axis * s)[None, :], (axis * s)[:, None])
    return SimpleNamespace(name="C5", px=px, py=py, field=field, x_axis=axis, y_axis=axis.copy())

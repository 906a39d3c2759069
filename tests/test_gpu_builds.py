"""The search-structure builds on their own (GPU): radix-sort based k-d tree build and esrg cell binning at sizes
around the sort's tile / warp boundaries, with duplicate keys, empty regions and clamped points."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from grid_interpolation_b200 import interpolation as ip  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def same(a, b):
    assert np.array_equal(np.asarray(a), np.asarray(b))


def check_tree(points_2n, faithful=False):
    order = ip.kd_build(np.ascontiguousarray(points_2n), faithful=faithful) - 1
    ref = orc.kd_build(points_2n)
    stack = [(ref.root, 0, order.size)]
    while stack:                                   # oracle pointer tree vs implicit array (kd_tree.f90:54-68)
        node, lo, hi = stack.pop()
        if node < 0:
            assert lo == hi
            continue
        m = lo + (hi - lo + 1) // 2 - 1
        assert order[m] == ref.index[node] - 1
        stack.append((ref.left[node], lo, m)); stack.append((ref.right[node], m + 1, hi))
    assert sorted(order.tolist()) == list(range(order.size))


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 511, 513, 4095, 4096, 4097, 8193, 100_003])
def test_kd_build_sizes(n):
    rng = np.random.default_rng(n)
    check_tree(rng.uniform(-50.0, 1000.0, (2, n)))


def test_kd_build_negative_and_tiny_coordinates():
    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.normal(0.0, 1e-200, (2, 700)), rng.normal(0.0, 1e200, (2, 700)),
                          rng.uniform(-1.0, 1.0, (2, 700))], axis=1)
    check_tree(pts)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 33, 513, 4097, 100_003])
def test_kd_build_quickselect_passes_on_distinct_coordinates(n):
    """the build that evaluates the reference's Lomuto passes in parallel gives the same (canonical) tree"""
    rng = np.random.default_rng(n)
    check_tree(rng.uniform(-50.0, 1000.0, (2, n)), faithful=True)


@pytest.mark.parametrize("case", ["lattice", "few_values", "all_equal", "one_axis_constant", "duplicates", "zeros",
                                  "float32_cast"])
def test_kd_build_tied_coordinates(case):
    """Tied coordinates: the reference's tree depends on its quickselect's swap history (equal keys stay right of the
    pivot, kd_tree.f90:129); the library detects the ties and reproduces the passes -- node by node the oracle's tree."""
    rng = np.random.default_rng(5)
    if case == "lattice":
        gx, gy = np.meshgrid(np.arange(0, 61, 2.0), np.arange(0, 45, 2.0))
        pts = np.stack([gx.ravel(), gy.ravel()])[:, rng.permutation(gx.size)]
    elif case == "few_values":
        pts = rng.integers(0, 5, (2, 3000)).astype(np.float64)
    elif case == "all_equal":
        pts = np.full((2, 777), 3.25)
    elif case == "one_axis_constant":
        pts = np.stack([rng.uniform(0, 9, 2500), np.full(2500, -1.5)])
    elif case == "duplicates":
        base = rng.uniform(0.0, 100.0, (2, 5000))
        pts = np.concatenate([base, base[:, rng.integers(0, 5000, 1500)]], axis=1)[:, rng.permutation(6500)]
    elif case == "zeros":
        pts = np.where(rng.uniform(size=(2, 900)) < 0.5, 0.0, -0.0) * np.where(rng.uniform(size=(2, 900)) < 0.1, 1.0, 1.0)
        pts[:, ::7] = rng.normal(size=pts[:, ::7].shape)
    else:
        pts = rng.uniform(0.0, 4095.0, (2, 200_000)).astype(np.float32).astype(np.float64)
    check_tree(pts)


@pytest.mark.parametrize("n,lx,ly", [(1, 5, 4), (40, 3, 3), (4097, 64, 48), (20_000, 300, 200), (9_000, 1500, 1100)])
def test_esrg_binning_sizes(n, lx, ly):
    """dense (many points per cell), sparse (long empty stretches of the table) and clamped points"""
    rng = np.random.default_rng(n + lx)
    x = np.arange(float(lx)); y = np.arange(float(ly))
    pts = np.column_stack([rng.uniform(-3.0, lx + 2.0, n), rng.uniform(-3.0, ly + 2.0, n)])
    if n > 1000:                                   # a tight cluster and a big hole
        pts[: n // 3] = rng.normal([lx * 0.8, ly * 0.2], 0.7, (n // 3, 2))
    idx, ptr = ip.assign_points_to_cells(pts, x, y, 1.0)
    oi, op, _ = orc.assign_points_to_cells(pts, x, y, 1.0)
    same(idx, oi)
    same(ptr[:-1].reshape(ly, lx) + 1, op)
    assert ptr[-1] == n


@pytest.mark.parametrize("switch", ["GI_KD_NO_SMEM_FINISH", "GI_KD_NO_SERIAL_FINISH", "GI_KD_NO_COOP",
                                    "GI_KD_NO_COOP,GI_KD_NO_FUSED_SCAN"])
def test_kd_build_alternative_paths(switch, tmp_path):
    """The A/B switches of the k-d builds (launch-per-level canonical build, rounds down to the leaves, host-driven
    rounds, six-launch rounds) are read once per process: each is exercised in a child process on a random set (distinct
    coordinates and a forced faithful build), a lattice (ties) and a set large enough for the host-driven rounds, and
    must give the oracle's tree node by node."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import test_gpu_builds as t
        rng = np.random.default_rng(3)
        gx, gy = np.meshgrid(np.arange(60.0), np.arange(45.0))
        lattice = np.stack([gx.ravel(), gy.ravel()])[:, rng.permutation(2700)]
        t.check_tree(rng.uniform(0, 100, (2, 5000)))
        t.check_tree(rng.uniform(0, 100, (2, 5000)), faithful=True)
        t.check_tree(lattice)
        t.check_tree(np.round(rng.uniform(0, 50, (2, 200_000)), 2))
        print("ok")
    """) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    for name in switch.split(","):
        env[name] = "1"
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr

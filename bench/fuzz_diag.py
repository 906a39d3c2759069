"""Diagnostic for a failing fuzz case: regenerate (seed, trial) of tests/test_gpu_fuzz.py and show where esrg differs."""
import sys
import numpy as np
from grid_interpolation_b200 import interpolation as ip
from oracle import oracle as orc
seed, want_trial = int(sys.argv[1]), int(sys.argv[2])
n_max, lx_max, ly_max = (30_000, 300, 250) if seed >= 100 else (400, 40, 30)
rng = np.random.default_rng(1000 + seed)
for trial in range(want_trial + 1):
    R = np.float64 if rng.uniform() < 0.6 else np.float32
    n = int(rng.integers(3, n_max)); mode = int(rng.integers(0, 5))
    lx = int(rng.integers(2, lx_max)); ly = int(rng.integers(2, ly_max))
    spac = float(rng.choice([1.0, 0.5, 2.0, 0.37]))
    x = (rng.uniform(-5, 5) + spac * np.arange(lx)).astype(R); y = (rng.uniform(-5, 5) + spac * np.arange(ly)).astype(R)
    ex = float(x[-1] - x[0]); ey = float(y[-1] - y[0])
    if mode == 0: pts = np.column_stack([rng.uniform(x[0] - spac, x[-1] + spac, n), rng.uniform(y[0] - spac, y[-1] + spac, n)])
    elif mode == 1: pts = np.column_stack([rng.normal(x[0] + ex / 2, ex / 8, n), rng.normal(y[0] + ey / 2, ey / 8, n)])
    elif mode == 2: pts = np.column_stack([rng.choice(x, n) + rng.choice([0, 0, spac / 2], n), rng.choice(y, n) + rng.choice([0, 0, spac / 2], n)])
    elif mode == 3: pts = np.column_stack([rng.uniform(x[0] - 3 * ex, x[-1] + 3 * ex, n), rng.uniform(y[0] - 3 * ey, y[-1] + 3 * ey, n)])
    else: pts = np.round(np.column_stack([rng.uniform(x[0], x[-1], n), rng.uniform(y[0], y[-1], n)]), 1)
    pts = pts.astype(R).astype(np.float64)
    dat = rng.normal(size=n).astype(R)
    k = int(rng.integers(1, min(n, 9) + 1)) if rng.uniform() < 0.8 else min(n, 32)
    if trial < want_trial:   # consume the rest of the trial's random stream like the test does
        rng.normal(size=(ly, lx)); rng.uniform(size=60); rng.uniform(size=60); rng.choice(x[:-1], 10); rng.choice(y[:-1], 10)
print(R.__name__, "mode", mode, "n", n, "k", k, lx, ly, spac)
want, w = orc.idw_esrg_nearest(pts, dat, x, y, spac, k, dtype=R, debug=True)
wi = np.moveaxis(w.nn_index, 0, -1)
iop, ptr = ip.assign_points_to_cells(pts.astype(R), x, y, spac) if R is np.float64 else (None, None)
for ro in (False, True):
    got, g = ip.idw_esrg_nearest_ex(pts, dat, x, y, spac, k, debug=True, dtype=R, reference_order=ro)
    bad = np.argwhere((g.nn_index != wi).any(-1))
    print("reference_order", ro, "targets differing:", len(bad), "values differing:", int((got != want).sum()),
          "exact-order targets", getattr(g, "exact_targets", None), "level_max", g.level_max, w.level_max)
    for iy, ix in bad[:5]:
        print("  target", ix, iy, "got", g.nn_index[iy, ix], np.sqrt(g.nn_dist2[iy, ix]), "want", wi[iy, ix], w.nn_dist[:, iy, ix])
        for i in set(g.nn_index[iy, ix].tolist() + wi[iy, ix].tolist()):
            print("     id", i, "xy", pts[i - 1], "target xy", x[ix], y[iy])

"""Diagnostic: regenerate (seed, trial) of tests/test_gpu_fuzz.py and show where barycentric differs per output order."""
import sys
import numpy as np
from scipy.spatial import Delaunay
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from grid_interpolation_b200 import interpolation as ip
from oracle import oracle as orc
import test_gpu_fuzz as t

seed, want_trial = int(sys.argv[1]), int(sys.argv[2])
captured = {}
real = orc.barycentric_interpolation
count = {"n": -1}
def spy(pts, dat, x, y, simp, nbr, dtype=np.float64, **kw):
    captured["last"] = (pts, dat, x, y, simp, nbr, dtype)
    return real(pts, dat, x, y, simp, nbr, dtype=dtype, **kw)
orc.barycentric_interpolation = spy
try:
    t.run_trials(seed, want_trial + 1, 400, 40, 30)
except AssertionError as e:
    print("assertion:", str(e)[:150])
orc.barycentric_interpolation = real
pts, dat, x, y, simp, nbr, R = captured["last"]
print("case:", np.dtype(R).name, "n", len(pts), "grid", len(x), "x", len(y), "triangles", len(simp))
want, w = real(pts, dat, x, y, simp, nbr, dtype=R, debug=True)
for order in ("F", "C"):
    for ro in (False, True):
        got, g = ip.barycentric_interpolation_ex(pts, dat, x, y, simp, nbr, dtype=R, debug=True, order=order, reference_order=ro)
        bad = np.argwhere(got != want)
        badt = np.argwhere((g.tri != w.tri) & ~w.mask)
        print(f"order {order} reference_order {ro}: values differ {len(bad)}, tri differ {len(badt)}, mask differ {(g.mask != w.mask).sum()}, fixups {g.fixups}")
        for iy, ix in badt[:6]:
            print("   target", ix, iy, (float(x[ix]), float(y[iy])), "got tri", g.tri[iy, ix], "want", w.tri[iy, ix], "values", got[iy, ix], want[iy, ix])

"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol that
include/gridinterp.h declares, the ctypes table covers exactly those symbols, argument errors surface as
f2py-style ValueErrors before any device work, and compute calls fail loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from grid_interpolation_b200 import _lib
from grid_interpolation_b200 import interpolation as ip

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gridinterp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    names = header_symbols()
    assert len(names) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in gridinterp.h but not exported"
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib().gi_version() == 200


def test_drop_in_module_name():
    from interpolation import interpolation as ip_fortran  # test_interpolation.py:27
    for fn in ("bilinear", "idw_kdtree", "idw_esrg_nearest", "barycentric_interpolation"):
        assert callable(getattr(ip_fortran, fn))


def test_signatures_match_f2py_kwlists():
    import inspect
    want = {
        "bilinear": ["x_axis", "y_axis", "data_in", "points"],
        "idw_kdtree": ["points", "data_in", "x_axis", "y_axis", "num_neighbours"],
        "idw_esrg_nearest": ["points", "data_in", "x_axis", "y_axis", "grid_spac", "num_neighbours"],
        "barycentric_interpolation": ["points", "data_in", "x_axis", "y_axis", "simplices", "neighbours"],
    }
    for name, args in want.items():
        params = [p.name for p in inspect.signature(getattr(ip, name)).parameters.values()
                  if p.kind == p.POSITIONAL_OR_KEYWORD]
        assert params == args


def test_shape_errors_raise_value_error_like_f2py():
    x, y = np.arange(5.0), np.arange(4.0)
    pts = np.random.rand(10, 2)
    with pytest.raises(ValueError):
        ip.bilinear(x, y, np.zeros((5, 4)), pts)                      # data_in must be (len_y, len_x)
    with pytest.raises(ValueError):
        ip.idw_kdtree(pts, np.zeros(10), x, y, 3)                     # points must be (2, n)
    with pytest.raises(ValueError):
        ip.idw_esrg_nearest(pts, np.zeros(9), x, y, 1.0, 3)           # data_in length
    with pytest.raises(ValueError):
        ip.barycentric_interpolation(pts, np.zeros(10), x, y, np.ones((3, 3), int), np.ones((4, 3), int))
    with pytest.raises(ValueError):
        ip.idw_esrg_nearest_ex(pts, np.zeros(10), x, y, 1.0, 3, rows=(3, 9))
    simp = np.ones((4, 3), np.int32)
    with pytest.raises(ValueError):
        ip.mesh_reorder(pts[:, :1], simp, simp)                       # points must be (n, 2)
    with pytest.raises(ValueError):
        ip.mesh_reorder(pts, simp[:, :2], simp)                       # simplices must be (t, 3)
    with pytest.raises(ValueError):
        ip.mesh_reorder(pts, simp, np.ones((5, 3), np.int32))         # neighbours must match simplices


def test_mesh_reorder_fails_loudly_without_a_gpu():
    if _lib.lib().gi_device_count() > 0:
        pytest.skip("a GPU is present")
    simp = np.array([[1, 2, 3]], np.int32)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        ip.mesh_reorder(np.random.rand(3, 2), simp, np.zeros((1, 3), np.int32))


def test_no_silent_cpu_fallback():
    """Without a CUDA device a compute call must raise, never return numbers."""
    if _lib.lib().gi_device_count() > 0:
        pytest.skip("a GPU is present")
    x, y = np.arange(5.0), np.arange(4.0)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        ip.bilinear(x, y, np.zeros((4, 5)), np.random.rand(10, 2))


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, link or call it."""
    pkg = os.path.join(ROOT, "grid_interpolation_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.\.?oracle|liboracle|orc_[a-z_]+\s*\(|oracle[/\\]")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), f"{f} references the oracle"

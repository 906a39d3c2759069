"""grid_interpolation_b200 -- B200-native point location + interpolation (drop-in for the f2py module of
ChristianSteger/grid_interpolation).

    from grid_interpolation_b200 import interpolation as ip_fortran      # or: from interpolation import interpolation
    data_ip = ip_fortran.barycentric_interpolation(points, data_in, x_axis, y_axis, simplices, neighbours)

The package is a thin ctypes layer over the CUDA shared library (csrc/, include/gridinterp.h).
"""
from . import interpolation  # noqa: F401
from ._lib import LIB_PATH, lib  # noqa: F401

__all__ = ["interpolation", "lib", "LIB_PATH"]

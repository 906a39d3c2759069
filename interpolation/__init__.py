"""Drop-in module name: ``from interpolation import interpolation as ip_fortran`` (test_interpolation.py:27,
development/idw_interp_esrg_dev.py:20) resolves to the CUDA-backed implementation."""
from grid_interpolation_b200 import interpolation  # noqa: F401

__all__ = ["interpolation"]

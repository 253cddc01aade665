"""tmlqcd_b200 - B200-native even/odd twisted-mass Wilson-Dirac operator and CG (tmLQCD hot path).

The product is the C-ABI shared library tmlqcd_b200/lib/libtmlqcd_b200.so (CUDA kernels for
sm_100a + C host layer, sources in tmlqcd_b200/csrc, headers in include/).  This Python
package is only a ctypes binding used by tests/ and bench.py; it has no compute path of its
own and raises if the CUDA library is missing.
"""
from .capi import Device, DropIn, lib_path, load, build  # noqa: F401

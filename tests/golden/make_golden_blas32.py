"""Generates tests/golden/ref_blas32_4x4x4x4.npz: the single-precision BLAS-1 of the UNMODIFIED reference (linalg/..._32.c,
operator/tm_operators_32.c:130 in oracle/_ref, one thread) on seeded float fields of VOLUME/2 = 128 sites.
Run in the build container only:  make -C oracle/ref_build && python tests/golden/make_golden_blas32.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refclient import Reference  # noqa: E402


def main():
    ref = Reference(4, 4, 4, 4, nthreads=1)
    L = ref.lib
    fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    f, i = C.c_float, C.c_int
    for name, res, args in (("ref_square_norm_32", f, [fp, i]), ("ref_scalar_prod_r_32", f, [fp, fp, i]),
                            ("ref_assign_add_mul_r_32", None, [fp, fp, f, i]), ("ref_assign_mul_add_r_32", None, [fp, f, fp, i]),
                            ("ref_diff_32", None, [fp, fp, fp, i]), ("ref_mul_r_32", None, [fp, f, fp, i]),
                            ("ref_assign_mul_add_mul_r_32", None, [fp, fp, f, f, i]), ("ref_gamma5_32", None, [fp, fp, i])):
        getattr(L, name).restype = res; getattr(L, name).argtypes = args
    rng = np.random.default_rng(32)
    n = 128
    r, s, s2 = (rng.normal(size=(n, 24)).astype(np.float32) for _ in range(3))
    c1, c2 = np.float32(0.37), np.float32(-1.21)
    out = {"r": r, "s": s, "s2": s2, "c1": c1, "c2": c2,
           "square_norm_32": np.float32(L.ref_square_norm_32(r, n)), "scalar_prod_r_32": np.float32(L.ref_scalar_prod_r_32(s, r, n))}
    x = r.copy(); L.ref_assign_add_mul_r_32(x, s, c1, n); out["assign_add_mul_r_32"] = x
    x = r.copy(); L.ref_assign_mul_add_r_32(x, c1, s, n); out["assign_mul_add_r_32"] = x
    x = np.zeros_like(r); L.ref_diff_32(x, s, s2, n); out["diff_32"] = x
    x = np.zeros_like(r); L.ref_mul_r_32(x, c1, s, n); out["mul_r_32"] = x
    x = r.copy(); L.ref_assign_mul_add_mul_r_32(x, s, c1, c2, n); out["assign_mul_add_mul_r_32"] = x
    x = np.zeros_like(r); L.ref_gamma5_32(x, s, n); out["gamma5_32"] = x
    np.savez_compressed(os.path.join(HERE, "ref_blas32_4x4x4x4.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()

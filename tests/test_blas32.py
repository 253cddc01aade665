"""Single-precision BLAS-1 of the mixed solvers (linalg/..._32.c, gamma5_32).
CPU: a numpy float32 restatement against values of the unmodified reference (golden fixture, bit-exact for the elementwise
routines) and, through the host emulation, the product's device functor.  GPU: the reference-named drop-in symbols with
host buffers, on VOLUME/2 and VOLUME sites."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "ref_blas32_4x4x4x4.npz")


def restate(g):
    """linalg/assign_add_mul_r_32.c:104, assign_mul_add_r_32.c:81, diff_32.c:39, mul_r_32.c:69, assign_mul_add_mul_r_32.c:37,
    operator/tm_operators_32.c:114-139 in float32 arithmetic"""
    r, s, s2, c1, c2 = g["r"], g["s"], g["s2"], np.float32(g["c1"]), np.float32(g["c2"])
    g5 = s.copy(); g5[:, 12:] = -g5[:, 12:]
    return {"assign_add_mul_r_32": r + c1 * s, "assign_mul_add_r_32": c1 * r + s, "diff_32": s - s2, "mul_r_32": c1 * s,
            "assign_mul_add_mul_r_32": c1 * r + c2 * s, "gamma5_32": g5,
            "square_norm_32": np.sum(r.astype(np.float64) ** 2), "scalar_prod_r_32": np.sum(s.astype(np.float64) * r.astype(np.float64))}


def test_restatement_against_reference_values():
    g = np.load(GOLD)
    exp = restate(g)
    for k in ("assign_add_mul_r_32", "assign_mul_add_r_32", "diff_32", "mul_r_32", "assign_mul_add_mul_r_32", "gamma5_32"):
        assert exp[k].dtype == np.float32 and np.array_equal(exp[k], g[k]), k
    # the reference sums in float with Kahan compensation: agrees with the double sum to float precision
    assert abs(float(g["square_norm_32"]) - exp["square_norm_32"]) <= 2e-7 * exp["square_norm_32"]
    assert abs(float(g["scalar_prod_r_32"]) - exp["scalar_prod_r_32"]) <= 2e-7 * np.sum(np.abs(g["s"].astype(np.float64) * g["r"]))


def test_device_functor_on_the_host():
    """EwBlas32 of tmb_kernels.cu compiled for the host, on the device layout [12][sites] (an elementwise functor: any
    consistent layout gives the same numbers; gamma5 needs the spin-major SoA order)"""
    from emul_client import load
    E = load()
    g = np.load(GOLD)
    n = 128
    soa = lambda a: np.ascontiguousarray(a.reshape(n, 12, 2).transpose(1, 0, 2)).reshape(-1)   # [12][n] complex
    aos = lambda v: np.ascontiguousarray(v.reshape(12, n, 2).transpose(1, 0, 2)).reshape(n, 24)
    r, s, s2 = soa(g["r"]), soa(g["s"]), soa(g["s2"])
    c1, c2 = float(g["c1"]), float(g["c2"])
    for op, name, inplace in ((0, "assign_add_mul_r_32", True), (1, "assign_mul_add_r_32", True), (2, "diff_32", False),
                              (3, "mul_r_32", False), (4, "assign_mul_add_mul_r_32", True), (5, "gamma5_32", False)):
        x = r.copy() if inplace else np.zeros_like(r)
        E.emul_blas32(op, x, s, s2, c1, c2, 12 * n)
        assert np.array_equal(aos(x), g[name]), name


@pytest.mark.gpu
@pytest.mark.parametrize("nparts", [1, 2])
def test_gpu_dropin_blas32(nparts):
    import tmlqcd_b200 as tm
    g = np.load(GOLD)
    D = tm.DropIn(4, 4, 4, 4)
    try:
        D.set_params(0.16, 0.0032)
        n = 128 * nparts  # VOLUME/2 or VOLUME sites: the golden fields twice for the full volume
        rep = lambda a: np.ascontiguousarray(np.concatenate([a] * nparts))
        r, s, s2 = rep(g["r"]), rep(g["s"]), rep(g["s2"])
        c1, c2 = float(g["c1"]), float(g["c2"])
        tol = lambda a, b: np.max(np.abs(a - b)) <= 2e-6 * np.max(np.abs(b))  # device code may fuse multiply-add
        x = r.copy(); D.assign_add_mul_r_32(x, s, c1, n); assert tol(x, rep(g["assign_add_mul_r_32"]))
        x = r.copy(); D.assign_mul_add_r_32(x, c1, s, n); assert tol(x, rep(g["assign_mul_add_r_32"]))
        x = np.zeros_like(r); D.diff_32(x, s, s2, n); assert np.array_equal(x, rep(g["diff_32"]))
        x = np.zeros_like(r); D.mul_r_32(x, c1, s, n); assert np.array_equal(x, rep(g["mul_r_32"]))
        x = r.copy(); D.assign_mul_add_mul_r_32(x, s, c1, c2, n); assert tol(x, rep(g["assign_mul_add_mul_r_32"]))
        x = np.zeros_like(r); D.gamma5_32(x, s, n); assert np.array_equal(x, rep(g["gamma5_32"]))
        sq = D.square_norm_32(r, n, 0); sp = D.scalar_prod_r_32(s, r, n, 0)
        assert abs(sq - nparts * float(g["square_norm_32"])) <= 1e-6 * nparts * float(g["square_norm_32"])
        assert abs(sp - nparts * float(g["scalar_prod_r_32"])) <= 1e-5 * nparts * abs(float(g["scalar_prod_r_32"])) + 1e-3
    finally:
        D.close()

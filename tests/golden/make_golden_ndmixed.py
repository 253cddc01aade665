"""Generates tests/golden/ref_ndmixed_4x4x4x4.npz with the UNMODIFIED reference (oracle/_ref, half-spinor build, one thread):
the single-precision two-flavour operator Qtm_pm_ndpsi_32 (operator/tm_operators_nd_32.c:215) and the mixed-precision doublet
solver rg_mixed_cg_her_nd (solver/rg_mixed_cg_her_nd.c:189, the RGMIXEDCG branch of invert_doublet_eo.c:145-150).
Groundwork for the next round: the B200 path serves that branch with the double-precision CG today.
Run in the build container only:  make -C oracle/ref_build && python tests/golden/make_golden_ndmixed.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refclient import Reference  # noqa: E402


def main():
    r = Reference(4, 4, 4, 4, nthreads=1, halfspinor=True)
    kappa, gmu, theta, nd, delta = 0.16, 0.0032, (1., 0., 0., 0.), (0.139, 0.15, 0.9), 0.1
    r.set_params(kappa, gmu, theta); r.set_nd_params(*nd)
    g = r.random_gauge(2024)
    assert r.lib.ref_init32() == 0
    r.lib.ref_update_gauge32()
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    r.lib.ref_rg_mixed_cg_her_nd.restype = C.c_int
    r.lib.ref_rg_mixed_cg_her_nd.argtypes = [dp] * 4 + [C.c_int, C.c_double, C.c_int, C.c_double]
    r.lib.ref_Qtm_pm_ndpsi_32.restype = None
    r.lib.ref_Qtm_pm_ndpsi_32.argtypes = [fp] * 4
    s, c = r.random_spinor_eo(), r.random_spinor_eo()
    s32, c32 = s.astype(np.float32), c.astype(np.float32)
    ls, lc = np.zeros_like(s32), np.zeros_like(s32)
    r.lib.ref_Qtm_pm_ndpsi_32(ls, lc, s32, c32)
    pu, pd = np.zeros_like(s), np.zeros_like(s)
    count = r.lib.ref_rg_mixed_cg_her_nd(pu, pd, s, c, 2000, 1e-20, 1, delta)
    np.savez_compressed(os.path.join(HERE, "ref_ndmixed_4x4x4x4.npz"), dims=np.array([4, 4, 4, 4]), kappa=kappa, gmu=gmu,
                        theta=np.array(theta), nd=np.array(nd), delta=delta, gauge=g, s=s, c=c, Qtm_pm_ndpsi_32_s=ls,
                        Qtm_pm_ndpsi_32_c=lc, eps_sq=1e-20, rel_prec=1, count=count, x_s=pu, x_c=pd)
    print("rg_mixed_cg_her_nd count", count)


if __name__ == "__main__":
    main()

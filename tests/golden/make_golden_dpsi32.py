"""Generates tests/golden/ref_dpsi32_4x4x4x4.npz: D_psi_32 (operator/D_psi.h:28) and Q_pm_psi_32
(operator/tm_operators_32.c:141) of the UNMODIFIED reference (oracle/_ref, half-spinor build, one thread) on the
lexicographic field `lex` of ref_4x4x4x4.npz rounded to float, same gauge field / kappa / mu / theta.
Run in the build container only:  make -C oracle/ref_build && python tests/golden/make_golden_dpsi32.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refclient import Reference  # noqa: E402


def main():
    base = np.load(os.path.join(HERE, "ref_4x4x4x4.npz"))
    dims = tuple(int(x) for x in base["dims"])
    r = Reference(*dims, nthreads=1, halfspinor=True)
    r.set_params(float(base["kappa"]), float(base["gmu"]), base["theta"])
    r.set_gauge(base["gauge"])
    assert r.lib.ref_init32() == 0
    r.lib.ref_update_gauge32()
    fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    for n in ("ref_D_psi_32", "ref_Q_pm_psi_32"):
        getattr(r.lib, n).restype = None; getattr(r.lib, n).argtypes = [fp, fp]
    lex32 = np.ascontiguousarray(base["lex"], dtype=np.float32)
    d = np.zeros_like(lex32); r.lib.ref_D_psi_32(d, lex32.copy())
    q = np.zeros_like(lex32); r.lib.ref_Q_pm_psi_32(q, lex32.copy())
    # sanity: the double-precision D_psi of the same build on the same (float-rounded) input
    dd = r.spinor(r.V); r.D_psi(dd, lex32.astype(np.float64))
    print("D_psi_32 vs D_psi rel:", np.linalg.norm(d - dd) / np.linalg.norm(dd))
    fn = os.path.join(HERE, "ref_dpsi32_4x4x4x4.npz")
    np.savez_compressed(fn, lex32=lex32, D_psi_32=d, Q_pm_psi_32=q)
    print("wrote", fn, os.path.getsize(fn), "bytes")


if __name__ == "__main__":
    main()

"""Generates tests/golden/ref_ops_4x4x4x4.npz: the remaining members of the operator families (symmetric
even/odd operators, Mee_psi / Mee_inv_psi, full-lattice M_minus_psi / D_dagg_psi, ND helpers) computed by
the UNMODIFIED reference (oracle/_ref) on the inputs of ref_4x4x4x4.npz (same gauge field, k, p, q, w).

Run in the build container only:  make -C oracle/ref_build && python tests/golden/make_golden_ops.py
"""
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
    r = Reference(*dims, nthreads=1)
    r.set_params(float(base["kappa"]), float(base["gmu"]), base["theta"])
    r.set_nd_params(*base["nd"])
    r.set_gauge(base["gauge"])
    k, p, q, w, lex = base["k"], base["p"], base["q"], base["w"], base["lex"]
    out = {}
    for name in ("Qtm_plus_sym_psi", "Qtm_minus_sym_psi", "Mtm_plus_sym_psi", "Mtm_minus_sym_psi", "Mtm_plus_sym_dagg_psi",
                 "Qtm_pm_sym_psi"):
        a = r.spinor(); getattr(r, name)(a, k.copy()); out[name] = a
    for name in ("M_minus_psi", "D_dagg_psi", "Q_plus_psi", "Q_minus_psi"):
        a = r.spinor(r.V); getattr(r, name)(a, lex.copy()); out[name] = a
    a = r.spinor(); r.Mee_psi(a, k, 0.37); out["Mee_psi"] = a
    a = r.spinor(); r.Mee_inv_psi(a, k, 0.37); out["Mee_inv_psi"] = a
    a = r.spinor(); r.mul_one_sub_mul_gamma5(a, k, p); out["mul_one_sub_mul_gamma5"] = a
    a = r.spinor(); r.mul_one_pm_imu_sub_mul(a, k, p, -1., r.Vh); out["mul_one_pm_imu_sub_mul"] = a
    a, b = r.spinor(), r.spinor(); r.M_minus_1_timesC(a, b, k, p); out["M_minus_1_timesC_e"], out["M_minus_1_timesC_o"] = a, b
    a, b = r.spinor(), r.spinor(); r.H_eo_tm_ndpsi(a, b, k, p, 1); out["H_eo_tm_ndpsi_s"], out["H_eo_tm_ndpsi_c"] = a, b
    a, b = r.spinor(), r.spinor(); r.M_oo_sub_g5_ndpsi(a, b, k, p, q, w, -0.139, -0.15)
    out["M_oo_sub_g5_ndpsi_s"], out["M_oo_sub_g5_ndpsi_c"] = a, b
    a = r.spinor(); r.mul_one_pm_iconst(a, k, 0.21, -1); out["mul_one_pm_iconst"] = a
    fn = os.path.join(HERE, "ref_ops_4x4x4x4.npz")
    np.savez_compressed(fn, **out)
    print("wrote", fn, os.path.getsize(fn), "bytes")


if __name__ == "__main__":
    main()

"""Generates tests/golden/ref_hmc_4x4x4x4.npz: fermion force, chronological guess and DET / DETRATIO
monomial results of the UNMODIFIED reference (oracle/_ref: deriv_Sb.c, solver/chrono_guess.c,
solver/monomial_solve.c, monomial/{monomial,det_monomial,detratio_monomial}.c compiled from
/root/reference by oracle/ref_build/Makefile) on its own RANLUX-generated inputs.

Run in the build container only (the GPU box has no /root/reference):
    make -C oracle/ref_build && python tests/golden/make_golden_hmc.py
Gauge: start_ranlux(1, 123456); random_gauge_field(repro=1).  Pseudo-fermion noise: the reference's
heatbath draws random_spinor_field_eo(w_fields[0], repro, RN_GAUSS); the same field is obtained here by
re-seeding ranlux before an explicit draw (seed 1000 + id), so that it can be handed to other
implementations.  kappa=0.16, g_mu=0.0032 (kappa2=0.16, g_mu2=0.032 for the ratio), theta=(1,0.3,0,0.7).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refclient import Reference  # noqa: E402

DIMS = (4, 4, 4, 4)
KAPPA, GMU, THETA = 0.16, 0.0032, (1.0, 0.3, 0.0, 0.7)
KAPPA2, GMU2 = 0.16, 0.032
FORCEPREC, ACCPREC = 1e-22, 1e-24
CG = 1
# (type, csg_N): DET = 0, DETRATIO = 1 (monomial.h:27-28)
MONOMIALS = [(0, 0), (1, 0), (0, 2), (1, 2)]


def main():
    r = Reference(*DIMS, nthreads=1)
    r.set_params(KAPPA, GMU, THETA)
    assert r.hmc_init() == 0
    out = {"dims": np.array(DIMS), "kappa": KAPPA, "gmu": GMU, "kappa2": KAPPA2, "gmu2": GMU2, "theta": np.array(THETA),
           "forceprec": FORCEPREC, "accprec": ACCPREC, "monomials": np.array(MONOMIALS)}
    out["gauge"] = r.random_gauge(123456)
    l, k = r.random_spinor_eo(), r.random_spinor_eo()
    out.update(l=l, k=k)
    for ieo in (0, 1):
        df = r.derivative()
        r.deriv_Sb(ieo, l, k, df, 0.7)
        out[f"deriv_Sb{ieo}"] = df
    ids = [r.mnl_add(t, KAPPA, GMU, KAPPA2, GMU2, CG, 2000, FORCEPREC, ACCPREC, n) for t, n in MONOMIALS]
    assert r.mnl_init() == 0
    for id in ids:
        r.start_ranlux(1, 1000 + id); eta = r.random_spinor_eo()
        r.start_ranlux(1, 1000 + id); r.mnl_heatbath(id)
        pf = r.spinor(); r.mnl_get_pf(id, pf)
        if MONOMIALS[id][0] == 0:  # det_heatbath: pf = Qtm_plus_psi(eta)
            chk = r.spinor(); r.Qtm_plus_psi(chk, eta); assert np.array_equal(chk, pf)
        out[f"m{id}_eta"], out[f"m{id}_pf"] = eta, pf
        out[f"m{id}_energy0"] = r.mnl_info(id)["energy0"]
        df = r.derivative()
        for call in range(3):  # three MD steps on the same gauge field: exercises the chronological guess
            r.mnl_derivative(id, df)
            out[f"m{id}_df{call}"] = df.copy()
            out[f"m{id}_iter1_{call}"] = r.mnl_info(id)["iter1"]
        out[f"m{id}_dH"] = r.mnl_acc(id)
        info = r.mnl_info(id)
        out[f"m{id}_iter0"], out[f"m{id}_csg_n"] = info["iter0"], info["csg_n"]
        print(id, MONOMIALS[id], {k_: v for k_, v in info.items()}, "dH", out[f"m{id}_dH"])
    fn = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_hmc_4x4x4x4.npz")
    np.savez_compressed(fn, **out)
    print("wrote", fn, os.path.getsize(fn), "bytes")


if __name__ == "__main__":
    main()

"""Generates tests/golden/ref_4x4x4x4.npz by running the UNMODIFIED reference (oracle/_ref,
compiled from /root/reference by oracle/ref_build/Makefile) on its own RANLUX-generated inputs.

Run in the build container only (the GPU box has no /root/reference):
    make -C oracle/ref_build && python tests/golden/make_golden.py
Inputs: start_ranlux(1, 123456); random_gauge_field(repro=1); random_spinor_field_eo(repro=1, RN_GAUSS)
exactly as benchmark.c:247-259.  kappa=0.16, g_mu=2*kappa*0.01, theta=(1, 0.3, 0, 0.7).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refclient import Reference  # noqa: E402

DIMS = (4, 4, 4, 4)
KAPPA, GMU, THETA = 0.16, 0.0032, (1.0, 0.3, 0.0, 0.7)
ND = (0.139, 0.15, 0.9)


def main():
    r = Reference(*DIMS, nthreads=1)
    r.set_params(KAPPA, GMU, THETA)
    r.set_nd_params(*ND)
    out = {"dims": np.array(DIMS), "kappa": KAPPA, "gmu": GMU, "theta": np.array(THETA), "nd": np.array(ND)}
    out["gauge"] = r.random_gauge(123456)
    k, p, q, w = (r.random_spinor_eo() for _ in range(4))
    out.update(k=k, p=p, q=q, w=w)
    out["eo2lexic"] = r.table("eo2lexic")
    out["hi"] = r.table("hi")
    ka = np.zeros(8); r.get_ka(ka); out["ka"] = ka
    for ieo in (0, 1):
        a = r.spinor(); r.Hopping_Matrix(ieo, a, k); out[f"hop{ieo}"] = a
        a = r.spinor(); r.tm_times_Hopping_Matrix(ieo, a, k, 0.9, -0.2); out[f"tm_times{ieo}"] = a
        a = r.spinor(); r.tm_sub_Hopping_Matrix(ieo, a, p, k, 1.0, 0.3); out[f"tm_sub{ieo}"] = a
    for name in ("Qtm_pm_psi", "Qtm_plus_psi", "Qtm_minus_psi", "Mtm_plus_psi", "Mtm_minus_psi"):
        a = r.spinor(); getattr(r, name)(a, k); out[name] = a
    a, b = r.spinor(), r.spinor(); r.M_full(a, b, k, p); out["M_full_e"], out["M_full_o"] = a, b
    lex = r.spinor(r.V); r.convert_eo_to_lexic(lex, k, p); out["lex"] = lex
    d = r.spinor(r.V); r.D_psi(d, lex); out["D_psi"] = d
    out["square_norm_k"] = r.square_norm(k, r.Vh)
    out["scalar_prod_kp"] = r.scalar_prod_r(k, p, r.Vh)
    x = r.spinor(); out["cg_iters"] = r.cg_her(x, k, 1000, 1e-20, 1); out["cg_x"] = x
    en, on = r.spinor(), r.spinor()
    out["invert_iters"] = r.invert_eo_cg(en, on, k, p, 1e-20, 1000, 1); out["invert_en"], out["invert_on"] = en, on
    for name in ("Qtm_ndpsi", "Qtm_dagger_ndpsi", "Qtm_pm_ndpsi"):
        a, b = r.spinor(), r.spinor(); getattr(r, name)(a, b, k, p); out[name + "_s"], out[name + "_c"] = a, b
    a, b = r.spinor(), r.spinor()
    out["cg_nd_iters"] = r.cg_her_nd(a, b, k, p, 1000, 1e-18, 1); out["cg_nd_s"], out["cg_nd_c"] = a, b
    A = [r.spinor() for _ in range(4)]
    out["invert_doublet_iters"] = r.invert_doublet_eo_cg(*A, k, p, q, w, 1e-18, 1000, 1)
    for n, v in zip(("ens", "ons", "enc", "onc"), A):
        out["invert_doublet_" + n] = v
    fn = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_4x4x4x4.npz")
    np.savez_compressed(fn, **out)
    print("wrote", fn, os.path.getsize(fn), "bytes")


if __name__ == "__main__":
    main()

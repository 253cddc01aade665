"""GPU parity of the remaining reference-named entry points (SURVEY 8a rows a13, a15, a16, a18, a25, a27, a29,
a31 and the HMC symbols of 8f) through the drop-in layer with host buffers, against results of the UNMODIFIED
reference (tests/golden/ref_ops_4x4x4x4.npz, ref_hmc_4x4x4x4.npz; generators committed beside them)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-13
CG, MIXEDCG, RGMIXEDCG = 1, 13, 14


def _gold(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


@pytest.fixture()
def dropin():
    import tmlqcd_b200 as tm
    base = _gold("ref_4x4x4x4.npz")
    D = tm.DropIn(*[int(x) for x in base["dims"]])
    D.set_params(float(base["kappa"]), float(base["gmu"]), base["theta"])
    D.set_nd_params(*base["nd"])
    D.set_gauge(base["gauge"])
    yield D, base
    D.close()


def test_operator_family_members_vs_reference(dropin):
    D, base = dropin
    ops = _gold("ref_ops_4x4x4x4.npz")
    k, p, q, w, lex = (np.array(base[n]) for n in ("k", "p", "q", "w", "lex"))
    out = D.spinor()
    for name in ("Qtm_plus_sym_psi", "Qtm_minus_sym_psi", "Mtm_plus_sym_psi", "Mtm_minus_sym_psi", "Mtm_plus_sym_dagg_psi",
                 "Qtm_pm_sym_psi"):
        getattr(D, name)(out, k); assert rel_l2(out, ops[name]) <= TOL, name
    for name, ref in (("Qtm_plus_sym_psi_nocom", "Qtm_plus_sym_psi"), ("Mtm_plus_sym_psi_nocom", "Mtm_plus_sym_psi"),
                      ("Mtm_minus_sym_psi_nocom", "Mtm_minus_sym_psi")):
        getattr(D, name)(out, k); assert rel_l2(out, ops[ref]) <= TOL, name
    for name, ref in (("Qtm_plus_psi_nocom", "Qtm_plus_psi"), ("Mtm_plus_psi_nocom", "Mtm_plus_psi"), ("Qtm_pm_psi_nocom", "Qtm_pm_psi")):
        getattr(D, name)(out, k); assert rel_l2(out, base[ref]) <= TOL, name
    outl = D.spinor(D.V)
    for name in ("M_minus_psi", "D_dagg_psi", "Q_plus_psi", "Q_minus_psi"):
        getattr(D, name)(outl, lex); assert rel_l2(outl, ops[name]) <= TOL, name
    D.Mee_psi(out, k, 0.37); assert rel_l2(out, ops["Mee_psi"]) <= TOL
    D.Mee_inv_psi(out, k, 0.37); assert rel_l2(out, ops["Mee_inv_psi"]) <= TOL
    D.mul_one_sub_mul_gamma5(out, k, p); assert rel_l2(out, ops["mul_one_sub_mul_gamma5"]) <= TOL
    D.mul_one_pm_imu_sub_mul(out, k, p, -1., D.Vh); assert rel_l2(out, ops["mul_one_pm_imu_sub_mul"]) <= TOL
    a, b = D.spinor(), D.spinor()
    D.M_minus_1_timesC(a, b, k, p)
    assert rel_l2(a, ops["M_minus_1_timesC_e"]) <= TOL and rel_l2(b, ops["M_minus_1_timesC_o"]) <= TOL
    D.H_eo_tm_ndpsi(a, b, k, p, 1)
    assert rel_l2(a, ops["H_eo_tm_ndpsi_s"]) <= TOL and rel_l2(b, ops["H_eo_tm_ndpsi_c"]) <= TOL
    D.M_oo_sub_g5_ndpsi(a, b, k, p, q, w, -0.139, -0.15)
    assert rel_l2(a, ops["M_oo_sub_g5_ndpsi_s"]) <= TOL and rel_l2(b, ops["M_oo_sub_g5_ndpsi_c"]) <= TOL
    D.mul_one_pm_iconst(out, k, 0.21, -1); assert rel_l2(out, ops["mul_one_pm_iconst"]) <= TOL
    # single-precision lexicographic operators (operator/D_psi.h:28, tm_operators_32.c:141) vs the unmodified reference's
    # float results: both sides round differently, north_star's single-precision tolerance is 1e-5
    g32 = _gold("ref_dpsi32_4x4x4x4.npz")
    lex32 = np.array(g32["lex32"]); o32 = np.zeros_like(lex32)
    D.D_psi_32(o32, lex32); assert rel_l2(o32.astype(np.float64), g32["D_psi_32"].astype(np.float64)) <= 1e-5
    assert rel_l2(o32.astype(np.float64), ops_D_psi(D, lex32)) <= 1e-5
    D.Q_pm_psi_32(o32, lex32); assert rel_l2(o32.astype(np.float64), g32["Q_pm_psi_32"].astype(np.float64)) <= 1e-5


def ops_D_psi(D, lex32):
    """the double-precision D_psi of the same library on the float-rounded input"""
    out = D.spinor(D.V); D.D_psi(out, np.ascontiguousarray(lex32, dtype=np.float64)); return out


def test_host_memory_helpers_and_precision_conversion(dropin):
    D, base = dropin
    k, p = np.array(base["k"]), np.array(base["p"])
    z = k.copy(); D.zero_spinor_field(z, D.Vh); assert not z.any()
    f = np.zeros((D.Vh, 24), dtype=np.float32)
    D.assign_to_32(f, k, D.Vh); assert np.array_equal(f, k.astype(np.float32))
    back = D.spinor(); D.assign_to_64(back, f, D.Vh); assert np.array_equal(back, f.astype(np.float64))
    acc = p.copy(); D.addto_32(acc, f, D.Vh); assert rel_l2(acc, p + f.astype(np.float64)) <= 1e-15
    sf = C.POINTER(C.c_void_p)()
    assert D.init_solver_field(C.byref(sf), D.Vh, 3) == 0
    assert sf[1] - sf[0] == D.Vh * 192 and sf[2] - sf[1] == D.Vh * 192 and sf[3] == sf[0]  # solver_field.c:63-66
    D.finalize_solver(sf, 3)


def test_solvers_with_reference_signatures(dropin):
    import tmlqcd_b200 as tm
    D, base = dropin
    k = np.array(base["k"])
    sp = tm.capi.SolverParams(); sp.mcg_delta = 5e-5
    xr = base["cg_x"]  # reference cg_her(k, 1e-20 relative)
    for solver in (CG, MIXEDCG, RGMIXEDCG):
        x = D.spinor()
        it = D.solve_degenerate(x, k, sp, 2000, 1e-20, 1, D.Vh, D.fptr("Qtm_pm_psi"), solver)
        assert it > 0 and rel_l2(x, xr) <= 1e-8, solver
        if solver == CG:
            assert abs(it - int(base["cg_iters"])) <= 1
    x = D.spinor()
    it = D.rg_mixed_cg_her(x, k, sp, 2000, 1e-20, 1, D.Vh, D.fptr("Qtm_pm_psi"), D.fptr("Qtm_pm_psi_32"))
    assert it > 0 and rel_l2(x, xr) <= 1e-8


def test_inversions_with_reference_signatures(dropin):
    """invert_eo (17 arguments, invert_eo.c:83-89) and invert_doublet_eo (invert_doublet_eo.c:68-76) with host buffers
    against the solutions of the unmodified reference: argument order of the four outputs and four sources included"""
    import tmlqcd_b200 as tm
    D, base = dropin
    k, p, q, w = (np.array(base[n]) for n in ("k", "p", "q", "w"))
    sp = tm.capi.SolverParams()
    en, on = D.spinor(), D.spinor()
    it = D.invert_eo(en, on, k, p, 1e-20, 1000, CG, 1, 0, 1, 0, None, sp, 0, 0, 0, 18)
    assert abs(it - int(base["invert_iters"])) <= 1
    assert rel_l2(en, base["invert_en"]) <= 1e-8 and rel_l2(on, base["invert_on"]) <= 1e-8
    outs = [D.spinor() for _ in range(4)]
    it = D.invert_doublet_eo(*outs, k, p, q, w, 1e-18, 1000, CG, 1, sp, 0, 0, 18)
    assert abs(it - int(base["invert_doublet_iters"])) <= 1
    for o_, name in zip(outs, ("ens", "ons", "enc", "onc")):
        assert rel_l2(o_, base["invert_doublet_" + name]) <= 1e-8, name
    # the RGMIXEDCG branch (invert_doublet_eo.c:145-149): same solution, to the solve's precision
    sp.mcg_delta = 0.1
    outs = [D.spinor() for _ in range(4)]
    it = D.invert_doublet_eo(*outs, k, p, q, w, 1e-18, 1000, RGMIXEDCG, 1, sp, 0, 0, 18)
    assert it > 0
    for o_, name in zip(outs, ("ens", "ons", "enc", "onc")):
        assert rel_l2(o_, base["invert_doublet_" + name]) <= 1e-7, name
    # rg_mixed_cg_her_nd and Qtm_pm_ndpsi_32 with the reference's signatures, against the unmodified reference's results
    g = _gold("ref_ndmixed_4x4x4x4.npz")
    D.set_params(float(g["kappa"]), float(g["gmu"]), g["theta"]); D.set_nd_params(*g["nd"]); D.set_gauge(g["gauge"])
    s_, c_ = np.array(g["s"]), np.array(g["c"])
    l1, l2 = np.zeros((D.Vh, 24), dtype=np.float32), np.zeros((D.Vh, 24), dtype=np.float32)
    D.Qtm_pm_ndpsi_32(l1, l2, s_.astype(np.float32), c_.astype(np.float32))
    assert rel_l2(l1.astype(np.float64), g["Qtm_pm_ndpsi_32_s"].astype(np.float64)) <= 1e-5
    assert rel_l2(l2.astype(np.float64), g["Qtm_pm_ndpsi_32_c"].astype(np.float64)) <= 1e-5
    sp.mcg_delta = float(g["delta"])
    xu, xd = D.spinor(), D.spinor()
    it = D.rg_mixed_cg_her_nd(xu, xd, s_, c_, sp, 2000, float(g["eps_sq"]), int(g["rel_prec"]), D.Vh, D.fptr("Qtm_pm_ndpsi"), D.fptr("Qtm_pm_ndpsi_32"))
    assert it > 0 and rel_l2(xu, g["x_s"]) <= 1e-8 and rel_l2(xd, g["x_c"]) <= 1e-8


def test_invert_eo_remaining_branches(dropin):
    """invert_eo's MIXEDCG and RGMIXEDCG branches (invert_eo.c:234-249) and its branch WITHOUT even/odd preconditioning
    (:364-558: cg_her on Q_pm_psi over VOLUME sites, the source as initial guess, then Q_minus_psi): all of them solve the
    system of the CG branch, whose solution by the unmodified reference is in the fixture.  Then cg_her(N = VOLUME,
    f = Q_pm_psi) - the recurrence on (even, odd) pairs of device fields - against the generic path (same recurrence, f
    through its host-pointer entry point) and against the defining equation."""
    import tmlqcd_b200 as tm
    D, base = dropin
    k, p, lex = (np.array(base[n]) for n in ("k", "p", "lex"))
    sp = tm.capi.SolverParams(); sp.mcg_delta = 5e-5
    for solver, eo, prec in ((MIXEDCG, 1, 1e-20), (RGMIXEDCG, 1, 1e-20), (CG, 0, 1e-24)):
        en, on = D.spinor(), D.spinor()
        it = D.invert_eo(en, on, k, p, prec, 3000, solver, 1, 0, eo, 0, None, sp, 0, 0, 0, 18)
        assert it > 0, (solver, eo)
        assert rel_l2(en, base["invert_en"]) <= 1e-7 and rel_l2(on, base["invert_on"]) <= 1e-7, (solver, eo)
    x1, x2, out = D.spinor(D.V), D.spinor(D.V), D.spinor(D.V)
    it1 = D.cg_her(x1, lex, 3000, 1e-20, 1, D.V, D.fptr("Q_pm_psi"))
    raw = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)(("Q_pm_psi", D.lib))
    cb = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)(lambda l_, k_: raw(l_, k_))
    it2 = D.cg_her(x2, lex, 3000, 1e-20, 1, D.V, C.cast(cb, C.c_void_p))
    assert it1 > 0 and abs(it1 - it2) <= 1 and rel_l2(x1, x2) <= 1e-9
    D.Q_pm_psi(out, x1)
    assert rel_l2(out, lex) <= 1e-9
    # solve_degenerate hands f == Q_pm_psi on VOLUME sites to the same CG (monomial_solve.c:149)
    x3 = D.spinor(D.V)
    it3 = D.solve_degenerate(x3, lex, sp, 3000, 1e-20, 1, D.V, D.fptr("Q_pm_psi"), CG)
    assert it3 == it1 and rel_l2(x3, x1) <= 1e-12


def _hmc_dropin():
    import tmlqcd_b200 as tm
    gold = _gold("ref_hmc_4x4x4x4.npz")
    D = tm.DropIn(*[int(x) for x in gold["dims"]])
    D.set_params(float(gold["kappa"]), float(gold["gmu"]), gold["theta"])
    D.set_gauge(gold["gauge"])
    return D, gold


def test_hmc_deriv_Sb_with_reference_signature():
    """deriv_Sb(ieo, l, k, hf, factor) with host buffers accumulates into the caller's derivative array"""
    D, gold = _hmc_dropin()
    try:
        l, k = np.array(gold["l"]), np.array(gold["k"])
        for ieo in (0, 1):
            df = np.zeros((D.V, 4, 8))
            hf = D.hamiltonian_field(df)
            D.deriv_Sb(ieo, l, k, C.byref(hf), 0.7)
            assert rel_l2(df, gold[f"deriv_Sb{ieo}"]) <= TOL
            D.deriv_Sb(ieo, l, k, C.byref(hf), -0.7)  # accumulates into the caller's array
            assert np.abs(df).max() <= 1e-12
    finally:
        D.close()


def test_hmc_chrono_guess_with_reference_signatures():
    """chronological guess on host fields: the exact solution in the history reproduces itself"""
    import tmlqcd_b200 as tm
    D, gold = _hmc_dropin()
    try:
        k = np.array(gold["k"])
        N = 2
        hist = [D.spinor() for _ in range(N)]
        v = (C.c_void_p * N)(*[h.ctypes.data for h in hist])
        idx, n = (C.c_int * N)(), C.c_int(0)
        x = D.spinor()
        sp = tm.capi.SolverParams()
        assert D.solve_degenerate(x, k, sp, 2000, 1e-24, 0, D.Vh, D.fptr("Qtm_pm_psi"), CG) > 0
        D.chrono_add_solution(x, v, idx, N, C.byref(n), D.Vh)
        assert n.value == 1 and abs(np.linalg.norm(hist[0]) - 1) < 1e-14
        trial = D.spinor()
        assert D.chrono_guess(trial, k, v, idx, N, n.value, D.Vh, D.fptr("Qtm_pm_psi")) == 0
        assert rel_l2(trial, x) <= 1e-9
    finally:
        D.close()


def test_hmc_monomials_with_reference_signatures():
    """det_* / detratio_* through the hbfunction / derivativefunction / accfunction signatures (monomial.h:125-127)"""
    import tmlqcd_b200 as tm
    D, gold = _hmc_dropin()
    try:
        etas = {}

        def rng(ptr, repro, rn_type):
            assert repro == 1 and rn_type == 0
            np.ctypeslib.as_array(ptr, shape=(D.Vh * 24,))[:] = etas["cur"].reshape(-1)
        cb = tm.capi.RANDOM_SPINOR_FN(rng)
        D.tmb_dropin_set_random_spinor_field_eo(cb)
        names = {0: ("det_heatbath", "det_derivative", "det_acc"), 1: ("detratio_heatbath", "detratio_derivative", "detratio_acc")}
        for id, (typ, csg_N) in enumerate(gold["monomials"]):
            assert D.tmb_dropin_register_monomial(id, int(typ), float(gold["kappa"]), float(gold["gmu"]), float(gold["kappa2"]),
                                                  float(gold["gmu2"]), CG, 2000, float(gold["forceprec"]), float(gold["accprec"]),
                                                  int(csg_N)) == 0
            hb, der, acc = names[int(typ)]
            etas["cur"] = np.array(gold[f"m{id}_eta"])
            df = np.zeros((D.V, 4, 8)); hf = D.hamiltonian_field(df)
            getattr(D, hb)(id, C.byref(hf))
            for call in range(3):
                getattr(D, der)(id, C.byref(hf))
                assert rel_l2(df, gold[f"m{id}_df{call}"]) <= 1e-9, (id, call)
            dH = getattr(D, acc)(id, C.byref(hf))
            assert abs(dH - float(gold[f"m{id}_dH"])) <= 1e-7
            e0, e1, i0, i1 = C.c_double(), C.c_double(), C.c_int(), C.c_int()
            assert D.tmb_dropin_monomial_info(id, C.byref(e0), C.byref(e1), C.byref(i0), C.byref(i1)) == 0
            assert abs(e0.value / float(gold[f"m{id}_energy0"]) - 1) <= 1e-13
            assert abs(i1.value - int(gold[f"m{id}_iter1_2"])) <= 3
        # the caller's globals are untouched by the monomials (mnl_backup_restore_globals)
        assert D.glob("g_mu").value == float(gold["gmu"]) and D.glob("g_kappa").value == float(gold["kappa"])
    finally:
        D.close()

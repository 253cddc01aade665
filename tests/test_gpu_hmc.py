"""GPU parity of the HMC side through the C ABI: fermion force deriv_Sb, complex
BLAS-1, chronological guess, solve_degenerate (CG / MIXEDCG / RGMIXEDCG) and the DET / DETRATIO monomials,
against the CPU oracle and the golden fixture of the unmodified reference.  Tolerances: relative L2
<= 1e-13 for direct kernels in double; results that sit behind a CG solve inherit the solve's residual."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, random_gauge, random_spinor, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-13
KAPPA, GMU = 0.16, 0.0032
CG, MIXEDCG, RGMIXEDCG = 1, 13, 14
GOLD = os.path.join(ROOT, "tests", "golden", "ref_hmc_4x4x4x4.npz")


def _setup(oracle_lib, dims, theta, seed=7):
    import tmlqcd_b200 as tm
    rng = np.random.default_rng(seed)
    o = oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g)
    o.set_params(KAPPA, GMU, theta)
    d = tm.Device(*dims)
    d.set_params(KAPPA, GMU, theta)
    d.gauge_upload(g)
    return rng, o, d, g


CASES = [((4, 4, 4, 4), (0., 0., 0., 0.)), ((8, 4, 6, 8), (1., 0.3, 0., 0.7)), ((6, 10, 2, 6), (1., 0., 0., 0.))]


@pytest.mark.parametrize("dims,theta", CASES)
@pytest.mark.parametrize("loopback", [0, 1, 2])
def test_deriv_Sb(oracle_lib, dims, theta, loopback):
    rng, o, d, g = _setup(oracle_lib, dims, theta)
    try:
        if loopback:
            d.ck(d.lib.tmb_comm_loopback(loopback))
            d.gauge_upload(g)
        l, k = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        dl, dk = d.field(l), d.field(k)
        df0 = rng.normal(size=(o.V, 4, 8))
        for ieo in (0, 1):
            exp = df0.copy(); o.deriv_Sb(ieo, l, k, exp, 0.7)
            d.derivative_upload(df0)
            d.call("deriv_Sb", ieo, dl, dk, 0.7)
            got = d.derivative_download()
            assert rel_l2(got - df0, exp - df0) <= TOL
        d.call("derivative_zero")
        assert np.abs(d.derivative_download()).max() == 0.
    finally:
        d.close()


def test_deriv_Sb_golden_reference():
    import tmlqcd_b200 as tm
    gold = np.load(GOLD)
    d = tm.Device(*[int(x) for x in gold["dims"]])
    try:
        d.set_params(float(gold["kappa"]), float(gold["gmu"]), gold["theta"])
        d.gauge_upload(gold["gauge"])
        dl, dk = d.field(gold["l"]), d.field(gold["k"])
        for ieo in (0, 1):
            d.call("derivative_zero")
            d.call("deriv_Sb", ieo, dl, dk, 0.7)
            assert rel_l2(d.derivative_download(), gold[f"deriv_Sb{ieo}"]) <= TOL
    finally:
        d.close()


def test_complex_blas_and_chrono(oracle_lib):
    rng, o, d, g = _setup(oracle_lib, (8, 4, 6, 8), (1., 0.3, 0., 0.7))
    try:
        a, b = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        da, db = d.field(a), d.field(b)
        ca, cb = a.reshape(-1).view(np.complex128), b.reshape(-1).view(np.complex128)
        re, im = C.c_double(), C.c_double()
        d.ck(d.lib.tmb_scalar_prod(da, db, C.byref(re), C.byref(im)))
        exp = np.vdot(ca, cb)  # sum conj(a) b  (linalg/scalar_prod_body.c)
        assert abs(complex(re.value, im.value) - exp) <= 1e-12 * abs(exp) + 1e-10
        c = 0.3 - 0.8j
        d.call("assign_add_mul", da, db, c.real, c.imag)
        assert rel_l2(d.download(da).reshape(-1).view(np.complex128), ca + c * cb) <= 1e-15
        d.call("assign_diff_mul", da, db, c.real, c.imag)
        assert rel_l2(d.download(da).reshape(-1).view(np.complex128), ca) <= 1e-15
        d.call("mul", da, c.real, c.imag, db)
        assert rel_l2(d.download(da).reshape(-1).view(np.complex128), c * cb) <= 1e-15
        # chronological guess: with the exact solution in the history the guess solves the system
        N = 3
        hist = (C.c_void_p * N)(*[d.field() for _ in range(N)])
        idx, n = (C.c_int * N)(), C.c_int(0)
        x = d.field()
        assert d.call("cg_her", x, db, 3000, 1e-24, 0) > 0
        d.ck(d.lib.tmb_chrono_add_solution(d.field(random_spinor(rng, o.Vh)), hist, idx, N, C.byref(n)))
        d.ck(d.lib.tmb_chrono_add_solution(x, hist, idx, N, C.byref(n)))
        assert n.value == 2 and list(idx)[:2] == [0, 1]
        trial, chk = d.field(), d.field()
        d.ck(d.lib.tmb_chrono_guess(trial, db, hist, idx, N, n.value, 0))
        d.call("Qtm_pm_psi", chk, trial)
        assert rel_l2(d.download(chk), b) <= 1e-9
        assert rel_l2(d.download(trial), d.download(x)) <= 1e-9
        # ring buffer behaviour of chrono_add_solution (chrono_guess.c:62-69)
        for _ in range(3):
            d.ck(d.lib.tmb_chrono_add_solution(x, hist, idx, N, C.byref(n)))
        assert n.value == 3 and sorted(idx) == [0, 1, 2]
    finally:
        d.close()


@pytest.mark.parametrize("solver", [CG, MIXEDCG, RGMIXEDCG])
def test_solve_degenerate(oracle_lib, solver):
    rng, o, d, g = _setup(oracle_lib, (8, 8, 8, 8), (1., 0., 0., 0.))
    try:
        b = random_spinor(rng, o.Vh)
        db, dx, chk = d.field(b), d.field(), d.field()
        xr = o.spinor(); itr = o.cg_her(xr, b, 5000, 1e-20, 1)
        it = d.call("solve_degenerate", dx, db, 5000, 1e-20, 1, solver)
        assert it > 0
        if solver == CG:
            assert abs(it - itr) <= 1
        else:  # float inner solver: "at most 1e-5 for any single-precision inner solver" is met by far
            assert it < 3 * itr
        assert rel_l2(d.download(dx), xr) <= 1e-8
        d.call("Qtm_pm_psi", chk, dx)
        assert rel_l2(d.download(chk), b) <= 2e-10
    finally:
        d.close()


def _run_monomials(d, gold, solver, tol_df, tol_pf):
    d.ck(d.lib.tmb_monomial_clear())
    d.ck(d.lib.tmb_set_relative_precision_flag(0))
    for id, (typ, csg_N) in enumerate(gold["monomials"]):
        mid = d.lib.tmb_monomial_add(int(typ), float(gold["kappa"]), float(gold["gmu"]), float(gold["kappa2"]),
                                     float(gold["gmu2"]), solver, 2000, float(gold["forceprec"]), float(gold["accprec"]), int(csg_N))
        assert mid == id, d.lib.tmb_last_error()
        e0 = C.c_double()
        eta = d.field(gold[f"m{id}_eta"])
        d.ck(d.lib.tmb_monomial_heatbath(id, eta, C.byref(e0)))
        assert abs(e0.value / float(gold[f"m{id}_energy0"]) - 1) <= 1e-13
        assert rel_l2(d.download(d.lib.tmb_monomial_pf(id)), gold[f"m{id}_pf"]) <= tol_pf
        d.call("derivative_zero")
        for call in range(3):
            d.ck(d.lib.tmb_monomial_derivative(id))
            assert rel_l2(d.derivative_download(), gold[f"m{id}_df{call}"]) <= tol_df, (id, call)
            if solver == CG:
                assert abs(d.monomial_info(id)["iter1"] - int(gold[f"m{id}_iter1_{call}"])) <= 1 + call, (id, call)
        dH = C.c_double()
        d.ck(d.lib.tmb_monomial_acc(id, C.byref(dH)))
        assert abs(dH.value - float(gold[f"m{id}_dH"])) <= 1e-7
        info = d.monomial_info(id)
        assert info["csg_n"] == int(gold[f"m{id}_csg_n"])
        if solver == CG:
            assert abs(info["iter0"] - int(gold[f"m{id}_iter0"])) <= 2
    d.ck(d.lib.tmb_monomial_clear())


@pytest.mark.parametrize("solver", [CG, RGMIXEDCG])
def test_monomials_golden_reference(solver):
    """det / detratio heatbath, derivative (3 MD steps, with and without chronological guess), acc against
    the unmodified reference's results (tests/golden/ref_hmc_4x4x4x4.npz)"""
    import tmlqcd_b200 as tm
    gold = np.load(GOLD)
    d = tm.Device(*[int(x) for x in gold["dims"]])
    try:
        d.set_params(float(gold["kappa"]), float(gold["gmu"]), gold["theta"])
        d.gauge_upload(gold["gauge"])
        _run_monomials(d, gold, solver, 1e-9, 1e-10)
        # the monomials restore the caller's kappa / mu (mnl_backup_restore_globals, monomial.c:679-708)
        k = d.field(gold["k"]); out = d.field()
        d.call("Hopping_Matrix", 0, out, k)
        assert np.isfinite(d.download(out)).all()
    finally:
        d.close()


@pytest.mark.parametrize("tloop,zloop", [(1, 0), (0, 2), (2, 2), (1, 1)])
def test_monomials_vs_oracle_loopback(oracle_lib, tloop, zloop):
    """a larger lattice, theta != 0, through the T-split (loopback) halo path of hop and deriv_Sb, through the Z-split path
    (faces pushed / copied, fix-up of the hops and of the force) and through both at once"""
    dims, theta = (8, 4, 6, 8), (1., 0.3, 0., 0.7)
    rng, o, d, g = _setup(oracle_lib, dims, theta)
    try:
        if tloop:
            d.ck(d.lib.tmb_comm_loopback(tloop))
        if zloop:
            d.ck(d.lib.tmb_comm_loopback_z(zloop))
        d.gauge_upload(g)
        o.mnl_clear(); d.ck(d.lib.tmb_monomial_clear())
        for id, (typ, csg_N) in enumerate(((0, 2), (1, 1))):
            args = (typ, 0.15, 0.01, 0.15, 0.05, CG, 3000, 1e-20, 1e-22, csg_N)
            assert o.mnl_add(*args) == id and d.lib.tmb_monomial_add(*args) == id
            eta = random_spinor(rng, o.Vh)
            e0 = C.c_double()
            d.ck(d.lib.tmb_monomial_heatbath(id, d.field(eta), C.byref(e0)))
            assert abs(e0.value / o.mnl_heatbath(id, eta) - 1) <= 1e-13
            pf = o.spinor(); o.mnl_get_pf(id, pf)
            assert rel_l2(d.download(d.lib.tmb_monomial_pf(id)), pf) <= 1e-9
            dfo = o.derivative(); d.call("derivative_zero")
            for call in range(3):
                o.mnl_derivative(id, dfo); d.ck(d.lib.tmb_monomial_derivative(id))
                assert rel_l2(d.derivative_download(), dfo) <= 1e-8, (id, call)
                assert abs(d.monomial_info(id)["iter1"] - o.mnl_info(id)["iter1"]) <= 1 + call
            dH = C.c_double(); d.ck(d.lib.tmb_monomial_acc(id, C.byref(dH)))
            assert abs(dH.value - o.mnl_acc(id)) <= 1e-7
        o.mnl_clear()
    finally:
        d.close()


def test_full_size_force_and_doublet_properties(oracle_lib):
    """BASELINE sizes: deriv_Sb at 24^3x48 (linearity, anti-symmetry of the accumulation, oracle on the full lattice),
    the fused two-flavour Qtm_pm_ndpsi against the unfused composition (T-split loopback path) at 32^3x64,
    hermiticity / positivity of Qtm_pm_ndpsi and a cg_her_nd solve checked by its true residual."""
    import tmlqcd_b200 as tm
    dims = (48, 24, 24, 24)
    rng = np.random.default_rng(9)
    V = int(np.prod(dims)); Vh = V // 2
    g = random_gauge(rng, V)
    d = tm.Device(*dims)
    try:
        d.set_params(KAPPA, GMU, (1., 0., 0., 0.))
        d.gauge_upload(g)
        l1, l2, k = (random_spinor(rng, Vh) for _ in range(3))
        dl1, dl2, dk, ds = d.field(l1), d.field(l2), d.field(k), d.field(l1 + 0.5 * l2)

        def force(ieo, fl, fk, factor):
            d.call("derivative_zero"); d.call("deriv_Sb", ieo, fl, fk, factor); return d.derivative_download()
        f1, f2, fs = force(0, dl1, dk, 1.0), force(0, dl2, dk, 1.0), force(0, ds, dk, 1.0)
        assert rel_l2(fs, f1 + 0.5 * f2) <= TOL                      # anti-linear in l ... with real coefficients: linear
        d.call("deriv_Sb", 0, ds, dk, -1.0)                           # accumulates: the same call with -factor cancels
        assert np.abs(d.derivative_download()).max() <= 1e-12 * np.abs(fs).max()
        o = oracle_lib.Oracle(*dims)
        o.set_gauge(g); o.set_params(KAPPA, GMU, (1., 0., 0., 0.))
        exp = o.derivative(); o.deriv_Sb(0, l1, k, exp, 1.0)
        assert rel_l2(f1, exp) <= TOL
    finally:
        d.close()
    dims = (64, 32, 32, 32)
    V = int(np.prod(dims)); Vh = V // 2
    g = random_gauge(rng, V)
    d = tm.Device(*dims)
    try:
        d.set_params(KAPPA, GMU)
        d.ck(d.lib.tmb_set_nd(0.139, 0.15, 0.9))
        d.gauge_upload(g)
        a, b = random_spinor(rng, Vh), random_spinor(rng, Vh)
        da, db, ls, lc, ms, mc = d.field(a), d.field(b), d.field(), d.field(), d.field(), d.field()
        d.call("Qtm_pm_ndpsi", ls, lc, da, db)
        fused = d.download(ls), d.download(lc)
        lhs = d.reduce("scalar_prod_r", da, ls) + d.reduce("scalar_prod_r", db, lc)
        assert lhs > 0                                                # positive
        d.call("Qtm_pm_ndpsi", ms, mc, db, da)                        # hermitian: <(b,a), Q (a,b)> == <Q (b,a), (a,b)>
        x = d.reduce("scalar_prod_r", db, ls) + d.reduce("scalar_prod_r", da, lc)
        y = d.reduce("scalar_prod_r", ms, da) + d.reduce("scalar_prod_r", mc, db)
        assert abs(x - y) <= 1e-11 * abs(lhs)
        # cg_her_nd: true residual by the operator itself
        d.call("field_zero", ms); d.call("field_zero", mc)
        it = d.call("cg_her_nd", ms, mc, da, db, 5000, 1e-16, 1)
        assert it > 0
        d.call("Qtm_pm_ndpsi", ls, lc, ms, mc)
        rr = np.sum((d.download(ls) - a) ** 2) + np.sum((d.download(lc) - b) ** 2)
        assert rr <= 4e-16 * (np.sum(a ** 2) + np.sum(b ** 2))
        # the same operator through the unfused composition (8 hops + sweeps: what the T-split path runs)
        d.ck(d.lib.tmb_comm_loopback(1)); d.gauge_upload(g)
        d.call("Qtm_pm_ndpsi", ls, lc, da, db)
        assert rel_l2(d.download(ls), fused[0]) <= TOL and rel_l2(d.download(lc), fused[1]) <= TOL
    finally:
        d.close()

"""GPU parity tests: the CUDA path, called through the C ABI (include/tmlqcd_b200.h and the
reference-named drop-in symbols of include/tmlqcd_b200_dropin.h), against the CPU oracle on the
same seeded inputs.  Tolerance from BASELINE.json north_star: relative L2 <= 1e-13 in double,
CG iteration count within +-1 of the reference recurrence.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import random_gauge, random_spinor, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-13
KAPPA, GMU = 0.16, 0.0032
ND = (0.139, 0.15, 0.9)


def _setup(oracle_lib, dims, theta, seed=7):
    import tmlqcd_b200 as tm
    rng = np.random.default_rng(seed)
    o = oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g)
    o.set_params(KAPPA, GMU, theta)
    o.set_nd_params(*ND)
    d = tm.Device(*dims)
    d.set_params(KAPPA, GMU, theta)
    d.ck(d.lib.tmb_set_nd(*ND))
    d.gauge_upload(g)
    return rng, o, d, g


CASES = [((4, 4, 4, 4), (0., 0., 0., 0.)), ((8, 4, 6, 8), (1., 0.3, 0., 0.7)), ((6, 10, 2, 6), (1., 0., 0., 0.)),
         ((8, 8, 8, 8), (0., 0., 0., 0.)), ((2, 6, 4, 4), (1., 0., 0.5, 0.)),  # T = 2, every slice is a boundary slice
         # lattices whose LY*LZ/2 is a multiple of 32: the hopping kernels traverse them in 2 x 2 x 32 CTA tiles (tmb_geom.h)
         ((4, 4, 8, 16), (1., 0.2, 0., 0.)), ((6, 2, 16, 4), (1., 0., 0., 0.3)), ((2, 4, 8, 8), (1., 0., 0., 0.))]


@pytest.mark.parametrize("dims,theta", CASES)
@pytest.mark.parametrize("loopback", [0, 1, 2])
def test_hopping_and_epilogues(oracle_lib, dims, theta, loopback):
    rng, o, d, g = _setup(oracle_lib, dims, theta)
    try:
        if loopback:  # the T-split halo/boundary kernels, this rank being its own neighbour
            d.ck(d.lib.tmb_comm_loopback(loopback))
            d.gauge_upload(g)
        k, p = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        dk, dp, dl = d.field(k), d.field(p), d.field()
        for ieo in (0, 1):
            exp = o.spinor()
            o.Hopping_Matrix(ieo, exp, k)
            d.call("Hopping_Matrix", ieo, dl, dk)
            assert rel_l2(d.download(dl), exp) <= TOL
            o.tm_times_Hopping_Matrix(ieo, exp, k, 0.9, -0.2)
            d.call("tm_times_Hopping_Matrix", ieo, dl, dk, 0.9, -0.2)
            assert rel_l2(d.download(dl), exp) <= TOL
            o.tm_sub_Hopping_Matrix(ieo, exp, p, k, 1.0, 0.3)
            d.call("tm_sub_Hopping_Matrix", ieo, dl, dp, dk, 1.0, 0.3)
            assert rel_l2(d.download(dl), exp) <= TOL
    finally:
        d.close()


@pytest.mark.parametrize("dims", [(8, 8, 8, 8), (4, 4, 8, 16), (6, 2, 16, 4), (2, 4, 8, 8)])
@pytest.mark.parametrize("loopback", [0, 1, 2])
def test_cta_tile_traversal_changes_nothing(oracle_lib, dims, loopback):
    """tmb_set_tile(1) (default) against tmb_set_tile(0): operators bit for bit, solvers with the same iteration counts -
    one rank, halo buffers and the peer-mode kernel (whose boundary CTAs are found by a different rule when tiled);
    double, float and 12-real links, the one- and the two-flavour kernels"""
    rng, o, d, g = _setup(oracle_lib, dims, (1., 0.1, 0., 0.4))
    try:
        if loopback:
            d.ck(d.lib.tmb_comm_loopback(loopback))
            d.gauge_upload(g)
        k, p = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        dk, dp, dl, dm = d.field(k), d.field(p), d.field(), d.field()
        gk, gl = d.field32(k.astype(np.float32)), d.field32()
        res = {}
        for tile in (0, 1):
            d.ck(d.lib.tmb_set_tile(tile))
            r = []
            for comp in (18, 12):
                d.ck(d.lib.tmb_set_compression(comp))
                for ieo in (0, 1):
                    d.call("Hopping_Matrix", ieo, dl, dk); r.append(d.download(dl))
                    d.call("tm_sub_Hopping_Matrix", ieo, dl, dp, dk, 1.0, 0.3); r.append(d.download(dl))
                    d.call("Hopping_Matrix_32", ieo, gl, gk); r.append(d.download32(gl))
                d.call("Qtm_pm_psi", dl, dk); r.append(d.download(dl))
            d.ck(d.lib.tmb_set_compression(18))
            d.call("Qtm_pm_ndpsi", dl, dm, dk, dp); r.append(d.download(dl)); r.append(d.download(dm))
            d.call("field_zero", dl)
            it = d.call("cg_her", dl, dk, 2000, 1e-20, 1)
            x = d.download(dl)
            d.call("field_zero", dl); d.call("field_zero", dm)
            itnd = d.call("cg_her_nd", dl, dm, dk, dp, 2000, 1e-20, 1)
            res[tile] = (r, it, x, itnd, d.download(dl))
        for a, b in zip(res[0][0], res[1][0]):
            assert np.array_equal(a, b)
        assert abs(res[0][1] - res[1][1]) <= 1 and abs(res[0][3] - res[1][3]) <= 1
        assert rel_l2(res[1][2], res[0][2]) <= 1e-12 and rel_l2(res[1][4], res[0][4]) <= 1e-12
        exp = o.spinor()
        o.Hopping_Matrix(1, exp, k)
        assert rel_l2(res[1][0][3], exp) <= TOL      # (18-real, ieo = 1, plain hop) against the oracle
    finally:
        d.close()


@pytest.mark.parametrize("variant", range(0, 11))
@pytest.mark.parametrize("hints,xblock", [(1, 0), (0, 0), (1, 2)])
def test_hopping_kernel_variants(oracle_lib, variant, hints, xblock):
    """every tuning variant of the plain kernel computes the same thing"""
    rng, o, d, g = _setup(oracle_lib, (4, 8, 6, 8), (1., 0., 0.5, 0.))
    try:
        d.ck(d.lib.tmb_set_tuning(variant, hints, xblock))
        k = random_spinor(rng, o.Vh)
        dk, dl = d.field(k), d.field()
        exp = o.spinor()
        o.Hopping_Matrix(1, exp, k)
        d.call("Hopping_Matrix", 1, dl, dk)
        assert rel_l2(d.download(dl), exp) <= TOL
    finally:
        d.close()


@pytest.mark.parametrize("hints", [1, 0])
def test_residency_448_every_epilogue(oracle_lib, hints):
    """variant 10 (64 threads x 7 CTAs per SM) serves every epilogue of the double kernel, the fused dot included"""
    rng, o, d, g = _setup(oracle_lib, (4, 8, 6, 8), (1., 0., 0.5, 0.))
    try:
        d.ck(d.lib.tmb_set_tuning(10, hints, 0))
        k, p = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        dk, dp, dl = d.field(k), d.field(p), d.field()
        exp = o.spinor()
        o.Hopping_Matrix(1, exp, k); d.call("Hopping_Matrix", 1, dl, dk)
        assert rel_l2(d.download(dl), exp) <= TOL
        o.tm_times_Hopping_Matrix(0, exp, k, 0.3, -0.7); d.call("tm_times_Hopping_Matrix", 0, dl, dk, 0.3, -0.7)
        assert rel_l2(d.download(dl), exp) <= TOL
        o.tm_sub_Hopping_Matrix(1, exp, p, k, 1.0, 0.2); d.call("tm_sub_Hopping_Matrix", 1, dl, dp, dk, 1.0, 0.2)
        assert rel_l2(d.download(dl), exp) <= TOL
        o.Qtm_pm_psi(exp, k); d.call("Qtm_pm_psi", dl, dk)
        assert rel_l2(d.download(dl), exp) <= TOL
        xr = o.spinor()
        itr = o.cg_her(xr, k, 2000, 1e-20, 1)
        d.call("field_zero", dl)
        it = d.call("cg_her", dl, dk, 2000, 1e-20, 1)
        assert abs(it - itr) <= 1 and rel_l2(d.download(dl), xr) <= 1e-9
    finally:
        d.close()


@pytest.mark.parametrize("loopback", [0, 1, 2])
def test_cg_pro_from_second_hop(oracle_lib, loopback):
    """<p, Q+ Q- p> taken as |Q- p|^2 in the epilogue of the second hop (default) against the operand dot product in the
    last hop (tmb_set_overlap bit 4): same iteration count, same solution, double and mixed precision"""
    rng, o, d, g = _setup(oracle_lib, (8, 4, 6, 8), (1., 0.3, 0., 0.7))
    try:
        if loopback:
            d.ck(d.lib.tmb_comm_loopback(loopback)); d.gauge_upload(g)
        k = random_spinor(rng, o.Vh)
        dk, dx = d.field(k), d.field()
        xr = o.spinor()
        itr = o.cg_her(xr, k, 2000, 1e-22, 1)
        out = {}
        for flags in (0, 16, 32, 64, 4, 68):  # 32: x / r update as a separate sweep; 64: <p,Ap> finished on the side stream next to the third hop; 4: no CUDA graphs
            d.ck(d.lib.tmb_set_overlap(flags))
            d.call("field_zero", dx)
            it = d.call("cg_her", dx, dk, 2000, 1e-22, 1)
            out[flags] = (it, d.download(dx))
            assert abs(it - itr) <= 1 and rel_l2(out[flags][1], xr) <= 1e-10
            d.call("field_zero", dx)
            itm = d.call("mixed_cg_her", dx, dk, 2000, 1e-22, 1)
            assert itm > 0 and rel_l2(d.download(dx), xr) <= 1e-8
        assert out[0][0] == out[16][0] and rel_l2(out[0][1], out[16][1]) <= 1e-12
        assert out[0][0] == out[32][0] and rel_l2(out[0][1], out[32][1]) <= 1e-12
        for f in (64, 4, 68):  # the same sums in the same order wherever they are finished: identical solutions
            assert out[0][0] == out[f][0] and np.array_equal(out[0][1], out[f][1])
    finally:
        d.ck(d.lib.tmb_set_overlap(0))
        d.close()


def test_automatic_residency_at_16x16x16x32(oracle_lib):
    """the default (-1) takes the 448-thread kernels at 65536 sites per parity: same numbers"""
    rng, o, d, g = _setup(oracle_lib, (32, 16, 16, 16), (1., 0., 0., 0.))
    try:
        k = random_spinor(rng, o.Vh)
        dk, dl = d.field(k), d.field()
        exp = o.spinor()
        o.Qtm_pm_psi(exp, k); d.call("Qtm_pm_psi", dl, dk)
        assert rel_l2(d.download(dl), exp) <= TOL
    finally:
        d.close()


@pytest.mark.parametrize("flags", [1, 2, 3])
@pytest.mark.parametrize("loopback", [0, 1, 2])
def test_overlap_flags(oracle_lib, flags, loopback):
    """programmatic dependent launch (bit 0) and L2 gauge prefetch (bit 1) change scheduling, not results"""
    rng, o, d, g = _setup(oracle_lib, (8, 4, 6, 8), (1., 0.3, 0., 0.7))
    try:
        if loopback:
            d.ck(d.lib.tmb_comm_loopback(loopback))
            d.gauge_upload(g)
        d.ck(d.lib.tmb_set_overlap(flags))
        k = random_spinor(rng, o.Vh)
        dk, dl, dx = d.field(k), d.field(), d.field()
        exp = o.spinor()
        for ieo in (0, 1):
            o.Hopping_Matrix(ieo, exp, k); d.call("Hopping_Matrix", ieo, dl, dk)
            assert rel_l2(d.download(dl), exp) <= TOL
        # back-to-back dependent launches: each hop consumes the previous one's output
        d.call("Hopping_Matrix", 0, dl, dk); d.call("Hopping_Matrix", 1, dx, dl); d.call("Hopping_Matrix", 0, dl, dx)
        e1, e2 = o.spinor(), o.spinor()
        o.Hopping_Matrix(0, e1, k); o.Hopping_Matrix(1, e2, e1); o.Hopping_Matrix(0, e1, e2)
        assert rel_l2(d.download(dl), e1) <= TOL
        o.Qtm_pm_psi(exp, k); d.call("Qtm_pm_psi", dl, dk)
        assert rel_l2(d.download(dl), exp) <= TOL
        xr = o.spinor(); itr = o.cg_her(xr, k, 2000, 1e-22, 1)
        it = d.call("cg_her", dx, dk, 2000, 1e-22, 1)
        assert abs(it - itr) <= 1 and rel_l2(d.download(dx), xr) <= 1e-10
    finally:
        d.close()


@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (16, 4, 4, 4), (34, 2, 4, 4), (6, 10, 2, 6)])
def test_pipelined_host_hopping(oracle_lib, dims):
    """tmb_Hopping_Matrix_host: chunked upload / compute / download gives the same field"""
    rng, o, d, g = _setup(oracle_lib, dims, (1., 0., 0.5, 0.))
    try:
        k = random_spinor(rng, o.Vh)
        out, exp = np.zeros_like(k), o.spinor()
        for ieo in (0, 1):
            d.call("Hopping_Matrix_host", ieo, out, k, 0, 1., 0.)
            o.Hopping_Matrix(ieo, exp, k)
            assert rel_l2(out, exp) <= TOL
            d.call("Hopping_Matrix_host", ieo, out, k, 1, 0.9, -0.2)
            o.tm_times_Hopping_Matrix(ieo, exp, k, 0.9, -0.2)
            assert rel_l2(out, exp) <= TOL
    finally:
        d.close()


@pytest.mark.parametrize("dims,theta", CASES[:3])
def test_composite_operators(oracle_lib, dims, theta):
    rng, o, d, g = _setup(oracle_lib, dims, theta)
    try:
        k, p = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        dk, dp, dl, dm = d.field(k), d.field(p), d.field(), d.field()
        for name in ("Qtm_pm_psi", "Qtm_plus_psi", "Qtm_minus_psi", "Mtm_plus_psi", "Mtm_minus_psi"):
            exp = o.spinor()
            getattr(o, name)(exp, k)
            d.call(name, dl, dk)
            assert rel_l2(d.download(dl), exp) <= TOL, name
        e1, e2 = o.spinor(), o.spinor()
        o.M_full(e1, e2, k, p)
        d.call("M_full", dl, dm, dk, dp)
        assert rel_l2(d.download(dl), e1) <= TOL and rel_l2(d.download(dm), e2) <= TOL
        o.Q_full(e1, e2, k, p)
        d.call("Q_full", dl, dm, dk, dp)
        assert rel_l2(d.download(dl), e1) <= TOL and rel_l2(d.download(dm), e2) <= TOL
        # D_psi on a lexicographic field == M_full after the eo permutation (qphix_test_Dslash.c:233 recipe)
        lex = random_spinor(rng, o.V)
        explex = o.spinor(o.V)
        o.D_psi(explex, lex)
        d.upload_lexic(dk, dp, lex)
        d.call("D_psi_eo", dl, dm, dk, dp)
        assert rel_l2(d.download_lexic(dl, dm), explex) <= TOL
        # in-place Qtm_minus_psi(l, l) as invert_eo.c:270 uses it
        exp = o.spinor()
        o.Qtm_minus_psi(exp, k)
        d.upload(dk, k)
        d.call("Qtm_minus_psi", dk, dk)
        assert rel_l2(d.download(dk), exp) <= TOL
    finally:
        d.close()


def test_blas1_and_diagonal(oracle_lib):
    dims = (4, 6, 4, 8)
    rng, o, d, g = _setup(oracle_lib, dims, (0., 0., 0., 0.))
    try:
        a, b, c = (random_spinor(rng, o.Vh) for _ in range(3))
        da, db, dc = d.field(a), d.field(b), d.field(c)
        n = o.Vh
        assert abs(d.reduce("square_norm", da) / o.square_norm(a, n) - 1) < 1e-14
        assert abs(d.reduce("scalar_prod_r", da, db) - o.scalar_prod_r(a, b, n)) < 1e-11
        x = a.copy(); o.assign_add_mul_r(x, b, 0.37, n); d.call("assign_add_mul_r", da, db, 0.37)
        assert rel_l2(d.download(da), x) <= TOL
        y = x.copy(); o.assign_mul_add_r(y, -1.3, b, n); d.call("assign_mul_add_r", da, -1.3, db)
        assert rel_l2(d.download(da), y) <= TOL
        z = y.copy(); e = o.assign_mul_add_r_and_square(z, 0.6, c, n)
        got = d.reduce("assign_mul_add_r_and_square", da, 0.6, dc)
        assert rel_l2(d.download(da), z) <= TOL and abs(got / e - 1) < 1e-14
        w = o.spinor(); o.diff(w, b, c, n); d.call("diff", da, db, dc); assert rel_l2(d.download(da), w) <= TOL
        o.add(w, b, c, n); d.call("add", da, db, dc); assert rel_l2(d.download(da), w) <= TOL
        o.mul_r(w, 2.5, b, n); d.call("mul_r", da, 2.5, db); assert rel_l2(d.download(da), w) <= TOL
        o.gamma5(w, b, n); d.call("gamma5", da, db); assert np.array_equal(d.download(da), w)
        d.call("assign", da, db); assert np.array_equal(d.download(da), b)
        for sign in (+1., -1.):
            o.assign_mul_one_pm_imu_inv(w, b, sign, n); d.call("assign_mul_one_pm_imu_inv", da, db, sign)
            assert rel_l2(d.download(da), w) <= TOL
            o.assign_mul_one_pm_imu(w, b, sign, n); d.call("assign_mul_one_pm_imu", da, db, sign)
            assert rel_l2(d.download(da), w) <= TOL
            o.mul_one_pm_imu_sub_mul_gamma5(w, b, c, sign); d.call("mul_one_pm_imu_sub_mul_gamma5", da, db, dc, sign)
            assert rel_l2(d.download(da), w) <= TOL
            o.mul_one_pm_imu_sub_mul(w, b, c, sign, n); d.call("mul_one_pm_imu_sub_mul", da, db, dc, sign)
            assert rel_l2(d.download(da), w) <= TOL
    finally:
        d.close()


@pytest.mark.parametrize("dims,theta,loopback", [((8, 4, 4, 4), (0., 0., 0., 0.), 0), ((8, 8, 8, 8), (1., 0., 0., 0.), 0),
                                                 ((8, 4, 6, 8), (1., 0.3, 0., 0.7), 1)])
def test_cg_and_invert_eo(oracle_lib, dims, theta, loopback):
    rng, o, d, g = _setup(oracle_lib, dims, theta)
    try:
        if loopback:
            d.ck(d.lib.tmb_comm_loopback(loopback))
            d.gauge_upload(g)
        q = random_spinor(rng, o.Vh)
        x_ref = o.spinor()
        it_ref = o.cg_her(x_ref, q, 2000, 1e-22, 1)
        dq, dx = d.field(q), d.field()
        it = d.call("cg_her", dx, dq, 2000, 1e-22, 1)
        assert it_ref > 0 and abs(it - it_ref) <= 1, (it, it_ref)
        assert rel_l2(d.download(dx), x_ref) <= 1e-10  # both solved to 1e-11 relative residual
        # absolute precision + non-convergence return value (cg_her.c:141)
        assert d.call("cg_her", dx, dq, 3, 1e-30, 0) == -1
        # invert_eo CG branch
        E, O = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        en_r, on_r = o.spinor(), o.spinor()
        it_ref = o.invert_eo_cg(en_r, on_r, E, O, 1e-22, 2000, 1)
        dE, dO, dEn, dOn = d.field(E), d.field(O), d.field(), d.field()
        it = d.call("invert_eo", dEn, dOn, dE, dO, 1e-22, 2000, 1)
        assert abs(it - it_ref) <= 1
        en, on = d.download(dEn), d.download(dOn)
        assert rel_l2(en, en_r) <= 1e-10 and rel_l2(on, on_r) <= 1e-10
        # the reference's own end-to-end check: |M_full x - b|^2 with the CPU operator (operator.c:358-384)
        r1, r2 = o.spinor(), o.spinor()
        o.M_full(r1, r2, en, on)
        res = np.linalg.norm(r1 - E) ** 2 + np.linalg.norm(r2 - O) ** 2
        assert res <= 1e-18 * (np.linalg.norm(E) ** 2 + np.linalg.norm(O) ** 2)
    finally:
        d.close()


@pytest.mark.parametrize("dims,variant", [((4, 4, 6, 8), 0), ((4, 4, 6, 8), 1), ((2, 6, 2, 6), 1), ((4, 4, 6, 8), 2), ((2, 6, 2, 6), 2)])
def test_nd_doublet(oracle_lib, dims, variant):
    """variant 2: two flavour groups of warps per CTA (default), 0: both flavours in one thread, 1: lane-paired two-flavour kernel;
    2x6x2x6: 72 sites per parity - the last warp of the lane-paired kernel is half filled, the last CTA of the warp-grouped one too"""
    rng, o, d, g = _setup(oracle_lib, dims, (1., 0., 0., 0.))
    try:
        d.ck(d.lib.tmb_set_hop2_variant(variant))
        s, c, q, w = (random_spinor(rng, o.Vh) for _ in range(4))
        ds, dc, dls, dlc = d.field(s), d.field(c), d.field(), d.field()
        for name in ("Qtm_ndpsi", "Qtm_dagger_ndpsi", "Qtm_pm_ndpsi"):
            e1, e2 = o.spinor(), o.spinor()
            getattr(o, name)(e1, e2, s, c)
            d.call(name, dls, dlc, ds, dc)
            assert rel_l2(d.download(dls), e1) <= TOL and rel_l2(d.download(dlc), e2) <= TOL, name
        e1, e2 = o.spinor(), o.spinor()
        it_ref = o.cg_her_nd(e1, e2, s, c, 2000, 1e-20, 1)
        d.call("field_zero", dls); d.call("field_zero", dlc)
        it = d.call("cg_her_nd", dls, dlc, ds, dc, 2000, 1e-20, 1)
        assert abs(it - it_ref) <= 1
        assert rel_l2(d.download(dls), e1) <= 1e-9 and rel_l2(d.download(dlc), e2) <= 1e-9
        A = [o.spinor() for _ in range(4)]
        it_ref = o.invert_doublet_eo_cg(*A, s, c, q, w, 1e-20, 2000, 1)
        dq, dw = d.field(q), d.field(w)
        outs = [d.field() for _ in range(4)]
        # argument order: Even_new_s, Odd_new_s, Even_new_c, Odd_new_c, Even_s, Odd_s, Even_c, Odd_c
        it = d.call("invert_doublet_eo", *outs, ds, dc, dq, dw, 1e-20, 2000, 1)
        assert abs(it - it_ref) <= 1
        for f, a in zip(outs, A):
            assert rel_l2(d.download(f), a) <= 1e-9
    finally:
        d.close()


@pytest.mark.parametrize("loopback", [0, 1, 2])
@pytest.mark.parametrize("compression", [18, 12])
def test_nd_two_flavour_kernel_every_mode(oracle_lib, loopback, compression):
    """hop_kernel with NFL = 2 (default two-flavour path): plain, halo-buffer and peer-mode T split, 18- and 12-real links;
    the CG takes <p, A p> from the second launch (invmaxev^2 |Qhat^dagger p|^2): same counts as the oracle's cg_her_nd"""
    rng, o, d, g = _setup(oracle_lib, (8, 4, 6, 8), (1., 0.3, 0., 0.7))
    try:
        if loopback:
            d.ck(d.lib.tmb_comm_loopback(loopback)); d.gauge_upload(g)
        d.ck(d.lib.tmb_set_compression(compression))
        d.ck(d.lib.tmb_set_hop2_variant(2))  # the automatic choice would take round 1's kernel on one rank with 18-real links
        s, c = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        ds, dc, dls, dlc = d.field(s), d.field(c), d.field(), d.field()
        for name in ("Qtm_ndpsi", "Qtm_dagger_ndpsi", "Qtm_pm_ndpsi"):
            e1, e2 = o.spinor(), o.spinor()
            getattr(o, name)(e1, e2, s, c)
            d.call(name, dls, dlc, ds, dc)
            assert rel_l2(d.download(dls), e1) <= TOL and rel_l2(d.download(dlc), e2) <= TOL, name
        d.call("Qtm_pm_ndpsi", ds, dc, ds, dc)  # l may alias k (tm_operators_nd.c:185)
        assert rel_l2(d.download(ds), e1) <= TOL and rel_l2(d.download(dc), e2) <= TOL
        d.upload(ds, s); d.upload(dc, c)
        e1, e2 = o.spinor(), o.spinor()
        it_ref = o.cg_her_nd(e1, e2, s, c, 2000, 1e-20, 1)
        it = d.call("cg_her_nd", dls, dlc, ds, dc, 2000, 1e-20, 1)
        assert abs(it - it_ref) <= 1
        assert rel_l2(d.download(dls), e1) <= 1e-9 and rel_l2(d.download(dlc), e2) <= 1e-9
    finally:
        d.close()


@pytest.mark.parametrize("loopback", [0, 2])
def test_rg_mixed_cg_her_nd_vs_reference(oracle_lib, loopback):
    """Qtm_pm_ndpsi_32 and rg_mixed_cg_her_nd (the RGMIXEDCG branch of invert_doublet_eo.c:145-149) against the unmodified
    reference's results (tests/golden/ref_ndmixed_4x4x4x4.npz): float operator <= 1e-5, count within +-1 outer / a few inner
    iterations of the reference's (float rounding decides single inner steps), solution to the solve's precision"""
    import tmlqcd_b200 as tm
    gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_ndmixed_4x4x4x4.npz"))
    dims = tuple(int(x) for x in gold["dims"])
    d = tm.Device(*dims)
    try:
        d.set_params(float(gold["kappa"]), float(gold["gmu"]), gold["theta"])
        d.ck(d.lib.tmb_set_nd(*[float(x) for x in gold["nd"]]))
        if loopback:
            d.ck(d.lib.tmb_comm_loopback(loopback))
        d.gauge_upload(gold["gauge"])
        s, c = np.array(gold["s"]), np.array(gold["c"])
        a32, b32, l32, m32 = d.field32(s.astype(np.float32)), d.field32(c.astype(np.float32)), d.field32(), d.field32()
        d.call("Qtm_pm_ndpsi_32", l32, m32, a32, b32)
        assert rel_l2(d.download32(l32).astype(np.float64), gold["Qtm_pm_ndpsi_32_s"].astype(np.float64)) <= 1e-5
        assert rel_l2(d.download32(m32).astype(np.float64), gold["Qtm_pm_ndpsi_32_c"].astype(np.float64)) <= 1e-5
        d.ck(d.lib.tmb_set_mcg_delta(float(gold["delta"])))
        ds, dc, dx, dy = d.field(s), d.field(c), d.field(), d.field()
        it = d.call("rg_mixed_cg_her_nd", dx, dy, ds, dc, 2000, float(gold["eps_sq"]), int(gold["rel_prec"]))
        assert it > 0 and abs(it - int(gold["count"])) <= 3, (it, int(gold["count"]))
        assert rel_l2(d.download(dx), gold["x_s"]) <= 1e-8 and rel_l2(d.download(dy), gold["x_c"]) <= 1e-8
        # the true residual of the returned solution, with the double-precision operator
        d.call("Qtm_pm_ndpsi", dx, dy, dx, dy)
        rs, rc = d.download(dx) - s, d.download(dy) - c
        assert np.sum(rs ** 2) + np.sum(rc ** 2) <= float(gold["eps_sq"]) * (np.sum(s ** 2) + np.sum(c ** 2)) * 1.01
        # invert_doublet_eo with solver_flag RGMIXEDCG against the oracle's CG solution
        o = oracle_lib.Oracle(*dims)
        o.set_gauge(gold["gauge"]); o.set_params(float(gold["kappa"]), float(gold["gmu"]), gold["theta"]); o.set_nd_params(*gold["nd"])
        rng = np.random.default_rng(4)
        q, w = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        A = [o.spinor() for _ in range(4)]
        o.invert_doublet_eo_cg(*A, s, c, q, w, 1e-20, 2000, 1)
        dq, dw = d.field(q), d.field(w)
        outs = [d.field() for _ in range(4)]
        it = d.call("invert_doublet_eo_solver", *outs, ds, dc, dq, dw, 1e-20, 2000, 1, 14)
        assert it > 0
        for f, a in zip(outs, A):
            assert rel_l2(d.download(f), a) <= 1e-8
    finally:
        d.close()


@pytest.mark.parametrize("loopback", [0, 1, 2])
def test_host_pointer_hop_pipeline(oracle_lib, loopback):
    """Hopping_Matrix / tm_times_Hopping_Matrix with HOST buffers: chunked full-duplex pipeline replayed as a cached CUDA
    graph (one rank) or with the boundary slices behind a face exchange (split T, here against itself); pageable numpy
    buffers are page-locked on first sight; a parameter change (g_mu, kappa, theta) must not replay a stale graph"""
    import tmlqcd_b200 as tm
    dims = (8, 8, 8, 8)
    rng = np.random.default_rng(17)
    o = oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g)
    D = tm.DropIn(*dims)
    try:
        if loopback:
            assert D.lib.tmb_comm_loopback(loopback) == 0
        D.set_gauge(g)
        k = random_spinor(rng, o.Vh); l = np.zeros_like(k); exp = o.spinor()
        for theta, kappa in (((1., 0.3, 0., 0.7), KAPPA), ((0., 0., 0., 0.), 0.12)):
            o.set_params(kappa, GMU, theta); D.set_params(kappa, GMU, theta)
            for rep in range(3):  # first call captures, the others replay
                for ieo in (0, 1):
                    o.Hopping_Matrix(ieo, exp, k); l[:] = 0; D.Hopping_Matrix(ieo, l, k)
                    assert rel_l2(l, exp) <= TOL, (theta, rep, ieo)
            o.tm_times_Hopping_Matrix(1, exp, k, 0.9, -0.2); D.tm_times_Hopping_Matrix(1, l, k, 0.9, -0.2)
            assert rel_l2(l, exp) <= TOL
        k2 = random_spinor(rng, o.Vh)  # another buffer, other contents: its own graph
        o.Hopping_Matrix(0, exp, k2); D.Hopping_Matrix(0, l, k2); assert rel_l2(l, exp) <= TOL
        for nch in (1, 3, 8):
            assert D.lib.tmb_set_host_chunks(nch) == 0
            o.Hopping_Matrix(1, exp, k); D.Hopping_Matrix(1, l, k); assert rel_l2(l, exp) <= TOL, nch
        assert D.lib.tmb_set_host_chunks(0) == 0
    finally:
        D.close()


def test_reductions_at_48x48x48x96_against_compensated_sums():
    """The reference's square_norm / scalar_prod_r are Kahan sums (linalg/scalar_prod_r.c:159-194, square_norm.c); the device
    sums in a fixed tree (thread -> warp -> CTA -> one CTA over the partials).  At the largest BASELINE volume (5.3 M sites per
    parity, 127 M doubles) both must agree with an extended-precision sum far below what changes a CG iteration count."""
    import tmlqcd_b200 as tm
    dims = (96, 48, 48, 48)
    d = tm.Device(*dims)
    try:
        rng = np.random.default_rng(96)
        a = rng.normal(scale=np.sqrt(0.5), size=(d.Vh, 24)); b = rng.normal(scale=np.sqrt(0.5), size=(d.Vh, 24))
        b += 0.25 * a  # a scalar product that does not average to zero
        da, db = d.field(a), d.field(b)
        n2 = d.reduce("square_norm", da); sp = d.reduce("scalar_prod_r", da, db)
        ld = np.longdouble
        n2_ref = float(np.sum(np.square(a, dtype=ld), dtype=ld)); sp_ref = float(np.sum(np.multiply(a, b, dtype=ld), dtype=ld))
        assert abs(n2 / n2_ref - 1) <= 1e-14, abs(n2 / n2_ref - 1)
        assert abs(sp / sp_ref - 1) <= 1e-13, abs(sp / sp_ref - 1)
        # the float fields of the mixed solvers accumulate in double on the device
        a32 = d.field32(a.astype(np.float32))
        n32 = d.reduce("square_norm_32", a32)
        n32_ref = float(np.sum(np.square(a.astype(np.float32), dtype=ld), dtype=ld))
        assert abs(n32 / n32_ref - 1) <= 1e-13
    finally:
        d.close()


@pytest.mark.parametrize("zmode", [1, 2])  # 1: packed faces + copy (the NCCL path's stand-in), 2: faces pushed through peer memory + flags
@pytest.mark.parametrize("tloop", [0, 1, 2])
def test_z_split_against_itself(oracle_lib, tloop, zmode):
    """Second split direction (Z) with this rank as its own z neighbour (the real two-slab arithmetic is checked on the CPU
    through the device code, tests/test_device_code_emul.py::test_z_split_face_exchange_and_fixup, and on several GPUs by
    scripts/mgpu_parity.py --grid): face pack, exchange, fix-up, un-fused solver reductions, alone and on top of the T split"""
    rng, o, d, g = _setup(oracle_lib, (8, 4, 6, 8), (1., 0.3, 0., 0.7))
    try:
        if tloop:
            d.ck(d.lib.tmb_comm_loopback(tloop))
        d.ck(d.lib.tmb_comm_loopback_z(zmode)); d.gauge_upload(g)
        k, p = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        dk, dp, dl, dx = d.field(k), d.field(p), d.field(), d.field()
        exp = o.spinor()
        for ieo in (0, 1):
            o.Hopping_Matrix(ieo, exp, k); d.call("Hopping_Matrix", ieo, dl, dk); assert rel_l2(d.download(dl), exp) <= TOL
            o.tm_times_Hopping_Matrix(ieo, exp, k, 0.9, -0.2); d.call("tm_times_Hopping_Matrix", ieo, dl, dk, 0.9, -0.2)
            assert rel_l2(d.download(dl), exp) <= TOL
            o.tm_sub_Hopping_Matrix(ieo, exp, p, k, 1.0, 0.3); d.call("tm_sub_Hopping_Matrix", ieo, dl, dp, dk, 1.0, 0.3)
            assert rel_l2(d.download(dl), exp) <= TOL
        o.Qtm_pm_psi(exp, k); d.call("Qtm_pm_psi", dl, dk); assert rel_l2(d.download(dl), exp) <= TOL
        e1, e2 = o.spinor(), o.spinor(); o.M_full(e1, e2, k, p)
        dm = d.field(); d.call("M_full", dl, dm, dk, dp)
        assert rel_l2(d.download(dl), e1) <= TOL and rel_l2(d.download(dm), e2) <= TOL
        k32 = k.astype(np.float32); dk32, dl32 = d.field32(k32), d.field32()
        o.Hopping_Matrix(1, exp, k); d.call("Hopping_Matrix_32", 1, dl32, dk32)
        assert rel_l2(d.download32(dl32).astype(np.float64), exp) <= 1e-5
        xr = o.spinor(); itr = o.cg_her(xr, k, 2000, 1e-22, 1)
        it = d.call("cg_her", dx, dk, 2000, 1e-22, 1)
        assert abs(it - itr) <= 1 and rel_l2(d.download(dx), xr) <= 1e-10
        it = d.call("mixed_cg_her", dx, dk, 2000, 1e-22, 1)
        assert it > 0 and rel_l2(d.download(dx), xr) <= 1e-9
        en, on = o.spinor(), o.spinor(); itr = o.invert_eo_cg(en, on, k, p, 1e-22, 2000, 1)
        dEn, dOn = d.field(), d.field()
        it = d.call("invert_eo", dEn, dOn, dk, dp, 1e-22, 2000, 1)
        assert abs(it - itr) <= 1 and rel_l2(d.download(dEn), en) <= 1e-10 and rel_l2(d.download(dOn), on) <= 1e-10
        o.set_nd_params(*ND); d.ck(d.lib.tmb_set_nd(*ND))
        es, ec = o.spinor(), o.spinor(); o.Qtm_pm_ndpsi(es, ec, k, p)
        dls, dlc = d.field(), d.field(); d.call("Qtm_pm_ndpsi", dls, dlc, dk, dp)
        assert rel_l2(d.download(dls), es) <= TOL and rel_l2(d.download(dlc), ec) <= TOL
        if zmode == 2:  # the pushes count on their own: consecutive pushes alternate the halo buffers whatever the T path takes
            c0, z0, c1, z1 = C.c_uint(), C.c_uint(), C.c_uint(), C.c_uint()
            d.ck(d.lib.tmb_comm_sequence_counts(C.byref(c0), C.byref(z0)))
            d.call("Qtm_pm_psi", dl, dk)
            d.ck(d.lib.tmb_comm_sequence_counts(C.byref(c1), C.byref(z1)))
            assert z1.value - z0.value == 4 and c1.value - c0.value == (4 if tloop == 2 else 0)
        # host-pointer hop on a split Z: plain upload / compute / download
        out = np.zeros_like(k); d.call("Hopping_Matrix_host", 0, out, k, 0, 1., 0.)
        o.Hopping_Matrix(0, exp, k); assert rel_l2(out, exp) <= TOL
        # the fermion force: slab kernels + fix-up of the z links of the last-z sites (deriv_Sb.c:402-649, xchange_2fields)
        df = o.derivative(); o.deriv_Sb(0, k, p, df, 0.7); o.deriv_Sb(1, p, k, df, -0.4)
        d.call("derivative_zero"); d.call("deriv_Sb", 0, dk, dp, 0.7); d.call("deriv_Sb", 1, dp, dk, -0.4)
        assert rel_l2(d.derivative_download(), df) <= TOL
        # T-split-only entry points refuse loudly
        import tmlqcd_b200 as tm
        plaq = C.c_double(0.)
        assert d.lib.tmb_measure_plaquette(C.byref(plaq)) < 0 and "split Z" in d.lib.tmb_last_error().decode()
    finally:
        d.close()


def test_dropin_reference_symbols(oracle_lib):
    """the reference-named entry points with host buffers (what a tmLQCD executable links)"""
    import tmlqcd_b200 as tm
    dims, theta = (8, 4, 4, 6), (1., 0., 0.25, 0.)
    rng = np.random.default_rng(3)
    o = oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g); o.set_params(KAPPA, GMU, theta); o.set_nd_params(*ND)
    D = tm.DropIn(*dims)
    try:
        D.set_params(KAPPA, GMU, theta); D.set_nd_params(*ND); D.set_gauge(g)
        k, p = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        l, exp = D.spinor(), o.spinor()
        for ieo in (0, 1):
            D.Hopping_Matrix(ieo, l, k); o.Hopping_Matrix(ieo, exp, k); assert rel_l2(l, exp) <= TOL
            D.tm_times_Hopping_Matrix(ieo, l, k, 0.9, -0.2); o.tm_times_Hopping_Matrix(ieo, exp, k, 0.9, -0.2)
            assert rel_l2(l, exp) <= TOL
            D.tm_sub_Hopping_Matrix(ieo, l, p, k, 1.0, 0.3); o.tm_sub_Hopping_Matrix(ieo, exp, p, k, 1.0, 0.3)
            assert rel_l2(l, exp) <= TOL
        for name in ("Qtm_pm_psi", "Qtm_plus_psi", "Qtm_minus_psi", "Mtm_plus_psi", "Mtm_minus_psi"):
            getattr(D, name)(l, k); getattr(o, name)(exp, k); assert rel_l2(l, exp) <= TOL, name
        # callers flip the global g_mu around calls (tm_operators.c:382-386): re-read at every call
        D.glob("g_mu").value = -GMU; o.set_params(KAPPA, -GMU, theta)
        D.Qtm_plus_psi(l, k); o.Qtm_plus_psi(exp, k); assert rel_l2(l, exp) <= TOL
        D.glob("g_mu").value = GMU; o.set_params(KAPPA, GMU, theta)
        # ... and call boundary() with another kappa per monomial (detratio_monomial.c:57-59)
        D.boundary(0.12); o.set_params(0.12, GMU, theta)
        D.Hopping_Matrix(0, l, k); o.Hopping_Matrix(0, exp, k); assert rel_l2(l, exp) <= TOL
        D.boundary(KAPPA); o.set_params(KAPPA, GMU, theta)
        # gauge dirty flag (g_update_gauge_copy)
        g2 = random_gauge(rng, o.V); D.set_gauge(g2); o.set_gauge(g2)
        D.Hopping_Matrix(1, l, k); o.Hopping_Matrix(1, exp, k); assert rel_l2(l, exp) <= TOL
        assert D.glob("g_update_gauge_copy", C.c_int).value == 0
        # lexicographic D_psi and BLAS-1
        lex, outl, expl = random_spinor(rng, o.V), D.spinor(o.V), o.spinor(o.V)
        D.D_psi(outl, lex); o.D_psi(expl, lex); assert rel_l2(outl, expl) <= TOL
        D.Q_pm_psi(outl, lex); o.Q_pm_psi(expl, lex); assert rel_l2(outl, expl) <= TOL
        assert abs(D.square_norm(k, o.Vh, 1) / o.square_norm(k, o.Vh) - 1) < 1e-14
        assert abs(D.square_norm(lex, o.V, 1) / o.square_norm(lex, o.V) - 1) < 1e-14
        assert abs(D.scalar_prod_r(k, p, o.Vh, 1) - o.scalar_prod_r(k, p, o.Vh)) < 1e-11
        a1, a2 = k.copy(), k.copy()
        D.assign_add_mul_r(a1, p, 0.4, o.Vh); o.assign_add_mul_r(a2, p, 0.4, o.Vh); assert rel_l2(a1, a2) <= TOL
        e1 = D.assign_mul_add_r_and_square(a1, -0.7, p, o.Vh, 1); e2 = o.assign_mul_add_r_and_square(a2, -0.7, p, o.Vh)
        assert rel_l2(a1, a2) <= TOL and abs(e1 / e2 - 1) < 1e-14
        # solvers: cg_her dispatches on the identity of f exactly like monomial_solve.c:134
        x, xr = D.spinor(), o.spinor()
        it = D.cg_her(x, k, 2000, 1e-22, 1, o.Vh, D.fptr("Qtm_pm_psi")); itr = o.cg_her(xr, k, 2000, 1e-22, 1)
        assert abs(it - itr) <= 1 and rel_l2(x, xr) <= 1e-10
        # any other f takes the generic path: same recurrence, f applied through its host-pointer entry point
        raw = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)(("Qtm_pm_psi", D.lib))
        cb = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)(lambda l_, k_: raw(l_, k_))
        x[:] = 0; it2 = D.cg_her(x, k, 2000, 1e-22, 1, o.Vh, C.cast(cb, C.c_void_p))
        assert abs(it2 - itr) <= 1 and rel_l2(x, xr) <= 1e-10
        en, on, enr, onr = D.spinor(), D.spinor(), o.spinor(), o.spinor()
        it = D.invert_eo(en, on, k, p, 1e-22, 2000, 1, 1, 0, 1, 0, None, tm.capi.SolverParams(), 0, 0, 0, 18)
        itr = o.invert_eo_cg(enr, onr, k, p, 1e-22, 2000, 1)
        assert abs(it - itr) <= 1 and rel_l2(en, enr) <= 1e-10 and rel_l2(on, onr) <= 1e-10
        # ND
        ls, lc, es, ec = D.spinor(), D.spinor(), o.spinor(), o.spinor()
        D.Qtm_pm_ndpsi(ls, lc, k, p); o.Qtm_pm_ndpsi(es, ec, k, p)
        assert rel_l2(ls, es) <= TOL and rel_l2(lc, ec) <= TOL
        ls[:] = 0; lc[:] = 0; es[:] = 0; ec[:] = 0
        it = D.cg_her_nd(ls, lc, k, p, 2000, 1e-20, 1, o.Vh, D.fptr("Qtm_pm_ndpsi")); itr = o.cg_her_nd(es, ec, k, p, 2000, 1e-20, 1)
        assert abs(it - itr) <= 1 and rel_l2(ls, es) <= 1e-9
    finally:
        D.close()


def test_tmLQCD_facade(oracle_lib):
    """include/tmLQCD.h entry points: lexicographic source in, propagator out (lib_wrapper.c:242-279)"""
    import tmlqcd_b200 as tm
    lib = tm.load()
    dims = (8, 4, 4, 4)
    rng = np.random.default_rng(11)
    o = oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g); o.set_params(KAPPA, GMU, (1., 0., 0., 0.))
    lib.tmLQCD_b200_set_lattice(*dims)
    lib.tmLQCD_b200_set_theta(1., 0., 0., 0.)
    op = lib.tmLQCD_b200_add_operator(KAPPA, GMU, 1e-22, 2000, 1)
    assert lib.tmLQCD_invert_init(0, None, 0, 0) == 0
    try:
        gf = C.POINTER(C.c_double)()
        assert lib.tmLQCD_get_gauge_field_pointer(C.byref(gf)) == 0
        C.memmove(gf, g.ctypes.data, g.nbytes)
        src = random_spinor(rng, o.V)
        prop = np.zeros_like(src)
        assert lib.tmLQCD_invert(prop, src, op, 0) == 0
        it, rp = C.c_int(0), C.c_double(0.)
        lib.tmLQCD_b200_get_solver_info(op, C.byref(it), C.byref(rp))
        # oracle: same pipeline
        E, O, En, On = o.spinor(), o.spinor(), o.spinor(), o.spinor()
        o.convert_lexic_to_eo(E, O, src)
        itr = o.invert_eo_cg(En, On, E, O, 1e-22, 2000, 1)
        exp = o.spinor(o.V)
        o.convert_eo_to_lexic(exp, En * (2 * KAPPA), On * (2 * KAPPA))
        assert abs(it.value - itr) <= 1 and rel_l2(prop, exp) <= 1e-10
        assert rp.value <= 1e-18 * np.linalg.norm(src) ** 2
        assert lib.tmLQCD_read_gauge(0) == -1  # no ./conf.0000 here: refused with a message (lib_wrapper.c:218-221)
    finally:
        assert lib.tmLQCD_finalise() == 0


def test_tmLQCD_facade_solver_keys(oracle_lib, tmp_path, monkeypatch):
    """the operator block's Solver / UseEvenOdd / mcgdelta / SolverRelativePrecision keys of invert.input (read_input.l:1108-1139,
    :967-974, :835-838, :824-833; defaults operator.c:102-125) reach invert_eo's branches through tmLQCD_invert (op_invert,
    operator.c:349-355): every implemented combination gives the propagator of the CG branch; an unimplemented solver is
    refused at tmLQCD_invert_init instead of silently running CG"""
    import tmlqcd_b200 as tm
    lib = tm.load()
    dims = (4, 4, 4, 6)
    rng = np.random.default_rng(12)
    o = oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g); o.set_params(KAPPA, GMU, (1., 0., 0., 0.))
    monkeypatch.chdir(tmp_path)
    blocks = [("cg", "yes", ""), ("mixedcg", "yes", ""), ("rgmixedcg", "yes", "  mcgdelta = 1.e-3\n"), ("cg", "no", "")]
    text = f"T = {dims[0]}\nLX = {dims[1]}\nLY = {dims[2]}\nLZ = {dims[3]}\nThetaT = 1.\n"
    for solver, eo, extra in blocks:
        text += (f"BeginOperator TMWILSON\n  2KappaMu = {GMU}\n  kappa = {KAPPA}\n  Solver = {solver}\n  UseEvenOdd = {eo}\n{extra}"
                 "  SolverPrecision = 1.e-20\n  MaxSolverIterations = 3000\n  SolverRelativePrecision = yes\nEndOperator\n")
    (tmp_path / "invert.input").write_text(text)
    assert lib.tmLQCD_invert_init(0, None, 0, 0) == 0
    try:
        lp = (C.c_uint * 7)()  # tmLQCD_lat_params: LX, LY, LZ, T, nstore, nsave, no_operators (include/tmLQCD.h:37-44)
        assert lib.tmLQCD_get_lat_params(C.cast(lp, C.c_void_p)) == 0 and lp[6] == len(blocks) and lp[3] == dims[0]
        gf = C.POINTER(C.c_double)()
        assert lib.tmLQCD_get_gauge_field_pointer(C.byref(gf)) == 0
        C.memmove(gf, g.ctypes.data, g.nbytes)
        src = random_spinor(rng, o.V)
        E, O, En, On = o.spinor(), o.spinor(), o.spinor(), o.spinor()
        o.convert_lexic_to_eo(E, O, src)
        itr = o.invert_eo_cg(En, On, E, O, 1e-20, 3000, 1)
        exp = o.spinor(o.V)
        o.convert_eo_to_lexic(exp, En * (2 * KAPPA), On * (2 * KAPPA))
        for op, (solver, eo, _) in enumerate(blocks):
            prop = np.zeros_like(src)
            assert lib.tmLQCD_invert(prop, src, op, 0) == 0
            it, rp = C.c_int(0), C.c_double(0.)
            lib.tmLQCD_b200_get_solver_info(op, C.byref(it), C.byref(rp))
            assert it.value > 0 and rel_l2(prop, exp) <= 1e-7, (solver, eo, it.value)
            assert rp.value <= 1e-14 * np.linalg.norm(src) ** 2, (solver, eo, rp.value)
            if op == 0:
                assert abs(it.value - itr) <= 1
        # the programmatic form of the same keys; combinations invert_eo does not implement are refused
        assert lib.tmLQCD_b200_set_operator_solver(0, 13, 1, 0.) == 0
        assert lib.tmLQCD_b200_set_operator_solver(0, 13, 0, 0.) == -1 and lib.tmLQCD_b200_set_operator_solver(0, 2, 1, 0.) == -1
        assert lib.tmLQCD_b200_set_operator_solver(99, 1, 1, 0.) == -1
    finally:
        assert lib.tmLQCD_finalise() == 0
    (tmp_path / "invert.input").write_text(text.replace("Solver = mixedcg", "Solver = bicgstab"))
    assert lib.tmLQCD_invert_init(0, None, 0, 0) == -1
    assert lib.tmLQCD_b200_set_operator_solver(0, 1, 1, 0.) == -1  # nothing of the refused file stays behind


def test_full_size_properties(oracle_lib):
    """BASELINE config 2 size (24^3 x 48): size-independent properties + a sampled oracle check."""
    import tmlqcd_b200 as tm
    dims = (48, 24, 24, 24)
    rng = np.random.default_rng(5)
    V = int(np.prod(dims)); Vh = V // 2
    g = random_gauge(rng, V)
    d = tm.Device(*dims)
    try:
        d.set_params(KAPPA, GMU, (0., 0., 0., 0.))
        d.gauge_upload(g)
        a, b = random_spinor(rng, Vh), random_spinor(rng, Vh)
        da, db, dl, dm = d.field(a), d.field(b), d.field(), d.field()
        # linearity: H(a + 0.5 b) == H a + 0.5 H b
        d.call("Hopping_Matrix", 0, dl, da); Ha = d.download(dl)
        d.call("Hopping_Matrix", 0, dl, db); Hb = d.download(dl)
        d.upload(dm, a + 0.5 * b); d.call("Hopping_Matrix", 0, dl, dm)
        assert rel_l2(d.download(dl), Ha + 0.5 * Hb) <= TOL
        # g5-hermiticity of the hopping term: <b, g5 H_oe g5 a>_odd == <H_eo b, a>  (H_oe^dag = g5 H_eo g5)
        d.call("Qtm_pm_psi", dl, da); d.call("Qtm_pm_psi", dm, db)
        lhs = d.reduce("scalar_prod_r", db, dl); rhs = d.reduce("scalar_prod_r", dm, da)
        assert abs(lhs - rhs) <= 1e-12 * abs(lhs)  # Qtm_pm_psi is hermitian
        assert d.reduce("scalar_prod_r", da, dl) > 0  # ... and positive
        # oracle on the full lattice (a few seconds on the host)
        o = oracle_lib.Oracle(*dims)
        o.set_gauge(g); o.set_params(KAPPA, GMU, (0., 0., 0., 0.))
        exp = o.spinor(); o.Hopping_Matrix(0, exp, a)
        assert rel_l2(Ha, exp) <= TOL
        exp2 = o.spinor(); o.Qtm_pm_psi(exp2, a)
        assert rel_l2(d.download(dl), exp2) <= TOL
    finally:
        d.close()


def test_two_gpu_T_split_nccl():
    """real NCCL halos: 2 ranks, global 16x8x8x8, vs the oracle on the global lattice (scripts/mgpu_parity.py)"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the gloo test in test_multirank_gloo.py covers the logic on CPU)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "scripts", "mgpu_parity.py"),
                        "8x8x8x8"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MGPU PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_single_precision_operator_and_mixed_cg(oracle_lib):
    """row a31: Hopping_Matrix_32 / Qtm_pm_psi_32 (<= 1e-5 of the double operator, north_star) and
    mixed_cg_her (float inner CG, double defect correction) reaching the double residual"""
    import tmlqcd_b200 as tm
    for dims, theta, loopback in (((8, 4, 6, 8), (1., 0.3, 0., 0.7), 0), ((8, 8, 8, 8), (1., 0., 0., 0.), 1)):
        rng, o, d, g = _setup(oracle_lib, dims, theta)
        try:
            if loopback:
                d.ck(d.lib.tmb_comm_loopback(1))
                d.gauge_upload(g)
            k = random_spinor(rng, o.Vh)
            k32 = k.astype(np.float32)
            dk32, dl32 = d.field32(k32), d.field32()
            exp = o.spinor()
            for ieo in (0, 1):
                o.Hopping_Matrix(ieo, exp, k); d.call("Hopping_Matrix_32", ieo, dl32, dk32)
                assert rel_l2(d.download32(dl32).astype(np.float64), exp) <= 1e-5
            o.Qtm_pm_psi(exp, k); d.call("Qtm_pm_psi_32", dl32, dk32)
            assert rel_l2(d.download32(dl32).astype(np.float64), exp) <= 1e-5
            # precision conversion round trip
            dk, dt = d.field(k), d.field()
            d.call("assign_to_32", dl32, dk); d.call("assign_to_64", dt, dl32)
            assert rel_l2(d.download(dt), k32.astype(np.float64)) == 0.0
            # mixed CG vs the double CG of the oracle
            xr = o.spinor(); itr = o.cg_her(xr, k, 2000, 1e-22, 1)
            dx = d.field()
            it = d.call("mixed_cg_her", dx, dk, 2000, 1e-22, 1)
            x = d.download(dx)
            assert it > 0 and it <= 1.3 * itr + 10, (it, itr)
            assert rel_l2(x, xr) <= 1e-9
            r = o.spinor(); o.Qtm_pm_psi(r, x)
            assert np.linalg.norm(r - k) ** 2 <= 1e-22 * np.linalg.norm(k) ** 2 * 1.01
            # invert_eo with solver_flag == MIXEDCG through the reference-named symbol
            E, O = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
            enr, onr = o.spinor(), o.spinor(); o.invert_eo_cg(enr, onr, E, O, 1e-22, 2000, 1)
            dE, dO, dEn, dOn = d.field(E), d.field(O), d.field(), d.field()
            it = d.call("invert_eo_mixed", dEn, dOn, dE, dO, 1e-22, 2000, 1)
            assert it > 0 and rel_l2(d.download(dEn), enr) <= 1e-9 and rel_l2(d.download(dOn), onr) <= 1e-9
        finally:
            d.close()
    # reference-named symbols with host buffers
    D = tm.DropIn(8, 4, 4, 4)
    try:
        rng = np.random.default_rng(21)
        o = oracle_lib.Oracle(8, 4, 4, 4)
        g = random_gauge(rng, o.V)
        o.set_gauge(g); o.set_params(KAPPA, GMU, (1., 0., 0., 0.))
        D.set_params(KAPPA, GMU, (1., 0., 0., 0.)); D.set_gauge(g)
        k = random_spinor(rng, o.Vh); k32 = k.astype(np.float32); l32 = np.zeros_like(k32); exp = o.spinor()
        D.Hopping_Matrix_32(1, l32, k32); o.Hopping_Matrix(1, exp, k); assert rel_l2(l32.astype(np.float64), exp) <= 1e-5
        D.Qtm_pm_psi_32(l32, k32); o.Qtm_pm_psi(exp, k); assert rel_l2(l32.astype(np.float64), exp) <= 1e-5
        x, xr = D.spinor(), o.spinor()
        itr = o.cg_her(xr, k, 2000, 1e-22, 1)
        it = D.mixed_cg_her(x, k, tm.capi.SolverParams(), 2000, 1e-22, 1, o.Vh, D.fptr("Qtm_pm_psi"), D.fptr("Qtm_pm_psi_32"))
        assert it > 0 and rel_l2(x, xr) <= 1e-9
        en, on, enr, onr = D.spinor(), D.spinor(), o.spinor(), o.spinor()
        o.invert_eo_cg(enr, onr, k, xr, 1e-22, 2000, 1)
        it = D.invert_eo(en, on, k, xr, 1e-22, 2000, 13, 1, 0, 1, 0, None, tm.capi.SolverParams(), 0, 0, 0, 18)  # MIXEDCG
        assert it > 0 and rel_l2(en, enr) <= 1e-9 and rel_l2(on, onr) <= 1e-9
    finally:
        D.close()


@pytest.mark.parametrize("loopback", [0, 1, 2])
def test_gauge_compression_12(oracle_lib, loopback):
    """CompressionType 12 (two rows streamed, third rebuilt): same results to 1e-13, refused for non-SU(3) links"""
    import tmlqcd_b200 as tm
    rng, o, d, g = _setup(oracle_lib, (8, 4, 6, 8), (1., 0.3, 0., 0.7))
    try:
        if loopback:
            d.ck(d.lib.tmb_comm_loopback(loopback))
            d.gauge_upload(g)
        d.ck(d.lib.tmb_set_compression(12))
        k, p = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
        dk, dp, dl, dx = d.field(k), d.field(p), d.field(), d.field()
        exp = o.spinor()
        for ieo in (0, 1):
            o.Hopping_Matrix(ieo, exp, k); d.call("Hopping_Matrix", ieo, dl, dk); assert rel_l2(d.download(dl), exp) <= TOL
            o.tm_sub_Hopping_Matrix(ieo, exp, p, k, 1.0, 0.3); d.call("tm_sub_Hopping_Matrix", ieo, dl, dp, dk, 1.0, 0.3)
            assert rel_l2(d.download(dl), exp) <= TOL
        o.Qtm_pm_psi(exp, k); d.call("Qtm_pm_psi", dl, dk); assert rel_l2(d.download(dl), exp) <= TOL
        xr = o.spinor(); itr = o.cg_her(xr, k, 2000, 1e-22, 1)
        it = d.call("cg_her", dx, dk, 2000, 1e-22, 1)
        assert abs(it - itr) <= 1 and rel_l2(d.download(dx), xr) <= 1e-10
        it = d.call("mixed_cg_her", dx, dk, 2000, 1e-22, 1)  # float inner solve on compressed float links
        assert it > 0 and rel_l2(d.download(dx), xr) <= 1e-9
        # a new gauge field keeps the mode; a non-unitary field is refused loudly
        g2 = random_gauge(rng, o.V); d.gauge_upload(g2); o.set_gauge(g2)
        o.Hopping_Matrix(0, exp, k); d.call("Hopping_Matrix", 0, dl, dk); assert rel_l2(d.download(dl), exp) <= TOL
        bad = g2.copy(); bad[5, 2, 13] += 1e-9
        with pytest.raises(tm.capi.TmbError, match="compression refused"):
            d.gauge_upload(bad)
        d.ck(d.lib.tmb_set_compression(18))
        d.gauge_upload(bad); o.set_gauge(bad)
        o.Hopping_Matrix(0, exp, k); d.call("Hopping_Matrix", 0, dl, dk); assert rel_l2(d.download(dl), exp) <= TOL
    finally:
        d.close()

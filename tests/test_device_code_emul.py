"""CPU-only logic tests of the PRODUCT's device code: tmb_site.cuh / tmb_geom.h / the layout functors
of tmb_kernels.cu are __host__ __device__ and are compiled here for the host (tests/emul/), then
compared with the oracle.  Covers what cannot be checked on this GPU-less box otherwise: SoA layout
conversion, closed-form neighbour arithmetic vs the reference's g_hi table, the spin projector
tables and epilogues, the T-slab halo (loopback) path and the x-blocked traversal."""
import os

import numpy as np
import pytest

from conftest import ROOT, random_gauge, random_spinor, rel_l2
from emul_client import Emul

KAPPA, GMU = 0.16, 0.0032


def ka_of(kappa, theta, ext):
    x = np.array(theta) * 3.14159265358979 / np.array(ext, dtype=float)
    return np.stack([kappa * np.cos(x), kappa * np.sin(x)], axis=1).reshape(-1)


@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (4, 6, 4, 8), (2, 2, 2, 2), (6, 2, 10, 4), (2, 2, 2, 4), (10, 2, 4, 6),
                                  (4, 8, 2, 2), (2, 10, 6, 4), (12, 4, 2, 8), (2, 4, 12, 2)])
def test_closed_form_geometry_matches_reference_tables(oracle_lib, dims):
    e = Emul(*dims)
    o = oracle_lib.Oracle(*dims)
    e2l = np.zeros(e.V, dtype=np.int32)
    e.E.emul_eo2lexic(e2l, *dims)
    assert np.array_equal(e2l, o.eo2lexic())
    # neighbour arithmetic (incl. the wrap-around of extents of 2) against a table built from the reference's orderings:
    # lexic ix = ((t LX + x) LY + y) LZ + z (geometry_eo.c:290), eo-sub index from g_lexic2eosub
    T, LX, LY, LZ = dims
    l2e = o.lexic2eosub()
    for par in (0, 1):
        nb = np.zeros(e.Vh * 8, dtype=np.int32)
        e.E.emul_neighbours(nb, par, *dims)
        ix = e2l[par * e.Vh:(par + 1) * e.Vh].astype(np.int64)
        z = ix % LZ; y = (ix // LZ) % LY; x = (ix // (LZ * LY)) % LX; t = ix // (LZ * LY * LX)
        exp = np.zeros((e.Vh, 8), dtype=np.int64)
        for mu, (c, ext) in enumerate(((t, T), (x, LX), (y, LY), (z, LZ))):
            for d, sh in ((0, +1), (1, -1)):
                cc = [t, x, y, z]
                cc[mu] = (c + sh) % ext
                exp[:, 2 * mu + d] = l2e[((cc[0] * LX + cc[1]) * LY + cc[2]) * LZ + cc[3]]
        assert np.array_equal(nb.reshape(e.Vh, 8), exp), par
    if dims == (4, 4, 4, 4):  # the reference's own g_hi table (golden fixture, geometry_eo.c:1470-1536)
        hi = np.load(os.path.join(ROOT, "tests", "golden", "ref_4x4x4x4.npz"))["hi"]
        for par in (0, 1):
            nb = np.zeros(e.Vh * 8, dtype=np.int32)
            e.E.emul_neighbours(nb, par, *dims)
            assert np.array_equal(nb.reshape(e.Vh, 8), hi[par * e.Vh:(par + 1) * e.Vh, 1::2])


@pytest.mark.parametrize("dims,theta", [((4, 4, 4, 4), (0., 0., 0., 0.)), ((4, 6, 4, 8), (1., 0.3, 0., 0.7)),
                                        ((2, 4, 2, 6), (1., 0., 0., 0.)), ((6, 2, 10, 4), (0., 0.5, 0.5, 1.))])
def test_hop_all_modes_and_halo_loopback(oracle_lib, dims, theta):
    rng = np.random.default_rng(5)
    e, o = Emul(*dims), oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g); o.set_params(KAPPA, GMU, theta)
    ka = ka_of(KAPPA, theta, dims)
    U = e.pack_gauge(g)
    k, p = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
    sk, sp = e.pack(k), e.pack(p)
    assert np.array_equal(e.unpack(sk), k)
    up, dn = e.pack_halo(sk)
    halo = (dn, up, e.pack_gauge_halo(U))  # loopback: halo_up <- own send_dn, halo_dn <- own send_up
    for par in (0, 1):
        exp = o.spinor()
        for h in (None, halo):
            o.Hopping_Matrix(par, exp, k)
            assert rel_l2(e.unpack(e.hop(par, sk, U, ka, 0, halo=h)), exp) < 1e-14
            o.tm_times_Hopping_Matrix(par, exp, k, 0.9, -0.2)
            assert rel_l2(e.unpack(e.hop(par, sk, U, ka, 1, (0.9, -0.2), halo=h)), exp) < 1e-14
            o.tm_sub_Hopping_Matrix(par, exp, p, k, 1.0, 0.3)
            assert rel_l2(e.unpack(e.hop(par, sk, U, ka, 2, (1.0, 0.3), sp, halo=h)), exp) < 1e-14
            hk = o.spinor(); o.Hopping_Matrix(par, hk, k)
            zp = o.spinor(); o.assign_mul_one_pm_imu(zp, p, +1., o.Vh)  # (1 + i mu g5) p
            got = e.unpack(e.hop(par, sk, U, ka, 3, (1.0, GMU), sp, halo=h))
            assert rel_l2(got, zp - hk) < 1e-14  # MODE 3: the M_full / D_psi row
        # peer mode: the halo the copy CTAs PULL out of the neighbours' fields equals what pack + send/recv deliver
        hu, hd = e.pull_halo(sk, sk)
        assert np.array_equal(hu, dn) and np.array_equal(hd, up)


def test_golden_hopping_through_device_code():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_4x4x4x4.npz"))
    dims = tuple(int(x) for x in gold["dims"])
    e = Emul(*dims)
    U, sk, sp = e.pack_gauge(gold["gauge"]), e.pack(gold["k"]), e.pack(gold["p"])
    for par in (0, 1):
        assert rel_l2(e.unpack(e.hop(par, sk, U, gold["ka"], 0)), gold[f"hop{par}"]) < 1e-14
        assert rel_l2(e.unpack(e.hop(par, sk, U, gold["ka"], 1, (0.9, -0.2))), gold[f"tm_times{par}"]) < 1e-14
        assert rel_l2(e.unpack(e.hop(par, sk, U, gold["ka"], 2, (1.0, 0.3), sp)), gold[f"tm_sub{par}"]) < 1e-14


def test_lexic_permutation_and_elementwise_functors(oracle_lib):
    dims = (4, 6, 4, 8)
    rng = np.random.default_rng(8)
    e, o = Emul(*dims), oracle_lib.Oracle(*dims)
    o.set_params(KAPPA, 0.37)
    lex = random_spinor(rng, o.V)
    ev, od = np.zeros(24 * e.Vh), np.zeros(24 * e.Vh)
    e.E.emul_pack_lexic(ev, od, lex.reshape(-1), *dims)
    E_, O_ = o.spinor(), o.spinor()
    o.convert_lexic_to_eo(E_, O_, lex)
    assert np.array_equal(e.unpack(ev), E_) and np.array_equal(e.unpack(od), O_)
    back = np.zeros(24 * e.V)
    e.E.emul_unpack_lexic(back, ev, od, *dims)
    assert np.array_equal(back.reshape(o.V, 24), lex)
    a, b, c, f = (random_spinor(rng, o.Vh) for _ in range(4))
    out, exp = np.zeros(24 * e.Vh), o.spinor()
    e.E.emul_gamma5(out, e.pack(a), e.Vh); o.gamma5(exp, a, o.Vh); assert np.array_equal(e.unpack(out), exp)
    nrm = 1. / (1. + 0.37 ** 2)
    e.E.emul_diag(out, e.pack(a), nrm, -nrm * 0.37, e.Vh); o.assign_mul_one_pm_imu_inv(exp, a, +1., o.Vh)
    assert rel_l2(e.unpack(out), exp) < 1e-15
    e.E.emul_diag_sub(out, e.pack(a), e.pack(b), 1., -0.37, 1, e.Vh); o.mul_one_pm_imu_sub_mul_gamma5(exp, a, b, -1.)
    assert rel_l2(e.unpack(out), exp) < 1e-15
    e.E.emul_diag_sub(out, e.pack(a), e.pack(b), 1., 0.37, 0, e.Vh); o.mul_one_pm_imu_sub_mul(exp, a, b, +1., o.Vh)
    assert rel_l2(e.unpack(out), exp) < 1e-15
    o1, o2, e1, e2 = np.zeros(24 * e.Vh), np.zeros(24 * e.Vh), o.spinor(), o.spinor()
    e.E.emul_nd_mee_inv(o1, o2, e.pack(a), e.pack(b), 0.139, 0.15, e.Vh); o.M_ee_inv_ndpsi(e1, e2, a, b, 0.139, 0.15)
    assert rel_l2(e.unpack(o1), e1) < 1e-15 and rel_l2(e.unpack(o2), e2) < 1e-15
    e.E.emul_nd_moo_sub_g5(o1, o2, e.pack(a), e.pack(b), e.pack(c), e.pack(f), -0.139, -0.15, e.Vh)
    o.M_oo_sub_g5_ndpsi(e1, e2, a, b, c, f, -0.139, -0.15)
    assert rel_l2(e.unpack(o1), e1) < 1e-15 and rel_l2(e.unpack(o2), e2) < 1e-15


@pytest.mark.parametrize("dims", [(8, 8, 8, 8), (4, 4, 8, 16), (6, 2, 16, 4), (2, 4, 8, 8), (12, 6, 24, 24), (4, 8, 6, 8), (3, 4, 8, 8)])
def test_cta_tile_traversal(dims):
    """The 2 x 2 x 32 CTA tiles of the hopping kernels (tmb_geom.h): a bijection of the sites; every warp keeps 32
    consecutive sites; a CTA holds one 32-run at 2 time-slices x 2 x-planes; with the odd slice shift of the peer mode the
    boundary slices T-1 and 0 share their CTAs (the kernel's `touches` rule finds exactly those)."""
    e = Emul(*dims)
    T, LX, LY, LZ = dims
    P = LY * LZ // 2
    for tshift in (0, (T // 2) | 1):
        site, ts = np.zeros(e.Vh, dtype=np.int32), np.zeros(e.Vh, dtype=np.int32)
        ok = e.E.emul_tile_perm(site, ts, *dims, tshift)
        assert ok == int(P % 32 == 0 and T % 2 == 0 and LX % 2 == 0)
        if not ok:
            continue
        assert e.Vh % 128 == 0
        assert np.array_equal(np.sort(site), np.arange(e.Vh))
        assert np.array_equal(ts, site // (LX * P))
        w = site.reshape(-1, 4, 32)
        assert np.all(np.diff(w, axis=2) == 1) and np.all(w[:, :, 0] % 32 == 0)      # warps: aligned consecutive runs
        t, x, off = w[:, :, 0] // (LX * P), (w[:, :, 0] // P) % LX, w[:, :, 0] % P
        assert np.all(off == off[:, :1])                                              # one run per CTA
        assert np.all(x[:, 1] == x[:, 0] + 1) and np.all(x[:, 2] == x[:, 0]) and np.all(x[:, 3] == x[:, 1]) and np.all(x[:, 0] % 2 == 0)
        assert np.all(t[:, 1] == t[:, 0]) and np.all(t[:, 2] == (t[:, 0] + 1) % T) and np.all(t[:, 3] == t[:, 2])
        tl = ts.reshape(-1, 128)[:, 0]
        touches = (tl == 0) | (tl >= T - 2)                                           # hop_kernel's boundary-CTA rule
        has_boundary = np.any((t == 0) | (t == T - 1), axis=1)
        assert np.array_equal(touches, has_boundary)
        if tshift & 1 and T > 2:
            assert np.all(t[touches][:, 0] == T - 1) and np.all(t[touches][:, 2] == 0)   # one tile layer holds both


def test_host_pipeline_chunk_schedule():
    """the chunk sizes of the pipelined host-pointer Hopping_Matrix (tmb_host_chunk_schedule, tmb_geom.h): they cover the
    slices exactly, respect the minimum size except for the remainder, stay few, start with a small chunk and end with the
    smallest; the measured schedule of 24^3x48 is pinned"""
    e = Emul(4, 4, 4, 4)
    buf = np.zeros(64, dtype=np.int32)
    assert [int(x) for x in buf[:e.E.emul_host_chunk_schedule(48, 1, buf)]] == [2, 6, 8, 8, 8, 8, 4, 2, 1, 1]
    for nt in list(range(1, 130)) + [256, 510, 1024]:
        for small in (1, 2, 3, 5, 16):
            n = e.E.emul_host_chunk_schedule(nt, small, buf)
            sz = [int(x) for x in buf[:n]]
            assert 1 <= n <= 8 + int(np.log2(nt)) + 1 and sum(sz) == nt and min(sz) >= 1
            assert all(x >= min(small, nt) for x in sz[:-1])     # only the last chunk may be a remainder below `small`
            if nt >= 24 * small:
                assert sz[0] <= sz[1] <= sz[2] and sz[-1] <= sz[-2] <= sz[-3] and max(sz) == (nt + 5) // 6


@pytest.mark.parametrize("dims,xb", [((4, 8, 4, 6), 2), ((4, 8, 4, 6), 4), ((6, 6, 2, 4), 3)])
def test_xblock_traversal_is_a_permutation(dims, xb):
    e = Emul(*dims)
    perm = np.zeros(e.Vh, dtype=np.int32)
    e.E.emul_xblock_perm(perm, *dims, xb)
    assert np.array_equal(np.sort(perm), np.arange(e.Vh))


def test_compressed_links_and_single_precision_instantiations(oracle_lib):
    """12-real link reconstruction (CFG bit 1) and the float2 instantiation of the same site code"""
    dims, theta = (4, 6, 4, 8), (1., 0.3, 0., 0.7)
    rng = np.random.default_rng(12)
    e, o = Emul(*dims), oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g); o.set_params(KAPPA, GMU, theta)
    ka = ka_of(KAPPA, theta, dims)
    U = e.pack_gauge(g)
    U12 = np.zeros(8 * 6 * 2 * e.Vh); e.E.emul_compress12(U12, U, e.Vh, 8)
    k = random_spinor(rng, o.Vh); sk = e.pack(k)
    up, dn = e.pack_halo(sk)
    Uh = e.pack_gauge_halo(U); Uh12 = np.zeros(2 * 6 * 2 * e.S); e.E.emul_compress12(Uh12, Uh, e.S, 2)
    z = np.zeros(2)
    for par in (0, 1):
        exp = o.spinor(); o.Hopping_Matrix(par, exp, k)
        out = np.zeros(24 * e.Vh)
        assert e.E.emul_hop12(par, out, sk, U12, z, z, z, *dims, ka, 0) == 0
        assert rel_l2(e.unpack(out), exp) < 1e-14
        assert e.E.emul_hop12(par, out, sk, U12, dn, up, Uh12, *dims, ka, 1) == 0
        assert rel_l2(e.unpack(out), exp) < 1e-14
        outf = np.zeros(24 * e.Vh, dtype=np.float32)
        assert e.E.emul_hop_f(par, outf, sk.astype(np.float32), U.astype(np.float32), *dims, ka) == 0
        assert rel_l2(e.unpack(outf.astype(np.float64)), exp) < 1e-6


@pytest.mark.parametrize("dims,theta", [((4, 4, 4, 4), (0., 0., 0., 0.)), ((4, 6, 4, 8), (1., 0.3, 0., 0.7)), ((2, 4, 2, 6), (1., 0., 0., 0.))])
def test_fermion_force_gather_and_halo_loopback(oracle_lib, dims, theta):
    """deriv_Sb as the device computes it (gather over link owners, tmb_deriv_site) vs the oracle's scatter
    restatement of deriv_Sb.c:402-649, with and without the T-halo path (loopback)"""
    rng = np.random.default_rng(21)
    e, o = Emul(*dims), oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g); o.set_params(KAPPA, GMU, theta)
    ka = ka_of(KAPPA, theta, dims)
    U = e.pack_gauge(g)
    l, k = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
    sl, sk = e.pack(l), e.pack(k)
    df0 = rng.normal(size=(o.V, 4, 8))
    for ieo in (0, 1):
        exp = df0.copy(); o.deriv_Sb(ieo, l, k, exp, 0.7)
        assert rel_l2(e.deriv(ieo, sl, sk, U, ka, df0, 0.7) - df0, exp - df0) < 1e-14
        halo = e.pack_deriv_halo(sk, sl)  # loopback: this rank is its own upper neighbour
        assert rel_l2(e.deriv(ieo, sl, sk, U, ka, df0, 0.7, halo=halo) - df0, exp - df0) < 1e-14


def test_golden_fermion_force_through_device_code():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_hmc_4x4x4x4.npz"))
    dims = tuple(int(x) for x in gold["dims"])
    e = Emul(*dims)
    ka = ka_of(float(gold["kappa"]), gold["theta"], dims)
    U, sl, sk = e.pack_gauge(gold["gauge"]), e.pack(gold["l"]), e.pack(gold["k"])
    for ieo in (0, 1):
        got = e.deriv(ieo, sl, sk, U, ka, np.zeros((e.V, 4, 8)), 0.7)
        assert rel_l2(got, gold[f"deriv_Sb{ieo}"]) < 1e-14


@pytest.mark.parametrize("dims,theta", [((4, 4, 4, 4), (0., 0., 0., 0.)), ((4, 6, 2, 8), (1., 0.3, 0., 0.7))])
def test_two_flavour_hop_and_fused_flavour_mixing(oracle_lib, dims, theta):
    """tmb_hop_site2 + the epilogues of hop2_kernel (host emulation), composed exactly as qtm_pm_nd() in tmb_capi.cu
    composes its four launches, against the oracle's Qtm_pm_ndpsi (tm_operators_nd.c:195-239) and its pieces"""
    rng = np.random.default_rng(9)
    e, o = Emul(*dims), oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    mubar, epsbar, invmaxev = 0.139, 0.15, 0.9
    o.set_gauge(g); o.set_params(KAPPA, GMU, theta); o.set_nd_params(mubar, epsbar, invmaxev)
    ka = ka_of(KAPPA, theta, dims)
    U = e.pack_gauge(g)
    ks, kc = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
    sks, skc = e.pack(ks), e.pack(kc)
    # mode 0: two plain hops with one gauge stream
    h0, h1 = e.hop2(1, sks, skc, U, ka)
    exp = o.spinor()
    o.Hopping_Matrix(1, exp, ks); assert rel_l2(e.unpack(h0), exp) < 1e-14
    o.Hopping_Matrix(1, exp, kc); assert rel_l2(e.unpack(h1), exp) < 1e-14
    # mode 1 against H_eo + M_ee_inv_ndpsi
    a0, a1 = e.hop2(0, skc, sks, U, ka, mode=1, mu=mubar, eps=epsbar)
    hc, hs = o.spinor(), o.spinor()
    o.Hopping_Matrix(0, hc, kc); o.Hopping_Matrix(0, hs, ks)
    m0, m1 = o.spinor(), o.spinor()
    o.M_ee_inv_ndpsi(m0, m1, hc, hs, mubar, epsbar)
    assert rel_l2(e.unpack(a0), m0) < 1e-14 and rel_l2(e.unpack(a1), m1) < 1e-14
    # the four-launch composition of Qtm_pm_ndpsi
    b0, b1 = e.hop2(1, a0, a1, U, ka, mode=2, p0=skc, p1=sks, mu=-mubar, eps=-epsbar, scale=1.)
    a0, a1 = e.hop2(0, b0, b1, U, ka, mode=1, mu=-mubar, eps=epsbar)
    ls, lc = e.hop2(1, a1, a0, U, ka, mode=2, p0=b1, p1=b0, mu=-mubar, eps=-epsbar, scale=invmaxev * invmaxev)
    es, ec = o.spinor(), o.spinor()
    o.Qtm_pm_ndpsi(es, ec, ks, kc)
    assert rel_l2(e.unpack(ls), es) < 1e-13 and rel_l2(e.unpack(lc), ec) < 1e-13


@pytest.mark.parametrize("gdims,nz", [((4, 4, 4, 8), 2), ((2, 4, 6, 12), 3), ((4, 2, 4, 4), 2)])
def test_z_split_face_exchange_and_fixup(oracle_lib, gdims, nz):
    """Second split direction (Z): every z slab runs the UNCHANGED hopping code as if it were periodic in z, then the fix-up
    replaces the wrapped term of its face sites by the term built from the neighbour slab's packed faces and z-links
    (tmb_site.cuh).  All epilogues, against the oracle on the global lattice."""
    T, LX, LY, LZ = gdims
    LZl = LZ // nz
    rng = np.random.default_rng(41)
    theta = (1., 0., 0.3, 0.7)
    o = oracle_lib.Oracle(*gdims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g); o.set_params(KAPPA, GMU, theta)
    ka = ka_of(KAPPA, theta, gdims)  # global extents in the phases
    k, p = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
    e = Emul(T, LX, LY, LZl)
    rows, Sz = T * LX * LY, T * LX * LY // 2
    gs = g.reshape(T, LX, LY, LZ, 4, 18)
    slab_f = lambda f, s: np.ascontiguousarray(f.reshape(rows, LZ // 2, 24)[:, s * (LZl // 2):(s + 1) * (LZl // 2), :]).reshape(-1, 24)
    U = [e.pack_gauge(np.ascontiguousarray(gs[:, :, :, s * LZl:(s + 1) * LZl]).reshape(-1, 4, 18)) for s in range(nz)]
    Uzh = []
    for s in range(nz):
        out = np.zeros(36 * Sz); e.E.emul_pack_gauge_zhalo(out, U[s], T, LX, LY, LZl); Uzh.append(out)
    for par in (0, 1):
        sk = [e.pack(slab_f(k, s)) for s in range(nz)]
        sp = [e.pack(slab_f(p, s)) for s in range(nz)]
        faces = []
        for s in range(nz):
            up, dn = np.zeros(12 * Sz), np.zeros(12 * Sz)
            e.E.emul_pack_zfaces(up, dn, sk[s], T, LX, LY, LZl, 1 - par)
            faces.append((up, dn))
        for mode, cf, ref in ((0, (1., 0.), lambda x: o.Hopping_Matrix(par, x, k)),
                              (1, (0.9, -0.2), lambda x: o.tm_times_Hopping_Matrix(par, x, k, 0.9, -0.2)),
                              (2, (1.0, 0.3), lambda x: o.tm_sub_Hopping_Matrix(par, x, p, k, 1.0, 0.3)),
                              (3, (1.0, GMU), None)):
            exp = o.spinor()
            if ref is not None:
                ref(exp)
            else:
                hk = o.spinor(); o.Hopping_Matrix(par, hk, k)
                zp = o.spinor(); o.assign_mul_one_pm_imu(zp, p, +1., o.Vh)
                exp = zp - hk
            for s in range(nz):
                out = e.hop(par, sk[s], U[s], ka, mode, cf, sp[s] if mode >= 2 else None)
                wrong = rel_l2(e.unpack(out), slab_f(exp, s))
                e.E.emul_zfix(mode, out, sk[s], U[s], faces[(s + 1) % nz][1], faces[(s - 1) % nz][0], Uzh[(s - 1) % nz],
                              T, LX, LY, LZl, par, np.asarray(ka, dtype=np.float64), cf[0], cf[1])
                assert rel_l2(e.unpack(out), slab_f(exp, s)) < 1e-14, (par, mode, s)
                assert wrong > 1e-3  # the un-fixed slab result really differs


@pytest.mark.parametrize("gdims,nz", [((4, 4, 4, 8), 2), ((2, 4, 6, 12), 3), ((4, 2, 4, 4), 2)])
def test_z_split_fermion_force_fixup(oracle_lib, gdims, nz):
    """deriv_Sb on z slabs: every slab runs the unchanged force code as if it were periodic in z, then the fix-up of the z links
    owned by its last-z sites takes the half-spinors of the slab above (the `dn` faces of the z-face pack of k and of l).
    Against the oracle's deriv_Sb on the global lattice, both parities."""
    T, LX, LY, LZ = gdims
    LZl = LZ // nz
    rng = np.random.default_rng(43)
    theta = (1., 0., 0.3, 0.7)
    o = oracle_lib.Oracle(*gdims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g); o.set_params(KAPPA, GMU, theta)
    ka = ka_of(KAPPA, theta, gdims)
    l, k = random_spinor(rng, o.Vh), random_spinor(rng, o.Vh)
    df0 = rng.normal(size=(o.V, 4, 8))
    e = Emul(T, LX, LY, LZl)
    rows, Sz = T * LX * LY, T * LX * LY // 2
    gs = g.reshape(T, LX, LY, LZ, 4, 18)
    slab_f = lambda f, s: np.ascontiguousarray(f.reshape(rows, LZ // 2, 24)[:, s * (LZl // 2):(s + 1) * (LZl // 2), :]).reshape(-1, 24)
    slab_d = lambda d, s: np.ascontiguousarray(d.reshape(rows, LZ, 4, 8)[:, s * LZl:(s + 1) * LZl]).reshape(-1, 4, 8)
    U = [e.pack_gauge(np.ascontiguousarray(gs[:, :, :, s * LZl:(s + 1) * LZl]).reshape(-1, 4, 18)) for s in range(nz)]
    for ieo in (0, 1):
        exp = df0.copy(); o.deriv_Sb(ieo, l, k, exp, 0.7)
        sl = [e.pack(slab_f(l, s)) for s in range(nz)]
        sk = [e.pack(slab_f(k, s)) for s in range(nz)]
        faces = []
        for s in range(nz):  # what slab s sends DOWN: the D = 6 projection of its first-z sites, of k (parity 1 - ieo) and of l (parity ieo)
            junk, fk, fl = np.zeros(12 * Sz), np.zeros(12 * Sz), np.zeros(12 * Sz)
            e.E.emul_pack_zfaces(junk, fk, sk[s], T, LX, LY, LZl, 1 - ieo)
            e.E.emul_pack_zfaces(junk, fl, sl[s], T, LX, LY, LZl, ieo)
            faces.append((fk, fl))
        for s in range(nz):
            d0 = slab_d(df0, s)
            dev = np.zeros(64 * e.Vh); e.E.emul_pack_deriv(dev, np.ascontiguousarray(d0).reshape(-1), T, LX, LY, LZl)
            e.E.emul_deriv(ieo, sl[s], sk[s], U[s], dev, np.zeros(2), T, LX, LY, LZl, np.asarray(ka, dtype=np.float64), 0.7, 0)
            out = np.zeros(32 * e.V); e.E.emul_unpack_deriv(out, dev, T, LX, LY, LZl)
            wrong = rel_l2(out.reshape(-1, 4, 8) - d0, slab_d(exp, s) - d0)
            fk, fl = faces[(s + 1) % nz]
            e.E.emul_deriv_zfix(ieo, sl[s], sk[s], U[s], dev, fk, fl, T, LX, LY, LZl, np.asarray(ka, dtype=np.float64), 0.7)
            e.E.emul_unpack_deriv(out, dev, T, LX, LY, LZl)
            assert wrong > 1e-3  # the slab alone is NOT right: the fix-up is what makes it so
            assert rel_l2(out.reshape(-1, 4, 8) - d0, slab_d(exp, s) - d0) < 1e-14, (ieo, s)


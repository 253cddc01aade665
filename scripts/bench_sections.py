#!/usr/bin/env python
"""Extra measurement sections of bench.py, each on its own lattice in its own process (the reference keeps its
lattice in C globals: one lattice per process), printing ONE JSON object on stdout.

    python scripts/bench_sections.py nd  [TxLXxLYxLZ]   BASELINE configs[3]: invert_doublet_eo (Qtm_pm_ndpsi CG), 32^3x64
    python scripts/bench_sections.py small [TxLXxLYxLZ] BASELINE configs[0]: benchmark (Hopping_Matrix + D_psi), 8^4
    python scripts/bench_sections.py hmc [TxLXxLYxLZ]   BASELINE configs[4]: det + detratio monomials, 16^3x32
                                                        (heatbath, derivative = inversion + deriv_Sb force, acc)

GPU numbers: device-resident through the C ABI, wall clock around the call (the calls synchronise).  CPU numbers:
the unmodified reference (oracle/_ref, half-spinor OpenMP build where the operator allows it) on all host cores,
bounded samples.  Parameters: kappa = 0.16, mu = 0.01 (g_mu = 2 kappa mu), random SU(3) hot start; ND: 2KappaMubar =
0.139, 2KappaEpsbar = 0.15 (sample-input/sample-cg.input:32-33); HMC: det with 2KappaMu = 0.032 (heavier), detratio
(0.0032 / 0.032), CG, forceprec 1e-14, accprec 1e-18, csg history 2.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
KAPPA, MU = 0.16, 0.01
GMU = 2 * KAPPA * MU
CG = 1


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def gauge_and_ref(dims, nthreads, halfspinor):
    """reference ranlux hot start when oracle/_ref is there (also the CPU baseline object), else numpy QR"""
    from oracle import refclient
    if refclient.available(halfspinor=halfspinor):
        ref = refclient.Reference(*dims, nthreads=nthreads, halfspinor=halfspinor)
        return ref.random_gauge(123456), ref, "reference ranlux random_gauge_field(seed 123456)"
    from conftest import random_gauge
    V = int(np.prod(dims))
    return random_gauge(np.random.default_rng(123456), V), None, "numpy QR random SU(3) (seed 123456)"


def section_nd(dims):
    import tmlqcd_b200 as tm
    from conftest import random_spinor
    ncores = os.cpu_count() or 1
    g, ref, how = gauge_and_ref(dims, ncores, True)
    mubar, epsbar = 0.139, 0.15
    d = tm.Device(*dims)
    d.set_params(KAPPA, GMU)
    d.ck(d.lib.tmb_set_nd(mubar, epsbar, 1.0))
    d.gauge_upload(g)
    rng = np.random.default_rng(5)
    src = [random_spinor(rng, d.Vh) for _ in range(4)]
    f = [d.field(s) for s in src] + [d.field() for _ in range(4)]
    eps_sq, maxit = 1e-14, 5000
    # operator alone: Qtm_pm_ndpsi = 8 hops + flavour mixing; variant 2 = the default two-flavour kernel (two flavour
    # groups of warps per CTA), variant 0 = round 1's kernel (both flavours in one thread)
    n = 20
    t_var = {}
    for variant in (0, 2):
        d.ck(d.lib.tmb_set_hop2_variant(variant))
        d.call("Qtm_pm_ndpsi", f[4], f[5], f[0], f[1]); d.ck(d.lib.tmb_sync())
        d.timer_start()
        for _ in range(n):
            d.lib.tmb_Qtm_pm_ndpsi(f[4], f[5], f[0], f[1])
        t_var[variant] = d.timer_stop() * 1e-3 / n
    d.ck(d.lib.tmb_set_hop2_variant(-1))  # automatic: round 1's kernel here (one rank, 18-real links, double)
    t_op = min(t_var.values())
    # the same in single precision (Qtm_pm_ndpsi_32, the operator of rg_mixed_cg_her_nd's inner loops)
    f32 = [d.field32(src[0].astype(np.float32)), d.field32(src[1].astype(np.float32)), d.field32(), d.field32()]
    t_var32 = {}
    for variant in (0, 2):
        d.ck(d.lib.tmb_set_hop2_variant(variant))
        d.call("Qtm_pm_ndpsi_32", f32[2], f32[3], f32[0], f32[1]); d.ck(d.lib.tmb_sync())
        d.timer_start()
        for _ in range(n):
            d.lib.tmb_Qtm_pm_ndpsi_32(f32[2], f32[3], f32[0], f32[1])
        t_var32[variant] = d.timer_stop() * 1e-3 / n
    d.ck(d.lib.tmb_set_hop2_variant(-1))
    t_op32 = min(t_var32.values())
    it = d.call("invert_doublet_eo", f[4], f[5], f[6], f[7], f[0], f[1], f[2], f[3], eps_sq, maxit, 1)  # warm-up
    d.call("field_zero", f[5]); d.call("field_zero", f[7])
    d.ck(d.lib.tmb_sync())
    t0 = time.perf_counter()
    it = d.call("invert_doublet_eo", f[4], f[5], f[6], f[7], f[0], f[1], f[2], f[3], eps_sq, maxit, 1)
    d.ck(d.lib.tmb_sync())
    t_solve = time.perf_counter() - t0
    its, err, t_cg = d.solver_stats()
    # RGMIXEDCG (invert_doublet_eo.c:145-149): rg_mixed_cg_her_nd, float inner loops, default mcg_delta (operator.c:125)
    x_cg = d.download(f[5])
    d.ck(d.lib.tmb_set_mcg_delta(5.0e-5))
    d.call("invert_doublet_eo_solver", f[4], f[5], f[6], f[7], f[0], f[1], f[2], f[3], eps_sq, maxit, 1, 14)  # warm-up
    d.call("field_zero", f[5]); d.call("field_zero", f[7])
    d.ck(d.lib.tmb_sync())
    t0 = time.perf_counter()
    it_rg = d.call("invert_doublet_eo_solver", f[4], f[5], f[6], f[7], f[0], f[1], f[2], f[3], eps_sq, maxit, 1, 14)
    d.ck(d.lib.tmb_sync())
    t_rg = time.perf_counter() - t0
    _, err_rg, _ = d.solver_stats()
    isp, idp, iou = C.c_int(), C.c_int(), C.c_int()
    d.lib.tmb_solver_stats_rg(C.byref(isp), C.byref(idp), C.byref(iou))
    x_rg = d.download(f[5])
    out = {"workload": "BASELINE configs[3]: invert_doublet_eo, non-degenerate doublet CG on Qtm_pm_ndpsi, %dx%dx%dx%d (TxLXxLYxLZ)" % dims,
           "gauge": how, "2KappaMubar": mubar, "2KappaEpsbar": epsbar, "eps_sq": eps_sq, "rel_prec": 1,
           "iterations": it, "time_to_solution_s": t_solve, "cg_loop_s": t_cg, "final_rr": err,
           "Qtm_pm_ndpsi_us": 1e6 * t_op, "Qtm_pm_ndpsi_us_one_thread_two_flavours": 1e6 * t_var[0],
           "Qtm_pm_ndpsi_us_two_flavour_warp_groups": 1e6 * t_var[2], "Qtm_pm_ndpsi_32_us": 1e6 * t_op32,
           "Qtm_pm_ndpsi_32_us_one_thread_two_flavours": 1e6 * t_var32[0], "Qtm_pm_ndpsi_32_us_two_flavour_warp_groups": 1e6 * t_var32[2],
           "Qtm_pm_ndpsi_32_hbm_gbs_effective": 4224.0 * d.Vh / t_op32 / 1e9,
           "rgmixed": {"count": it_rg, "time_to_solution_s": t_rg, "true_rr": err_rg, "inner_sp": isp.value, "inner_dp": idp.value,
                       "outer": iou.value, "mcg_delta": 5.0e-5, "speedup_vs_cg": t_solve / t_rg,
                       "odd_solution_rel_l2_vs_cg": float(np.linalg.norm(x_rg - x_cg) / np.linalg.norm(x_cg))},
           "Qtm_pm_ndpsi_algorithmic_bytes_per_site": 8448,
           "Qtm_pm_ndpsi_bytes_how": "4 two-flavour launches: 2 x (1152 links + 4 x 192 spinors) + 2 x (1152 + 6 x 192); "
                                     "the reference's call sequence (8 hops + 5 sweeps) moves 17664",
           "Qtm_pm_ndpsi_hbm_gbs_effective": 8448.0 * d.Vh / t_op / 1e9,
           "Qtm_pm_ndpsi_hbm_gbs_at_8x1536_B_per_site": 8 * 1536.0 * d.Vh / t_op / 1e9,
           "gflops_1320_per_hop": 8 * 1320.0 * d.Vh / t_op / 1e9}
    d.close()
    fix = os.path.join(ROOT, "tests", "golden", "ref_nd_count_%dx%dx%dx%d.json" % dims)
    if os.path.exists(fix):  # counts of the unmodified reference's invert_doublet_eo on these very inputs (tests/golden/make_golden_nd_count.py)
        j = json.load(open(fix))
        out["cpu_reference_counts"] = {"CG": j["CG"]["iterations"], "RGMIXEDCG": j["RGMIXEDCG"]["iterations"],
                                       "CG_seconds": j["CG"]["seconds"], "RGMIXEDCG_seconds": j["RGMIXEDCG"]["seconds"], "threads": j["threads"],
                                       "how": "unmodified invert_doublet_eo.c (oracle/_ref) on the same gauge field and sources, committed fixture"}
        out["iterations_match_reference"] = bool(abs(it - j["CG"]["iterations"]) <= 1)
    if ref is not None:
        ref.set_params(KAPPA, GMU); ref.set_nd_params(mubar, epsbar, 1.0)
        a, b = ref.spinor(), ref.spinor()
        ref.Qtm_pm_ndpsi(a, b, src[0], src[1])
        nrep = 2
        t0 = time.perf_counter()
        for _ in range(nrep):
            ref.Qtm_pm_ndpsi(a, b, src[0], src[1])
        t_ref = (time.perf_counter() - t0) / nrep
        out["cpu_reference"] = {"Qtm_pm_ndpsi_s": t_ref, "cores": ref.nthreads, "kind": "reference",
                                "sample": f"{nrep} applications of Qtm_pm_ndpsi, half-spinor OpenMP build of the unmodified reference",
                                "time_to_solution_s_est": t_ref * (max(it, 0) + 1),
                                "est_how": "Qtm_pm_ndpsi time x (iterations + 1); lower bound, BLAS-1 of cg_her_nd not included"}
    return out


def section_hmc(dims):
    import tmlqcd_b200 as tm
    from conftest import random_spinor
    ncores = os.cpu_count() or 1
    # half-spinor OpenMP build: the fastest generic-C variant of the CG's Hopping_Matrix; its deriv_Sb is the same code
    g, ref, how = gauge_and_ref(dims, ncores, True)
    forceprec, accprec, csgN, maxit = 1e-14, 1e-18, 2, 5000
    mons = [(0, KAPPA, 10 * GMU, KAPPA, 10 * GMU), (1, KAPPA, GMU, KAPPA, 10 * GMU)]  # det(heavy), detratio(light/heavy)
    d = tm.Device(*dims)
    d.set_params(KAPPA, GMU)
    d.gauge_upload(g)
    d.ck(d.lib.tmb_set_relative_precision_flag(0))
    rng = np.random.default_rng(11)
    if ref is not None:
        # both arms see the SAME pseudo-fermion noise: what the reference's heatbath draws after start_ranlux(1, 1000 + id)
        # (random_spinor_field_eo(w_fields[0], repro, RN_GAUSS), det_monomial.c:177, detratio_monomial.c:224)
        etas = []
        for id in range(len(mons)):
            ref.start_ranlux(1, 1000 + id)
            etas.append(ref.random_spinor_eo())
    else:
        etas = [random_spinor(rng, d.Vh) for _ in mons]
    out = {"workload": "BASELINE configs[4]: det + detratio monomials (heatbath, derivative = CG inversion + deriv_Sb force, acc), "
                       "%dx%dx%dx%d (TxLXxLYxLZ), fields resident in HBM" % dims,
           "gauge": how, "forceprec": forceprec, "accprec": accprec, "csg_N": csgN, "solver": "CG", "monomials": []}
    V = d.V
    # force kernel alone
    l, k = d.field(etas[0]), d.field(etas[1])
    d.call("derivative_zero"); d.call("deriv_Sb", 0, l, k, 1.0); d.ck(d.lib.tmb_sync())
    n = 200
    d.timer_start()
    for _ in range(n):
        d.lib.tmb_deriv_Sb(0, l, k, 1.0); d.lib.tmb_deriv_Sb(1, k, l, 1.0)
    ms = d.timer_stop()
    t_force = ms * 1e-3 / (2 * n)
    out["deriv_Sb"] = {"us_per_call": 1e6 * t_force, "algorithmic_bytes_per_link_owner_site": 1280,
                       "hbm_gbs_effective": 1280.0 * V / t_force / 1e9, "sites_per_call": V}
    tot_gpu = 0.
    for id, (typ, k1, m1, k2, m2) in enumerate(mons):
        assert d.lib.tmb_monomial_add(typ, k1, m1, k2, m2, CG, maxit, forceprec, accprec, csgN) == id
        eta = d.field(etas[id])
        e0, dH = C.c_double(), C.c_double()
        rec = {"type": "DET" if typ == 0 else "DETRATIO"}
        d.ck(d.lib.tmb_sync()); t0 = time.perf_counter()
        d.ck(d.lib.tmb_monomial_heatbath(id, eta, C.byref(e0)))
        rec["heatbath_s"] = time.perf_counter() - t0
        d.call("derivative_zero")
        ts = []
        for call in range(3):
            d.ck(d.lib.tmb_sync()); t0 = time.perf_counter()
            d.ck(d.lib.tmb_monomial_derivative(id))
            ts.append(time.perf_counter() - t0)
        rec["derivative_s"] = ts
        d.ck(d.lib.tmb_sync()); t0 = time.perf_counter()
        d.ck(d.lib.tmb_monomial_acc(id, C.byref(dH)))
        rec["acc_s"] = time.perf_counter() - t0
        rec.update(d.monomial_info(id)); rec["dH"] = dH.value
        tot_gpu += rec["heatbath_s"] + sum(ts) + rec["acc_s"]
        out["monomials"].append(rec)
    out["total_s"] = tot_gpu
    df_gpu = d.derivative_download()
    d.close()
    if ref is not None:
        assert ref.hmc_init() == 0
        ref.set_params(KAPPA, GMU)
        ids = [ref.mnl_add(typ, k1, m1, k2, m2, CG, maxit, forceprec, accprec, csgN) for typ, k1, m1, k2, m2 in mons]
        assert ref.mnl_init() == 0
        tot = 0.
        recs = []
        for id in ids:
            ref.start_ranlux(1, 1000 + id)  # the heatbath draws the field the device arm was given
            t0 = time.perf_counter(); ref.mnl_heatbath(id); th = time.perf_counter() - t0
            df = ref.derivative()
            ts = []
            for call in range(3):
                t0 = time.perf_counter(); ref.mnl_derivative(id, df); ts.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); ref.mnl_acc(id); ta = time.perf_counter() - t0
            recs.append({"heatbath_s": th, "derivative_s": ts, "acc_s": ta, **ref.mnl_info(id)})
            tot += th + sum(ts) + ta
        out["cpu_reference"] = {"total_s": tot, "monomials": recs, "cores": ref.nthreads, "kind": "reference",
                                "sample": "the same sequence (heatbath, 3 derivatives, acc per monomial) by the unmodified reference "
                                          "(half-spinor OpenMP build: det_monomial.c, detratio_monomial.c, deriv_Sb.c, chrono_guess.c, cg_her.c)"}
        out["speedup_vs_cpu_reference"] = tot / tot_gpu
        cmp = []
        for a, b in zip(out["monomials"], recs):
            cmp.append({"energy0_rel_diff": abs(a["energy0"] - b["energy0"]) / abs(b["energy0"]),
                        "energy1_rel_diff": abs(a["energy1"] - b["energy1"]) / abs(b["energy1"]),
                        "iter0": [a["iter0"], b["iter0"]], "iter1": [a["iter1"], b["iter1"]]})
        out["same_noise_comparison"] = cmp
        # the force itself: three accumulated derivative calls of the last monomial, device against reference
        out["derivative_rel_l2_vs_cpu_reference"] = float(np.linalg.norm(df_gpu - df) / np.linalg.norm(df))
    out["derivative_norm"] = float(np.linalg.norm(df_gpu))
    return out


def section_small(dims):
    """BASELINE configs[0]: the reference's `benchmark` program (benchmark.c:262-327: Hopping_Matrix EO+OE pairs with
    even_odd_flag, D_psi without) on 8^4.  At 2048 sites per parity the GPU is launch-latency bound, not HBM bound."""
    import tmlqcd_b200 as tm
    from conftest import random_spinor
    g, ref, how = gauge_and_ref(dims, 1, False)  # scalar full-spinor build, one thread: "scalar --disable-mpi CPU build"
    d = tm.Device(*dims)
    d.set_params(KAPPA, GMU)
    d.gauge_upload(g)
    rng = np.random.default_rng(3)
    V, Vh = d.V, d.Vh
    f = [d.field(random_spinor(rng, Vh)) for _ in range(2)] + [d.field() for _ in range(2)]
    n = 2000
    for _ in range(50):
        d.lib.tmb_Hopping_Matrix(0, f[2], f[0]); d.lib.tmb_Hopping_Matrix(1, f[3], f[2])
    d.timer_start()
    for _ in range(n):
        d.lib.tmb_Hopping_Matrix(0, f[2], f[0]); d.lib.tmb_Hopping_Matrix(1, f[3], f[2])
    t_pair = d.timer_stop() * 1e-3 / n
    for _ in range(50):
        d.lib.tmb_D_psi_eo(f[2], f[3], f[0], f[1])
    d.timer_start()
    for _ in range(n):
        d.lib.tmb_D_psi_eo(f[2], f[3], f[0], f[1])
    t_dpsi = d.timer_stop() * 1e-3 / n
    out = {"workload": "BASELINE configs[0]: benchmark (Hopping_Matrix + D_psi), %dx%dx%dx%d (TxLXxLYxLZ), device-resident fields" % dims,
           "gauge": how, "Hopping_Matrix_pair_us": 1e6 * t_pair, "Hopping_Matrix_gflops_1320": V * 1320.0 / t_pair / 1e9,
           "Hopping_Matrix_gflops_1608": V * 1608.0 / t_pair / 1e9,
           "D_psi_us": 1e6 * t_dpsi, "D_psi_gflops_1680": V * 1680.0 / t_dpsi / 1e9,
           "note": "2048 sites per parity = 16 CTAs: launch-latency bound (2 launches per pair and per D_psi), not an HBM measurement"}
    # the same through the reference-named symbols with host buffers (PCIe copies inside the timing)
    D = tm.DropIn(*dims)
    D.set_params(KAPPA, GMU); D.set_gauge(g)
    hk, h1, h2 = random_spinor(rng, Vh), np.zeros((Vh, 24)), np.zeros((Vh, 24))
    P, Q = np.zeros((V, 24)), random_spinor(rng, V)
    D.Hopping_Matrix(0, h1, hk); D.D_psi(P, Q)
    m = 200
    t0 = time.perf_counter()
    for _ in range(m):
        D.Hopping_Matrix(0, h1, hk); D.Hopping_Matrix(1, h2, h1)
    out["dropin_Hopping_Matrix_pair_us"] = 1e6 * (time.perf_counter() - t0) / m
    t0 = time.perf_counter()
    for _ in range(m):
        D.D_psi(P, Q)
    out["dropin_D_psi_us"] = 1e6 * (time.perf_counter() - t0) / m
    d.close()
    if ref is not None:
        ref.set_params(KAPPA, GMU)
        nrep = 2000
        th = ref.lib.ref_bench_hopping(nrep) / nrep
        td = ref.lib.ref_bench_D_psi(nrep) / (2 * nrep)
        out["cpu_reference"] = {"kind": "reference", "cores": 1,
                                "sample": f"{nrep} EO+OE pairs and {2 * nrep} D_psi applications, scalar full-spinor build of the unmodified reference, 1 thread (benchmark.c recipe)",
                                "Hopping_Matrix_pair_us": 1e6 * th, "Hopping_Matrix_gflops_1608": V * 1608.0 / th / 1e9,
                                "D_psi_us": 1e6 * td, "D_psi_gflops_1680": V * 1680.0 / td / 1e9}
    return out


if __name__ == "__main__":
    which = sys.argv[1]
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if which == "nd":
        dims = tuple(int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "64x32x32x32").split("x"))
        res = section_nd(dims)
    elif which == "hmc":
        dims = tuple(int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "32x16x16x16").split("x"))
        res = section_hmc(dims)
    elif which == "small":
        dims = tuple(int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "8x8x8x8").split("x"))
        res = section_small(dims)
    else:
        raise SystemExit("usage: bench_sections.py nd|hmc|small [TxLXxLYxLZ]")
    print(json.dumps(res), file=real_stdout, flush=True)

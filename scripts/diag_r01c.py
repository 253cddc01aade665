#!/usr/bin/env python
"""Round 1, third session diagnostics (one GPU):
  nd    two-flavour hop kernel variants (0 both flavours per thread, 1 lane-paired): Qtm_pm_ndpsi time at 32^3x64
  e2e   the host-pointer invert_eo of bench.py, phase by phase (g_debug_level = 2 prints upload / solve / download)
  chunk host-pointer Hopping_Matrix against the number of pipeline chunks
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tmlqcd_b200 as tm
from conftest import random_gauge, random_spinor

what = sys.argv[1]
KAPPA, GMU = 0.16, 0.0032


def timeit(d, fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    d.ck(d.lib.tmb_sync())
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    d.ck(d.lib.tmb_sync())
    return (time.perf_counter() - t0) / n


if what == "nd":
    dims = (64, 32, 32, 32)
    rng = np.random.default_rng(1)
    d = tm.Device(*dims)
    d.set_params(KAPPA, GMU); d.ck(d.lib.tmb_set_nd(0.139, 0.15, 1.0))
    d.gauge_upload(random_gauge(rng, d.V))
    f = [d.field(random_spinor(rng, d.Vh)) for _ in range(2)] + [d.field() for _ in range(4)]
    res = {}
    for v in (0, 1, 0, 1):
        d.ck(d.lib.tmb_set_hop2_variant(v))
        t = timeit(d, lambda: d.call("Qtm_pm_ndpsi", f[2], f[3], f[0], f[1]), n=50)
        print(f"hop2 variant {v}: Qtm_pm_ndpsi {1e6 * t:8.1f} us  -> {8448.0 * d.Vh / t / 1e9:7.1f} GB/s effective (8448 B/site)", flush=True)
        res[v] = d.download(f[2])
    print("variants agree: rel", np.linalg.norm(res[0] - res[1]) / np.linalg.norm(res[1]))
    d.close()
elif what == "quant":
    # wave quantisation: 384 (variant 0) vs 448 (variant 10) resident threads per SM, hop pair and CG iteration
    for dims in ((32, 16, 16, 16), (16, 16, 16, 16), (48, 24, 24, 24), (8, 8, 8, 8)):
        rng = np.random.default_rng(1)
        d = tm.Device(*dims)
        d.set_params(KAPPA, GMU)
        d.gauge_upload(random_gauge(rng, d.V))
        E, O = d.field(random_spinor(rng, d.Vh)), d.field(random_spinor(rng, d.Vh))
        En, On, W = d.field(), d.field(), d.field()
        for v in (0, 10, -1, 0, 10):
            d.ck(d.lib.tmb_set_tuning(v, 1, 0))
            n = 500
            for _ in range(20):
                d.lib.tmb_Hopping_Matrix(0, W, E); d.lib.tmb_Hopping_Matrix(1, En, W)
            d.timer_start()
            for _ in range(n):
                d.lib.tmb_Hopping_Matrix(0, W, E); d.lib.tmb_Hopping_Matrix(1, En, W)
            us = d.timer_stop() * 1e3 / (2 * n)
            best = 1e9
            for rep in range(3):
                d.call("field_zero", On); d.ck(d.lib.tmb_sync())
                t0 = time.perf_counter()
                it = d.call("invert_eo", En, On, E, O, 1e-22, 5000, 1)
                best = min(best, time.perf_counter() - t0)
            print(f"{dims} variant {v:2d}: {us:7.2f} us/hop ({1536.0 * d.Vh / us / 1e3:7.1f} GB/s), invert_eo {it} it. {1e6 * best / it:7.2f} us/iteration", flush=True)
        d.close()
elif what == "restart":
    # fixed cost of a solve that converges at once (HMC derivative calls with a good chronological guess): CUDA-graph
    # capture + instantiation per solve against plain launches (tmb_set_overlap bit 2)
    for dims in ((32, 16, 16, 16), (48, 24, 24, 24)):
        rng = np.random.default_rng(1)
        d = tm.Device(*dims)
        d.set_params(KAPPA, GMU)
        d.gauge_upload(random_gauge(rng, d.V))
        k, x = d.field(random_spinor(rng, d.Vh)), d.field()
        it = d.call("cg_her", x, k, 5000, 1e-20, 1)
        for flags in (0, 4, 0, 4):
            d.ck(d.lib.tmb_set_overlap(flags))
            ts = []
            for rep in range(6):
                d.ck(d.lib.tmb_sync())
                t0 = time.perf_counter()
                it2 = d.call("cg_her", x, k, 5000, 1e-18, 1)  # x already solves to 1e-20
                ts.append(time.perf_counter() - t0)
            print(f"{dims} first solve {it} it.; restart from the solution: {it2} it., flags {flags}: min {1e6 * min(ts):8.1f} us  all {[round(1e6 * t) for t in ts]}", flush=True)
        d.close()
elif what == "hints":
    # small lattices fit (partly) in the 126 MB L2: is the evict-first policy on the gauge stream still right there?
    for dims in ((8, 8, 8, 8), (16, 8, 8, 8), (16, 16, 16, 16), (32, 16, 16, 16), (24, 24, 24, 24)):
        rng = np.random.default_rng(1)
        d = tm.Device(*dims)
        d.set_params(KAPPA, GMU)
        d.gauge_upload(random_gauge(rng, d.V))
        E, O = d.field(random_spinor(rng, d.Vh)), d.field(random_spinor(rng, d.Vh))
        En, On, W = d.field(), d.field(), d.field()
        for hints in (1, 0, 1, 0):
            d.ck(d.lib.tmb_set_tuning(-1, hints, 0))
            n = 500
            for _ in range(20):
                d.lib.tmb_Hopping_Matrix(0, W, E); d.lib.tmb_Hopping_Matrix(1, En, W)
            d.timer_start()
            for _ in range(n):
                d.lib.tmb_Hopping_Matrix(0, W, E); d.lib.tmb_Hopping_Matrix(1, En, W)
            us = d.timer_stop() * 1e3 / (2 * n)
            best = 1e9
            for rep in range(3):
                d.call("field_zero", On); d.ck(d.lib.tmb_sync())
                t0 = time.perf_counter()
                it = d.call("invert_eo", En, On, E, O, 1e-22, 5000, 1)
                best = min(best, time.perf_counter() - t0)
            print(f"{dims} hints {hints}: {us:7.2f} us/hop ({1536.0 * d.Vh / us / 1e3:7.1f} GB/s), invert_eo {it} it. {1e6 * best / it:7.2f} us/iteration", flush=True)
        d.close()
elif what == "selfnorm":
    for dims in ((48, 24, 24, 24), (32, 16, 16, 16)):
        rng = np.random.default_rng(1)
        d = tm.Device(*dims)
        d.set_params(KAPPA, GMU)
        d.gauge_upload(random_gauge(rng, d.V))
        E, O = d.field(random_spinor(rng, d.Vh)), d.field(random_spinor(rng, d.Vh))
        En, On = d.field(), d.field()
        for flags in (16, 0, 16, 0):
            d.ck(d.lib.tmb_set_overlap(flags))
            for solver in ("invert_eo", "invert_eo_mixed"):
                best = 1e9
                for rep in range(4):
                    d.call("field_zero", On); d.ck(d.lib.tmb_sync())
                    t0 = time.perf_counter()
                    it = d.call(solver, En, On, E, O, 1e-22, 5000, 1)
                    best = min(best, time.perf_counter() - t0)
                print(f"{dims} overlap flags {flags:2d} {solver:16s}: {it} iterations, {1e3 * best:8.3f} ms ({1e6 * best / it:7.2f} us/iteration)", flush=True)
        d.close()
elif what == "cgpf":
    # CG time-to-solution with / without the L2 prefetch of the epilogue operands (tmb_set_overlap bit 3)
    dims = (48, 24, 24, 24)
    rng = np.random.default_rng(1)
    d = tm.Device(*dims)
    d.set_params(KAPPA, GMU)
    d.gauge_upload(random_gauge(rng, d.V))
    E, O = d.field(random_spinor(rng, d.Vh)), d.field(random_spinor(rng, d.Vh))
    En, On, W = d.field(), d.field(), d.field()
    for flags in (0, 8, 0, 8):
        d.ck(d.lib.tmb_set_overlap(flags))
        best = 1e9
        for rep in range(4):
            d.call("field_zero", On); d.ck(d.lib.tmb_sync())
            t0 = time.perf_counter()
            it = d.call("invert_eo", En, On, E, O, 1e-22, 5000, 1)
            best = min(best, time.perf_counter() - t0)
        tq = timeit(d, lambda: d.call("Qtm_pm_psi", W, E), n=200)
        print(f"overlap flags {flags}: invert_eo {it} iterations, best {1e3 * best:8.3f} ms ({1e6 * best / it:7.2f} us/iteration); Qtm_pm_psi {1e6 * tq:7.2f} us", flush=True)
    d.close()
elif what in ("e2e", "chunk"):
    dims = (48, 24, 24, 24)
    rng = np.random.default_rng(1)
    d = tm.Device(*dims)
    d.set_params(KAPPA, GMU)
    g = random_gauge(rng, d.V)
    d.gauge_upload(g)
    Vh = d.Vh

    def pinned(shape):
        n = int(np.prod(shape))
        p = d.lib.tmb_host_alloc(n * 8)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,)).reshape(shape)
    D = tm.DropIn(*dims)
    D.set_params(KAPPA, GMU)
    D.set_gauge(g)
    if what == "chunk":
        hk, h1 = pinned((Vh, 24)), pinned((Vh, 24))
        hk[:] = random_spinor(rng, Vh)
        D.Hopping_Matrix(0, h1, hk)
        for nch in (4, 6, 8, 12, 16, 24, 48, 8):
            d.ck(d.lib.tmb_set_host_chunks(nch))
            t = timeit(d, lambda: D.Hopping_Matrix(0, h1, hk), n=10)
            print(f"host-pointer Hopping_Matrix, {nch:2d} chunks: {1e3 * t:7.3f} ms  ({Vh * 192 / t / 1e9:5.1f} GB/s each way)", flush=True)
    else:
        E, O = random_spinor(rng, Vh), random_spinor(rng, Vh)
        dE, dO, dEn, dOn = d.field(E), d.field(O), d.field(), d.field()
        for _ in range(2):
            d.call("field_zero", dOn); d.ck(d.lib.tmb_sync())
            t0 = time.perf_counter()
            it = d.call("invert_eo", dEn, dOn, dE, dO, 1e-14, 5000, 1)
            print("device invert_eo", it, "iterations", time.perf_counter() - t0, "s; stats", d.solver_stats(), flush=True)
        hE, hO, hEn, hOn = (pinned((Vh, 24)) for _ in range(4))
        hE[:] = E; hO[:] = O
        C.c_int.in_dll(D.lib, "g_debug_level").value = 2
        sp = tm.capi.SolverParams()
        for _ in range(3):
            hOn[:] = 0
            t0 = time.perf_counter()
            it = D.invert_eo(hEn, hOn, hE, hO, 1e-14, 5000, 1, 1, 0, 1, 0, None, sp, 0, 0, 0, 18)
            print("drop-in invert_eo", it, "iterations", time.perf_counter() - t0, "s; stats", d.solver_stats(), flush=True)
    d.close()

#!/usr/bin/env python
"""A/B of the CTA tile traversal of the hopping kernels (tmb_set_tile, tmb_geom.h: 2 x 2 x 32 tiles against 128 consecutive
sites; memory layout and results unchanged).  One GPU; per lattice: the plain hop in double / float / 12-real links, burst
(200 pairs) and sustained (>= 0.6 s), the same through the peer-mode kernel (loop-back: this rank is its own T neighbour), the
CG (iterations must not change), and the two-flavour operator.  Prints one JSON line per lattice.
    python scripts/tile_ab.py [TxLXxLYxLZ ...]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import tmlqcd_b200 as tm
from bench import numpy_gauge, measured_peaks

PEAK = measured_peaks()[0]


def timed(d, fn, n):
    d.ck(d.lib.tmb_sync()); d.timer_start()
    for _ in range(n):
        fn()
    return d.timer_stop() / n * 1e3  # us per call


def burst_and_sustained(d, fn, per_call_sites_bytes):
    timed(d, fn, 20)
    b = timed(d, fn, 200)
    n = int(max(400, 0.6e6 / b))
    s = timed(d, fn, n)
    return {"burst_us": round(b, 2), "sustained_us": round(s, 2), "frac_burst": round(per_call_sites_bytes / b / 1e3 / PEAK, 4),
            "frac_sustained": round(per_call_sites_bytes / s / 1e3 / PEAK, 4)}


def run(dims):
    rng = np.random.default_rng(5)
    g = numpy_gauge(dims, 11)
    out = {"lattice_TxLXxLYxLZ": list(dims)}
    for loop in (0, 2):
        d = tm.Device(*dims)
        d.set_params(0.16, 0.01)
        if loop:
            d.ck(d.lib.tmb_comm_loopback(loop))
        d.gauge_upload(g)
        Vh = d.Vh
        src = rng.normal(size=(Vh, 24))
        f0, f1, f2 = d.field(src), d.field(), d.field()
        g0, g1, g2 = d.field32(src.astype(np.float32)), d.field32(), d.field32()
        key = "peer_loopback" if loop else "single"
        res = {}
        ref = {}
        E, O = d.field(src), d.field(rng.normal(size=(Vh, 24)))
        for tile in (0, 1):
            d.ck(d.lib.tmb_set_tile(tile))
            r = {}
            def pair():
                d.lib.tmb_Hopping_Matrix(0, f1, f0); d.lib.tmb_Hopping_Matrix(1, f2, f1)
            def pair32():
                d.lib.tmb_Hopping_Matrix_32(0, g1, g0); d.lib.tmb_Hopping_Matrix_32(1, g2, g1)
            x = burst_and_sustained(d, pair, 2 * 1536.0 * Vh); r["hop_f64"] = x
            x = burst_and_sustained(d, pair32, 2 * 768.0 * Vh); r["hop_f32"] = x
            if not loop:
                d.ck(d.lib.tmb_set_compression(12))
                r["hop_f64_12real"] = burst_and_sustained(d, pair, 2 * 1152.0 * Vh)
                d.ck(d.lib.tmb_set_compression(18))
            # results must be bit-identical
            d.lib.tmb_Hopping_Matrix(1, f1, f0); h = d.download(f1)
            d.call("Qtm_pm_psi", f2, f0); q = d.download(f2)
            if tile == 0:
                ref["h"], ref["q"] = h, q
            else:
                r["bit_identical_to_linear"] = bool(np.array_equal(h, ref["h"]) and np.array_equal(q, ref["q"]))
            # CG: same sources for both traversals
            En, On = d.field(), d.field()
            best = None
            for _ in range(2):
                d.call("field_zero", On)
                it = d.call("invert_eo", En, On, E, O, 1e-14, 5000, 1)
                _, rr, loop_s = d.solver_stats()
                best = loop_s if best is None else min(best, loop_s)
            r["cg"] = {"iterations": it, "cg_loop_s": round(best, 5), "ms_per_iteration": round(1e3 * best / it, 4), "final_rr": rr}
            d.free(En, On)
            res["tile" if tile else "linear"] = r
        out[key] = res
        if not loop and dims[0] * dims[1] * dims[2] * dims[3] <= 2 * 32 ** 3 * 64:
            # two-flavour operator (K6a on one rank)
            d.ck(d.lib.tmb_set_nd(0.139, 0.15, 0.9))
            nd = [d.field(rng.normal(size=(Vh, 24))) for _ in range(4)]
            nd32 = [d.field32(rng.normal(size=(Vh, 24)).astype(np.float32)) for _ in range(4)]
            ndr = {}
            keep = None
            for tile in (0, 1):
                d.ck(d.lib.tmb_set_tile(tile))
                a = burst_and_sustained(d, lambda: d.lib.tmb_Qtm_pm_ndpsi(nd[2], nd[3], nd[0], nd[1]), 8448.0 * Vh)
                b = burst_and_sustained(d, lambda: d.lib.tmb_Qtm_pm_ndpsi_32(nd32[2], nd32[3], nd32[0], nd32[1]), 4224.0 * Vh)
                o = d.download(nd[2])
                if tile == 0:
                    keep = o
                ndr["tile" if tile else "linear"] = {"Qtm_pm_ndpsi": a, "Qtm_pm_ndpsi_32": b}
            ndr["bit_identical_to_linear"] = bool(np.array_equal(o, keep))
            out["two_flavour"] = ndr
        d.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    lat = sys.argv[1:] or ["48x24x24x24", "12x48x48x48", "64x32x32x32"]
    for s in lat:
        run(tuple(int(x) for x in s.split("x")))

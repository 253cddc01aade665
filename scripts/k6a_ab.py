#!/usr/bin/env python
"""Tuning experiments on one GPU: (1) L2 bulk prefetch of the gauge rows a wave AHEAD (tmb_set_overlap bit 1 +
tmb_set_prefetch_distance) for the one-field hop and for the one-thread-two-flavours kernel K6a, (2) K6a with 3 CTAs per SM
(168 registers, spills; tmb_set_hop2_variant(3)).  Burst (200 calls) and sustained (>= 0.6 s).  One JSON line per lattice."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmlqcd_b200 as tm
from bench import numpy_gauge, measured_peaks
from tile_ab import burst_and_sustained


def run(dims):
    rng = np.random.default_rng(5)
    d = tm.Device(*dims)
    d.set_params(0.16, 0.01); d.ck(d.lib.tmb_set_nd(0.139, 0.15, 0.9))
    d.gauge_upload(numpy_gauge(dims, 11))
    Vh = d.Vh
    f = [d.field(rng.normal(size=(Vh, 24))) for _ in range(4)]
    f32 = [d.field32(rng.normal(size=(Vh, 24)).astype(np.float32)) for _ in range(4)]
    out = {"lattice_TxLXxLYxLZ": list(dims), "hop_f64": {}, "hop_f32": {}, "Qtm_pm_ndpsi": {}, "Qtm_pm_ndpsi_32": {}}
    def pair():
        d.lib.tmb_Hopping_Matrix(0, f[2], f[0]); d.lib.tmb_Hopping_Matrix(1, f[3], f[2])
    def pair32():
        d.lib.tmb_Hopping_Matrix_32(0, f32[2], f32[0]); d.lib.tmb_Hopping_Matrix_32(1, f32[3], f32[2])
    def nd():
        d.lib.tmb_Qtm_pm_ndpsi(f[2], f[3], f[0], f[1])
    def nd32():
        d.lib.tmb_Qtm_pm_ndpsi_32(f32[2], f32[3], f32[0], f32[1])
    nd(); ref = d.download(f[2]).copy()
    for pf, dist in [(0, 0), (1, 0), (1, 148), (1, 444), (1, 888)]:
        d.ck(d.lib.tmb_set_overlap(2 if pf else 0)); d.ck(d.lib.tmb_set_prefetch_distance(dist))
        key = f"prefetch{dist}" if pf else "baseline"
        out["hop_f64"][key] = burst_and_sustained(d, pair, 2 * 1536.0 * Vh)
        out["hop_f32"][key] = burst_and_sustained(d, pair32, 2 * 768.0 * Vh)
    for variant in (0, 3):
        d.ck(d.lib.tmb_set_hop2_variant(variant))
        for pf, dist in [(0, 0), (1, 0), (1, 148), (1, 296), (1, 592)]:
            d.ck(d.lib.tmb_set_overlap(2 if pf else 0)); d.ck(d.lib.tmb_set_prefetch_distance(dist))
            key = f"variant{variant}_" + (f"prefetch{dist}" if pf else "baseline")
            out["Qtm_pm_ndpsi"][key] = burst_and_sustained(d, nd, 8448.0 * Vh)
            assert np.array_equal(d.download(f[2]), ref), key
            if variant == 0:
                out["Qtm_pm_ndpsi_32"][key] = burst_and_sustained(d, nd32, 4224.0 * Vh)
    d.ck(d.lib.tmb_set_overlap(0)); d.ck(d.lib.tmb_set_prefetch_distance(0)); d.ck(d.lib.tmb_set_hop2_variant(-1))
    d.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for s in (sys.argv[1:] or ["64x32x32x32", "48x24x24x24"]):
        run(tuple(int(x) for x in s.split("x")))

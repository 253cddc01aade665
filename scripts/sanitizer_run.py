#!/usr/bin/env python
"""One pass over every kernel family on a tiny lattice: an all-kernel smoke that prints the same numbers for the
three hop paths, and the workload to put under compute-sanitizer where that tool is allowed (it is closed on the
development pool: `compute-sanitizer --tool memcheck|racecheck python scripts/sanitizer_run.py`):
plain, T-split halo (loopback 1) and peer-mode (loopback 2) paths; double, float, 12-real links; CG, mixed CG, ND doublet,
fermion force, a det monomial, plaquette, host-pointer pipelined hop, lexicographic D_psi."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tmlqcd_b200 as tm
from conftest import random_gauge, random_spinor

dims = (4, 4, 6, 8)
for loopback in (0, 1, 2):
    rng = np.random.default_rng(3)
    d = tm.Device(*dims)
    d.set_params(0.16, 0.0032, (1., 0., 0.3, 0.))
    d.ck(d.lib.tmb_set_nd(0.139, 0.15, 0.9))
    if loopback:
        d.ck(d.lib.tmb_comm_loopback(loopback))
    g = random_gauge(rng, d.V)
    d.gauge_upload(g)
    k, p = d.field(random_spinor(rng, d.Vh)), d.field(random_spinor(rng, d.Vh))
    l, m, x, y = d.field(), d.field(), d.field(), d.field()
    for ieo in (0, 1):
        d.call("Hopping_Matrix", ieo, l, k)
        d.call("tm_times_Hopping_Matrix", ieo, l, k, 0.3, 0.1)
        d.call("tm_sub_Hopping_Matrix", ieo, l, p, k, 1.0, 0.2)
    d.call("Qtm_pm_psi", l, k); d.call("M_full", l, m, k, p); d.call("D_psi_eo", l, m, k, p)
    it = d.call("cg_her", x, k, 500, 1e-18, 1)
    d.call("field_zero", x); d.call("field_zero", y)
    it2 = d.call("invert_eo", l, x, k, p, 1e-18, 500, 1)
    itm = d.call("mixed_cg_her", x, k, 500, 1e-18, 1)
    itr = d.call("rg_mixed_cg_her", x, k, 500, 1e-18, 1)
    d.call("Qtm_pm_ndpsi", l, m, k, p)
    d.call("field_zero", x); d.call("field_zero", y)
    itn = d.call("cg_her_nd", x, y, k, p, 500, 1e-16, 1)
    d.ck(d.lib.tmb_set_compression(12)); d.call("Hopping_Matrix", 0, l, k); d.call("field_zero", x); d.call("cg_her", x, k, 500, 1e-18, 1)
    d.ck(d.lib.tmb_set_compression(18))
    for v in (10, -1):
        d.ck(d.lib.tmb_set_tuning(v, -1, 0)); d.call("Qtm_pm_psi", l, k)
    d.call("derivative_zero"); d.call("deriv_Sb", 0, k, p, 0.7); d.call("deriv_Sb", 1, p, k, -0.4)
    df = d.derivative_download()
    assert d.lib.tmb_monomial_add(1, 0.16, 0.0032, 0.16, 0.032, 1, 500, 1e-16, 1e-18, 2) == 0
    e0 = C.c_double(); d.ck(d.lib.tmb_monomial_heatbath(0, k, C.byref(e0)))
    for _ in range(2):
        d.ck(d.lib.tmb_monomial_derivative(0))
    dH = C.c_double(); d.ck(d.lib.tmb_monomial_acc(0, C.byref(dH)))
    pl = C.c_double(); d.ck(d.lib.tmb_measure_plaquette(C.byref(pl)))
    if not loopback:
        hk, hl = random_spinor(rng, d.Vh), np.zeros((d.Vh, 24))
        d.ck(d.lib.tmb_set_host_chunks(4))
        d.ck(d.lib.tmb_Hopping_Matrix_host(0, hl.ctypes.data_as(C.c_void_p), hk.ctypes.data_as(C.c_void_p), 0, 1., 0.))
    print(f"loopback {loopback}: cg {it} invert {it2} mixed {itm} rg {itr} nd {itn} |df| {np.linalg.norm(df):.6e} plaq {pl.value / (6 * d.V):.6f} dH {dH.value:.2e}", flush=True)
    d.close()
print("sanitizer_run done")

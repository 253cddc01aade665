#!/usr/bin/env python
"""Where does the time of the host-pointer (drop-in) entry points go?  Times field upload/download
(PCIe), the pipelined host-pointer Hopping_Matrix and repeated invert_eo calls through the drop-in."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import tmlqcd_b200 as tm
from conftest import random_gauge, random_spinor

dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "48x24x24x24").split("x"))
rng = np.random.default_rng(1)
d = tm.Device(*dims)
D = tm.DropIn(*dims)
D.set_params(0.16, 0.0032)
g = random_gauge(rng, d.V)
D.set_gauge(g)


def pinned(shape):
    n = int(np.prod(shape))
    p = d.lib.tmb_host_alloc(n * 8)
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,)).reshape(shape)


def T(label, fn, reps=5):
    ts = []
    for _ in range(reps):
        d.ck(d.lib.tmb_sync())
        t0 = time.perf_counter(); fn(); d.ck(d.lib.tmb_sync()); ts.append(time.perf_counter() - t0)
    print(f"{label:50s} min {1e3 * min(ts):9.3f} ms   all {[round(1e3 * t, 2) for t in ts]}", flush=True)
    return min(ts)


a, b, c = C.c_double(), C.c_double(), C.c_double()
d.ck(d.lib.tmb_measure_pcie_gbs(64 << 20, 20, C.byref(a), C.byref(b), C.byref(c)))
print(f"host link (pinned, 64 MB copies): H2D {a.value:.1f} GB/s, D2H {b.value:.1f} GB/s, both at once {c.value:.1f} GB/s per direction", flush=True)
hk, h1, h2, h3 = (pinned((d.Vh, 24)) for _ in range(4))
hk[:] = random_spinor(rng, d.Vh); h1[:] = random_spinor(rng, d.Vh)
f0, f1 = d.field(), d.field()
mb = d.Vh * 192 / 1e6
t = T("field_upload (pinned)", lambda: d.upload(f0, hk)); print(f"   -> {mb / t / 1e3:.1f} GB/s")
t = T("field_download (pinned, fresh numpy out)", lambda: d.download(f0))
t = T("tmb_field_download (pinned out)", lambda: d.ck(d.lib.tmb_field_download(h2.ctypes.data_as(C.c_void_p), f0))); print(f"   -> {mb / t / 1e3:.1f} GB/s")
pg = np.array(hk)
t = T("field_upload (pageable)", lambda: d.upload(f0, pg)); print(f"   -> {mb / t / 1e3:.1f} GB/s")
T("drop-in Hopping_Matrix (first: gauge upload)", lambda: D.Hopping_Matrix(0, h2, hk), reps=1)
t = T("drop-in Hopping_Matrix (pinned)", lambda: D.Hopping_Matrix(0, h2, hk)); print(f"   -> {mb / t / 1e3:.1f} GB/s each way")
pgk, pgl = np.array(hk), np.zeros_like(hk)
t = T("drop-in Hopping_Matrix (pageable numpy buffers, page-locked on first sight)", lambda: D.Hopping_Matrix(0, pgl, pgk)); print(f"   -> {mb / t / 1e3:.1f} GB/s each way")
def pair():
    D.Hopping_Matrix(0, h2, hk); D.Hopping_Matrix(1, h3, h2)
t = T("drop-in EO+OE pair (pinned)", pair, reps=8); print(f"   -> {2 * mb / t / 1e3:.1f} GB/s each way")
for nch in (4, 8, 16, 24, 48, 0):
    d.ck(d.lib.tmb_set_host_chunks(nch))
    t = T(f"drop-in Hopping_Matrix (pinned), {nch} chunks", lambda: D.Hopping_Matrix(0, h2, hk)); print(f"   -> {mb / t / 1e3:.1f} GB/s each way")
d.ck(d.lib.tmb_set_host_chunks(0))
t = T("device Hopping_Matrix", lambda: d.lib.tmb_Hopping_Matrix(0, f1, f0))
hk2 = pinned((d.Vh, 24)); hk2[:] = hk
sp = tm.capi.SolverParams()
def solve():
    h3[:] = 0
    return D.invert_eo(h2, h3, hk, h1, 1e-14, 5000, 1, 1, 0, 1, 0, None, sp, 0, 0, 0, 18)
T("drop-in invert_eo (pinned), repeated", solve, reps=4)
print("stats", d.solver_stats())
dE, dO, dEn, dOn = d.field(hk), d.field(h1), d.field(), d.field()
def solve_dev():
    d.call("field_zero", dOn)
    return d.call("invert_eo", dEn, dOn, dE, dO, 1e-14, 5000, 1)
T("device invert_eo, repeated", solve_dev, reps=4)
print("stats", d.solver_stats())
d.close()

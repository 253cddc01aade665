#!/usr/bin/env python
"""The pipelined host-pointer Hopping_Matrix (tmb_Hopping_Matrix_host): the timeline of one call (upload / kernels / download
completion per chunk, tmb_host_hop_timeline) and the time per call for a list of explicit chunk schedules
(tmb_set_host_chunk_sizes).  usage: pipe_diag.py [TxLXxLYxLZ]"""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmlqcd_b200 as tm
from bench import numpy_gauge

dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "48x24x24x24").split("x"))
T = dims[0]
d = tm.Device(*dims); lib = d.lib
d.set_params(0.16, 0.01)
d.gauge_upload(numpy_gauge(dims, 3))


def pinned(n):
    p = lib.tmb_host_alloc(n * 8)
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,))


hk, hl = pinned(d.Vh * 24), pinned(d.Vh * 24)
hk[:] = np.random.default_rng(1).normal(size=d.Vh * 24)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
a, b, c = C.c_double(), C.c_double(), C.c_double()
d.ck(lib.tmb_measure_pcie_gbs(64 << 20, 20, C.byref(a), C.byref(b), C.byref(c)))
print(f"link: H2D {a.value:.1f}  D2H {b.value:.1f}  duplex {c.value:.1f} GB/s per direction", flush=True)


def time_call(reps=10):
    ts = []
    for _ in range(reps):
        d.ck(lib.tmb_sync()); t0 = time.perf_counter()
        d.ck(lib.tmb_Hopping_Matrix_host(0, vp(hl), vp(hk), 0, 0., 0.))
        ts.append(time.perf_counter() - t0)
    return 1e6 * min(ts), 1e6 * float(np.median(ts))


def timeline(label):
    out = np.zeros(3 * 400)
    n = lib.tmb_host_hop_timeline(0, vp(hl), vp(hk), vp(out), 400)
    assert n > 0, n
    rows = out[:3 * n].reshape(n, 3)
    print(f"timeline [{label}] (us since start):")
    for kind, name in ((0, "upload done  "), (1, "kernels done "), (2, "download done"), (3, "end          ")):
        print("  " + name + " " + " ".join(f"{int(r[1])}:{r[2]:.0f}" for r in rows if int(r[0]) == kind))


ref = None
res = {}
scheds = {"auto": [], "5x8+7+1": [5] * 8 + [7, 1], "6x7+5+1": [6] * 7 + [5, 1], "2,4,6x6,5,1": [2, 4] + [6] * 6 + [5, 1],
          "8x5,7,1": [8] * 5 + [7, 1], "1,2,4,8x4,4,2,2,1": [1, 2, 4, 8, 8, 8, 8, 4, 2, 2, 1], "4x12": [4] * 12, "3x16": [3] * 16,
          "12,12,12,8,3,1": [12, 12, 12, 8, 3, 1], "16,16,8,4,2,1,1": [16, 16, 8, 4, 2, 1, 1], "2,6,8x4,4,2,1,1": [2, 6, 8, 8, 8, 8, 4, 2, 1, 1],
          "8x6": [8] * 6, "4,8x5,4": [4] + [8] * 5 + [4], "2,2,4,8x4,4,2,1,1": [2, 2, 4, 8, 8, 8, 8, 4, 2, 1, 1], "24,12,6,3,2,1": [24, 12, 6, 3, 2, 1]}
for name, sz in scheds.items():
    if sz and sum(sz) != T:
        continue
    arr = (C.c_int * max(1, len(sz)))(*sz)
    d.ck(lib.tmb_set_host_chunk_sizes(arr, len(sz)))
    time_call(3)
    mn, med = time_call(12)
    out = hl.copy()
    if ref is None:
        ref = out
    ok = bool(np.array_equal(out, ref))
    res[name] = {"min_us": round(mn, 1), "median_us": round(med, 1), "same_result": ok}
    print(f"{name:24s} min {mn:8.1f} us  median {med:8.1f} us  same result {ok}", flush=True)
d.ck(lib.tmb_set_host_chunk_sizes((C.c_int * 1)(0), 0))
timeline("auto")
for name in ("1,2,4,8x4,4,2,2,1",):
    sz = scheds[name]
    arr = (C.c_int * max(1, len(sz)))(*sz)
    d.ck(lib.tmb_set_host_chunk_sizes(arr, len(sz)))
    timeline(name)
print(json.dumps(res))
d.close()

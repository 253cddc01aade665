#!/usr/bin/env python
"""A/B of CG time per iteration between builds of the library: python scripts/ab_cg.py lib1.so lib2.so ...
Each library is loaded in its own subprocess (raw ctypes, only symbols every build of this round has) on the same
box, the same lattice and inputs, several alternating rounds."""
import ctypes as C
import subprocess
import sys
import time

import numpy as np

if len(sys.argv) > 2 and sys.argv[1] == "--one":
    lib = C.CDLL(sys.argv[2])
    dims = tuple(int(x) for x in sys.argv[3].split("x"))
    T, LX, LY, LZ = dims
    V = T * LX * LY * LZ; Vh = V // 2
    vp, d, i = C.c_void_p, C.c_double, C.c_int
    lib.tmb_field_alloc.restype = vp
    lib.tmb_last_error.restype = C.c_char_p
    def ck(rc):
        if rc < 0:
            raise RuntimeError(lib.tmb_last_error().decode())
        return rc
    ck(lib.tmb_init(T, LX, LY, LZ, 0))
    theta = (d * 4)(0., 0., 0., 0.)
    ck(lib.tmb_set_boundary(d(0.16), theta)); ck(lib.tmb_set_mu(d(0.0032)))
    rng = np.random.default_rng(1)
    a = rng.normal(size=(V * 4, 3, 3)) + 1j * rng.normal(size=(V * 4, 3, 3))
    q, r = np.linalg.qr(a)
    dg = np.diagonal(r, axis1=1, axis2=2)
    q = q * (dg / np.abs(dg))[:, None, :]
    q = q / np.linalg.det(q)[:, None, None] ** (1. / 3.)
    g = np.ascontiguousarray(q.reshape(-1, 9)).view(np.float64).reshape(V, 4, 18)
    ck(lib.tmb_gauge_upload(g.ctypes.data_as(vp)))
    F = [vp(lib.tmb_field_alloc()) for _ in range(4)]
    for f in F[:2]:
        h = np.ascontiguousarray(rng.normal(scale=np.sqrt(0.5), size=(Vh, 24)))
        ck(lib.tmb_field_upload(f, h.ctypes.data_as(vp)))
    lib.tmb_invert_eo.argtypes = [vp, vp, vp, vp, d, i, i]
    best = 1e9
    for rep in range(5):
        ck(lib.tmb_field_zero(F[3])); ck(lib.tmb_sync())
        t0 = time.perf_counter()
        it = lib.tmb_invert_eo(F[2], F[3], F[0], F[1], 1e-22, 5000, 1)
        best = min(best, time.perf_counter() - t0)
    print(f"{sys.argv[2]:60s} {sys.argv[3]}: {it} iterations, {1e6 * best / it:8.2f} us/iteration", flush=True)
    lib.tmb_finalize()
    sys.exit(0)

libs = sys.argv[1:]
for dims in ("48x24x24x24", "32x16x16x16"):
    for rnd in range(2):
        for lp in libs:
            subprocess.run([sys.executable, __file__, "--one", lp, dims])

#!/usr/bin/env python
"""A few launches of the kernels outside the headline hop: deriv_kernel (fermion force), hop2_kernel (two-flavour
hop with fused ND epilogues), the peer-mode hop against itself, cdot/caxpy.  For ncu captures (profiles/)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tmlqcd_b200 as tm
from conftest import random_gauge, random_spinor
dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "48x24x24x24").split("x"))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rng = np.random.default_rng(3)
d = tm.Device(*dims)
d.set_params(0.16, 0.0032); d.ck(d.lib.tmb_set_nd(0.139, 0.15, 1.0))
d.gauge_upload(random_gauge(rng, d.V))
f = [d.field(random_spinor(rng, d.Vh)) for _ in range(2)] + [d.field() for _ in range(2)]
d.call("derivative_zero")
for _ in range(n):
    d.call("deriv_Sb", 0, f[0], f[1], 1.0); d.call("deriv_Sb", 1, f[1], f[0], 1.0)
    d.call("Qtm_pm_ndpsi", f[2], f[3], f[0], f[1])
d.ck(d.lib.tmb_sync())
d.ck(d.lib.tmb_comm_loopback(2)); d.gauge_upload(random_gauge(rng, d.V))
for _ in range(n):
    d.call("Hopping_Matrix", 0, f[2], f[0]); d.call("Hopping_Matrix", 1, f[3], f[2])
d.ck(d.lib.tmb_sync())
print("ok launches", d.lib.tmb_launch_count())
d.close()

#!/usr/bin/env python
"""Two applications of Qtm_pm_ndpsi (K6a: hop2_kernel MODE 1 and MODE 2) for an ncu capture with warp-state sections."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmlqcd_b200 as tm
from bench import numpy_gauge
dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "48x24x24x24").split("x"))
rng = np.random.default_rng(3)
d = tm.Device(*dims)
d.set_params(0.16, 0.0032); d.ck(d.lib.tmb_set_nd(0.139, 0.15, 1.0))
d.gauge_upload(numpy_gauge(dims, 4))
f = [d.field(rng.normal(size=(d.Vh, 24))) for _ in range(4)]
for _ in range(2):
    d.lib.tmb_Qtm_pm_ndpsi(f[2], f[3], f[0], f[1])
    d.lib.tmb_Hopping_Matrix(0, f[2], f[0])
d.ck(d.lib.tmb_sync()); print("ok", d.lib.tmb_launch_count()); d.close()

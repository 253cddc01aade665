#!/usr/bin/env python
"""A few launches of the hopping kernels with the linear CTA traversal (tmb_set_tile(0)) and then with the 2 x 2 x 32 tiles
(tmb_set_tile(1)) for one ncu pass: duration, DRAM bytes, L1 hit rate, L2 -> SM sectors.  usage: profile_tile.py [TxLXxLYxLZ]"""
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
src = [rng.normal(size=(d.Vh, 24)) for _ in range(2)]
f = [d.field(s) for s in src] + [d.field() for _ in range(2)]
f32 = [d.field32(s.astype(np.float32)) for s in src] + [d.field32() for _ in range(2)]
for tile in (0, 1):
    d.ck(d.lib.tmb_set_tile(tile))
    for _ in range(2):
        d.lib.tmb_Hopping_Matrix(0, f[2], f[0]); d.lib.tmb_Hopping_Matrix(1, f[3], f[2])
    for _ in range(2):
        d.lib.tmb_Hopping_Matrix_32(0, f32[2], f32[0]); d.lib.tmb_Hopping_Matrix_32(1, f32[3], f32[2])
    d.ck(d.lib.tmb_set_compression(12))
    for _ in range(2):
        d.lib.tmb_Hopping_Matrix(0, f[2], f[0]); d.lib.tmb_Hopping_Matrix(1, f[3], f[2])
    d.ck(d.lib.tmb_set_compression(18))
    for _ in range(2):
        d.lib.tmb_Qtm_pm_ndpsi(f[2], f[3], f[0], f[1])
    d.lib.tmb_Qtm_pm_ndpsi_32(f32[2], f32[3], f32[0], f32[1])
    it = d.call("cg_her", f[2], f[0], 3, 1e-30, 1)
d.ck(d.lib.tmb_sync())
print("ok launches", d.lib.tmb_launch_count())
d.close()

#!/usr/bin/env python
"""A few hops on a Z-split slab with this rank as its own z neighbour (tmb_comm_loopback_z(2): faces pushed through peer
memory), for an ncu launch list: hop kernel, face pack, flag kernel, fix-up.  usage: profile_zsplit.py [TxLXxLYxLZ]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmlqcd_b200 as tm
from bench import numpy_gauge
dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "24x48x48x24").split("x"))
rng = np.random.default_rng(3)
d = tm.Device(*dims)
d.set_params(0.16, 0.0032)
d.ck(d.lib.tmb_comm_loopback_z(2))
d.gauge_upload(numpy_gauge(dims, 4))
f = [d.field(rng.normal(size=(d.Vh, 24))), d.field(), d.field()]
for _ in range(3):
    d.lib.tmb_Hopping_Matrix(0, f[1], f[0]); d.lib.tmb_Hopping_Matrix(1, f[2], f[1])
d.ck(d.lib.tmb_sync())
d.timer_start()
for _ in range(50):
    d.lib.tmb_Hopping_Matrix(0, f[1], f[0]); d.lib.tmb_Hopping_Matrix(1, f[2], f[1])
ms = d.timer_stop()
print(f"z loop-back (peer push): {1e3 * ms / 100:.1f} us per hop", flush=True)
d.timer_start()
for _ in range(50):
    d.lib.tmb_Hopping_Matrix_nocom(0, f[1], f[0]); d.lib.tmb_Hopping_Matrix_nocom(1, f[2], f[1])
ms = d.timer_stop()
print(f"nocom: {1e3 * ms / 100:.1f} us per hop", flush=True)
d.close()

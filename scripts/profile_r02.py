#!/usr/bin/env python
"""A few launches of each kernel family for the round-2 ncu captures (profiles/r02_*): the double hop (18- and 12-real links),
the float hop, the two-flavour hop in its three forms (one thread carrying both flavours, two flavour groups of warps per CTA in
double and in float), the fermion force, the CG's fused sweep.  usage: profile_r02.py [TxLXxLYxLZ] [launches per kernel]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tmlqcd_b200 as tm
from conftest import random_gauge, random_spinor
dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "48x24x24x24").split("x"))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(3)
d = tm.Device(*dims)
d.set_params(0.16, 0.0032); d.ck(d.lib.tmb_set_nd(0.139, 0.15, 1.0))
d.gauge_upload(random_gauge(rng, d.V))
src = [random_spinor(rng, d.Vh) for _ in range(2)]
f = [d.field(s) for s in src] + [d.field() for _ in range(2)]
f32 = [d.field32(s.astype(np.float32)) for s in src] + [d.field32() for _ in range(2)]
for _ in range(n):   # 1. Hopping_Matrix, double, 18-real links
    d.lib.tmb_Hopping_Matrix(0, f[2], f[0]); d.lib.tmb_Hopping_Matrix(1, f[3], f[2])
for _ in range(n):   # 2. Hopping_Matrix_32
    d.lib.tmb_Hopping_Matrix_32(0, f32[2], f32[0]); d.lib.tmb_Hopping_Matrix_32(1, f32[3], f32[2])
d.ck(d.lib.tmb_set_compression(12))
for _ in range(n):   # 3. 12-real links, double
    d.lib.tmb_Hopping_Matrix(0, f[2], f[0]); d.lib.tmb_Hopping_Matrix(1, f[3], f[2])
d.ck(d.lib.tmb_set_compression(18))
for variant in (0, 2):  # 4. two-flavour hop: hop2_kernel (variant 0), hop_kernel<.., NFL = 2> (variant 2)
    d.ck(d.lib.tmb_set_hop2_variant(variant))
    for _ in range(n):
        d.lib.tmb_Qtm_pm_ndpsi(f[2], f[3], f[0], f[1])
d.ck(d.lib.tmb_set_hop2_variant(-1))
for _ in range(n):   # 5. float two-flavour hop
    d.lib.tmb_Qtm_pm_ndpsi_32(f32[2], f32[3], f32[0], f32[1])
d.call("derivative_zero")
for _ in range(n):   # 6. fermion force
    d.lib.tmb_deriv_Sb(0, f[0], f[1], 1.0); d.lib.tmb_deriv_Sb(1, f[1], f[0], 1.0)
it = d.call("cg_her", f[2], f[0], 6, 1e-30, 1)  # 7. six CG iterations: fused hops + sweeps
d.ck(d.lib.tmb_sync())
print("ok launches", d.lib.tmb_launch_count(), "cg", it)
d.close()

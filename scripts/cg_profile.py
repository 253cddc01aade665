#!/usr/bin/env python
"""a short invert_eo solve at 24^3x48 for `ncu --metrics gpu__time_duration.sum` launch lists"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tmlqcd_b200 as tm
from conftest import random_gauge, random_spinor
dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "48x24x24x24").split("x"))
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 12
mixed = len(sys.argv) > 3 and sys.argv[3] == "mixed"
rng = np.random.default_rng(1)
d = tm.Device(*dims)
d.set_params(0.16, 0.0032)
d.gauge_upload(random_gauge(rng, d.V))
E, O = d.field(random_spinor(rng, d.Vh)), d.field(random_spinor(rng, d.Vh))
En, On = d.field(), d.field()
it = d.call("invert_eo_mixed" if mixed else "invert_eo", En, On, E, O, 1e-14, maxit, 1)
print("iterations", it, d.solver_stats())
d.close()

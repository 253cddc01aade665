#!/usr/bin/env python
"""Small lattices (configs[0] 8^4, configs[4] 16^3x32, 16^4): CG loop time per iteration with and without programmatic dependent
launch of the hopping kernels (tmb_set_overlap bit 0), graph replay on."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmlqcd_b200 as tm
from bench import numpy_gauge
for dims in ((8, 8, 8, 8), (16, 16, 16, 16), (32, 16, 16, 16), (32, 24, 24, 24)):
    rng = np.random.default_rng(1)
    d = tm.Device(*dims)
    d.set_params(0.16, 0.0008)
    d.gauge_upload(numpy_gauge(dims, 2))
    E, O = d.field(rng.normal(size=(d.Vh, 24))), d.field(rng.normal(size=(d.Vh, 24)))
    En, On = d.field(), d.field()
    row = {"lattice": list(dims)}
    for flags in (0, 1, 0, 1):
        d.ck(d.lib.tmb_set_overlap(flags))
        best = None
        for rep in range(4):
            d.call("field_zero", On)
            it = d.call("invert_eo", En, On, E, O, 1e-16, 5000, 1)
            _, rr, s = d.solver_stats()
            best = s if best is None else min(best, s)
        key = "pdl" if flags else "plain"
        v = 1e6 * best / it
        row[key] = round(min(v, row.get(key, 1e9)), 2); row["iterations"] = it
    print(json.dumps(row), flush=True)
    d.close()

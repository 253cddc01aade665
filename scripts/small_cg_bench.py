#!/usr/bin/env python
"""CG time-to-solution on small lattices (BASELINE configs[0] 8^4 and configs[4] 16^3x32 sizes),
with and without CUDA-graph replay of the iteration chunk (tmb_set_overlap bit 2 disables it)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tmlqcd_b200 as tm
from conftest import random_gauge, random_spinor
out = []
for dims in ((8, 8, 8, 8), (32, 16, 16, 16)):
    rng = np.random.default_rng(1)
    d = tm.Device(*dims)
    d.set_params(0.16, 0.0032)
    d.gauge_upload(random_gauge(rng, d.V))
    E, O = d.field(random_spinor(rng, d.Vh)), d.field(random_spinor(rng, d.Vh))
    En, On = d.field(), d.field()
    for solver in ("invert_eo", "invert_eo_mixed"):
        for flags, name in ((4, "launches"), (0, "graph")):
            d.ck(d.lib.tmb_set_overlap(flags))
            best = 1e9
            for rep in range(4):
                d.call("field_zero", On)
                d.ck(d.lib.tmb_sync())
                t0 = time.perf_counter()
                it = d.call(solver, En, On, E, O, 1e-14, 5000, 1)
                best = min(best, time.perf_counter() - t0)
            out.append({"lattice": dims, "solver": solver, "mode": name, "count": it, "seconds": best,
                        "us_per_iteration": 1e6 * best / max(it, 1)})
            print(out[-1], flush=True)
    d.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "small_cg.json"), "w"), indent=1)

#!/usr/bin/env python
"""A/B of where the CG finishes <p, A p>: the last CTA of the second hop (default) against a one-CTA kernel on the side stream
next to the third hop (tmb_set_overlap bit 6).  One GPU (also peer-mode loop-back), double and mixed precision."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmlqcd_b200 as tm
from bench import numpy_gauge

out = {}
for lat in (sys.argv[1:] or ["48x24x24x24", "32x16x16x16", "12x48x48x48"]):
    dims = tuple(int(x) for x in lat.split("x"))
    rng = np.random.default_rng(5)
    g = numpy_gauge(dims, 11)
    for loop in (0, 2):
        d = tm.Device(*dims)
        d.set_params(0.16, 0.002)
        if loop:
            d.ck(d.lib.tmb_comm_loopback(loop))
        d.gauge_upload(g)
        E, O = d.field(rng.normal(size=(d.Vh, 24))), d.field(rng.normal(size=(d.Vh, 24)))
        En, On = d.field(), d.field()
        r = {}
        for flags in (0, 64, 0, 64):
            d.ck(d.lib.tmb_set_overlap(flags))
            best = None
            for _ in range(3):
                d.call("field_zero", On)
                it = d.call("invert_eo", En, On, E, O, 1e-14, 5000, 1)
                _, rr, s = d.solver_stats()
                best = s if best is None else min(best, s)
            x = d.download(On)
            d.call("field_zero", On)
            import time
            d.ck(d.lib.tmb_sync()); t0 = time.perf_counter()
            itm = d.call("invert_eo_mixed", En, On, E, O, 1e-14, 5000, 1)
            sm = time.perf_counter() - t0
            key = "side" if flags == 64 else "in_kernel"
            prev = r.get(key)
            r[key] = {"iterations": it, "cg_loop_s": round(min(best, prev["cg_loop_s"]) if prev else best, 6), "final_rr": rr, "mixed_count": itm, "mixed_s": round(min(sm, prev["mixed_s"]) if prev else sm, 6)}
            r[key]["ms_per_iteration"] = round(1e3 * r[key]["cg_loop_s"] / it, 4)
            r.setdefault("x", {})[key] = x
        r["identical_solutions"] = bool(np.array_equal(r["x"]["side"], r["x"]["in_kernel"])); del r["x"]
        out[lat + ("/peer_loopback" if loop else "")] = r
        print(lat, "peer loop-back" if loop else "single", json.dumps(r), flush=True)
        d.close()
print(json.dumps(out))

"""Generates tests/golden/ref_plaquette_4x4x4x4.json: measure_plaquette (measure_gauge_action.c:46) of the UNMODIFIED
reference (oracle/_ref, one thread) on the gauge fields of the committed fixtures ref_io_4x4x4x4.npz (start_ranlux(1, 2024))
and ref_4x4x4x4.npz.  Run in the build container only:
    make -C oracle/ref_build && python tests/golden/make_golden_plaquette.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refclient import Reference  # noqa: E402


def main():
    r = Reference(4, 4, 4, 4, nthreads=1)
    r.set_params(0.16, 0.0032)
    out = {}
    for name, key in (("ref_io_4x4x4x4.npz", "gauge"), ("ref_4x4x4x4.npz", "gauge")):
        g = np.load(os.path.join(HERE, name))[key]
        r.set_gauge(np.ascontiguousarray(g))
        out[name] = {"measure_plaquette": r.lib.ref_measure_plaquette(), "hex": float(r.lib.ref_measure_plaquette()).hex()}
    json.dump(out, open(os.path.join(HERE, "ref_plaquette_4x4x4x4.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()

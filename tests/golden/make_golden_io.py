"""Generates tests/golden/conf_ref_4x4x4x4.0000 (ILDG gauge configuration, 64 bit) and
tests/golden/prop_ref_4x4x4x4.inverted (SciDAC propagator file, 32 bit) with the UNMODIFIED reference's io/ code
(io/gauge_write.c, io/spinor_write*.c, io/dml.c ... in oracle/_ref) over the stand-in LIME layer
(oracle/ref_build/stubs/lime_standin.c; c-lime itself is not available), plus ref_io_4x4x4x4.npz with the fields
that were written.  Run in the build container only:
    make -C oracle/ref_build && python tests/golden/make_golden_io.py
Gauge: start_ranlux(1, 2024); random_gauge_field.  Propagator: two random_spinor_field_eo draws.
"""
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
    g = r.random_gauge(2024)
    conf = os.path.join(HERE, "conf_ref_4x4x4x4.0000")
    assert r.lib.ref_write_gauge(conf.encode(), 64, 0.5872, 17) == 0
    e, o = r.random_spinor_eo(), r.random_spinor_eo()
    prop = os.path.join(HERE, "prop_ref_4x4x4x4.inverted")
    assert r.lib.ref_write_propagator(prop.encode(), e, o, 32, 1e-19, 123) == 0
    np.savez_compressed(os.path.join(HERE, "ref_io_4x4x4x4.npz"), gauge=g, even=e, odd=o)
    print("wrote", conf, os.path.getsize(conf), prop, os.path.getsize(prop))


if __name__ == "__main__":
    main()

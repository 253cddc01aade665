"""Generates tests/golden/ref_nd_count_64x32x32x32.json: iteration counts, final residuals and wall time of the UNMODIFIED
reference's invert_doublet_eo (invert_doublet_eo.c compiled as it is into oracle/_ref, half-spinor OpenMP build) at the full
BASELINE configs[3] size, 32^3 x 64, on the inputs scripts/bench_sections.py's `nd` section uses: reference RANLUX hot start
(seed 123456), numpy Gaussian sources (seed 5), kappa = 0.16, 2 kappa mu = 0.0032, 2KappaMubar = 0.139, 2KappaEpsbar = 0.15,
phmc_invmaxev = 1, eps_sq = 1e-14 relative: solver_flag CG (cg_her_nd) and RGMIXEDCG (rg_mixed_cg_her_nd, mcg_delta = 5e-5,
the default of operator.c:125).  Takes minutes of CPU time: run in the build container only:
    make -C oracle/ref_build && python tests/golden/make_golden_nd_count.py"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refclient import Reference  # noqa: E402

DIMS = (64, 32, 32, 32)
KAPPA, GMU, MUBAR, EPSBAR, EPS_SQ, MAXIT, DELTA = 0.16, 0.0032, 0.139, 0.15, 1e-14, 5000, 5.0e-5


def main():
    r = Reference(*DIMS, nthreads=os.cpu_count() or 1, halfspinor=True)
    r.set_params(KAPPA, GMU); r.set_nd_params(MUBAR, EPSBAR, 1.0)
    r.random_gauge(123456)
    assert r.lib.ref_init32() == 0
    r.lib.ref_update_gauge32()
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    r.lib.ref_invert_doublet_eo.restype = C.c_int
    r.lib.ref_invert_doublet_eo.argtypes = [dp] * 8 + [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double]
    rng = np.random.default_rng(5)
    src = [rng.normal(scale=np.sqrt(0.5), size=(r.Vh, 24)) for _ in range(4)]  # Even_s, Odd_s, Even_c, Odd_c
    out = {"dims_TxLXxLYxLZ": list(DIMS), "kappa": KAPPA, "gmu": GMU, "mubar": MUBAR, "epsbar": EPSBAR, "eps_sq": EPS_SQ,
           "rel_prec": 1, "mcg_delta": DELTA, "threads": r.nthreads,
           "inputs": "reference ranlux random_gauge_field(seed 123456); numpy default_rng(5) normal(scale sqrt(1/2)) x 4"}
    for name, flag in (("CG", 1), ("RGMIXEDCG", 14)):
        sol = [r.spinor() for _ in range(4)]
        t0 = time.perf_counter()
        it = r.lib.ref_invert_doublet_eo(sol[0], sol[1], sol[2], sol[3], src[0], src[1], src[2], src[3], EPS_SQ, MAXIT, 1, flag, DELTA)
        out[name] = {"iterations": int(it), "seconds": time.perf_counter() - t0,
                     "solution_norm_sq": [float(np.sum(s ** 2)) for s in sol]}
        print(name, out[name], flush=True)
    json.dump(out, open(os.path.join(HERE, "ref_nd_count_64x32x32x32.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

"""Link-time replacement: a tmLQCD-style main program plus UNMODIFIED reference
callers - invert_eo.c, invert_doublet_eo.c, solver/monomial_solve.c, monomial/{monomial,det_monomial,detratio_monomial}.c,
start.c, geometry_eo.c, init/ - compiled from /root/reference by oracle/ref_build/Makefile (target `linktime`) and linked
against the product library INSTEAD of the reference's operator/, linalg/, solver/cg_her*, deriv_Sb, chrono_guess objects.
The executable defines the reference's globals itself (INIT_GLOBALS + global.h): the test passes only if the library really
reads the executable's g_mu / ka0..3 / g_gauge_field / g_update_gauge_copy (symbol interposition) and if the function
pointers the reference's code hands down (`f == Qtm_pm_psi`, solver/monomial_solve.c:134) compare equal inside the library.
Expected values: results of the unmodified reference (tests/golden/ref_4x4x4x4.npz, ref_hmc_4x4x4x4.npz).
CPU: against the host stand-in of the device ABI (tests/stubdev); GPU: against libtmlqcd_b200.so."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_l2

REFDIR = os.path.join(ROOT, "oracle", "_ref")
CG, MIXEDCG, RGMIXEDCG = 1, 13, 14


def _build(which):
    exe = os.path.join(REFDIR, "linktime_" + which)
    if os.path.isdir("/root/reference"):
        if which == "stub":
            r = subprocess.run(["bash", os.path.join(ROOT, "tests", "stubdev", "build.sh")], capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr
        elif not os.path.exists(os.path.join(ROOT, "tmlqcd_b200", "lib", "libtmlqcd_b200.so")):
            pytest.skip("product library not built")
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle", "ref_build"), os.path.join("..", "_ref", "linktime_" + which)],
                           capture_output=True, text=True)
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    return exe


def _run(exe, tmp_path, tag, dims, par, gauge, srcs):
    inp, out = str(tmp_path / f"in_{tag}.bin"), str(tmp_path / f"out_{tag}.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("4i", *dims))
        f.write(np.asarray(par, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(gauge, dtype=np.float64).tobytes())
        for s in srcs:
            f.write(np.ascontiguousarray(s, dtype=np.float64).tobytes())
    r = subprocess.run([exe, inp, out], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "linktime: done" in r.stdout, (r.stdout + r.stderr)[-3000:]
    V = int(np.prod(dims)); Vh = V // 2
    buf = open(out, "rb").read()
    pos = 0

    def take_int():
        nonlocal pos
        v = struct.unpack_from("i", buf, pos)[0]; pos += 4
        return v

    def take_f64(n=None):
        nonlocal pos
        if n is None:
            v = struct.unpack_from("d", buf, pos)[0]; pos += 8
            return v
        a = np.frombuffer(buf, dtype=np.float64, count=n, offset=pos).copy(); pos += 8 * n
        return a
    res = {"invert_eo": (take_int(), take_f64(Vh * 24).reshape(Vh, 24), take_f64(Vh * 24).reshape(Vh, 24))}
    for name in ("invert_eo_rg", "invert_eo_no_eo"):
        res[name] = (take_int(), take_f64(Vh * 24).reshape(Vh, 24), take_f64(Vh * 24).reshape(Vh, 24))
    for name in ("doublet_cg", "doublet_rg"):
        res[name] = (take_int(), [take_f64(Vh * 24).reshape(Vh, 24) for _ in range(4)])
    for name in ("sd_cg", "sd_mixed", "sd_rg"):
        res[name] = (take_int(), take_f64(Vh * 24).reshape(Vh, 24))
    for id in range(2):
        m = {"energy0": take_f64(), "iter1": [], "df": []}
        for _ in range(3):
            m["iter1"].append(take_int()); m["df"].append(take_f64(V * 32).reshape(V, 4, 8))
        m["dH"], m["iter0"] = take_f64(), take_int()
        res[f"mnl{id}"] = m
    assert pos == len(buf)
    return res


def _check(exe, tmp_path, tol_solve, tol_force):
    base = np.load(os.path.join(ROOT, "tests", "golden", "ref_4x4x4x4.npz"))
    dims = [int(x) for x in base["dims"]]
    nd = [float(x) for x in base["nd"]]
    par = [float(base["kappa"]), float(base["gmu"]), *[float(t) for t in base["theta"]], *nd, 0.16, 0.032, 0.1]
    r = _run(exe, tmp_path, "ops", dims, par, base["gauge"], [base[n] for n in ("k", "p", "q", "w")])
    # (1) invert_eo.c unmodified on the library's operators and cg_her
    it, en, on = r["invert_eo"]
    assert abs(it - int(base["invert_iters"])) <= 1
    assert rel_l2(en, base["invert_en"]) <= tol_solve and rel_l2(on, base["invert_on"]) <= tol_solve
    # (1b) its RGMIXEDCG branch and its branch without even/odd preconditioning (cg_her on Q_pm_psi over VOLUME sites): the same system
    for name in ("invert_eo_rg", "invert_eo_no_eo"):
        it, en, on = r[name]
        assert it > 0, name
        assert rel_l2(en, base["invert_en"]) <= 1e-7 and rel_l2(on, base["invert_on"]) <= 1e-7, name
    # (2) invert_doublet_eo.c unmodified: CG -> cg_her_nd, RGMIXEDCG -> rg_mixed_cg_her_nd of the library
    it, sol = r["doublet_cg"]
    assert abs(it - int(base["invert_doublet_iters"])) <= 1
    for s, name in zip(sol, ("ens", "ons", "enc", "onc")):
        assert rel_l2(s, base["invert_doublet_" + name]) <= tol_solve, name
    it, sol = r["doublet_rg"]
    assert it > 0
    for s, name in zip(sol, ("ens", "ons", "enc", "onc")):
        assert rel_l2(s, base["invert_doublet_" + name]) <= 1e-7, name
    # (3) solver/monomial_solve.c unmodified: f == Qtm_pm_psi reaches the library's cg_her / mixed_cg_her / rg_mixed_cg_her
    it, x = r["sd_cg"]
    assert abs(it - int(base["cg_iters"])) <= 1 and rel_l2(x, base["cg_x"]) <= tol_solve
    for name in ("sd_mixed", "sd_rg"):
        it, x = r[name]
        assert it > 0 and rel_l2(x, base["cg_x"]) <= 1e-8, name
    # (4) det_monomial.c / detratio_monomial.c unmodified over the library's operators, chrono_guess, deriv_Sb
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_hmc_4x4x4x4.npz"))
    dims = [int(x) for x in gold["dims"]]
    par = [float(gold["kappa"]), float(gold["gmu"]), *[float(t) for t in gold["theta"]], 0., 0., 1., float(gold["kappa2"]), float(gold["gmu2"]), 0.1]
    z = np.zeros_like(gold["l"])
    r = _run(exe, tmp_path, "hmc", dims, par, gold["gauge"], [gold["l"], gold["k"], z + gold["l"], z + gold["k"]])
    for id, gid in ((0, 2), (1, 3)):  # DET and DETRATIO with a chronological history of 2: the fixture's monomials 2 and 3
        m = r[f"mnl{id}"]
        assert abs(m["energy0"] / float(gold[f"m{gid}_energy0"]) - 1) <= 1e-13
        for call in range(3):
            assert rel_l2(m["df"][call], gold[f"m{gid}_df{call}"]) <= tol_force, (gid, call)
            assert abs(m["iter1"][call] - int(gold[f"m{gid}_iter1_{call}"])) <= 1 + call
        assert abs(m["iter0"] - int(gold[f"m{gid}_iter0"])) <= 2
        assert abs(m["dH"] - float(gold[f"m{gid}_dH"])) <= 1e-7


def test_unmodified_reference_callers_link_against_the_host_layer(tmp_path):
    """CPU: the product's C host layer (tmb_dropin.c) under the unmodified callers, device ABI served by the stand-in"""
    _check(_build("stub"), tmp_path, 1e-10, 1e-9)


@pytest.mark.gpu
def test_unmodified_reference_callers_link_against_the_library(tmp_path):
    """GPU: the same executable recipe against libtmlqcd_b200.so"""
    _check(_build("b200"), tmp_path, 1e-8, 1e-8)

"""measure_plaquette (measure_gauge_action.c:46-106; printed by tmLQCD_read_gauge, wrapper/lib_wrapper.c:232-235).
CPU: the oracle restatement against values of the unmodified reference (golden + live), the product's device site
function (host emulation) against the oracle, including a two-slab T split with the first-slice link halo.
GPU: tmb_measure_plaquette / the drop-in symbol against the oracle, plain and through the T-split halo path."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import ROOT, random_gauge

GOLD = os.path.join(ROOT, "tests", "golden")


def test_oracle_against_reference_values(oracle_lib):
    gold = json.load(open(os.path.join(GOLD, "ref_plaquette_4x4x4x4.json")))
    o = oracle_lib.Oracle(4, 4, 4, 4)
    for name, rec in gold.items():
        g = np.ascontiguousarray(np.load(os.path.join(GOLD, name))["gauge"])
        o.set_gauge(g)
        assert o.measure_plaquette() == float.fromhex(rec["hex"]), name  # bit-exact: same Kahan sum, same order
    # unit links: every plaquette is the identity -> 6 V planes, tr/3 = 1
    unit = np.zeros((256, 4, 18)); unit[:, :, [0, 8, 16]] = 1.
    o.set_gauge(unit)
    assert o.measure_plaquette() == 6 * 256


def test_oracle_against_live_reference(oracle_lib, ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref not built here")
    from oracle.refclient import Reference
    dims = (4, 6, 4, 8)
    r = Reference(*dims, nthreads=3)
    g = r.random_gauge(99)
    o = oracle_lib.Oracle(*dims); o.set_gauge(g)
    a, b = r.lib.ref_measure_plaquette(), o.measure_plaquette()
    assert abs(a - b) <= 1e-14 * abs(a)  # OpenMP partial sums in the reference


@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (4, 6, 2, 8), (2, 2, 2, 2), (6, 2, 10, 4)])
def test_device_site_function_and_T_split(oracle_lib, dims):
    from emul_client import Emul
    rng = np.random.default_rng(3)
    o = oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g)
    ref = o.measure_plaquette()
    e = Emul(*dims)
    U = e.pack_gauge(g)
    assert abs(e.plaquette(U) - ref) <= 1e-13 * abs(ref)
    # periodic single rank through the halo path: the "rank above" is this rank
    assert abs(e.plaquette(U, e.pack_gauge_first_slice(U)) - ref) <= 1e-13 * abs(ref)
    # two ranks: the global lattice (2T, LX, LY, LZ) cut into two slabs; t is the slowest lexicographic index
    T, LX, LY, LZ = dims
    og = oracle_lib.Oracle(2 * T, LX, LY, LZ)
    gg = random_gauge(rng, og.V)
    og.set_gauge(gg)
    slabs = [np.ascontiguousarray(gg[r * o.V:(r + 1) * o.V]) for r in range(2)]
    Us = [e.pack_gauge(s) for s in slabs]
    tot = sum(e.plaquette(Us[r], e.pack_gauge_first_slice(Us[(r + 1) % 2])) for r in range(2))
    assert abs(tot - og.measure_plaquette()) <= 1e-13 * abs(tot)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 6, 8), (2, 6, 4, 4), (16, 8, 8, 8)])
@pytest.mark.parametrize("loopback", [0, 1])
def test_gpu_plaquette(oracle_lib, dims, loopback):
    import tmlqcd_b200 as tm
    rng = np.random.default_rng(11)
    o = oracle_lib.Oracle(*dims)
    g = random_gauge(rng, o.V)
    o.set_gauge(g)
    ref = o.measure_plaquette()
    d = tm.Device(*dims)
    try:
        if loopback:
            d.ck(d.lib.tmb_comm_loopback(loopback))
        d.gauge_upload(g)
        res = C.c_double(0.)
        d.ck(d.lib.tmb_measure_plaquette(C.byref(res)))
        assert abs(res.value - ref) <= 1e-13 * abs(ref)
        unit = np.zeros((o.V, 4, 18)); unit[:, :, [0, 8, 16]] = 1.
        d.gauge_upload(unit)
        d.ck(d.lib.tmb_measure_plaquette(C.byref(res)))
        assert res.value == 6 * o.V
    finally:
        d.close()


@pytest.mark.gpu
def test_gpu_dropin_symbol(oracle_lib):
    import tmlqcd_b200 as tm
    dims = (4, 4, 4, 4)
    gold = json.load(open(os.path.join(GOLD, "ref_plaquette_4x4x4x4.json")))["ref_io_4x4x4x4.npz"]["measure_plaquette"]
    D = tm.DropIn(*dims)
    try:
        D.set_params(0.16, 0.0032)
        D.set_gauge(np.ascontiguousarray(np.load(os.path.join(GOLD, "ref_io_4x4x4x4.npz"))["gauge"]))
        gf = C.c_void_p.in_dll(D.lib, "g_gauge_field")
        val = D.lib.measure_plaquette(gf)
        assert abs(val - gold) <= 1e-13 * abs(gold)
    finally:
        D.close()

"""The product's HOST layer on the CPU: tmlqcd_b200/csrc/tmb_dropin.c (the reference-named symbols: argument order,
scratch use, VOLUME/2 | VOLUME splitting, the globals it pushes down, the tmLQCD.h facade with its lexicographic
conversion and 2 kappa normalisation, error returns) linked against a host stand-in for the device-level C ABI
(tests/stubdev/tmb_stub.c, which hands every operator to the CPU oracle).  The test bodies are the GPU tests'
own - run here on the stand-in, on a B200 on the CUDA library - and the expected values are those of the unmodified
reference (tests/golden) or of the oracle.  TEST INFRASTRUCTURE: the stand-in is never loaded by the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_l2

STUB = os.path.join(ROOT, "tests", "stubdev", "libtmb_dropin_stub.so")


@pytest.fixture(scope="module")
def stub_lib():
    import tmlqcd_b200.capi as capi
    r = subprocess.run(["bash", os.path.join(ROOT, "tests", "stubdev", "build.sh")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lib = C.CDLL(STUB)
    for name, (res, args) in capi.DROPIN_API.items():
        f = getattr(lib, name)  # every reference-named symbol of the product's host layer is in the test library
        f.restype = res; f.argtypes = args
    for name in ("tmb_last_error",):
        getattr(lib, name).restype = C.c_char_p
    return lib


@pytest.fixture()
def on_stub(stub_lib, monkeypatch):
    """tm.DropIn / tm.load bound to the stand-in library for the duration of one test"""
    import tmlqcd_b200 as tm
    import tmlqcd_b200.capi as capi

    class StubDropIn(capi.DropIn):
        def __init__(self, T, LX, LY, LZ, device=0):
            self.lib = stub_lib
            self.dims = (T, LX, LY, LZ); self.V = T * LX * LY * LZ; self.Vh = self.V // 2
            assert stub_lib.tmb_dropin_init(T, LX, LY, LZ, device) == 0, stub_lib.tmb_last_error()

    monkeypatch.setattr(tm, "DropIn", StubDropIn)
    monkeypatch.setattr(tm, "load", lambda: stub_lib)
    return StubDropIn


def _gold(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


@pytest.fixture()
def dropin(on_stub):
    base = _gold("ref_4x4x4x4.npz")
    D = on_stub(*[int(x) for x in base["dims"]])
    D.set_params(float(base["kappa"]), float(base["gmu"]), base["theta"])
    D.set_nd_params(*base["nd"])
    D.set_gauge(base["gauge"])
    yield D, base
    D.close()


def test_operator_family_members_vs_reference(dropin):
    import test_gpu_dropin_ops as g
    g.test_operator_family_members_vs_reference(dropin)


def test_host_memory_helpers_and_precision_conversion(dropin):
    import test_gpu_dropin_ops as g
    g.test_host_memory_helpers_and_precision_conversion(dropin)


def test_solvers_with_reference_signatures(dropin):
    import test_gpu_dropin_ops as g
    g.test_solvers_with_reference_signatures(dropin)


def test_inversions_with_reference_signatures(dropin):
    import test_gpu_dropin_ops as g
    g.test_inversions_with_reference_signatures(dropin)


def test_invert_eo_remaining_branches(dropin):
    import test_gpu_dropin_ops as g
    g.test_invert_eo_remaining_branches(dropin)


def test_invert_eo_branches_against_the_live_reference(dropin, ref_available):
    """the unmodified invert_eo.c (oracle/_ref, half-spinor build for the float operator) run here with the same flags:
    RGMIXEDCG with even/odd preconditioning, CG without it - solutions, and the iteration count of the branch whose
    arithmetic the stand-in shares (the full-lattice CG)"""
    from oracle import refclient
    import tmlqcd_b200 as tm
    if not ref_available or not refclient.available(halfspinor=True):
        pytest.skip("oracle/_ref not built")
    D, base = dropin
    k, p = np.array(base["k"]), np.array(base["p"])
    ref = refclient.Reference(*[int(x) for x in base["dims"]], nthreads=2, halfspinor=True)
    ref.set_gauge(base["gauge"]); ref.set_params(float(base["kappa"]), float(base["gmu"]), base["theta"])
    assert ref.init32() == 0
    ref.update_gauge32()
    sp = tm.capi.SolverParams(); sp.mcg_delta = 5e-5
    for solver, eo, prec in ((14, 1, 1e-20), (1, 0, 1e-24)):
        enr, onr = ref.spinor(), ref.spinor()
        itr = ref.invert_eo_flags(enr, onr, k.copy(), p.copy(), prec, 3000, 1, solver, eo, 5e-5)
        en, on = D.spinor(), D.spinor()
        it = D.invert_eo(en, on, k, p, prec, 3000, solver, 1, 0, eo, 0, None, sp, 0, 0, 0, 18)
        assert itr > 0 and it > 0
        assert rel_l2(en, enr) <= 1e-7 and rel_l2(on, onr) <= 1e-7, (solver, eo)
        if not eo:
            assert abs(it - itr) <= 1, (it, itr)


def test_reference_symbols_globals_and_dirty_flag(on_stub, oracle_lib):
    import test_gpu_parity as g
    g.test_dropin_reference_symbols(oracle_lib)


def test_tmLQCD_facade(on_stub, oracle_lib, tmp_path, monkeypatch):
    import test_gpu_parity as g
    monkeypatch.chdir(tmp_path)  # tmLQCD_read_gauge(0) looks for ./conf.0000
    g.test_tmLQCD_facade(oracle_lib)


def test_tmLQCD_facade_solver_keys(on_stub, oracle_lib, tmp_path, monkeypatch):
    import test_gpu_parity as g
    g.test_tmLQCD_facade_solver_keys(oracle_lib, tmp_path, monkeypatch)


def test_fermion_force_accumulates_into_the_callers_array(on_stub):
    gold = _gold("ref_hmc_4x4x4x4.npz")
    D = on_stub(*[int(x) for x in gold["dims"]])
    try:
        D.set_params(float(gold["kappa"]), float(gold["gmu"]), gold["theta"])
        D.set_gauge(gold["gauge"])
        l, k = np.array(gold["l"]), np.array(gold["k"])
        for ieo in (0, 1):
            df = np.zeros((D.V, 4, 8))
            hf = D.hamiltonian_field(df)
            D.deriv_Sb(ieo, l, k, C.byref(hf), 0.7)
            assert rel_l2(df, gold[f"deriv_Sb{ieo}"]) <= 1e-13
            D.deriv_Sb(ieo, l, k, C.byref(hf), -0.7)
            assert np.abs(df).max() <= 1e-12
    finally:
        D.close()


def test_monomials_with_reference_signatures(on_stub):
    import test_gpu_dropin_ops as g
    g.test_hmc_monomials_with_reference_signatures()


def test_chrono_guess_with_reference_signatures(on_stub):
    import test_gpu_dropin_ops as g
    g.test_hmc_chrono_guess_with_reference_signatures()


def test_blas32_and_plaquette_symbols(on_stub):
    g32 = _gold("ref_blas32_4x4x4x4.npz")
    io = _gold("ref_io_4x4x4x4.npz")
    import json
    plaq = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_plaquette_4x4x4x4.json")))["ref_io_4x4x4x4.npz"]
    D = on_stub(4, 4, 4, 4)
    try:
        D.set_params(0.16, 0.0032)
        D.set_gauge(np.ascontiguousarray(io["gauge"]))
        val = D.lib.measure_plaquette(C.c_void_p.in_dll(D.lib, "g_gauge_field"))
        assert val == float.fromhex(plaq["hex"])
        for nparts in (1, 2):  # VOLUME/2 and VOLUME sites
            n = 128 * nparts
            rep = lambda a: np.ascontiguousarray(np.concatenate([a] * nparts))
            r, s, s2 = rep(g32["r"]), rep(g32["s"]), rep(g32["s2"])
            c1, c2 = float(g32["c1"]), float(g32["c2"])
            x = r.copy(); D.assign_add_mul_r_32(x, s, c1, n); assert np.array_equal(x, rep(g32["assign_add_mul_r_32"]))
            x = r.copy(); D.assign_mul_add_r_32(x, c1, s, n); assert np.array_equal(x, rep(g32["assign_mul_add_r_32"]))
            x = np.zeros_like(r); D.diff_32(x, s, s2, n); assert np.array_equal(x, rep(g32["diff_32"]))
            x = np.zeros_like(r); D.mul_r_32(x, c1, s, n); assert np.array_equal(x, rep(g32["mul_r_32"]))
            x = r.copy(); D.assign_mul_add_mul_r_32(x, s, c1, c2, n); assert np.array_equal(x, rep(g32["assign_mul_add_mul_r_32"]))
            x = np.zeros_like(r); D.gamma5_32(x, s, n); assert np.array_equal(x, rep(g32["gamma5_32"]))
            assert abs(D.square_norm_32(r, n, 0) - nparts * float(g32["square_norm_32"])) <= 1e-6 * nparts * float(g32["square_norm_32"])
    finally:
        D.close()


def test_fatal_conditions_terminate_like_the_reference(stub_lib, tmp_path):
    """operators are void and exit(1) with a message (D_psi_body.c:267-272 style); checked in a child process"""
    code = f"""
import ctypes as C, numpy as np
lib = C.CDLL({STUB!r})
a = np.zeros((128, 24))
lib.Hopping_Matrix(0, a.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p))
"""
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert r.returncode == 1 and "tmb_dropin_init has not been called" in r.stderr
    code = f"""
import ctypes as C, numpy as np
lib = C.CDLL({STUB!r})
assert lib.tmb_dropin_init(4, 4, 4, 4, 0) == 0
a = np.zeros((128, 24))
lib.square_norm.restype = C.c_double
lib.square_norm(a.ctypes.data_as(C.c_void_p), 100, 0)
"""
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert r.returncode == 1 and "neither VOLUME/2 nor VOLUME" in r.stderr


@pytest.mark.parametrize("name,args", [("benchmark_b200", ["4", "4", "4", "4"]), ("invert_b200", [])])
def test_c_host_programs_run_on_the_stand_in(stub_lib, tmp_path, name, args):
    """examples/*.c (plain C on the two public headers) linked against the stand-in library instead of the CUDA
    library: their own consistency checks - host-pointer vs device-level Hopping_Matrix, D_psi vs M_full, the invert_eo
    residual, the ILDG round trip through tmLQCD_read_gauge, D_psi(prop)/(2 kappa) = source, the propagator file -
    exercise the C ABI from C and the whole host layer end to end"""
    exe = tmp_path / name
    d = os.path.dirname(STUB)
    cmd = ["gcc", "-std=gnu99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", name + ".c"),
           "-o", str(exe), "-L", d, "-ltmb_dropin_stub", f"-Wl,-rpath,{d}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)] + args, capture_output=True, text=True, cwd=str(tmp_path), timeout=900)
    assert r.returncode == 0 and "# all checks passed" in r.stdout, r.stdout[-2500:] + r.stderr[-1500:]
    if name == "invert_b200":
        assert "# The computed plaquette value is" in r.stdout and os.path.exists(tmp_path / "prop_b200.0000.00.00.inverted")


@pytest.mark.parametrize("name,args", [("invert_b200", []), ("benchmark_b200", ["4", "4", "4", "4"])])
def test_host_layer_under_address_and_ub_sanitizers(tmp_path, name, args):
    """the same programs with the whole C host layer (tmb_dropin.c, tmb_io.c) compiled in, under ASan + UBSan"""
    exe = tmp_path / (name + "_san")
    src = [os.path.join(ROOT, "examples", name + ".c"), os.path.join(ROOT, "tmlqcd_b200", "csrc", "tmb_dropin.c"),
           os.path.join(ROOT, "tmlqcd_b200", "csrc", "tmb_io.c"), os.path.join(ROOT, "tests", "stubdev", "tmb_stub.c"),
           os.path.join(ROOT, "oracle", "tmoracle.c")]
    cmd = ["gcc", "-std=gnu99", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fno-strict-aliasing",
           "-ffp-contract=off", "-I", os.path.join(ROOT, "include")] + src + ["-o", str(exe), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("no sanitizer runtime with this gcc")
    assert r.returncode == 0, r.stderr[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    env.pop("LD_PRELOAD", None)
    r = subprocess.run([str(exe)] + args, capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=900)
    assert r.returncode == 0 and "# all checks passed" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]

"""ILDG gauge configurations and SciDAC propagator files (tmlqcd_b200/csrc/tmb_io.c, host-only code of the product
library) against files written by the UNMODIFIED reference's io/ code (tests/golden/conf_ref_4x4x4x4.0000,
prop_ref_4x4x4x4.inverted; generator make_golden_io.py) and - where oracle/_ref exists - against the live
reference in both directions.  The LIME container layer is a restatement of the published format on both sides
(c-lime is not available): what is pinned here is everything inside the records.  CPU only: no GPU call is made."""
import ctypes as C
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
DIMS = (4, 4, 4, 4)


def lime_records(path):
    """independent little parser of the container: [(type, MB, ME, payload bytes)]"""
    out, data = [], open(path, "rb").read()
    pos = 0
    while pos < len(data):
        magic, ver, flags, n = struct.unpack(">IHHQ", data[pos:pos + 16])
        assert magic == 0x456789ab and ver == 1
        typ = data[pos + 16:pos + 144].split(b"\0")[0].decode()
        out.append((typ, flags >> 15 & 1, flags >> 14 & 1, data[pos + 144:pos + 144 + n]))
        pos += 144 + n + (8 - n % 8) % 8
    assert pos == len(data)
    return out


@pytest.fixture()
def io(tmp_path):
    import tmlqcd_b200 as tm
    lib = tm.load()
    T, LX, LY, LZ = DIMS
    V = T * LX * LY * LZ
    for n, v in (("T", T), ("L", LX), ("LX", LX), ("LY", LY), ("LZ", LZ), ("VOLUME", V), ("VOLUMEPLUSRAND", V), ("RAND", 0),
                 ("g_debug_level", 0), ("gauge_precision_read_flag", 64), ("g_disable_IO_checks", 0)):
        C.c_int.in_dll(lib, n).value = v
    C.c_double.in_dll(lib, "g_kappa").value = 0.16
    C.c_double.in_dll(lib, "g_mu").value = 0.0032
    gauge = np.zeros((V, 4, 18))
    rows = (C.c_void_p * V)(*[gauge.ctypes.data + ix * 4 * 144 for ix in range(V)])
    gf = C.POINTER(C.c_void_p).in_dll(lib, "g_gauge_field")
    old = C.cast(gf, C.c_void_p).value
    C.c_void_p.in_dll(lib, "g_gauge_field").value = C.addressof(rows)
    yield lib, gauge, rows, tmp_path
    C.c_void_p.in_dll(lib, "g_gauge_field").value = old


def test_read_reference_gauge_file(io):
    lib, gauge, rows, tmp = io
    ref = np.load(os.path.join(GOLD, "ref_io_4x4x4x4.npz"))
    conf = os.path.join(GOLD, "conf_ref_4x4x4x4.0000").encode()
    assert lib.read_gauge_field(conf, rows) == 0
    assert np.array_equal(gauge, ref["gauge"])
    assert C.c_int.in_dll(lib, "g_update_gauge_copy").value == 1
    recs = {t: p for t, _, _, p in lime_records(conf.decode())}
    assert f"{C.c_uint.in_dll(lib, 'GaugeInfo').value}" is not None  # GaugeInfo exported
    # a 32-bit read of a 64-bit file is refused like the reference does (gauge_read_binary.c:145-151)
    C.c_int.in_dll(lib, "gauge_precision_read_flag").value = 32
    assert lib.read_gauge_field(conf, rows) == -1
    C.c_int.in_dll(lib, "gauge_precision_read_flag").value = 64
    # corrupted payload -> checksum mismatch; truncated file; missing file
    raw = bytearray(open(conf, "rb").read())
    off = raw.find(b"ildg-binary-data") - 16 + 144 + 1000
    raw[off] ^= 0x40
    bad = str(tmp / "conf.bad"); open(bad, "wb").write(raw)
    assert lib.read_gauge_field(bad.encode(), rows) == -1
    open(bad, "wb").write(open(conf, "rb").read()[:60000])
    assert lib.read_gauge_field(bad.encode(), rows) == -1
    assert lib.read_gauge_field(b"/nonexistent/conf.0000", rows) == -1
    assert "suma" in recs["scidac-checksum"].decode()


@pytest.mark.parametrize("prec", [64, 32])
def test_written_gauge_file_matches_the_reference_record_for_record(io, prec):
    lib, gauge, rows, tmp = io
    ref = np.load(os.path.join(GOLD, "ref_io_4x4x4x4.npz"))
    gauge[:] = ref["gauge"]
    fn = str(tmp / f"conf.{prec}")
    lib.construct_paramsXlfInfo.restype = C.c_void_p
    xlf = lib.construct_paramsXlfInfo(0.5872, 17)
    assert lib.write_gauge_field(fn.encode(), prec, C.c_void_p(xlf)) == 0
    mine = lime_records(fn)
    assert [t for t, _, _, _ in mine] == ["xlf-info", "ildg-format", "ildg-binary-data", "scidac-checksum"]
    assert [(mb, me) for _, mb, me, _ in mine] == [(1, 1), (1, 0), (0, 0), (0, 1)]  # gauge_write.c:34-47
    if prec == 64:  # byte-identical to what the reference wrote, except the date/version text of xlf-info
        theirs = lime_records(os.path.join(GOLD, "conf_ref_4x4x4x4.0000"))
        assert [(t, mb, me) for t, mb, me, _ in mine] == [(t, mb, me) for t, mb, me, _ in theirs]
        for (t, _, _, a), (_, _, _, b) in zip(mine[1:], theirs[1:]):
            assert a == b, t
        assert mine[0][3].split(b"\n")[:3] == theirs[0][3].split(b"\n")[:3]  # plaquette, trajectory nr, beta/kappa/mu line
    # and back
    gauge[:] = 0
    C.c_int.in_dll(lib, "gauge_precision_read_flag").value = prec
    assert lib.read_gauge_field(fn.encode(), rows) == 0
    C.c_int.in_dll(lib, "gauge_precision_read_flag").value = 64
    exp = ref["gauge"] if prec == 64 else ref["gauge"].astype(np.float32).astype(np.float64)
    assert np.array_equal(gauge, exp)


def test_propagator_files(io):
    lib, gauge, rows, tmp = io
    ref = np.load(os.path.join(GOLD, "ref_io_4x4x4x4.npz"))
    Vh = int(np.prod(DIMS)) // 2
    e, o = np.zeros((Vh, 24)), np.zeros((Vh, 24))
    prop = os.path.join(GOLD, "prop_ref_4x4x4x4.inverted")
    assert lib.read_spinor(e, o, prop.encode(), 0) == 0
    assert np.array_equal(e, ref["even"].astype(np.float32).astype(np.float64))
    assert np.array_equal(o, ref["odd"].astype(np.float32).astype(np.float64))
    assert lib.read_spinor(e, o, prop.encode(), 1) == -5  # no second scidac-binary-data record (spinor_read.c:75-78)
    fn = str(tmp / "prop.out")

    class GaugeInfo(C.Structure):  # io/params.h:98-104
        _fields_ = [("plaquetteEnergy", C.c_double), ("gaugeRead", C.c_int), ("suma", C.c_uint), ("sumb", C.c_uint),
                    ("xlfInfo", C.c_void_p), ("ildg_data_lfn", C.c_void_p)]
    gi = GaugeInfo.in_dll(lib, "GaugeInfo")  # the state of the generating process: no gauge file had been read
    gi.xlfInfo = None; gi.ildg_data_lfn = None; gi.suma = 0; gi.sumb = 0
    assert lib.tmb_write_propagator(fn.encode(), np.ascontiguousarray(ref["even"]), np.ascontiguousarray(ref["odd"]), 32, 1e-19, 123, b"CG", 0) == 0
    mine, theirs = lime_records(fn), lime_records(prop)
    assert [(t, mb, me) for t, mb, me, _ in mine] == [(t, mb, me) for t, mb, me, _ in theirs]
    for (t, _, _, a), (_, _, _, b) in zip(mine, theirs):
        if t != "inverter-info":  # carries a date and the package version
            assert a == b, t
    assert lib.tmb_write_propagator(fn.encode(), np.ascontiguousarray(ref["even"]), np.ascontiguousarray(ref["odd"]), 64, 1e-19, 123, b"CG", 0) == 0
    assert lib.read_spinor(e, o, fn.encode(), 0) == 0
    assert np.array_equal(e, ref["even"]) and np.array_equal(o, ref["odd"])


def test_both_directions_against_the_live_reference(ref_available, tmp_path):
    if not ref_available:
        pytest.skip("oracle/_ref not built here (needs /root/reference); the golden files cover this box")
    code = f"""
import sys, ctypes as C, numpy as np
sys.path.insert(0, {ROOT!r})
from oracle.refclient import Reference
import tmlqcd_b200 as tm
dims=(6,4,4,4); T,LX,LY,LZ=dims; V=T*LX*LY*LZ  # cubic space: the reference sizes the record with L^3 (gauge_write.c:31)
r=Reference(*dims, nthreads=1); r.set_params(0.16,0.0032)
lib=tm.load()
for n,v in (("T",T),("L",LX),("LX",LX),("LY",LY),("LZ",LZ),("VOLUME",V),("VOLUMEPLUSRAND",V),("RAND",0)):
    C.c_int.in_dll(lib,n).value=v
C.c_double.in_dll(lib,"g_kappa").value=0.16
gauge=np.zeros((V,4,18)); rows=(C.c_void_p*V)(*[gauge.ctypes.data+ix*576 for ix in range(V)])
C.c_void_p.in_dll(lib,"g_gauge_field").value=C.addressof(rows)
g=r.random_gauge(31)
d={str(tmp_path)!r}
for prec in (64,32):
    exp = g if prec==64 else g.astype(np.float32).astype(np.float64)
    assert r.lib.ref_write_gauge((d+"/a").encode(),prec,0.6,3)==0
    C.c_int.in_dll(lib,"gauge_precision_read_flag").value=prec
    gauge[:]=0; assert lib.read_gauge_field((d+"/a").encode(),rows)==0 and np.array_equal(gauge,exp)
    gauge[:]=g
    lib.construct_paramsXlfInfo.restype=C.c_void_p
    assert lib.write_gauge_field((d+"/b").encode(),prec,C.c_void_p(lib.construct_paramsXlfInfo(C.c_double(0.6),3)))==0
    r.set_gauge(np.zeros_like(g)); assert r.lib.ref_read_gauge((d+"/b").encode(),prec)==0 and np.array_equal(r.get_gauge(),exp)
r.set_gauge(g)
e,o=r.random_spinor_eo(),r.random_spinor_eo()
for prec in (64,32):
    f=lambda a: a if prec==64 else a.astype(np.float32).astype(np.float64)
    assert r.lib.ref_write_propagator((d+"/p").encode(),e,o,prec,1e-20,5)==0
    a,b=np.zeros_like(e),np.zeros_like(o); assert lib.read_spinor(a,b,(d+"/p").encode(),0)==0 and np.array_equal(a,f(e)) and np.array_equal(b,f(o))
    assert lib.tmb_write_propagator((d+"/q").encode(),e,o,prec,1e-20,5,b"CG",0)==0
    a[:]=0; b[:]=0; assert r.lib.ref_read_spinor(a,b,(d+"/q").encode(),0)==0 and np.array_equal(a,f(e)) and np.array_equal(b,f(o))
print("OK")
"""
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout[-2000:] + res.stderr[-3000:]


@pytest.mark.parametrize("dims", [(4, 2, 6, 4), (2, 4, 4, 8)])
def test_non_cubic_round_trips_and_appended_propagators(dims, tmp_path):
    """lattices with LX != LY != LZ (the reference sizes its records with L^3, gauge_write.c:31, so only this library is
    on both sides here): gauge round trip at both precisions, three propagators appended to one file and read back by
    position (spinor_read.c:60-78), a position past the end refused"""
    import tmlqcd_b200 as tm
    lib = tm.load()
    T, LX, LY, LZ = dims
    V = T * LX * LY * LZ
    for n, v in (("T", T), ("L", LX), ("LX", LX), ("LY", LY), ("LZ", LZ), ("VOLUME", V), ("VOLUMEPLUSRAND", V), ("RAND", 0),
                 ("g_debug_level", 0), ("g_disable_IO_checks", 0)):
        C.c_int.in_dll(lib, n).value = v
    rng = np.random.default_rng(4)
    gauge = rng.normal(size=(V, 4, 18))
    keep = gauge.copy()
    rows = (C.c_void_p * V)(*[gauge.ctypes.data + ix * 4 * 144 for ix in range(V)])
    old = C.c_void_p.in_dll(lib, "g_gauge_field").value
    C.c_void_p.in_dll(lib, "g_gauge_field").value = C.addressof(rows)
    try:
        lib.construct_paramsXlfInfo.restype = C.c_void_p
        xlf = C.c_void_p(lib.construct_paramsXlfInfo(C.c_double(0.6), 1))
        for prec in (64, 32):
            fn = str(tmp_path / f"conf.{prec}").encode()
            gauge[:] = keep
            assert lib.write_gauge_field(fn, prec, xlf) == 0
            gauge[:] = 0
            C.c_int.in_dll(lib, "gauge_precision_read_flag").value = prec
            assert lib.read_gauge_field(fn, rows) == 0
            assert np.array_equal(gauge, keep if prec == 64 else keep.astype(np.float32).astype(np.float64))
        C.c_int.in_dll(lib, "gauge_precision_read_flag").value = 64
        # a file of another lattice with the SAME volume is refused by the ildg-format record (the reference only prints the
        # mismatch, gauge_read.c:172-180; this reader is stricter) ...
        C.c_int.in_dll(lib, "LX").value = LZ; C.c_int.in_dll(lib, "LZ").value = LX  # LX != LZ in both cases
        assert lib.read_gauge_field(str(tmp_path / "conf.64").encode(), rows) != 0
        C.c_int.in_dll(lib, "LX").value = LX; C.c_int.in_dll(lib, "LZ").value = LZ
        # ... and inconsistent globals (VOLUME != T LX LY LZ) before anything is written into the field
        C.c_int.in_dll(lib, "LY").value = LY + 2
        gauge[:] = 7.
        assert lib.read_gauge_field(str(tmp_path / "conf.64").encode(), rows) != 0 and np.all(gauge == 7.)
        C.c_int.in_dll(lib, "LY").value = LY
        props = [(rng.normal(size=(V // 2, 24)), rng.normal(size=(V // 2, 24))) for _ in range(3)]
        fn = str(tmp_path / "props").encode()
        for k, (e, o) in enumerate(props):
            assert lib.tmb_write_propagator(fn, e, o, 64, 1e-12, 10 + k, b"CG", 1 if k else 0) == 0
        a, b = np.zeros((V // 2, 24)), np.zeros((V // 2, 24))
        for k in (2, 0, 1):
            assert lib.read_spinor(a, b, fn, k) == 0
            assert np.array_equal(a, props[k][0]) and np.array_equal(b, props[k][1])
        assert lib.read_spinor(a, b, fn, 3) == -5
    finally:
        C.c_void_p.in_dll(lib, "g_gauge_field").value = old


def test_damaged_files_under_address_and_ub_sanitizers(tmp_path):
    """tests/io_sanitize/io_fuzz.c: tmb_io.c compiled with -fsanitize=address,undefined, 400 truncated / bit-flipped /
    garbage-overwritten configuration and propagator files; every read must come back without a memory error"""
    exe = tmp_path / "io_fuzz"
    cmd = ["gcc", "-std=gnu99", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "io_sanitize", "io_fuzz.c"), os.path.join(ROOT, "tmlqcd_b200", "csrc", "tmb_io.c"), "-o", str(exe), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("no sanitizer runtime with this gcc")
    assert r.returncode == 0, r.stderr[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    env.pop("LD_PRELOAD", None)
    r = subprocess.run([str(exe), "400"], capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=600)
    assert r.returncode == 0 and "IOFUZZ accepted" in r.stderr, r.stderr[-3000:]
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr
    refused = int(r.stderr.split("refused")[-1].split()[0])
    assert refused > 100  # most damage is detected (flips in free-text records are legitimately accepted)

"""The C host programs under examples/ use nothing but include/*.h and the shared library: they must compile
with a C compiler from the public headers alone (CPU test), refuse to run without a GPU (no CPU path), and
pass their own consistency checks on a B200 (GPU test): host-pointer vs device-resident Hopping_Matrix,
D_psi vs M_full, invert_eo residual, ILDG round trip + tmLQCD_invert + SciDAC propagator file."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

EX = os.path.join(ROOT, "examples")


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    import tmlqcd_b200 as tm
    tm.build()
    out = tmp_path_factory.mktemp("examples")
    lib = os.path.join(ROOT, "tmlqcd_b200", "lib")
    for name in ("benchmark_b200", "invert_b200"):
        cmd = ["gcc", "-std=gnu99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
               os.path.join(EX, name + ".c"), "-o", str(out / name), "-L", lib, "-ltmlqcd_b200", f"-Wl,-rpath,{lib}", "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    return out


def test_examples_compile_from_public_headers_and_refuse_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    for name in ("benchmark_b200", "invert_b200"):
        r = subprocess.run([str(built / name)], capture_output=True, text=True, cwd=str(built))
        assert r.returncode == 2 and "no CPU path" in r.stderr, (name, r.returncode, r.stderr[-300:])


@pytest.mark.gpu
def test_benchmark_program(built):
    r = subprocess.run([str(built / "benchmark_b200"), "16", "8", "8", "8"], capture_output=True, text=True, cwd=str(built), timeout=600)
    assert r.returncode == 0 and "# all checks passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "differ by 0.000000e+00" in r.stdout
    m = re.search(r"invert_eo: (\d+) iterations", r.stdout)
    assert m and 0 < int(m.group(1)) < 5000


@pytest.mark.gpu
def test_invert_program_with_files(built):
    r = subprocess.run([str(built / "invert_b200")], capture_output=True, text=True, cwd=str(built), timeout=600)
    assert r.returncode == 0 and "# all checks passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    m = re.search(r"# The computed plaquette value is ([0-9.e+-]+)\.", r.stdout)  # lib_wrapper.c:232-235
    assert m and 0.0 < float(m.group(1)) < 0.5  # hot start: small average plaquette
    assert os.path.exists(built / "conf.0000") and os.path.exists(built / "prop_b200.0000.00.00.inverted")
    # the files are LIME containers with the reference's record sequence (io/gauge_write.c:34-47, operator.c:532-605)
    from test_io_formats import lime_records
    assert [t for t, _, _, _ in lime_records(str(built / "conf.0000"))] == ["xlf-info", "ildg-format", "ildg-binary-data", "scidac-checksum"]
    types = [t for t, _, _, _ in lime_records(str(built / "prop_b200.0000.00.00.inverted"))]
    assert types[0] == "propagator-type" and "scidac-binary-data" in types and types[-1] == "scidac-checksum"

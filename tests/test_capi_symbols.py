"""The C-ABI library loads on a GPU-less box, exports every symbol the headers declare, and refuses
to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import tmlqcd_b200 as tm
    tm.build()
    return tm.load()


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set()
    for m in re.finditer(r"^\s*(?:extern\s+)?(?:const\s+)?(?:int|void|double|long long|char)\s*\*?\s*(\w+)\s*\(", txt, flags=re.M):
        names.add(m.group(1))
    return names


def test_every_declared_function_is_exported(lib):
    import tmlqcd_b200.capi as capi
    dev, drop = _declared("tmlqcd_b200.h"), _declared("tmlqcd_b200_dropin.h")
    assert len(dev) >= 60 and len(drop) >= 55
    out = subprocess.run(["nm", "-D", "--defined-only", capi.lib_path()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert not (dev - exported), f"declared in tmlqcd_b200.h but not exported: {sorted(dev - exported)}"
    assert not (drop - exported), f"declared in tmlqcd_b200_dropin.h but not exported: {sorted(drop - exported)}"
    # ... and the ctypes tables bind all of them
    assert dev <= set(capi.DEVICE_API), sorted(dev - set(capi.DEVICE_API))
    assert drop <= set(capi.DROPIN_API), sorted(drop - set(capi.DROPIN_API))
    for g in capi.DROPIN_GLOBALS:  # the reference's globals the path reads
        C.c_int.in_dll(lib, g)


def test_no_cpu_path(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.tmb_is_initialized() == 0
    assert lib.tmb_init(4, 4, 4, 4, 0) < 0
    assert b"no CPU path" in lib.tmb_last_error()
    assert lib.tmb_Hopping_Matrix(0, None, None) < 0  # every operator refuses before tmb_init
    assert lib.tmb_field_alloc() is None


def test_product_does_not_reference_the_oracle():
    """the shipped sources never include, link or load anything under oracle/ or tests/"""
    for root, _, files in os.walk(os.path.join(ROOT, "tmlqcd_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".c", ".h", ".py")) or f == "Makefile":
                txt = open(os.path.join(root, f)).read()
                assert "tmoracle" not in txt and "oracle/" not in txt and "libtmb_emul" not in txt, os.path.join(root, f)
    out = subprocess.run(["ldd", os.path.join(ROOT, "tmlqcd_b200", "lib", "libtmlqcd_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "emul" not in out


def test_sass_is_sm100a_with_128bit_policy_loads():
    so = os.path.join(ROOT, "tmlqcd_b200", "lib", "libtmlqcd_b200.so")
    elf = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in elf

import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _make(target_dir, *args):
    r = subprocess.run(["make", "-C", os.path.join(ROOT, target_dir), *args], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.fixture(scope="session")
def oracle_lib():
    """the plain-C restatement (oracle/tmoracle.c) -> oracle/libtmoracle.so"""
    _make("oracle")
    from oracle import oracleclient
    return oracleclient


@pytest.fixture(scope="session")
def ref_available():
    """oracle/_ref/*.so: the unmodified reference compiled from /root/reference (built where it exists)."""
    from oracle import refclient
    if not refclient.available() and os.path.isdir("/root/reference"):
        _make("oracle/ref_build")
    return refclient.available()


def random_su3_field(rng, n):
    """n random SU(3) matrices (QR of complex Gaussians, det fixed to 1) as [n,18] float64"""
    a = rng.normal(size=(n, 3, 3)) + 1j * rng.normal(size=(n, 3, 3))
    q, r = np.linalg.qr(a)
    d = np.diagonal(r, axis1=1, axis2=2)
    q = q * (d / np.abs(d))[:, None, :]
    det = np.linalg.det(q)
    q = q / det[:, None, None] ** (1.0 / 3.0)
    return np.ascontiguousarray(q.reshape(n, 9)).view(np.float64).reshape(n, 18)


def random_gauge(rng, V):
    return random_su3_field(rng, V * 4).reshape(V, 4, 18)


def random_spinor(rng, n):
    return rng.normal(scale=np.sqrt(0.5), size=(n, 24))


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))

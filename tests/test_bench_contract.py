"""bench.py's reference arm runs on the host (no GPU) and prints ONE JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_arm_json_line(ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref not built here")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--lattice", "8x8x8x8"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True and d["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # both arms build `config` with the same function and the same description of the inputs: the driver compares them
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(1, (8, 8, 8, 8), bench.REFERENCE_INPUTS_HOW, nt=1, nz=1)
    assert d["config"]["rank_grid_TxZ"] == [1, 1] and d["config"]["global_lattice_TxLXxLYxLZ"] == [8, 8, 8, 8]


def test_b200_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--lattice", "4x4x4x4",
                        "--skip-cpu", "--skip-sections"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def test_strong_scaling_inputs_do_not_depend_on_the_decomposition(oracle_lib):
    """bench.py --global-chunk-t: N = 1, 2, 4 ranks assemble the same GLOBAL gauge field and sources from per-chunk seeds,
    and the fields are a valid input of the path (SU(3) links: unit plaquette normalisation through the oracle)."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    T, LX, LY, LZ, ct = 8, 4, 4, 4, 2
    g1, s1, how = bench.chunked_fields((T, LX, LY, LZ), 0, ct)
    assert "chunks of 2" in how and g1.shape == (T * LX * LY * LZ, 4, 18) and s1[0].shape == (T * LX * LY * LZ // 2, 24)
    for world in (2, 4):
        parts = [bench.chunked_fields((T // world, LX, LY, LZ), r, ct) for r in range(world)]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), g1)
        for k in range(3):
            assert np.array_equal(np.concatenate([p[1][k] for p in parts]), s1[k])
    u = g1.reshape(-1, 9, 2)
    m = (u[..., 0] + 1j * u[..., 1]).reshape(-1, 3, 3)
    assert np.allclose(m @ m.conj().transpose(0, 2, 1), np.eye(3), atol=1e-13) and np.allclose(np.linalg.det(m), 1., atol=1e-13)
    o = oracle_lib.Oracle(T, LX, LY, LZ)
    o.set_gauge(np.ascontiguousarray(g1))
    assert abs(o.measure_plaquette()) < 6 * T * LX * LY * LZ  # finite, below the unit-gauge value

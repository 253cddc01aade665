#!/bin/bash
# session 3, call 1: GPU tests after the I/O + lane-paired two-flavour kernel changes, then diagnostics
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_h.log 2>&1; tail -3 gpurun_out/pytest_gpu_h.log
echo "== nd variants"; timeout 300 python scripts/diag_r01c.py nd 2>&1 | tail -8
echo "== e2e invert_eo phases"; timeout 300 python scripts/diag_r01c.py e2e 2>&1 | tail -14
echo "== host chunks"; timeout 300 python scripts/diag_r01c.py chunk 2>&1 | tail -10

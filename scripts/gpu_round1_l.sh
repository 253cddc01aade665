#!/bin/bash
mkdir -p gpurun_out
echo "== pytest: CG paths"; timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cg or residency or mixed or invert or compression or dropin or facade" > gpurun_out/pytest_gpu_l.log 2>&1; tail -4 gpurun_out/pytest_gpu_l.log
echo "== selfnorm timing"; timeout 400 python scripts/diag_r01c.py selfnorm 2>&1 | tail -18

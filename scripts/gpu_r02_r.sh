#!/bin/bash
# round 2, call r (1 GPU): the CG's side-stream finish of <p, A p> - tests and A/B
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin_ops.py tests/test_gpu_hmc.py -x -q -m gpu -k "cg or solver or invert or mixed or hmc or monomial or full_size" > gpurun_out/r02r_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02r_pytest.log
echo "== A/B"; timeout 900 python scripts/cg_side_ab.py > gpurun_out/r02r_cg_side_ab.log 2>&1; echo "rc=$?"; grep -v "^{" gpurun_out/r02r_cg_side_ab.log | cut -c1-700

#!/bin/bash
# round 2, call h (2 GPUs): shifted-output host pipeline (tests, timing), then the N = 2 parity + bench of call c
mkdir -p gpurun_out
echo "== pytest host pipeline + nd"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host or pipelined or nd or dropin" > gpurun_out/r02h_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02h_pytest.log
echo "== e2e diag"; timeout 600 python scripts/e2e_diag.py > gpurun_out/r02h_e2e_diag.log 2>&1; echo "rc=$?"; grep -A1 "host link\|drop-in Hopping_Matrix (pinned)\|drop-in EO+OE" gpurun_out/r02h_e2e_diag.log | cut -c1-200
bash scripts/gpu_r02_c.sh 2

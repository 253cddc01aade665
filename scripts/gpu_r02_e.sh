#!/bin/bash
# round 2, call e (1 GPU): re-check after the auto-pin default and chunking changes; two-flavour kernel with gauge links through L1
mkdir -p gpurun_out
echo "== pytest -m gpu"; SECONDS=0; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "rc=$? wall=${SECONDS}s"; tail -6 gpurun_out/r02e_pytest_gpu.log
echo "== e2e diag"; timeout 600 python scripts/e2e_diag.py > gpurun_out/r02e_e2e_diag.log 2>&1; echo "rc=$?"; grep -A1 "host link\|drop-in" gpurun_out/r02e_e2e_diag.log | cut -c1-200
echo "== nd section, default"; timeout 600 python scripts/bench_sections.py nd 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:v for k,v in d.items() if k.startswith('Qtm_pm_ndpsi_us') or k in ('iterations','time_to_solution_s','rgmixed','iterations_match_reference','cpu_reference_counts')})"
echo "== nd section, TMB_ND_L1=1"; TMB_ND_L1=1 timeout 600 python scripts/bench_sections.py nd 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:v for k,v in d.items() if k.startswith('Qtm_pm_ndpsi_us')})"

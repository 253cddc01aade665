#!/bin/bash
# round 2, call g (1 GPU): zero-copy host pipeline against the copy-engine pipeline; float two-flavour kernels; GPU tests
mkdir -p gpurun_out
echo "== pytest -m gpu"; SECONDS=0; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "rc=$? wall=${SECONDS}s"; tail -4 gpurun_out/r02g_pytest_gpu.log
echo "== e2e diag, copy engines"; timeout 600 python scripts/e2e_diag.py > gpurun_out/r02g_e2e_diag_ce.log 2>&1; echo "rc=$?"; grep -A1 "host link\|drop-in Hopping_Matrix (pinned)\|drop-in EO+OE" gpurun_out/r02g_e2e_diag_ce.log | cut -c1-200
for ct in 16 32 64 128; do
echo "== e2e diag, zero copy, $ct CTAs"; TMB_E2E_ZEROCOPY=1 TMB_E2E_CTAS=$ct timeout 600 python scripts/e2e_diag.py > gpurun_out/r02g_e2e_diag_zc$ct.log 2>&1; echo "rc=$?"; grep -A1 "drop-in Hopping_Matrix (pinned)\|drop-in EO+OE" gpurun_out/r02g_e2e_diag_zc$ct.log | cut -c1-200
done
echo "== nd section"; timeout 600 python scripts/bench_sections.py nd 2>/dev/null > gpurun_out/r02g_section_nd.json; python -c "
import json,sys; d=json.loads(open('gpurun_out/r02g_section_nd.json').read().strip().splitlines()[-1]); print({k:v for k,v in d.items() if k.startswith('Qtm_pm_ndpsi') and 'bytes' not in k or k in ('iterations','time_to_solution_s','rgmixed','iterations_match_reference')})"

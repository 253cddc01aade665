#!/bin/bash
# round 2, call d (1 GPU): host-pointer pipeline (graph cache, auto-pin), link-time test, e2e diagnostics, bench sections
mkdir -p gpurun_out
echo "== pytest (new tests first)"; SECONDS=0
timeout 900 python -m pytest tests/test_linktime.py "tests/test_gpu_parity.py::test_host_pointer_hop_pipeline" tests/test_gpu_dropin_ops.py -x -q -m gpu > gpurun_out/r02d_pytest_new.log 2>&1; echo "rc=$? wall=${SECONDS}s"; tail -15 gpurun_out/r02d_pytest_new.log
echo "== e2e diag"; timeout 600 python scripts/e2e_diag.py > gpurun_out/r02d_e2e_diag.log 2>&1; echo "rc=$?"; cat gpurun_out/r02d_e2e_diag.log | cut -c1-220
echo "== sections"
for s in nd hmc; do timeout 600 python scripts/bench_sections.py $s > gpurun_out/r02d_section_$s.json 2> gpurun_out/r02d_section_$s.err; echo "$s rc=$?"; done
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_section_nd.json').read().strip().splitlines()[-1])
print({k:v for k,v in d.items() if k not in ('workload','Qtm_pm_ndpsi_bytes_how','gauge')})
d=json.loads(open('gpurun_out/r02d_section_hmc.json').read().strip().splitlines()[-1])
print({k:v for k,v in d.items() if k in ('total_s','speedup_vs_cpu_reference','same_noise_comparison','derivative_rel_l2_vs_cpu_reference','deriv_Sb')})
PY
echo "== sustained sweep"; timeout 600 python bench.py --steps 20 --warmup 5 --sweep-sustained --skip-cpu --skip-cg --skip-e2e --skip-sections --skip-anchor --skip-parity 2>&1 >/dev/null | grep "variant\|sustained"
echo "== full pytest -m gpu"; SECONDS=0; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "rc=$? wall=${SECONDS}s"; tail -4 gpurun_out/r02d_pytest_gpu.log

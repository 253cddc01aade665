#!/bin/bash
mkdir -p gpurun_out
echo "== pytest: solver paths"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hmc.py tests/test_gpu_dropin_ops.py -x -q -m gpu -k "cg or mixed or invert or nd or monomial or solve or facade or dropin or hmc or chrono" > gpurun_out/pytest_gpu_m.log 2>&1; tail -4 gpurun_out/pytest_gpu_m.log
echo "== restart cost"; timeout 300 python scripts/diag_r01c.py restart 2>&1 | tail -8
echo "== hmc section"; timeout 300 python scripts/bench_sections.py hmc 2>gpurun_out/hmc.err > gpurun_out/r01c_section_hmc.json; python -c "
import json; d=json.load(open('gpurun_out/r01c_section_hmc.json')); print('total', d['total_s'], 'speedup', d.get('speedup_vs_cpu_reference')); [print(m['type'], m['heatbath_s'], m['derivative_s'], m['acc_s'], m['iter0'], m['iter1']) for m in d['monomials']]"

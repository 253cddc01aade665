#!/bin/bash
mkdir -p gpurun_out
SECONDS=0; python bench.py > gpurun_out/r01c_bench_default.json 2> gpurun_out/r01c_bench_default.err; echo "rc=$? wall=${SECONDS}s"; tail -1 gpurun_out/r01c_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r01c_bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print('roofline', d['roofline']); print('e2e', d['e2e']); print('cg', d['cg'])
print('cpu', d.get('cpu_baseline')); print('small', d.get('benchmark_8x8x8x8'))
print('nd', {k:v for k,v in d.get('nd',{}).items() if k in ('iterations','time_to_solution_s','Qtm_pm_ndpsi_us','error')}); print('hmc', {k:v for k,v in d.get('hmc',{}).items() if k in ('total_s','speedup_vs_cpu_reference','error')})
PY

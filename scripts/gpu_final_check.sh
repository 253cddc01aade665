#!/bin/bash
# what the driver runs at round end, plus the launch list of the bench command
mkdir -p gpurun_out
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== bench default"; /usr/bin/time -f "%e s wall" python bench.py > gpurun_out/r01b_bench_default.json 2> gpurun_out/r01b_bench_default.err; echo "rc=$?"; tail -1 gpurun_out/r01b_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01b_bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print('roofline', d['roofline']); print('e2e', d['e2e']); print('cg', d['cg'])
print('cpu', d.get('cpu_baseline')); print('nd', {k:v for k,v in d.get('nd',{}).items() if k in ('iterations','time_to_solution_s','Qtm_pm_ndpsi_us','error')}); print('hmc', {k:v for k,v in d.get('hmc',{}).items() if k in ('total_s','speedup_vs_cpu_reference','error')})
PY
echo "== bench reference arm"; python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/r01b_bench_reference.json 2>/dev/null; cat gpurun_out/r01b_bench_reference.json | cut -c1-400
echo "== ncu launch list"
CMD="python bench.py --steps 20 --warmup 3 --skip-cpu --skip-sections"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01b_launches_bench_full.csv $CMD > gpurun_out/r01b_ncu_list2.log 2>&1; echo "ncu rc=$?"

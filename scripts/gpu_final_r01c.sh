#!/bin/bash
# what the driver runs at round end, plus the launch list of the bench command and one ncu --set full of the hop kernel
mkdir -p gpurun_out
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final_r01c.log 2>&1; tail -3 gpurun_out/pytest_gpu_final_r01c.log
echo "== bench default"; SECONDS=0; python bench.py > gpurun_out/r01c_bench_default.json 2> gpurun_out/r01c_bench_default.err; echo "rc=$? wall=${SECONDS}s"; tail -1 gpurun_out/r01c_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r01c_bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print('roofline', d['roofline']); print('e2e', d['e2e']); print('cg', d['cg'])
print('cpu', d.get('cpu_baseline')); print('small', d.get('benchmark_8x8x8x8'))
print('nd', {k:v for k,v in d.get('nd',{}).items() if k in ('iterations','time_to_solution_s','Qtm_pm_ndpsi_us','error')}); print('hmc', {k:v for k,v in d.get('hmc',{}).items() if k in ('total_s','speedup_vs_cpu_reference','error')})
PY
echo "== bench reference arm"; python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/r01c_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r01c_bench_reference.json
echo "== ncu launch list"
CMD="python bench.py --steps 20 --warmup 3 --skip-cpu --skip-sections"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01c_launches_bench.csv $CMD > gpurun_out/r01c_ncu_list.log 2>&1; echo "ncu rc=$?"
CMD2="python bench.py --steps 10 --warmup 3 --skip-cpu --skip-cg --skip-e2e --skip-sections"
$CMD2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hop_kernel -s 8 -c 3 -f -o gpurun_out/r01c_hop $CMD2 > gpurun_out/r01c_ncu_hop.log 2>&1
echo "ncu hop rc=$?"; ls -la gpurun_out/r01c_hop.ncu-rep

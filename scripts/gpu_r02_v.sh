#!/bin/bash
# round 2, call v (1 GPU): the final tree - full GPU test suite, smoke(), a short bench line
mkdir -p gpurun_out
echo "== pytest -m gpu"; SECONDS=0; timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02v_pytest_gpu.log 2>&1; echo "rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/r02v_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (20 steps)"; SECONDS=0; timeout 600 python bench.py --steps 20 --warmup 5 --skip-sections > gpurun_out/r02v_bench_n1.json 2> gpurun_out/r02v_bench_n1.err; echo "rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02v_bench_n1.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, 'roofline', {k: d['roofline'].get(k) for k in ('frac', 'frac_sustained', 'avg_launch_us')})
print('e2e', {k: v for k, v in d['e2e'].items() if k in ('value', 'ms_per_step', 'frac_of_duplex_link')}, 'pageable', d['e2e'].get('pageable', {}).get('value'))
print('cg', {k: v for k, v in d['cg'].items() if k in ('iterations', 'cg_loop_s', 'ms_per_iteration', 'mixed_time_to_solution_s', 'cpu_reference_iterations')}, 'parity', d['parity'].get('ok'))
PY

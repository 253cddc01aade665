#!/bin/bash
# round 2, call n (1 GPU): final state - full GPU test suite, smoke(), the default bench line, the reference arm, and the
# one-GPU point of the strong-scaling series on 48^3x96 (configs[2])
mkdir -p gpurun_out
echo "== pytest -m gpu"; SECONDS=0; timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02n_pytest_gpu.log 2>&1; echo "rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/r02n_pytest_gpu.log
echo "== smoke"; SECONDS=0; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_smoke.log 2>&1; echo "rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/r02n_smoke.log
echo "== bench default"; SECONDS=0; timeout 900 python bench.py > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err; echo "rc=$? wall=${SECONDS}s"
echo "== bench reference arm"; SECONDS=0; timeout 900 python bench.py --impl reference > gpurun_out/r02n_bench_reference.json 2> gpurun_out/r02n_bench_reference.err; echo "rc=$? wall=${SECONDS}s"; tail -c 600 gpurun_out/r02n_bench_reference.json; echo
echo "== strong scaling N=1: 96x48x48x48"; SECONDS=0; timeout 900 python bench.py --lattice 96x48x48x48 --steps 20 --warmup 5 --skip-cpu --skip-sections --skip-anchor --skip-e2e > gpurun_out/r02n_strong_n1.json 2> gpurun_out/r02n_strong_n1.err; echo "rc=$? wall=${SECONDS}s"; tail -2 gpurun_out/r02n_strong_n1.err
python - <<'PY'
import json
for f in ('r02n_bench_n1', 'r02n_strong_n1'):
    try:
        d = json.loads(open('gpurun_out/' + f + '.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, {k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d['config'].get('lattice_TxLXxLYxLZ'))
    print('  roofline', {k: d['roofline'].get(k) for k in ('frac', 'frac_sustained', 'avg_launch_us', 'traffic')}, 'clocks', d.get('clocks'))
    if 'e2e' in d: print('  e2e', {k: v for k, v in d['e2e'].items() if k in ('value', 'ms_per_step', 'frac_of_duplex_link')})
    if 'cg' in d: print('  cg', {k: v for k, v in d['cg'].items() if k in ('iterations', 'cg_loop_s', 'ms_per_iteration', 'mixed_time_to_solution_s', 'mixed_count', 'cpu_reference_iterations', 'cpu_reference_time_to_solution_s')})
    if d.get('parity'): print('  parity ok', d['parity'].get('ok'), d['parity'].get('hop_rel_l2'), d['parity'].get('cg_iters'), d['parity'].get('cg_iters_ref'))
    if 'cpu_baseline' in d: print('  cpu', d['cpu_baseline'].get('value'), d['cpu_baseline'].get('cores'))
    for k in ('nd', 'hmc', 'sections'):
        if k in d: print('  ', k, str(d[k])[:400])
PY

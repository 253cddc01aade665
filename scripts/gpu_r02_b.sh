#!/bin/bash
# round 2, call b (1 GPU): GPU tests after the cross-rank-sum / sequence-number changes, new bench flow at N = 1
mkdir -p gpurun_out
echo "== pytest -m gpu"; SECONDS=0; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "rc=$? wall=${SECONDS}s"; tail -5 gpurun_out/r02b_pytest_gpu.log
echo "== bench driver style"; SECONDS=0; python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/r02b_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02b_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print('roofline', d['roofline']); print('e2e', d['e2e']); print('cg', d['cg'])
print('parity', d.get('parity')); print('anchor', d.get('weak_anchor')); print('cpu', d.get('cpu_baseline'))
PY

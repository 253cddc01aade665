#!/bin/bash
# round 2, call i (1 GPU): GPU tests after the CG-tail epilogue and the shared-memory links of the two-flavour kernel; their timings
mkdir -p gpurun_out
echo "== pytest -m gpu"; SECONDS=0; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02i_pytest_gpu.log 2>&1; rc=$?; echo "rc=$rc wall=${SECONDS}s"; tail -6 gpurun_out/r02i_pytest_gpu.log
[ $rc -eq 0 ] || exit 1
echo "== nd section"; timeout 600 python scripts/bench_sections.py nd 2>/dev/null > gpurun_out/r02i_section_nd.json; python -c "
import json,sys; d=json.loads(open('gpurun_out/r02i_section_nd.json').read().strip().splitlines()[-1]); print({k:v for k,v in d.items() if k.startswith('Qtm_pm_ndpsi') and 'bytes' not in k or k in ('iterations','time_to_solution_s','rgmixed','iterations_match_reference')})"
echo "== bench N=1 driver style"; python bench.py --steps 20 --warmup 5 --skip-sections > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e', d['e2e']); print('cg', d['cg']); print('parity ok', d['parity']['ok'], d['parity'].get('cg_iters_cpu_reference'))
PY
echo "== CG with the separate sweep (flag 32) for comparison"; python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import tmlqcd_b200 as tm
from conftest import random_gauge, random_spinor
rng=np.random.default_rng(1); d=tm.Device(48,24,24,24); d.set_params(0.16,0.0032); d.gauge_upload(random_gauge(rng,d.V))
E,O=d.field(random_spinor(rng,d.Vh)),d.field(random_spinor(rng,d.Vh)); En,On=d.field(),d.field()
for flags in (0,32,0,32):
    d.ck(d.lib.tmb_set_overlap(flags)); d.call("field_zero",On)
    it=d.call("invert_eo",En,On,E,O,1e-14,5000,1); d.call("field_zero",On)
    it=d.call("invert_eo",En,On,E,O,1e-14,5000,1); print("flags",flags,"iters",it,"cg loop s",d.solver_stats()[2])
    d.call("field_zero",On); it=d.call("invert_eo_mixed",En,On,E,O,1e-14,5000,1); print("   mixed count",it, "s", d.solver_stats()[2])
d.close()
PY

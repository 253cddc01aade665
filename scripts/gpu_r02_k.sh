#!/bin/bash
# round 2, call k (4 GPUs): the driver-style bench at N = 4 (T split) and the same per-GPU volume on a 2 x 2 (T x Z) grid
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 4"
echo "== bench N=4 (4x1)"; SECONDS=0; timeout 800 $TR --master-port 29561 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02k_bench_n4.json 2> gpurun_out/r02k_bench_n4.err; echo "rc=$? wall=${SECONDS}s"
echo "== bench N=4 (2x2)"; SECONDS=0; timeout 800 $TR --master-port 29562 bench.py --gpus 4 --nz 2 --lattice 24x48x48x24 --steps 20 --warmup 5 --skip-anchor > gpurun_out/r02k_bench_n4_grid2x2.json 2> gpurun_out/r02k_bench_n4_grid2x2.err; echo "rc=$? wall=${SECONDS}s"; grep -v "^\[\|^W\|^\*\|Setting OMP\|^#" gpurun_out/r02k_bench_n4_grid2x2.err | tail -5
python - <<'PY'
import json
for f in ('r02k_bench_n4','r02k_bench_n4_grid2x2'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
    print(f, {k:d[k] for k in ('value','ms_per_step','gpu_launches','peer_mode')}, d['config'].get('rank_grid_TxZ'), d['config'].get('global_lattice_TxLXxLYxLZ'))
    print('  roofline', {k:d['roofline'].get(k) for k in ('frac','frac_sustained','avg_launch_us')}); print('  comm', {k:v for k,v in d.get('comm',{}).items() if k!='how'})
    print('  e2e', {k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step','gbs_per_direction_per_gpu','frac_of_duplex_link')})
    print('  cg', {k:v for k,v in d['cg'].items() if k in ('iterations','cg_loop_s','ms_per_iteration','mixed_time_to_solution_s','mixed_count')})
    p=d['parity']; print('  parity ok', p['ok'], {k:{x:p[k][x] for x in ('hop_rel_l2','cg_iters','cg_iters_n1_device','ok')} for k in p if isinstance(p[k],dict) and 'hop_rel_l2' in p[k]})
    print('  anchor', d.get('weak_anchor'))
PY

#!/bin/bash
# round 2, call s: z faces through peer memory.  1 GPU: the loop-back tests.  4 GPUs ($1 = 4): 2 x 2, 1 x 4 and 1 x 2 grids against
# the oracle (peer push, TMB_ZPEER=0 = NCCL faces), then the bench on 2 x 2 grids (48^4 and 24^3x48)
N=${1:-1}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
[ $N -ge 4 ] || { echo "== pytest z split"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "z_split or cg_pro" > gpurun_out/r02s_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02s_pytest.log; }
if [ $N -ge 4 ]; then
par() { # name port nproc env args
  local name=$1 port=$2 np=$3 env=$4; shift 4
  echo "== $name"; env $env timeout 600 $TR --nproc-per-node $np --master-port $port scripts/mgpu_parity.py "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; grep -v "^\[\|^W\|^\*\|Setting OMP" gpurun_out/$name.log | tail -13
}
par r02s_mgpu_parity_grid2x2_zpeer 29581 4 X=1 8x8x8x8 --grid=2x2
if [ "$2" = "all" ]; then
par r02s_mgpu_parity_grid1x4_zpeer 29582 4 X=1 8x8x8x8 --grid=1x4
par r02s_mgpu_parity_grid1x2_zpeer 29583 2 X=1 8x8x8x8 --grid=1x2
fi
par r02s_mgpu_parity_grid2x2_znccl 29584 4 TMB_ZPEER=0 8x8x8x8 --grid=2x2
B="--nproc-per-node 4 bench.py --gpus 4 --steps 20 --warmup 5 --skip-anchor --nz 2"
echo "== bench 2x2 48^4"; SECONDS=0; timeout 800 $TR --master-port 29585 $B --lattice 24x48x48x24 > gpurun_out/r02s_bench_n4_grid2x2_48.json 2> gpurun_out/r02s_bench_n4_grid2x2_48.err; echo "rc=$? wall=${SECONDS}s"
echo "== bench 2x2 24^3x48"; SECONDS=0; timeout 800 $TR --master-port 29586 $B --lattice 24x24x24x12 > gpurun_out/r02s_bench_n4_grid2x2_24.json 2> gpurun_out/r02s_bench_n4_grid2x2_24.err; echo "rc=$? wall=${SECONDS}s"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02s_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, {k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'peer_mode')}, d['config'].get('rank_grid_TxZ'), d['config'].get('global_lattice_TxLXxLYxLZ'))
    print('  roofline', {k: d['roofline'].get(k) for k in ('frac', 'frac_sustained', 'avg_launch_us')}, 'comm', {k: v for k, v in d.get('comm', {}).items() if k != 'how'})
    print('  cg', {k: v for k, v in d['cg'].items() if k in ('iterations', 'cg_loop_s', 'ms_per_iteration', 'mixed_time_to_solution_s', 'mixed_count')})
    p = d['parity']; print('  parity ok', p['ok'], p.get('path'), p.get('hop_rel_l2'), p.get('cg_iters'), p.get('cg_iters_ref'), [k for k in p if isinstance(p[k], dict) and 'hop_rel_l2' in p[k]])
PY
fi

#!/bin/bash
echo "== pytest host pipeline"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_linktime.py -x -q -m gpu -k "pipelin or host_pointer or linktime" 2>&1 | tail -3
# round 2, call o (N GPUs, N = $1): strong-scaling points.  N = 4: 48^3x96 as 24x48^3 per GPU; 24^3x48 (configs[1]) as 4 x 1 and
# as 2 x 2 (T x Z).  N = 2: the driver-style bench and 48^3x96 as 48x48^3 per GPU.
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
run() { # name, port, args...
  local name=$1 port=$2; shift 2
  echo "== $name: $*"; SECONDS=0
  timeout 800 $TR --master-port $port bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "rc=$? wall=${SECONDS}s"
}
if [ "$N" = "4" ]; then
  run r02o_strong_48x48x48x96_n4 29571 --lattice 24x48x48x48 --skip-anchor
  run r02o_strong_24x24x24x48_n4_grid4x1 29572 --lattice 12x24x24x24 --skip-anchor
  run r02o_strong_24x24x24x48_n4_grid2x2 29573 --nz 2 --lattice 24x24x24x12 --skip-anchor
else
  run r02o_bench_n2 29574
  run r02o_strong_48x48x48x96_n2 29575 --lattice 48x48x48x48 --skip-anchor
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02o_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, {k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'peer_mode')}, d['config'].get('rank_grid_TxZ'), d['config'].get('global_lattice_TxLXxLYxLZ'))
    print('  roofline', {k: d['roofline'].get(k) for k in ('frac', 'frac_sustained', 'avg_launch_us')}, 'comm', {k: v for k, v in d.get('comm', {}).items() if k != 'how'})
    print('  e2e', {k: v for k, v in d['e2e'].items() if k in ('value', 'ms_per_step', 'frac_of_duplex_link')})
    print('  cg', {k: v for k, v in d['cg'].items() if k in ('iterations', 'cg_loop_s', 'ms_per_iteration', 'mixed_time_to_solution_s', 'mixed_count')})
    p = d['parity']; print('  parity ok', p['ok'], p.get('hop_rel_l2'), p.get('cg_iters'), p.get('cg_iters_ref'), 'n1 device cg_loop_s', p.get('n1_device', {}).get('cg_loop_s'))
    print('  anchor', d.get('weak_anchor'))
PY

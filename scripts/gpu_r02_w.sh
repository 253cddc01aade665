#!/bin/bash
# round 2, call w (2 GPUs): the final tree on the default multi-GPU path (T split, peer mode) - parity script and a bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
echo "== mgpu_parity"; timeout 300 $TR --master-port 29591 scripts/mgpu_parity.py 8x8x8x8 > gpurun_out/r02w_mgpu_parity_n2.log 2>&1; echo "rc=$?"; grep "MGPU\|peer mode\|cg_her\|deriv\|monomial" gpurun_out/r02w_mgpu_parity_n2.log | tail -8
echo "== bench"; SECONDS=0; timeout 400 $TR --master-port 29592 bench.py --gpus 2 --steps 20 --warmup 5 --skip-anchor > gpurun_out/r02w_bench_n2.json 2> gpurun_out/r02w_bench_n2.err; echo "rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02w_bench_n2.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches', 'peer_mode')}, 'roofline', {k: d['roofline'].get(k) for k in ('frac', 'avg_launch_us')})
print('e2e', {k: v for k, v in d['e2e'].items() if k in ('value', 'ms_per_step', 'frac_of_duplex_link')})
print('cg', {k: v for k, v in d['cg'].items() if k in ('iterations', 'cg_loop_s', 'ms_per_iteration', 'mixed_time_to_solution_s')}, 'parity', d['parity'].get('ok'), d['parity'].get('cg_iters'), d['parity'].get('cg_iters_ref'))
PY

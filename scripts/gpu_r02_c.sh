#!/bin/bash
# round 2, call c (N GPUs): multi-GPU parity (peer mode with cross-rank sums in the reduction finish, NCCL halos, peer hops
# with NCCL all-reduce) and the driver-style bench at N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
echo "== mgpu_parity peer"; timeout 600 $TR --master-port 29541 scripts/mgpu_parity.py 8x8x8x8 > gpurun_out/r02c_mgpu_parity_n${N}_peer.log 2>&1; echo "rc=$?"; grep -v "^\[\|^W\|^\*\|Setting OMP" gpurun_out/r02c_mgpu_parity_n${N}_peer.log | tail -16
echo "== mgpu_parity TMB_P2P=0"; TMB_P2P=0 timeout 600 $TR --master-port 29542 scripts/mgpu_parity.py 8x8x8x8 > gpurun_out/r02c_mgpu_parity_n${N}_nccl.log 2>&1; echo "rc=$?"; grep "MGPU\|peer mode\|cg_her\|FAIL" gpurun_out/r02c_mgpu_parity_n${N}_nccl.log | tail -6
echo "== mgpu_parity TMB_XRED=0"; TMB_XRED=0 timeout 600 $TR --master-port 29543 scripts/mgpu_parity.py 8x8x8x8 > gpurun_out/r02c_mgpu_parity_n${N}_peer_ncclsum.log 2>&1; echo "rc=$?"; grep "MGPU\|peer mode\|cg_her\|FAIL" gpurun_out/r02c_mgpu_parity_n${N}_peer_ncclsum.log | tail -6
echo "== bench --gpus $N driver style"; SECONDS=0; timeout 800 $TR --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02c_bench_n${N}.json 2> gpurun_out/r02c_bench_n${N}.err; echo "rc=$? wall=${SECONDS}s"; grep -v "^\[\|^W\|^\*\|Setting OMP\|^#" gpurun_out/r02c_bench_n${N}.err | tail -8
python - <<PY
import json
d=json.loads(open('gpurun_out/r02c_bench_n${N}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','peer_mode')}); print('roofline', {k:d['roofline'][k] for k in ('frac','frac_sustained','avg_launch_us','sustained_avg_launch_us')})
print('comm', d.get('comm')); print('e2e', d['e2e']); print('cg', d['cg'])
print('parity', json.dumps(d.get('parity'), indent=1)); print('anchor', d.get('weak_anchor'))
PY

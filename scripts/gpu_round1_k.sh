#!/bin/bash
# 2 GPUs: parity of the T-split paths (incl. the plaquette link halo) in peer mode and with NCCL halos, then the bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== N=2 parity, peer mode"; timeout 400 $TR --nproc-per-node 2 --master-port 29511 scripts/mgpu_parity.py 8x8x8x8 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -16
echo "== N=2 parity, NCCL halos"; TMB_P2P=0 timeout 400 $TR --nproc-per-node 2 --master-port 29512 scripts/mgpu_parity.py 8x8x8x8 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -4
echo "== N=2 bench (driver command line)"
timeout 900 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 500 --warmup 20 2> gpurun_out/r01c_bench_n2.err > gpurun_out/r01c_bench_n2.json
python -c "
import json; d=json.loads(open('gpurun_out/r01c_bench_n2.json').read().strip().splitlines()[-1]); print('us/hop', d['roofline']['avg_launch_us'], 'GFLOP/s', d['value'], 'peer', d.get('peer_mode'), 'cg', d['cg']['time_to_solution_s'], d['cg']['iterations'], 'mixed', d['cg'].get('mixed_time_to_solution_s'), d['roofline'].get('copy_gbs_sustained_this_run'))"
tail -2 gpurun_out/r01c_bench_n2.err

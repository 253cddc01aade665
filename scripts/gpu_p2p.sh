#!/bin/bash
# peer-mode hop (NVLink reads + flags, one launch per hop) vs NCCL halos: parity and throughput at N = $1 (default 2)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== loopback tests (1 GPU)"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hopping_and_epilogues or cg_and_invert or overlap or compression" 2>&1 | tail -3
echo "== N=$N parity, peer mode"; timeout 300 $TR --nproc-per-node $N --master-port 29511 scripts/mgpu_parity.py 8x8x8x8 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12
echo "== N=$N parity, NCCL halos"; TMB_P2P=0 timeout 300 $TR --nproc-per-node $N --master-port 29512 scripts/mgpu_parity.py 8x8x8x8 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3
for mode in 1 0; do
  echo "== N=$N bench TMB_P2P=$mode"
  TMB_P2P=$mode timeout 600 $TR --nproc-per-node $N --master-port 2951$mode bench.py --gpus $N --steps 500 --warmup 20 --skip-cpu 2> gpurun_out/bench_p2p${mode}_n$N.err > gpurun_out/bench_p2p${mode}_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_p2p${mode}_n$N.json')); print('us/hop', d['roofline']['avg_launch_us'], 'GFLOP/s', d['value'], 'peer', d.get('peer_mode'), 'cg', d['cg']['time_to_solution_s'], d['cg']['iterations'], 'mixed', d['cg'].get('mixed_time_to_solution_s'))"
  tail -2 gpurun_out/bench_p2p${mode}_n$N.err
done

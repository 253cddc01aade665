#!/bin/bash
# weak-scaling check at N GPUs: parity on a small lattice, then the bench line (peer mode and NCCL halos)
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== N=$N parity (peer mode)"; timeout 300 $TR --nproc-per-node $N --master-port 29541 scripts/mgpu_parity.py 4x8x8x8 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tail -9
for mode in ${MODES:-1 0}; do
  TMB_P2P=$mode timeout 900 $TR --nproc-per-node $N --master-port 2955$mode bench.py --gpus $N --steps 300 --warmup 20 --skip-cpu 2> gpurun_out/bench_scale_p2p${mode}_n$N.err > gpurun_out/bench_scale_p2p${mode}_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_scale_p2p${mode}_n$N.json')); print('N=$N TMB_P2P=$mode us/hop', round(d['roofline']['avg_launch_us'],1), 'GFLOP/s', round(d['value']), 'peer', d.get('peer_mode'), 'cg', d['cg']['iterations'], round(d['cg']['time_to_solution_s'],4), 'mixed', round(d['cg'].get('mixed_time_to_solution_s',0),4), d['clocks'])"
done

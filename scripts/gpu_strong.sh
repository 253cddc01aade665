#!/bin/bash
# strong scaling of BASELINE configs[2]: 48^3 x 96 global, T split over N ranks (local T = 96/N)
N=${1:-1}
EXTRA=${EXTRA:---global-chunk-t 12}   # the same global gauge field and sources for every N
TL=$((96 / N))
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 --lattice ${TL}x48x48x48 --steps 300 --warmup 20 --skip-cpu --skip-sections --skip-e2e $EXTRA 2> gpurun_out/r01c_strongcg_n1.err > gpurun_out/r01c_strongcg_n1.json
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 2957$N bench.py --gpus $N --lattice ${TL}x48x48x48 --steps 300 --warmup 20 --skip-cpu --skip-sections $EXTRA 2> gpurun_out/r01c_strongcg_n$N.err > gpurun_out/r01c_strongcg_n$N.json
fi
python -c "
import json; d=json.loads(open('gpurun_out/r01c_strongcg_n$N.json').read().strip().splitlines()[-1]); print('N=$N local', d['config']['lattice_TxLXxLYxLZ'], 'us/hop', round(d['roofline']['avg_launch_us'],1), 'GFLOP/s', round(d['value']), 'frac', round(d['roofline']['frac'],3), 'peer', d.get('peer_mode'), 'cg', d['cg']['iterations'], round(d['cg']['time_to_solution_s'],4), 'mixed', round(d['cg'].get('mixed_time_to_solution_s',0),4), d['clocks'])"
tail -3 gpurun_out/r01c_strongcg_n$N.err

#!/bin/bash
# the driver's scaling command line at N GPUs: both arms
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
SECONDS=0
timeout 900 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 3 2>gpurun_out/r01c_drv_ref_n$N.err | tail -1 | cut -c1-400
echo "reference arm wall=${SECONDS}s"; SECONDS=0
timeout 900 $TR bench.py --gpus $N --steps 500 --warmup 20 2> gpurun_out/r01c_drv_n$N.err > gpurun_out/r01c_drv_n$N.json; echo "rc=$? wall=${SECONDS}s"
python -c "
import json; d=json.loads(open('gpurun_out/r01c_drv_n$N.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('metric','value','n_gpus','ms_per_step','scaling','gpu_launches','peer_mode')}); print(d['config']['workload']); print('roofline', d['roofline']['frac'], 'cg', d['cg']['iterations'], d['cg']['time_to_solution_s'], 'e2e' in d, d['clocks'])"

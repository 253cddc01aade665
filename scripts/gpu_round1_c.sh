#!/bin/bash
# session C (2 GPUs): new tests, PDL/prefetch sweep, split overhead (loopback), NCCL overlap with priority stream
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu_c.log
echo "== overlap sweep 24^3x48"; timeout 600 python bench.py --sweep-overlap --steps 200 --warmup 10 --skip-cpu --skip-cg > gpurun_out/bench_c_sweep.json 2> gpurun_out/bench_c_sweep.err; echo "rc=$?"; grep overlap gpurun_out/bench_c_sweep.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c_sweep.json')); print('e2e', d.get('e2e'))"
echo "== N=1 48^3x12 plain / loopback / loopback+pdl"
for extra in "" "--loopback" "--loopback --overlap 1" "--overlap 3"; do
  timeout 600 python bench.py --gpus 1 --lattice 12x48x48x48 --steps 300 --warmup 20 --skip-cpu --skip-e2e --skip-cg $extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$extra', 'us/hop', d['roofline']['avg_launch_us'], 'GFLOP/s', d['value'])"
done
echo "== N=2 bench (priority comm stream)"
for extra in "" "--overlap 1"; do
timeout 900 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 500 --warmup 20 --skip-cpu $extra 2> gpurun_out/bench_c_n2.err | tee gpurun_out/bench_c_n2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2 $extra', 'us/hop', d['roofline']['avg_launch_us'], 'GFLOP/s', d['value'], 'cg', d['cg'])"
done
echo "== N=2 parity"; timeout 600 $TR --nproc-per-node 2 --master-port 29511 scripts/mgpu_parity.py 8x8x8x8 2>&1 | tail -2

#!/bin/bash
# session D (2 GPUs): fp32 operator + mixed CG tests, boundary kernel on the comm stream, N=2 scaling
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu_d.log
echo "== N=1 24^3x48"; timeout 900 python bench.py --steps 1000 --warmup 20 --skip-cpu > gpurun_out/bench_d_n1.json 2> gpurun_out/bench_d_n1.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_d_n1.json')); print('us/hop', d['roofline']['avg_launch_us'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cg', d['cg'])"
echo "== N=1 48^3x12 plain / loopback"
for extra in "" "--loopback"; do
  timeout 600 python bench.py --gpus 1 --lattice 12x48x48x48 --steps 300 --warmup 20 --skip-cpu --skip-e2e $extra 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$extra', 'us/hop', d['roofline']['avg_launch_us'], 'GFLOP/s', d['value'], 'cg', d['cg'])"
done
echo "== N=2 bench"
timeout 900 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 500 --warmup 20 --skip-cpu 2> gpurun_out/bench_d_n2.err | tail -1 > gpurun_out/bench_d_n2.json; python -c "
import json; d=json.load(open('gpurun_out/bench_d_n2.json')); print('N=2', 'us/hop', d['roofline']['avg_launch_us'], 'GFLOP/s', d['value'], 'cg', d['cg'])"
echo "== N=2 parity"; timeout 600 $TR --nproc-per-node 2 --master-port 29511 scripts/mgpu_parity.py 8x8x8x8 2>&1 | tail -2

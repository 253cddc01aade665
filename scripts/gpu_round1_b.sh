#!/bin/bash
# 2-GPU session: NCCL halo parity, weak-scaling bench (48^3x12 per GPU), same-local-volume 1-GPU reference point
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== mgpu parity N=2"; timeout 600 $TR --nproc-per-node 2 --master-port 29511 scripts/mgpu_parity.py 8x8x8x8 > gpurun_out/mgpu2.log 2>&1; echo "rc=$?"; grep -v "^\[W\|^W0\|Warning" gpurun_out/mgpu2.log | tail -12
echo "== mgpu parity N=2 thin slabs"; timeout 600 $TR --nproc-per-node 2 --master-port 29512 scripts/mgpu_parity.py 2x6x4x8 > gpurun_out/mgpu2b.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/mgpu2b.log
echo "== bench N=1 local 48^3x12"; timeout 600 python bench.py --gpus 1 --lattice 12x48x48x48 --steps 500 --warmup 20 --skip-cpu --skip-e2e > gpurun_out/bench_b_n1_48.json 2> gpurun_out/bench_b_n1_48.err; echo "rc=$?"; cat gpurun_out/bench_b_n1_48.json
echo "== bench N=2"; timeout 900 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 500 --warmup 20 --skip-cpu > gpurun_out/bench_b_n2.json 2> gpurun_out/bench_b_n2.err; echo "rc=$?"; cat gpurun_out/bench_b_n2.json; tail -5 gpurun_out/bench_b_n2.err
echo "== bench N=1 default"; timeout 900 python bench.py > gpurun_out/bench_b_n1.json 2> gpurun_out/bench_b_n1.err; echo "rc=$?"; cat gpurun_out/bench_b_n1.json

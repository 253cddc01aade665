#!/bin/bash
# session F (2 GPUs): epilogue load batching + fused reduction finish; CG launch list; N=2 check
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu_f.log
echo "== bench"; timeout 900 python bench.py --steps 1000 --warmup 20 --skip-cpu > gpurun_out/bench_f_n1.json 2> gpurun_out/bench_f_n1.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_f_n1.json')); print('us/hop', d['roofline']['avg_launch_us'], 'frac', d['roofline']['frac']); print('c12', d['compression12']['us_per_hop']); print('cg', d['cg'])"; tail -3 gpurun_out/bench_f_n1.err
echo "== N=2 bench"
timeout 900 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 500 --warmup 20 --skip-cpu 2> gpurun_out/bench_f_n2.err > gpurun_out/bench_f_n2.json; python -c "
import json; d=json.load(open('gpurun_out/bench_f_n2.json')); print('N=2', 'us/hop', d['roofline']['avg_launch_us'], 'GFLOP/s', d['value'], 'cg', d['cg'])"
echo "== N=2 parity"; timeout 600 $TR --nproc-per-node 2 --master-port 29511 scripts/mgpu_parity.py 8x8x8x8 2>&1 | tail -2
echo "== ncu CG launch list"
CMD="python scripts/cg_profile.py 48x24x24x24 12"
timeout 300 $CMD > gpurun_out/cg_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 100 --csv --log-file gpurun_out/cg_launches_f.csv $CMD > gpurun_out/cg_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/cg_plain.log

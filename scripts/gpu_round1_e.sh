#!/bin/bash
# session E (1 GPU): compression tests + numbers, CG launch list
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu_e.log
echo "== bench"; timeout 900 python bench.py --steps 1000 --warmup 20 --skip-cpu > gpurun_out/bench_e_n1.json 2> gpurun_out/bench_e_n1.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_e_n1.json')); print('us/hop', d['roofline']['avg_launch_us'], 'frac', d['roofline']['frac']); print('c12', d['compression12']); print('cg', d['cg'])"; tail -3 gpurun_out/bench_e_n1.err
echo "== ncu CG launch list"
CMD="python scripts/cg_profile.py 48x24x24x24 12"
timeout 300 $CMD > gpurun_out/cg_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/cg_launches.csv $CMD > gpurun_out/cg_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/cg_plain.log
CMD2="python scripts/cg_profile.py 48x24x24x24 12 mixed"
timeout 300 $CMD2 > gpurun_out/cgm_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/cgm_launches.csv $CMD2 > gpurun_out/cgm_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/cgm_plain.log

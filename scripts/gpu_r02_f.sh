#!/bin/bash
# round 2, call f (1 GPU): ncu evidence - launch list of the bench command, --set full of every kernel family
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --skip-cpu --skip-sections --skip-anchor --skip-parity"
echo "== plain bench"; $CMD > gpurun_out/r02f_bench_short.json 2> gpurun_out/r02f_bench_short.err && echo ok &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02f_ncu_list.log 2>&1
echo "ncu list rc=$?"
CMD3="python scripts/profile_r02.py 48x24x24x24 2"
$CMD3 > gpurun_out/r02f_profile_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:deriv_kernel\|hop2_kernel\|hop_kernel\|red_kernel\|ew_kernel -c 120 -f -o gpurun_out/r02_kernels $CMD3 > gpurun_out/r02f_ncu_kernels.log 2>&1
echo "ncu kernels rc=$?"; tail -3 gpurun_out/r02f_profile_plain.log
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# round-1 (second session) evidence: bench line, ncu launch list of the same command, ncu --set full of the top kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 40 --warmup 3 --skip-cpu --skip-sections"
echo "== plain bench"; $CMD > gpurun_out/r01b_bench_short.json 2> gpurun_out/r01b_bench_short.err && echo ok &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 500 --csv --log-file gpurun_out/r01b_launches_bench.csv $CMD > gpurun_out/r01b_ncu_list.log 2>&1
echo "ncu list rc=$?"
CMD2="python bench.py --steps 10 --warmup 3 --skip-cpu --skip-cg --skip-e2e --skip-sections"
$CMD2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hop_kernel -s 8 -c 3 -f -o gpurun_out/r01b_hop $CMD2 > gpurun_out/r01b_ncu_hop.log 2>&1
echo "ncu hop rc=$?"
CMD3="python scripts/profile_extra.py 48x24x24x24 4"
$CMD3 > gpurun_out/r01b_extra_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:deriv_kernel\|hop2_kernel\|hop_kernel -s 6 -c 12 -f -o gpurun_out/r01b_extra $CMD3 > gpurun_out/r01b_ncu_extra.log 2>&1
echo "ncu extra rc=$?"
ls -la gpurun_out/*.ncu-rep

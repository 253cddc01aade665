#!/bin/bash
# round 2, call f (1 GPU): ncu evidence - launch list of the bench command, per-kernel DRAM bytes / duration / registers /
# occupancy of every kernel family (a short metric list: two passes per launch), --set full of the headline hop kernel.
# Only CSV summaries are kept (a .ncu-rep of a hundred launches does not fit the 64 MiB that travel back).
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --skip-cpu --skip-sections --skip-anchor --skip-parity"
echo "== plain bench"; $CMD > gpurun_out/r02f_bench_short.json 2> gpurun_out/r02f_bench_short.err && echo ok &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02f_ncu_list.log 2>&1
echo "ncu list rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,l1tex__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
CMD3="python scripts/profile_r02.py 48x24x24x24 2"
$CMD3 > gpurun_out/r02f_profile_plain.log 2>&1 && ncu --metrics $M --clock-control none -k regex:deriv_kernel\|hop2_kernel\|hop_kernel\|red_kernel\|ew_kernel -c 130 --csv --log-file gpurun_out/r02_kernels_metrics.csv $CMD3 > gpurun_out/r02f_ncu_kernels.log 2>&1
echo "ncu kernels rc=$?"; tail -2 gpurun_out/r02f_profile_plain.log
CMD2="python bench.py --steps 6 --warmup 3 --skip-cpu --skip-cg --skip-e2e --skip-sections --skip-anchor --skip-parity"
$CMD2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hop_kernel -s 8 -c 2 -f -o /tmp/r02_hop $CMD2 > gpurun_out/r02f_ncu_hop.log 2>&1
echo "ncu hop rc=$?"
ncu -i /tmp/r02_hop.ncu-rep --page raw --csv > gpurun_out/r02_hop_kernel_ncu_raw.csv 2>/dev/null
ncu -i /tmp/r02_hop.ncu-rep --page details --csv > gpurun_out/r02_hop_kernel_ncu_details.csv 2>/dev/null
ls -la /tmp/r02_hop.ncu-rep gpurun_out/ | head -20; du -sh gpurun_out

#!/bin/bash
# first GPU session: parity tests, kernel-variant sweep, bench line, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench sweep" ; timeout 1500 python bench.py --sweep --steps 100 --warmup 10 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err ; echo "bench rc=$?" ; tail -45 gpurun_out/bench_a.err ; cat gpurun_out/bench_a.json
echo "== ncu"
CMD="python bench.py --steps 3 --warmup 3 --skip-cpu --skip-cg --skip-e2e"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hop_kernel -s 8 -c 2 -o gpurun_out/prof_hop $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -5 gpurun_out/ncu_full.log

#!/bin/bash
# session G (1 GPU): CUDA-graph CG chunk, small-lattice CG, full default bench
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_g.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu_g.log
echo "== small CG"; timeout 600 python scripts/small_cg_bench.py 2>&1 | tail -10
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench_g_n1.json 2> gpurun_out/bench_g_n1.err; echo "rc=$?"; cat gpurun_out/bench_g_n1.json; tail -3 gpurun_out/bench_g_n1.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/bench_g_ref.json 2> gpurun_out/bench_g_ref.err; echo "rc=$?"; cat gpurun_out/bench_g_ref.json

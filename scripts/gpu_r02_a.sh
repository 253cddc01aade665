#!/bin/bash
# round 2, call a: state of the tree on today's box - smoke, GPU tests, default bench
mkdir -p gpurun_out
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; SECONDS=0; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/r02a_pytest_gpu.log
echo "== bench default"; SECONDS=0; python bench.py > gpurun_out/r02a_bench_default.json 2> gpurun_out/r02a_bench_default.err; echo "rc=$? wall=${SECONDS}s"; tail -2 gpurun_out/r02a_bench_default.err
nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader

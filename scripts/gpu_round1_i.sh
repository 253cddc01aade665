#!/bin/bash
# session 3, call 2: C host examples, ND variants, epilogue-prefetch A/B in the CG, configs[0] section
mkdir -p gpurun_out
echo "== pytest (examples, nd, io, dropin)"; timeout 900 python -m pytest tests/test_c_host_examples.py tests/test_gpu_dropin_ops.py "tests/test_gpu_parity.py::test_nd_doublet" "tests/test_gpu_parity.py::test_tmLQCD_facade" -x -q -m gpu > gpurun_out/pytest_gpu_i.log 2>&1; tail -5 gpurun_out/pytest_gpu_i.log
echo "== CG epilogue prefetch"; timeout 300 python scripts/diag_r01c.py cgpf 2>&1 | tail -6
echo "== configs[0] section"; timeout 300 python scripts/bench_sections.py small 2>gpurun_out/small.err | tee gpurun_out/r01c_section_small.json; tail -3 gpurun_out/small.err

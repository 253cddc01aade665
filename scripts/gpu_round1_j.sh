#!/bin/bash
mkdir -p gpurun_out
echo "== pytest plaquette + examples + io"; timeout 900 python -m pytest tests/test_plaquette.py tests/test_c_host_examples.py tests/test_gpu_hmc.py -x -q -m gpu > gpurun_out/pytest_gpu_j.log 2>&1; tail -5 gpurun_out/pytest_gpu_j.log
echo "== configs[0] section (adaptive chunks)"; timeout 300 python scripts/bench_sections.py small 2>gpurun_out/small.err | tee gpurun_out/r01c_section_small.json | cut -c1-900

#!/bin/bash
# round 2, call q (1 GPU): host-pointer pipeline diagnostics (timeline, chunk schedules)
mkdir -p gpurun_out
timeout 600 python scripts/pipe_diag.py 48x24x24x24 > gpurun_out/r02q_pipe_diag.log 2>&1; echo "rc=$?"; cat gpurun_out/r02q_pipe_diag.log | cut -c1-1500

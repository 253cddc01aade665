#!/bin/bash
# round 2, last call (2 GPUs): the fermion force and the det monomial on a real 1 x 2 (Z-split) grid against the oracle
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29599 scripts/mgpu_parity.py 8x8x8x8 --grid=1x2 > gpurun_out/r02zz_mgpu_parity_grid1x2_force.log 2>&1; echo "rc=$?"
grep -v "^\[W\|^W1\|^\*\*\|Setting OMP" gpurun_out/r02zz_mgpu_parity_grid1x2_force.log | tail -16

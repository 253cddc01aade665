#!/bin/bash
# round 2, call j (N GPUs): Z split - single-GPU plumbing test, then real grids: 1 x 2 (Z only) on two GPUs, 2 x 2 on four
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== pytest z split + nd + cg"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "z_split or nd or cg_pro or host" > gpurun_out/r02j_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02j_pytest.log
echo "== grid 1x2"; timeout 600 $TR --nproc-per-node 2 --master-port 29551 scripts/mgpu_parity.py 8x8x8x8 --grid=1x2 > gpurun_out/r02j_mgpu_parity_grid1x2.log 2>&1; echo "rc=$?"; grep -v "^\[\|^W\|^\*\|Setting OMP" gpurun_out/r02j_mgpu_parity_grid1x2.log | tail -12
echo "== grid 1x2 TMB_P2P=0"; TMB_P2P=0 timeout 600 $TR --nproc-per-node 2 --master-port 29552 scripts/mgpu_parity.py 8x8x8x4 --grid=1x2 > gpurun_out/r02j_mgpu_parity_grid1x2_nccl.log 2>&1; echo "rc=$?"; grep "MGPU\|grid\|cg_her" gpurun_out/r02j_mgpu_parity_grid1x2_nccl.log | tail -5
if [ $N -ge 4 ]; then
echo "== grid 2x2"; timeout 600 $TR --nproc-per-node 4 --master-port 29553 scripts/mgpu_parity.py 8x8x8x8 --grid=2x2 > gpurun_out/r02j_mgpu_parity_grid2x2.log 2>&1; echo "rc=$?"; grep -v "^\[\|^W\|^\*\|Setting OMP" gpurun_out/r02j_mgpu_parity_grid2x2.log | tail -12
echo "== grid 2x2 TMB_P2P=0"; TMB_P2P=0 timeout 600 $TR --nproc-per-node 4 --master-port 29554 scripts/mgpu_parity.py 4x8x8x6 --grid=2x2 > gpurun_out/r02j_mgpu_parity_grid2x2_nccl.log 2>&1; echo "rc=$?"; grep "MGPU\|grid\|cg_her" gpurun_out/r02j_mgpu_parity_grid2x2_nccl.log | tail -5
echo "== grid 4x1 (T only, the default split)"; timeout 600 $TR --nproc-per-node 4 --master-port 29555 scripts/mgpu_parity.py 8x8x8x8 > gpurun_out/r02j_mgpu_parity_n4_peer.log 2>&1; echo "rc=$?"; grep "MGPU\|grid\|cg_her" gpurun_out/r02j_mgpu_parity_n4_peer.log | tail -5
fi

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "z_split" 2>&1 | tail -2
python scripts/profile_zsplit.py 24x48x48x24 2>&1 | tail -2
python scripts/profile_zsplit.py 24x24x24x12 2>&1 | tail -2

#!/bin/bash
# round 2, call x (1 GPU): last check of the pieces touched after call v (chunk schedule as a shared function, sequence-count hook)
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_linktime.py tests/test_gpu_dropin_ops.py -x -q -m gpu -k "z_split or pipelin or host_pointer or linktime or dropin_reference or facade or inversions" 2>&1 | tail -3

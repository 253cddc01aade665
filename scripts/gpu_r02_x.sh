#!/bin/bash
# round 2, call x (1 GPU): last checks of pieces touched late - chunk schedule as a shared function, sequence-count hook, the
# fermion force and the monomials on the Z split (loop-back)
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hmc.py -x -q -m gpu -k "z_split or monomials_vs_oracle_loopback or deriv_Sb" 2>&1 | tail -8

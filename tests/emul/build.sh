#!/bin/bash
# Builds the TEST-ONLY host emulation of the device site functions (see tmb_emul.cu).
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Xcompiler -fPIC -shared \
     -cudart shared -o libtmb_emul.so tmb_emul.cu

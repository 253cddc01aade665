#!/bin/bash
# Builds the TEST-ONLY host emulation of the device site functions (see tmb_emul.cu); skipped when up to date.
set -e
cd "$(dirname "$0")"
SRC=../../tmlqcd_b200/csrc
fresh=1
for f in tmb_emul.cu build.sh $SRC/tmb_hop.cuh $SRC/tmb_kernels.cu $SRC/tmb_force.cu $SRC/tmb_site.cuh $SRC/tmb_geom.h $SRC/tmb_kernels.h; do
  if [ ! -e libtmb_emul.so ] || [ "$f" -nt libtmb_emul.so ]; then fresh=0; fi
done
if [ $fresh = 1 ]; then exit 0; fi
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Xcompiler -fPIC -shared \
     -cudart shared -o libtmb_emul.so tmb_emul.cu

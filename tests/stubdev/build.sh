#!/bin/bash
# TEST ONLY: the product's host layer (tmb_dropin.c, tmb_io.c) over a host stand-in for the device-level C ABI
# (tmb_stub.c, which delegates to the CPU oracle).  See the header of tmb_stub.c.
set -e
cd "$(dirname "$0")"
SRC=../../tmlqcd_b200/csrc
fresh=1
for f in tmb_stub.c build.sh $SRC/tmb_dropin.c $SRC/tmb_io.c ../../oracle/tmoracle.c ../../include/tmlqcd_b200.h ../../include/tmlqcd_b200_dropin.h; do
  if [ ! -e libtmb_dropin_stub.so ] || [ "$f" -nt libtmb_dropin_stub.so ]; then fresh=0; fi
done
if [ $fresh = 1 ]; then exit 0; fi
gcc -std=gnu99 -O2 -fno-strict-aliasing -ffp-contract=off -fPIC -shared -Wall -Wl,-Bsymbolic-functions -o libtmb_dropin_stub.so \
    $SRC/tmb_dropin.c $SRC/tmb_io.c tmb_stub.c ../../oracle/tmoracle.c -lm

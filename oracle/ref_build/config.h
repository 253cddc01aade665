/* Hand-written config.h for compiling the UNMODIFIED tmLQCD reference sources
 * (read in place from /root/reference) into oracle/_ref/.  Replaces what
 * autoconf's configure.in -> config.h would generate (configure.in:92-1022);
 * autoconf is not available in this image.  Generic-C branch only: no SSE,
 * no MPI.  TM_USE_OMP / _USE_HALFSPINOR are switched from the Makefile.
 * TEST INFRASTRUCTURE ONLY - never linked into the product library. */
#ifndef TMB_REF_CONFIG_H
#define TMB_REF_CONFIG_H
#define ALIGN
#define ALIGN32
#define ALIGN_BASE 0x00
#define ALIGN_BASE32 0x00
#define _GAUGE_COPY 1
#define HAVE_CLOCK_GETTIME 1
#define PACKAGE_VERSION "ref-oracle"
#endif

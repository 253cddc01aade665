/* Stand-in for the header of the un-vendored third-party c-lime library (usqcd-software/c-lime,
 * "tested with 1.2.3", reference README:10), which is absent from /root/reference and this image.
 * solver/monomial_solve.c reaches it only through io/spinor.h -> io/selector.h:23 for type names in
 * prototypes; no LIME function is ever called on the scoped path.  Opaque types only.
 * TEST INFRASTRUCTURE ONLY (oracle/ref_build). */
#ifndef TMB_REF_STUB_LIME_H
#define TMB_REF_STUB_LIME_H
#include <stdint.h>
#include <stdio.h>
typedef uint64_t n_uint64_t;
typedef struct LimeReader LimeReader;
typedef struct LimeWriter LimeWriter;
typedef struct LimeRecordHeader LimeRecordHeader;
#endif

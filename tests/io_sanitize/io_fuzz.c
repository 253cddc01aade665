/* io_fuzz.c - TEST ONLY: the product's file code (tmlqcd_b200/csrc/tmb_io.c) built with AddressSanitizer + UBSan
 * against damaged files.  A valid ILDG configuration and a valid SciDAC propagator file are written, then truncated,
 * bit-flipped (mostly inside the record headers) or overwritten with garbage N times and handed to read_gauge_field /
 * read_spinor: every call must return (0 or an error code) without a memory error.  The harness defines the handful
 * of reference globals tmb_io.c reads; no GPU and no CUDA library is involved. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "tmlqcd_b200.h"
#include "tmlqcd_b200_dropin.h"
int T, L, LX, LY, LZ, VOLUME, RAND, VOLUMEPLUSRAND, g_update_gauge_copy, g_proc_id, g_debug_level = 0, g_nproc = 1, g_nproc_t = 1;
double g_kappa = 0.16, g_mu = 0.003, g_mubar, g_epsbar;
su3 **g_gauge_field;
static unsigned long long st = 0x9E3779B97F4A7C15ull;
static unsigned rnd(void) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (unsigned)(st >> 11); }
static unsigned char *slurp(const char *fn, long *n) {
  FILE *f = fopen(fn, "rb"); if (!f) exit(3);
  fseek(f, 0, SEEK_END); *n = ftell(f); rewind(f);
  unsigned char *b = malloc((size_t)*n);
  if (fread(b, 1, (size_t)*n, f) != (size_t)*n) exit(3);
  fclose(f); return b;
}
int main(int argc, char **argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 400;
  T = 4; L = LX = LY = LZ = 4; VOLUME = 256; VOLUMEPLUSRAND = 256;
  su3 *slab = calloc((size_t)VOLUME * 4, sizeof(su3)); g_gauge_field = calloc((size_t)VOLUME, sizeof(su3 *));
  for (int i = 0; i < VOLUME; i++) g_gauge_field[i] = slab + 4 * i;
  double *d = (double *)slab; for (int i = 0; i < VOLUME * 72; i++) d[i] = (double)rand() / RAND_MAX;
  paramsXlfInfo *x = construct_paramsXlfInfo(0.5, 3);
  if (write_gauge_field("fuzz_conf.ok", 64, x)) return 1;
  spinor *e = calloc(128, sizeof(spinor)), *o = calloc(128, sizeof(spinor));
  if (tmb_write_propagator("fuzz_prop.ok", e, o, 32, 1e-10, 12, "CG", 0)) return 2;
  if (read_gauge_field("fuzz_conf.ok", g_gauge_field) || read_spinor(e, o, "fuzz_prop.ok", 0)) return 4; /* intact files are accepted */
  long n, np;
  unsigned char *buf = slurp("fuzz_conf.ok", &n), *bp = slurp("fuzz_prop.ok", &np);
  int ok = 0, bad = 0;
  if (!freopen("/dev/null", "w", stdout)) return 5; /* the readers' diagnostics */
  for (int it = 0; it < iters; it++) {
    const int which = it & 1;
    const unsigned char *src = which ? bp : buf; const long len = which ? np : n;
    unsigned char *c = malloc((size_t)len); memcpy(c, src, (size_t)len);
    long newlen = len;
    const int mode = (int)(rnd() % 3);
    if (mode == 0) newlen = (long)(rnd() % (unsigned)len);
    else if (mode == 1) for (int k = 0; k < 1 + (int)(rnd() % 4); k++) c[rnd() % (unsigned)(rnd() % 2 ? 600 : len)] ^= (unsigned char)(1u << (rnd() % 8));
    else { const long p = (long)(rnd() % (unsigned)(len - 16)); for (int k = 0; k < 8; k++) c[p + k] = (unsigned char)rnd(); }
    FILE *f = fopen("fuzz.tmp", "wb"); fwrite(c, 1, (size_t)newlen, f); fclose(f); free(c);
    const int rc = which ? read_spinor(e, o, "fuzz.tmp", 0) : read_gauge_field("fuzz.tmp", g_gauge_field);
    if (rc == 0) ok++; else bad++;
  }
  fprintf(stderr, "IOFUZZ accepted %d refused %d\n", ok, bad);
  return 0;
}

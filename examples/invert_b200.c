/* invert_b200.c - the library facade of include/tmLQCD.h (wrapper/lib_wrapper.c) used from plain C the way an
 * external code (e.g. a contraction code) uses the reference's libwrapper: read "invert.input", read an ILDG
 * gauge configuration, invert a lexicographic point source, have the propagator written as a SciDAC file.
 *
 * Self-contained: the program first writes the invert.input and - with write_gauge_field - a hot-start
 * configuration conf.0000 into the working directory.  Checks (exit code 0 only if all pass):
 *   - the configuration read back by tmLQCD_read_gauge is bit-identical to the one written,
 *   - D_psi(propagator) / (2 kappa) reproduces the source (the facade normalises by 2 kappa, lib_wrapper.c:274),
 *   - the propagator file read by read_spinor equals the returned propagator to single precision.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "tmlqcd_b200.h"
#include "tmlqcd_b200_dropin.h"

static unsigned long long rng_state = 88172645463325252ull;
static double uniform(void) {
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (rng_state >> 11) * (1.0 / 9007199254740992.0) + 1e-17;
}
static double gauss(void) { return sqrt(-2. * log(uniform())) * cos(6.283185307179586 * uniform()); }
static void random_su3(su3 *u) {
  _Complex double a[3], b[3], s = 0;
  double n = 0;
  for (int i = 0; i < 3; i++) { a[i] = gauss() + I * gauss(); b[i] = gauss() + I * gauss(); n += creal(a[i] * conj(a[i])); }
  n = 1. / sqrt(n);
  for (int i = 0; i < 3; i++) { a[i] *= n; s += conj(a[i]) * b[i]; }
  n = 0;
  for (int i = 0; i < 3; i++) { b[i] -= s * a[i]; n += creal(b[i] * conj(b[i])); }
  n = 1. / sqrt(n);
  for (int i = 0; i < 3; i++) b[i] *= n;
  u->c00 = a[0]; u->c01 = a[1]; u->c02 = a[2];
  u->c10 = b[0]; u->c11 = b[1]; u->c12 = b[2];
  u->c20 = conj(a[1] * b[2] - a[2] * b[1]); u->c21 = conj(a[2] * b[0] - a[0] * b[2]); u->c22 = conj(a[0] * b[1] - a[1] * b[0]);
}

int main(void) {
  int fails = 0;
  FILE *f = fopen("invert.input", "w");
  if (!f) { perror("invert.input"); return 2; }
  fprintf(f, "# written by invert_b200.c\nT = 8\nL = 4\nkappa = 0.16\n2KappaMu = 0.0032\nThetaT = 1.0\n"
             "GaugeConfigInputFile = conf\nSourceFilename = prop_b200\n"
             "BeginOperator TMWILSON\n  kappa = 0.16\n  2KappaMu = 0.0032\n  SolverPrecision = 1e-18\n"
             "  MaxSolverIterations = 2000\n  UseRelativePrecision = yes\n  PropagatorPrecision = 32\nEndOperator\n"
             /* the keys that pick invert_eo's branch (read_input.l:1108-1139, :967-974; default: cg with even/odd preconditioning) */
             "BeginOperator TMWILSON\n  kappa = 0.16\n  2KappaMu = 0.0032\n  Solver = rgmixedcg\n  mcgdelta = 1.e-4\n"
             "  SolverPrecision = 1e-18\n  MaxSolverIterations = 2000\n  SolverRelativePrecision = yes\nEndOperator\n"
             "BeginOperator TMWILSON\n  kappa = 0.16\n  2KappaMu = 0.0032\n  Solver = cg\n  UseEvenOdd = no\n"
             "  SolverPrecision = 1e-18\n  MaxSolverIterations = 4000\n  SolverRelativePrecision = yes\nEndOperator\n");
  fclose(f);
  if (tmLQCD_invert_init(0, NULL, 1, 0) != 0) return 2;
  tmLQCD_lat_params lp;
  tmLQCD_get_lat_params(&lp);
  printf("# lattice %u x %u x %u x %u, %u operator(s)\n", lp.T, lp.LX, lp.LY, lp.LZ, lp.no_operators);

  /* a hot-start configuration, written as an ILDG file, wiped, and read back through the facade */
  double *gf = NULL;
  tmLQCD_get_gauge_field_pointer(&gf);
  su3 *links = (su3 *)gf;
  const size_t nlinks = (size_t)4 * VOLUME;
  for (size_t i = 0; i < nlinks; i++) random_su3(&links[i]);
  su3 *copy = (su3 *)malloc(nlinks * sizeof(su3));
  memcpy(copy, links, nlinks * sizeof(su3));
  paramsXlfInfo *xlf = construct_paramsXlfInfo(0.5, 0);
  if (write_gauge_field("conf.0000", 64, xlf) != 0) { fprintf(stderr, "write_gauge_field failed\n"); return 2; }
  free(xlf);
  memset(links, 0, nlinks * sizeof(su3));
  if (tmLQCD_read_gauge(0) != 0) return 2;
  if (memcmp(copy, links, nlinks * sizeof(su3)) != 0) { printf("# gauge field read back differs\n"); fails++; }
  else printf("# conf.0000 written and read back bit-identically (checksum %#x %#x)\n", GaugeInfo.checksum.suma, GaugeInfo.checksum.sumb);

  /* point source at the origin, spin 0 colour 0 (lexicographic, 24 doubles per site) */
  double *src = (double *)calloc((size_t)24 * VOLUME, sizeof(double)), *prop = (double *)calloc((size_t)24 * VOLUME, sizeof(double));
  src[0] = 1.;
  if (tmLQCD_invert(prop, src, 0, 1) != 0) return 2;
  int iters = 0; double reached = 0.;
  tmLQCD_b200_get_solver_info(0, &iters, &reached);
  printf("# inversion: %d iterations, squared residue %e\n", iters, reached);

  /* D_psi(prop) / (2 kappa) == source */
  spinor *chk = (spinor *)calloc((size_t)VOLUME, sizeof(spinor));
  D_psi(chk, (spinor *)prop);
  double d2 = 0.;
  for (size_t i = 0; i < (size_t)24 * VOLUME; i++) { const double x = ((double *)chk)[i] / (2. * g_kappa) - src[i]; d2 += x * x; }
  printf("# |D_psi(prop)/(2 kappa) - source|^2 = %e\n", d2);
  if (!(d2 <= 1e-16)) fails++;

  /* the other two operators of invert.input - reliable-update mixed CG, and CG without even/odd preconditioning - give the same propagator */
  double *prop2 = (double *)calloc((size_t)24 * VOLUME, sizeof(double));
  for (int op = 1; op < (int)lp.no_operators; op++) {
    if (tmLQCD_invert(prop2, src, op, 0) != 0) return 2;
    tmLQCD_b200_get_solver_info(op, &iters, &reached);
    double dd = 0., nn = 0.;
    for (size_t i = 0; i < (size_t)24 * VOLUME; i++) { const double x = prop2[i] - prop[i]; dd += x * x; nn += prop[i] * prop[i]; }
    printf("# operator %d (%s): %d iterations, squared residue %e, relative squared difference to operator 0: %e\n", op,
           op == 1 ? "Solver = rgmixedcg" : "UseEvenOdd = no", iters, reached, dd / nn);
    if (!(iters > 0 && dd / nn <= 1e-13)) fails++;
  }
  free(prop2);

  /* the propagator file (source.<nstore>.<ix>.<is>.inverted, operator.c:566) against the returned propagator */
  spinor *e = (spinor *)calloc((size_t)VOLUME / 2, sizeof(spinor)), *o = (spinor *)calloc((size_t)VOLUME / 2, sizeof(spinor));
  spinor *lex = (spinor *)calloc((size_t)VOLUME, sizeof(spinor));
  if (read_spinor(e, o, "prop_b200.0000.00.00.inverted", 0) != 0) { printf("# cannot read the propagator file\n"); fails++; }
  else {
    convert_eo_to_lexic(lex, e, o);
    double dd = 0., nn = 0.;
    for (size_t i = 0; i < (size_t)24 * VOLUME; i++) { const double x = ((double *)lex)[i] - prop[i]; dd += x * x; nn += prop[i] * prop[i]; }
    printf("# propagator file vs returned propagator: relative squared difference %e (single precision file)\n", dd / nn);
    if (!(dd / nn <= 1e-13)) fails++;
  }
  tmLQCD_finalise();
  printf(fails ? "# FAILED (%d checks)\n" : "# all checks passed\n", fails);
  return fails ? 1 : 0;
}

/* benchmark_b200.c - a plain-C host program on top of libtmlqcd_b200.so, with the flow of the reference's
 * `benchmark` executable (benchmark.c:127-327: lattice set-up, hot-start gauge field, boundary(kappa), Gaussian
 * spinor, warm-up, timed Hopping_Matrix EO+OE pairs, then D_psi), followed by an even/odd CG solve through
 * invert_eo and the reference's residual check |M_full x - b|^2 (operator.c:358-384).
 *
 * It includes ONLY the two public headers and calls the reference-named symbols with host pointers; the second
 * timing block shows the same pairs with device-resident fields through the device-level C ABI.
 * Exit code 0 only if every internal consistency check passes.
 *
 *   gcc -std=gnu99 -O2 -I../include benchmark_b200.c -L../tmlqcd_b200/lib -ltmlqcd_b200 -Wl,-rpath,... -lm
 *   ./benchmark_b200 [T LX LY LZ]      (default 16 8 8 8)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "tmlqcd_b200.h"
#include "tmlqcd_b200_dropin.h"

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static double uniform(void) { /* xorshift64*, (0,1) */
  rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
  return ((rng_state * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0) + 1e-17;
}
static double gauss(void) { return sqrt(-2. * log(uniform())) * cos(6.283185307179586 * uniform()); }
static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/* hot start: every link an independent random SU(3) (two Gaussian rows, Gram-Schmidt, third row = conj(r0 x r1)) */
static void random_su3(su3 *u) {
  _Complex double a[3], b[3], c[3], s = 0;
  double n = 0;
  for (int i = 0; i < 3; i++) { a[i] = gauss() + I * gauss(); b[i] = gauss() + I * gauss(); n += creal(a[i] * conj(a[i])); }
  n = 1. / sqrt(n);
  for (int i = 0; i < 3; i++) { a[i] *= n; s += conj(a[i]) * b[i]; }
  n = 0;
  for (int i = 0; i < 3; i++) { b[i] -= s * a[i]; n += creal(b[i] * conj(b[i])); }
  n = 1. / sqrt(n);
  for (int i = 0; i < 3; i++) b[i] *= n;
  c[0] = conj(a[1] * b[2] - a[2] * b[1]); c[1] = conj(a[2] * b[0] - a[0] * b[2]); c[2] = conj(a[0] * b[1] - a[1] * b[0]);
  u->c00 = a[0]; u->c01 = a[1]; u->c02 = a[2];
  u->c10 = b[0]; u->c11 = b[1]; u->c12 = b[2];
  u->c20 = c[0]; u->c21 = c[1]; u->c22 = c[2];
}
static void random_spinor_field(spinor *s, int n) { /* variance 1/2 per real component like gauss_vector (start.c:81-107) */
  double *d = (double *)s;
  for (size_t i = 0; i < (size_t)n * 24; i++) d[i] = gauss() * 0.7071067811865476;
}
static spinor *field(int n) {
  spinor *p = (spinor *)tmb_host_alloc((size_t)n * sizeof(spinor)); /* pinned; plain calloc works too, slower over PCIe */
  if (!p) { fprintf(stderr, "host allocation failed: %s\n", tmb_last_error()); exit(2); }
  memset(p, 0, (size_t)n * sizeof(spinor));
  return p;
}
static double sqdiff(const spinor *a, const spinor *b, int n) {
  const double *x = (const double *)a, *y = (const double *)b;
  double s = 0;
  for (size_t i = 0; i < (size_t)n * 24; i++) s += (x[i] - y[i]) * (x[i] - y[i]);
  return s;
}

int main(int argc, char **argv) {
  int t = 16, lx = 8, ly = 8, lz = 8, fails = 0;
  if (argc == 5) { t = atoi(argv[1]); lx = atoi(argv[2]); ly = atoi(argv[3]); lz = atoi(argv[4]); }
  if (tmb_dropin_init(t, lx, ly, lz, -1) != 0) { fprintf(stderr, "tmb_dropin_init: %s\n", tmb_last_error()); return 2; }
  printf("# The lattice size is %d x %d x %d x %d\n", T, LX, LY, LZ);
  g_kappa = 0.16; g_mu = 2. * g_kappa * 0.01;
  X0 = 1.; X1 = X2 = X3 = 0.; /* antiperiodic in time */
  boundary(g_kappa);
  for (int ix = 0; ix < VOLUME; ix++)
    for (int mu = 0; mu < 4; mu++) random_su3(&g_gauge_field[ix][mu]);
  g_update_gauge_copy = 1; /* the reference's dirty flag (start.c:506): the next operator call uploads the links */

  const int Vh = VOLUME / 2;
  spinor *k = field(Vh), *l = field(Vh), *m = field(Vh), *w = field(Vh);
  random_spinor_field(k, Vh);

  /* ---- Hopping_Matrix pairs with host pointers: benchmark.c:262-327 ---- */
  Hopping_Matrix(EO, l, k); Hopping_Matrix(OE, m, l); /* warm-up, gauge upload */
  int reps = 20;
  double t0 = now();
  for (int j = 0; j < reps; j++) { Hopping_Matrix(EO, l, k); Hopping_Matrix(OE, m, l); }
  double dt = (now() - t0) / reps;
  printf("# The following result is just to make sure that the calculation is not optimized away: %e\n", square_norm(m, Vh, 0));
  printf("# host pointers : %.1f us per EO+OE pair, %.1f Mflops (1608 flop/site)\n", 1e6 * dt, 1608. * VOLUME / dt / 1e6);

  /* ---- the same pair with device-resident fields (tmlqcd_b200.h) ---- */
  void *dk = tmb_field_alloc(), *dl = tmb_field_alloc(), *dm = tmb_field_alloc();
  if (!dk || !dl || !dm || tmb_field_upload(dk, (const double *)k) != 0) { fprintf(stderr, "%s\n", tmb_last_error()); return 2; }
  reps = 500;
  for (int j = 0; j < 10; j++) { tmb_Hopping_Matrix(EO, dl, dk); tmb_Hopping_Matrix(OE, dm, dl); }
  float ms = 0.f;
  tmb_timer_start();
  for (int j = 0; j < reps; j++) { tmb_Hopping_Matrix(EO, dl, dk); tmb_Hopping_Matrix(OE, dm, dl); }
  tmb_timer_stop(&ms);
  dt = 1e-3 * ms / reps;
  printf("# device fields : %.1f us per EO+OE pair, %.1f Mflops (1608 flop/site), %.1f GB/s at 1536 B/site\n", 1e6 * dt,
         1608. * VOLUME / dt / 1e6, 1536. * VOLUME / dt / 1e9);
  if (tmb_field_download((double *)w, dm) != 0) { fprintf(stderr, "%s\n", tmb_last_error()); return 2; }
  double d2 = sqdiff(w, m, Vh);
  printf("# host-pointer and device-resident results differ by %e (must be 0)\n", d2);
  if (d2 != 0.) fails++;

  /* ---- D_psi on a lexicographic field, checked against M_full on its (even, odd) halves ---- */
  spinor *P = field(VOLUME), *Q = field(VOLUME), *R = field(VOLUME);
  random_spinor_field(Q, VOLUME);
  D_psi(P, Q);
  reps = 10;
  t0 = now();
  for (int j = 0; j < reps; j++) D_psi(P, Q);
  dt = (now() - t0) / reps;
  printf("# D_psi, host pointers: %.1f us per application, %.1f Mflops (1680 flop/site)\n", 1e6 * dt, 1680. * VOLUME / dt / 1e6);
  convert_lexic_to_eo(k, l, Q);
  M_full(m, w, k, l);
  convert_eo_to_lexic(R, m, w);
  d2 = sqdiff(P, R, VOLUME) / square_norm(P, VOLUME, 0);
  printf("# |D_psi(Q) - M_full(Q)|^2 / |D_psi(Q)|^2 = %e\n", d2);
  if (!(d2 <= 1e-26)) fails++;

  /* ---- even/odd CG through invert_eo, then the reference's check |M x - b|^2 (operator.c:358-384) ---- */
  spinor *Even = field(Vh), *Odd = field(Vh), *En = field(Vh), *On = field(Vh);
  random_spinor_field(Even, Vh); random_spinor_field(Odd, Vh);
  solver_params_t sp; memset(&sp, 0, sizeof(sp));
  t0 = now();
  int iter = invert_eo(En, On, Even, Odd, 1e-20, 5000, TMB_SOLVER_CG, 1, 0, 1, 0, NULL, sp, 0, NO_EXT_INV, SLOPPY_DOUBLE, NO_COMPRESSION);
  dt = now() - t0;
  M_full(m, w, En, On);
  diff(m, m, Even, Vh); diff(w, w, Odd, Vh);
  const double res = square_norm(m, Vh, 0) + square_norm(w, Vh, 0), src = square_norm(Even, Vh, 0) + square_norm(Odd, Vh, 0);
  printf("# invert_eo: %d iterations in %.4f s, |M x - b|^2 / |b|^2 = %e\n", iter, dt, res / src);
  if (iter < 0 || !(res / src <= 1e-18)) fails++;

  tmb_dropin_finalize();
  printf(fails ? "# FAILED (%d checks)\n" : "# all checks passed\n", fails);
  return fails ? 1 : 0;
}

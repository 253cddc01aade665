/* tmb_stub.c - TEST INFRASTRUCTURE ONLY.
 *
 * A host stand-in for the DEVICE-LEVEL C ABI of include/tmlqcd_b200.h, so that the product's host layer
 * (tmlqcd_b200/csrc/tmb_dropin.c: the reference-named symbols, their argument conventions, scratch use, the
 * N = VOLUME/2 | VOLUME splitting, the globals it pushes, the tmLQCD.h facade with its lexicographic conversion and
 * 2 kappa normalisation) can be exercised in the GPU-less container.  "Device fields" are plain host arrays in the
 * reference's AoS layout and every operator is delegated to the CPU oracle (oracle/tmoracle.c, linked into the same
 * test library).  tests/stubdev/build.sh links THIS file with tmb_dropin.c and tmb_io.c into
 * tests/stubdev/libtmb_dropin_stub.so; nothing under tmlqcd_b200/ references it, it is never a fallback of the
 * product (whose tmb_init refuses to run without a CUDA device).
 *
 * The chronological guess is restated here on host fields (a few lines; the oracle keeps its own inside the monomials).
 */
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/tmlqcd_b200.h"

/* oracle/tmoracle.c */
int orc_init(int, int, int, int);
void orc_finalize(void);
void orc_set_hopping_phases(const double ka8[8], double gmu);
void orc_set_nd_params(double, double, double);
void orc_set_gauge(const double *);
void orc_Hopping_Matrix(int, double *, const double *);
void orc_tm_times_Hopping_Matrix(int, double *, const double *, double, double);
void orc_tm_sub_Hopping_Matrix(int, double *, const double *, const double *, double, double);
void orc_H_eo_tm_inv_psi(double *, const double *, int, double);
void orc_tm_sub_H_eo_gamma5(double *, const double *, const double *, int, double);
void orc_Qtm_pm_psi(double *, const double *);
void orc_Qtm_plus_psi(double *, const double *);
void orc_Qtm_minus_psi(double *, const double *);
void orc_Mtm_plus_psi(double *, const double *);
void orc_Mtm_minus_psi(double *, const double *);
void orc_M_full(double *, double *, const double *, const double *);
void orc_Q_full(double *, double *, const double *, const double *);
void orc_convert_eo_to_lexic(double *, const double *, const double *);
void orc_convert_lexic_to_eo(double *, double *, const double *);
int orc_cg_her(double *, const double *, int, double, int);
int orc_invert_eo_cg(double *, double *, const double *, const double *, double, int, int);
void orc_M_ee_inv_ndpsi(double *, double *, const double *, const double *, double, double);
void orc_M_oo_sub_g5_ndpsi(double *, double *, const double *, const double *, const double *, const double *, double, double);
void orc_Qtm_ndpsi(double *, double *, const double *, const double *);
void orc_Qtm_dagger_ndpsi(double *, double *, const double *, const double *);
void orc_Qtm_pm_ndpsi(double *, double *, const double *, const double *);
int orc_cg_her_nd(double *, double *, const double *, const double *, int, double, int);
int orc_invert_doublet_eo_cg(double *, double *, double *, double *, const double *, const double *, const double *, const double *, double, int, int);
void orc_deriv_Sb(int, const double *, const double *, double *, double);
double orc_measure_plaquette(void);

static struct {
  int up, T, LX, LY, LZ, V, Vh;
  double ka[8], mu;
  double *gauge, *df;
  int last_iters; double last_err;
  char err[256];
} S;
#define NEED() do { if (!S.up) { snprintf(S.err, sizeof(S.err), "tmb_init has not been called"); return -1; } } while (0)
#define NF ((size_t)S.Vh * 24)
typedef double _Complex cplx;

int tmb_init(int T, int LX, int LY, int LZ, int device) {
  (void)device;
  if (S.up) return (S.T == T && S.LX == LX && S.LY == LY && S.LZ == LZ) ? 0 : -2;
  if (orc_init(T, LX, LY, LZ) != 0) return -3;
  memset(&S, 0, sizeof(S));
  S.up = 1; S.T = T; S.LX = LX; S.LY = LY; S.LZ = LZ; S.V = T * LX * LY * LZ; S.Vh = S.V / 2;
  return 0;
}
int tmb_finalize(void) { if (S.up) { orc_finalize(); free(S.gauge); free(S.df); memset(&S, 0, sizeof(S)); } return 0; }
int tmb_is_initialized(void) { return S.up; }
const char *tmb_last_error(void) { return S.err; }
int tmb_volume_half(void) { return S.Vh; }
int tmb_set_hopping_phases(const double ka[8]) { NEED(); memcpy(S.ka, ka, sizeof(S.ka)); orc_set_hopping_phases(S.ka, S.mu); return 0; }
int tmb_set_mu(double gmu) { NEED(); S.mu = gmu; orc_set_hopping_phases(S.ka, S.mu); return 0; }
int tmb_set_nd(double a, double b, double c) { NEED(); orc_set_nd_params(a, b, c); return 0; }
int tmb_set_compression(int n) { NEED(); return (n == 18 || n == 12) ? 0 : -21; }
int tmb_set_mixcg(double e, int n) { (void)e; (void)n; NEED(); return 0; }
int tmb_set_mcg_delta(double d) { (void)d; NEED(); return 0; }
void orc_set_relative_precision_flag(int);
int tmb_set_relative_precision_flag(int f) { NEED(); orc_set_relative_precision_flag(f); return 0; }

void *tmb_field_alloc(void) { return S.up ? calloc(NF, sizeof(double)) : NULL; }
void *tmb_field32_alloc(void) { return S.up ? calloc(NF, sizeof(float)) : NULL; }
int tmb_field_free(void *f) { free(f); return 0; }
int tmb_field_zero(void *f) { NEED(); memset(f, 0, NF * sizeof(double)); return 0; }
void *tmb_host_alloc(size_t bytes) { return malloc(bytes); }
int tmb_host_free(void *p) { free(p); return 0; }
int tmb_host_register(void *p, size_t b) { (void)p; (void)b; return 0; }
int tmb_host_unregister(void *p) { (void)p; return 0; }
int tmb_sync(void) { return 0; }
#include <time.h>
static struct timespec t_start;
int tmb_timer_start(void) { clock_gettime(CLOCK_MONOTONIC, &t_start); return 0; }
int tmb_timer_stop(float *ms) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); *ms = (float)(1e3 * (t.tv_sec - t_start.tv_sec) + 1e-6 * (t.tv_nsec - t_start.tv_nsec)); return 0; }
int tmb_field_upload(void *f, const double *h) { NEED(); memcpy(f, h, NF * sizeof(double)); return 0; }
int tmb_field_download(double *h, const void *f) { NEED(); memcpy(h, f, NF * sizeof(double)); return 0; }
int tmb_field32_upload(void *f, const float *h) { NEED(); memcpy(f, h, NF * sizeof(float)); return 0; }
int tmb_field32_download(float *h, const void *f) { NEED(); memcpy(h, f, NF * sizeof(float)); return 0; }
int tmb_field_upload_lexic(void *e, void *o, const double *lex) { NEED(); orc_convert_lexic_to_eo(e, o, lex); return 0; }
int tmb_field_download_lexic(double *lex, const void *e, const void *o) { NEED(); orc_convert_eo_to_lexic(lex, e, o); return 0; }
int tmb_gauge_upload(const double *g) {
  NEED();
  if (!S.gauge) S.gauge = malloc((size_t)S.V * 72 * sizeof(double));
  memcpy(S.gauge, g, (size_t)S.V * 72 * sizeof(double));
  orc_set_gauge(S.gauge);
  return 0;
}
#define NEEDG() do { NEED(); if (!S.gauge) { snprintf(S.err, sizeof(S.err), "no gauge field on the device"); return -9; } } while (0)

int tmb_Hopping_Matrix(int ieo, void *l, const void *k) { NEEDG(); orc_Hopping_Matrix(ieo, l, k); return 0; }
int tmb_Hopping_Matrix_nocom(int ieo, void *l, const void *k) { return tmb_Hopping_Matrix(ieo, l, k); }
int tmb_comm_nranks(void) { return 1; }
int tmb_Hopping_Matrix_host(int ieo, double *l, const double *k, int mode, double cre, double cim) {
  NEEDG();
  if (mode == 0) orc_Hopping_Matrix(ieo, l, k); else if (mode == 1) orc_tm_times_Hopping_Matrix(ieo, l, k, cre, cim); else return -13;
  return 0;
}
int tmb_tm_times_Hopping_Matrix(int ieo, void *l, const void *k, double a, double b) { NEEDG(); orc_tm_times_Hopping_Matrix(ieo, l, k, a, b); return 0; }
int tmb_tm_sub_Hopping_Matrix(int ieo, void *l, const void *p, const void *k, double a, double b) { NEEDG(); orc_tm_sub_Hopping_Matrix(ieo, l, p, k, a, b); return 0; }
int tmb_H_eo_tm_inv_psi(void *l, const void *k, int ieo, double sign) { NEEDG(); orc_H_eo_tm_inv_psi(l, k, ieo, sign); return 0; }
int tmb_tm_sub_H_eo_gamma5(void *l, const void *p, const void *k, int ieo, double sign) { NEEDG(); orc_tm_sub_H_eo_gamma5(l, p, k, ieo, sign); return 0; }
/* the oracle's operators do not allow l == k; the device kernels do (site-local last step): go through a copy */
#define UNARY(name) int tmb_##name(void *l, const void *k) { NEEDG(); double *t = malloc(NF * sizeof(double)); memcpy(t, k, NF * sizeof(double)); orc_##name(l, t); free(t); return 0; }
UNARY(Qtm_pm_psi) UNARY(Qtm_plus_psi) UNARY(Qtm_minus_psi) UNARY(Mtm_plus_psi) UNARY(Mtm_minus_psi)
int tmb_M_full(void *en, void *on, const void *e, const void *o) { NEEDG(); orc_M_full(en, on, e, o); return 0; }
int tmb_Q_full(void *en, void *on, const void *e, const void *o) { NEEDG(); orc_Q_full(en, on, e, o); return 0; }
int tmb_D_psi_eo(void *en, void *on, const void *e, const void *o) { return tmb_M_full(en, on, e, o); }

/* l = (z on s0,s1 | conj z on s2,s3) k   and   l = [g5]((z | conj z) k - j) */
int tmb_diag(void *l, const void *k, double zre, double zim) {
  NEED();
  cplx *L = l; const cplx *K = k; const cplx z = zre + zim * I;
  for (int i = 0; i < S.Vh; i++) for (int c = 0; c < 12; c++) L[12 * i + c] = (c < 6 ? z : conj(z)) * K[12 * i + c];
  return 0;
}
int tmb_diag_sub(void *l, const void *k, const void *j, double zre, double zim, int g5) {
  NEED();
  cplx *L = l; const cplx *K = k, *J = j; const cplx z = zre + zim * I;
  for (int i = 0; i < S.Vh; i++) for (int c = 0; c < 12; c++) {
    const cplx v = (c < 6 ? z : conj(z)) * K[12 * i + c] - J[12 * i + c];
    L[12 * i + c] = (g5 && c >= 6) ? -v : v;
  }
  return 0;
}
int tmb_assign_mul_one_pm_imu_inv(void *l, const void *k, double sign) {
  const double nrm = 1. / (1. + S.mu * S.mu); return tmb_diag(l, k, nrm, (sign < 0. ? 1. : -1.) * nrm * S.mu);
}
int tmb_assign_mul_one_pm_imu(void *l, const void *k, double sign) { return tmb_diag(l, k, 1., (sign < 0. ? -1. : 1.) * S.mu); }
int tmb_mul_one_pm_imu_sub_mul_gamma5(void *l, const void *k, const void *j, double sign) { return tmb_diag_sub(l, k, j, 1., (sign < 0. ? -1. : 1.) * S.mu, 1); }
int tmb_mul_one_pm_imu_sub_mul(void *l, const void *k, const void *j, double sign) { return tmb_diag_sub(l, k, j, 1., (sign < 0. ? -1. : 1.) * S.mu, 0); }
int tmb_gamma5(void *l, const void *k) {
  NEED(); double *L = l; const double *K = k;
  for (int i = 0; i < S.Vh; i++) for (int c = 0; c < 24; c++) L[24 * i + c] = c < 12 ? K[24 * i + c] : -K[24 * i + c];
  return 0;
}

/* BLAS-1 */
int tmb_square_norm(const void *p, double *r) { NEED(); const double *P = p; double s = 0; for (size_t i = 0; i < NF; i++) s += P[i] * P[i]; *r = s; return 0; }
int tmb_scalar_prod_r(const void *a, const void *b, double *r) { NEED(); const double *A = a, *B = b; double s = 0; for (size_t i = 0; i < NF; i++) s += A[i] * B[i]; *r = s; return 0; }
int tmb_assign_add_mul_r(void *p, const void *q, double c) { NEED(); double *P = p; const double *Q = q; for (size_t i = 0; i < NF; i++) P[i] += c * Q[i]; return 0; }
int tmb_assign_mul_add_r(void *r, double c, const void *s) { NEED(); double *R = r; const double *Sv = s; for (size_t i = 0; i < NF; i++) R[i] = c * R[i] + Sv[i]; return 0; }
int tmb_assign_mul_add_r_and_square(void *r, double c, const void *s, double *res) { tmb_assign_mul_add_r(r, c, s); return tmb_square_norm(r, res); }
int tmb_diff(void *q, const void *r, const void *s) { NEED(); double *Q = q; const double *R = r, *Sv = s; for (size_t i = 0; i < NF; i++) Q[i] = R[i] - Sv[i]; return 0; }
int tmb_add(void *q, const void *r, const void *s) { NEED(); double *Q = q; const double *R = r, *Sv = s; for (size_t i = 0; i < NF; i++) Q[i] = R[i] + Sv[i]; return 0; }
int tmb_assign(void *r, const void *s) { NEED(); memmove(r, s, NF * sizeof(double)); return 0; }
int tmb_mul_r(void *r, double c, const void *s) { NEED(); double *R = r; const double *Sv = s; for (size_t i = 0; i < NF; i++) R[i] = c * Sv[i]; return 0; }

/* solvers */
int tmb_cg_her(void *P, const void *Q, int max_iter, double eps_sq, int rel_prec) { NEEDG(); return S.last_iters = orc_cg_her(P, Q, max_iter, eps_sq, rel_prec); }
int tmb_invert_eo(void *en, void *on, const void *e, const void *o, double prec, int max_iter, int rel_prec) {
  NEEDG(); return S.last_iters = orc_invert_eo_cg(en, on, e, o, prec, max_iter, rel_prec);
}
int tmb_solver_stats(int *it, double *err, double *sec) { if (it) *it = S.last_iters; if (err) *err = S.last_err; if (sec) *sec = 0.; return 0; }
/* the mixed solvers reach the same solution; the stand-in serves them with the double CG after zeroing the guess
 * like mixed_cg_her.c:108 does */
int tmb_mixed_cg_her(void *P, const void *Q, int m, double e, int r) { NEEDG(); memset(P, 0, NF * sizeof(double)); return tmb_cg_her(P, Q, m, e, r); }
int tmb_rg_mixed_cg_her(void *P, const void *Q, int m, double e, int r) { return tmb_cg_her(P, Q, m, e, r); }
int tmb_invert_eo_mixed(void *en, void *on, const void *e, const void *o, double p, int m, int r) { memset(on, 0, NF * sizeof(double)); return tmb_invert_eo(en, on, e, o, p, m, r); }
int tmb_invert_eo_rgmixed(void *en, void *on, const void *e, const void *o, double p, int m, int r) { return tmb_invert_eo_mixed(en, on, e, o, p, m, r); }
int tmb_solve_degenerate(void *P, const void *Q, int m, double e, int r, int solver) {
  if (solver != TMB_SOLVER_CG && solver != TMB_SOLVER_MIXEDCG && solver != TMB_SOLVER_RGMIXEDCG) { snprintf(S.err, sizeof(S.err), "solver %d not allowed", solver); return -32; }
  return tmb_cg_her(P, Q, m, e, r);
}

/* single precision */
static double *widen(const void *f32) { double *d = malloc(NF * sizeof(double)); const float *f = f32; for (size_t i = 0; i < NF; i++) d[i] = f[i]; return d; }
static void narrow(void *f32, const double *d) { float *f = f32; for (size_t i = 0; i < NF; i++) f[i] = (float)d[i]; }
int tmb_assign_to_32(void *f32, const void *f64) { NEED(); narrow(f32, f64); return 0; }
int tmb_assign_to_64(void *f64, const void *f32) { NEED(); double *d = f64; const float *f = f32; for (size_t i = 0; i < NF; i++) d[i] = f[i]; return 0; }
int tmb_Hopping_Matrix_32(int ieo, void *l, const void *k) { NEEDG(); double *a = widen(k), *b = malloc(NF * sizeof(double)); orc_Hopping_Matrix(ieo, b, a); narrow(l, b); free(a); free(b); return 0; }
int tmb_Qtm_pm_psi_32(void *l, const void *k) { NEEDG(); double *a = widen(k), *b = malloc(NF * sizeof(double)); orc_Qtm_pm_psi(b, a); narrow(l, b); free(a); free(b); return 0; }
int tmb_M_full_32(void *en, void *on, const void *e, const void *o, int g5) {
  NEEDG();
  double *a = widen(e), *b = widen(o), *c = malloc(NF * sizeof(double)), *d = malloc(NF * sizeof(double));
  if (g5) orc_Q_full(c, d, a, b); else orc_M_full(c, d, a, b);
  narrow(en, c); narrow(on, d); free(a); free(b); free(c); free(d);
  return 0;
}
int tmb_D_psi_eo_32(void *en, void *on, const void *e, const void *o) { return tmb_M_full_32(en, on, e, o, 0); }
int tmb_field32_upload_lexic(void *e, void *o, const float *lex) {
  NEED();
  double *l = malloc(2 * NF * sizeof(double)), *a = malloc(NF * sizeof(double)), *b = malloc(NF * sizeof(double));
  for (size_t i = 0; i < 2 * NF; i++) l[i] = lex[i];
  orc_convert_lexic_to_eo(a, b, l); narrow(e, a); narrow(o, b); free(l); free(a); free(b);
  return 0;
}
int tmb_field32_download_lexic(float *lex, const void *e, const void *o) {
  NEED();
  double *l = malloc(2 * NF * sizeof(double)), *a = widen(e), *b = widen(o);
  orc_convert_eo_to_lexic(l, a, b);
  for (size_t i = 0; i < 2 * NF; i++) lex[i] = (float)l[i];
  free(l); free(a); free(b);
  return 0;
}
int tmb_Qtm_pm_ndpsi_32(void *ls, void *lc, const void *ks, const void *kc) {
  NEEDG();
  double *a = widen(ks), *b = widen(kc), *c = malloc(NF * sizeof(double)), *d = malloc(NF * sizeof(double));
  orc_Qtm_pm_ndpsi(c, d, a, b); narrow(ls, c); narrow(lc, d); free(a); free(b); free(c); free(d);
  return 0;
}
/* the reliable-update solver reaches the CG's solution to the requested precision; the stand-in serves it with the oracle's CG */
int tmb_rg_mixed_cg_her_nd(void *Pup, void *Pdn, const void *Qup, const void *Qdn, int m, double e, int r) {
  NEEDG();
  memset(Pup, 0, NF * sizeof(double)); memset(Pdn, 0, NF * sizeof(double));
  return orc_cg_her_nd(Pup, Pdn, Qup, Qdn, m, e, r);
}
int tmb_blas32(int op, void *r, const void *s1, const void *s2, double c1d, double c2d) {
  NEED();
  float *R = r; const float *A = s1, *B = s2; const float c1 = (float)c1d, c2 = (float)c2d;
  for (size_t i = 0; i < NF; i++) {
    switch (op) {
      case 0: R[i] += c1 * A[i]; break;
      case 1: R[i] = c1 * R[i] + A[i]; break;
      case 2: R[i] = A[i] - B[i]; break;
      case 3: R[i] = c1 * A[i]; break;
      case 4: R[i] = c1 * R[i] + c2 * A[i]; break;
      case 5: R[i] = (i % 24 >= 12) ? -A[i] : A[i]; break;
      default: return -7;
    }
  }
  return 0;
}
int tmb_square_norm_32(const void *f, double *r) { NEED(); const float *F = f; double s = 0; for (size_t i = 0; i < NF; i++) s += (double)F[i] * F[i]; *r = s; return 0; }
int tmb_scalar_prod_r_32(const void *a, const void *b, double *r) { NEED(); const float *A = a, *B = b; double s = 0; for (size_t i = 0; i < NF; i++) s += (double)A[i] * B[i]; *r = s; return 0; }

/* non-degenerate doublet */
int tmb_M_ee_inv_ndpsi(void *ls, void *lc, const void *ks, const void *kc, double mu, double eps) { NEED(); orc_M_ee_inv_ndpsi(ls, lc, ks, kc, mu, eps); return 0; }
int tmb_M_oo_sub_g5_ndpsi(void *ls, void *lc, const void *ks, const void *kc, const void *js, const void *jc, double mu, double eps) {
  NEED(); orc_M_oo_sub_g5_ndpsi(ls, lc, ks, kc, js, jc, mu, eps); return 0;
}
#define ND(name) int tmb_##name(void *ls, void *lc, const void *ks, const void *kc) { \
  NEEDG(); double *a = malloc(NF * sizeof(double)), *b = malloc(NF * sizeof(double)); memcpy(a, ks, NF * sizeof(double)); memcpy(b, kc, NF * sizeof(double)); \
  orc_##name(ls, lc, a, b); free(a); free(b); return 0; }
ND(Qtm_ndpsi) ND(Qtm_dagger_ndpsi) ND(Qtm_pm_ndpsi)
int tmb_cg_her_nd(void *pu, void *pd, const void *qu, const void *qd, int m, double e, int r) { NEEDG(); return S.last_iters = orc_cg_her_nd(pu, pd, qu, qd, m, e, r); }
int tmb_invert_doublet_eo(void *ens, void *ons, void *enc, void *onc, const void *es, const void *os, const void *ec, const void *oc, double p, int m, int r) {
  NEEDG(); return S.last_iters = orc_invert_doublet_eo_cg(ens, ons, enc, onc, es, os, ec, oc, p, m, r);
}
/* RGMIXEDCG (rg_mixed_cg_her_nd) reaches the same solution as the CG to the requested precision; the stand-in serves both with the oracle's CG */
int tmb_invert_doublet_eo_solver(void *ens, void *ons, void *enc, void *onc, const void *es, const void *os, const void *ec, const void *oc, double p, int m, int r, int solver) {
  (void)solver; return tmb_invert_doublet_eo(ens, ons, enc, onc, es, os, ec, oc, p, m, r);
}

/* fermion force */
static int need_df(void) { if (!S.df) S.df = calloc((size_t)S.V * 32, sizeof(double)); return S.df ? 0 : -100; }
int tmb_derivative_zero(void) { NEED(); if (need_df()) return -100; memset(S.df, 0, (size_t)S.V * 32 * sizeof(double)); return 0; }
int tmb_derivative_upload(const double *h) { NEED(); if (need_df()) return -100; memcpy(S.df, h, (size_t)S.V * 32 * sizeof(double)); return 0; }
int tmb_derivative_download(double *h) { NEED(); if (need_df()) return -100; memcpy(h, S.df, (size_t)S.V * 32 * sizeof(double)); return 0; }
int tmb_deriv_Sb(int ieo, const void *l, const void *k, double factor) { NEEDG(); if (need_df()) return -100; orc_deriv_Sb(ieo, l, k, S.df, factor); return 0; }
int tmb_measure_plaquette(double *r) { NEEDG(); *r = orc_measure_plaquette(); return 0; }

/* chronological guess (solver/chrono_guess.c:43-172) on the stand-in's host fields: the ring-buffer bookkeeping,
 * Gram-Schmidt against the newest vector, G_ij = <v_i, A v_j>, b_i = <v_i, phi>, a plain Gaussian elimination for the
 * few-by-few system, trial = sum_i x_i v_i */
static cplx cdot(const double *a, const double *b) { const cplx *A = (const cplx *)a, *B = (const cplx *)b; cplx s = 0; for (int i = 0; i < S.Vh * 12; i++) s += conj(A[i]) * B[i]; return s; }
int tmb_chrono_add_solution(const void *trial, void *const *v, int *ia, int N, int *n) {
  NEED();
  if (N <= 0) return 0;
  int slot;
  if (*n < N) { ia[*n] = *n; *n += 1; slot = ia[*n - 1]; }
  else { for (int i = 1; i < N; i++) ia[i - 1] = ia[i]; ia[N - 1] = (N >= 2 ? (ia[N - 2] + 1) % N : 0); slot = ia[N - 1]; }
  double nrm; tmb_square_norm(trial, &nrm);
  return tmb_mul_r(v[slot], 1. / sqrt(nrm), trial);
}
int tmb_chrono_guess(void *trial, const void *phi, void *const *v, const int *ia, int N, int n, int op) {
  NEEDG();
  if (N <= 0 || n <= 0) { memset(trial, 0, NF * sizeof(double)); return 0; }
  if (n > 20 || op < 0 || op > 2) return -31;
  cplx G[20][20], b[20], x[20];
  for (int i = n - 2; i > -1; i--) { /* orthogonalise the older vectors against the newest (:105-117) */
    const cplx s = cdot(v[ia[n - 1]], v[ia[i]]);
    cplx *vi = (cplx *)v[ia[i]]; const cplx *vj = (const cplx *)v[ia[n - 1]];
    for (int k = 0; k < S.Vh * 12; k++) vi[k] -= s * vj[k];
  }
  double *t = malloc(NF * sizeof(double));
  for (int j = 0; j < n; j++) {
    if (op == TMB_OP_QTM_PM) orc_Qtm_pm_psi(t, v[ia[j]]); else if (op == TMB_OP_QTM_PLUS) orc_Qtm_plus_psi(t, v[ia[j]]); else orc_Qtm_minus_psi(t, v[ia[j]]);
    for (int i = 0; i <= j; i++) { G[i][j] = cdot(v[ia[i]], t); if (i != j) G[j][i] = conj(G[i][j]); }
    b[j] = cdot(v[ia[j]], phi);
  }
  free(t);
  for (int c = 0; c < n; c++) { /* Gaussian elimination with partial pivoting */
    int p = c; for (int r = c + 1; r < n; r++) if (cabs(G[r][c]) > cabs(G[p][c])) p = r;
    if (p != c) { for (int k = 0; k < n; k++) { cplx w = G[c][k]; G[c][k] = G[p][k]; G[p][k] = w; } cplx w = b[c]; b[c] = b[p]; b[p] = w; }
    for (int r = c + 1; r < n; r++) { const cplx f = G[r][c] / G[c][c]; for (int k = c; k < n; k++) G[r][k] -= f * G[c][k]; b[r] -= f * b[c]; }
  }
  for (int r = n - 1; r >= 0; r--) { cplx s = b[r]; for (int k = r + 1; k < n; k++) s -= G[r][k] * x[k]; x[r] = s / G[r][r]; }
  cplx *T = trial;
  for (int k = 0; k < S.Vh * 12; k++) { cplx s = 0; for (int i = 0; i < n; i++) s += x[i] * ((const cplx *)v[ia[i]])[k]; T[k] = s; }
  return 0;
}

/* DET / DETRATIO monomials: the oracle's restatement of det_monomial.c / detratio_monomial.c; the oracle sets kappa, mu and
 * the hopping phases per monomial from its own state (periodic or the theta given to orc_set_params), so the stand-in
 * re-installs the caller's phases after every call like mnl_backup_restore_globals does */
int orc_mnl_add(int, double, double, double, double, int, int, double, double, int);
void orc_mnl_clear(void);
double orc_mnl_heatbath(int, const double *);
void orc_mnl_derivative(int, double *);
double orc_mnl_acc(int);
void orc_mnl_info(int, double *, double *, int *, int *, int *);
void orc_set_relative_precision_flag(int);
void orc_set_theta(double, double, double, double);
extern double X0, X1, X2, X3; /* tmb_dropin.c: the reference's boundary angles */
int tmb_monomial_add(int type, double kappa, double mu, double kappa2, double mu2, int solver, int maxiter, double fp, double ap, int csg) {
  NEED(); return orc_mnl_add(type, kappa, mu, kappa2, mu2, solver, maxiter, fp, ap, csg);
}
int tmb_monomial_clear(void) { orc_mnl_clear(); return 0; }
int tmb_monomial_heatbath(int id, const void *g, double *e) {
  NEEDG(); orc_set_theta(X0, X1, X2, X3); const double v = orc_mnl_heatbath(id, g); if (e) *e = v; orc_set_hopping_phases(S.ka, S.mu); return 0;
}
int tmb_monomial_derivative(int id) {
  NEEDG(); if (need_df()) return -100; orc_set_theta(X0, X1, X2, X3); orc_mnl_derivative(id, S.df); orc_set_hopping_phases(S.ka, S.mu); return 0;
}
int tmb_monomial_acc(int id, double *dH) {
  NEEDG(); orc_set_theta(X0, X1, X2, X3); const double v = orc_mnl_acc(id); if (dH) *dH = v; orc_set_hopping_phases(S.ka, S.mu); return 0;
}
int tmb_monomial_info(int id, double *a, double *b, int *c, int *d, int *e) {
  NEED(); double e0, e1; int i0, i1, n; orc_mnl_info(id, &e0, &e1, &i0, &i1, &n);
  if (a) *a = e0;
  if (b) *b = e1;
  if (c) *c = i0;
  if (d) *d = i1;
  if (e) *e = n;
  return 0;
}

/* ref_api.c - thin driver over the UNMODIFIED tmLQCD reference objects.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled by oracle/ref_build/Makefile together with
 * reference sources read in place from /root/reference into oracle/_ref/*.so.
 * Used (a) to pin the C restatement in oracle/tmoracle.c, (b) to generate the
 * golden fixtures in tests/golden/, (c) as the "reference" CPU baseline of
 * bench.py.  The product library (tmlqcd_b200/csrc) never links or loads it.
 *
 * It does what benchmark.c:85-262 and tmlqcd_mpi_init (mpi_init.c:321-357) do by
 * hand for a single process: set T,LX,..,VOLUME, allocate, geometry(), boundary().
 * Every ref_* entry point takes plain double* buffers in the reference's own AoS
 * layouts (spinor = 24 doubles, su3 = 18 doubles) and forwards to the reference
 * function of the same name.
 */
#define INIT_GLOBALS
#include "config.h"
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <complex.h>
#ifdef TM_USE_OMP
#include <omp.h>
#endif
#include "global.h"
#include "su3.h"
#include "geometry_eo.h"
#include "boundary.h"
extern double X0, X1, X2, X3; /* read_input.h:65, defined in boundary.c:37 */
#include "start.h"
#include "ranlxd.h"
#include "gamma.h"
#include "gettime.h"
#include "update_backward_gauge.h"
#include "init/init_gauge_field.h"
#include "init/init_geometry_indices.h"
#include "init/init_spinor_field.h"
#include "init/init_dirac_halfspinor.h"
#include "init/init_openmp.h"
#include "operator/Hopping_Matrix.h"
#include "operator/Hopping_Matrix_nocom.h"
#include "operator/tm_times_Hopping_Matrix.h"
#include "operator/tm_sub_Hopping_Matrix.h"
#include "operator/tm_operators.h"
#include "operator/tm_operators_nd.h"
#include "operator/D_psi.h"
#include "linalg_eo.h"
#include "solver/matrix_mult_typedef.h"
#include "solver/matrix_mult_typedef_nd.h"
#include "solver/cg_her.h"
#include "solver/cg_her_nd.h"
#include "solver/solver_params.h"
#include "solver/mixed_cg_her.h"
#include "invert_eo.h"
#include "invert_doublet_eo.h"
#include "solver/solver_types.h"
#include "operator/Hopping_Matrix_32.h"
#include "operator/tm_operators_32.h"

/* phmc.c is not compiled (drags in the whole PHMC); tm_operators_nd.c only needs these
 * scalars from it (phmc.h). */
double phmc_invmaxev = 1.0;
double phmc_cheb_evmin, phmc_cheb_evmax, phmc_Cpol;
int phmc_dop_n_cheby;
double *phmc_dop_cheby_coef;
int phmc_ptilde_n_cheby;
double *phmc_ptilde_cheby_coef;
_Complex double *phmc_root;
double phmc_stilde_low, phmc_stilde_max;
int phmc_exact_poly;

static int ref_initialised = 0;
#define NSF 14 /* g_spinor_field slots: user 0..3, DUM_DERI 4..7, DUM_MATRIX 8..13 */

int ref_init(int t, int lx, int ly, int lz, int nthreads) {
  if (ref_initialised) {
    if (t == T && lx == LX && ly == LY && lz == LZ) return 0;
    fprintf(stderr, "ref_init: already initialised with another lattice\n");
    return -1;
  }
  T = t; L = lx; LX = lx; LY = ly; LZ = lz; T_global = t;
  N_PROC_T = N_PROC_X = N_PROC_Y = N_PROC_Z = 1;
  g_nproc = g_nproc_t = g_nproc_x = g_nproc_y = g_nproc_z = 1;
  g_proc_id = 0; g_cart_id = 0; g_stdio_proc = 0;
  for (int i = 0; i < 4; i++) g_proc_coords[i] = 0;
  VOLUME = T * LX * LY * LZ; RAND = 0; EDGES = 0; VOLUMEPLUSRAND = VOLUME;
  SPACEVOLUME = LX * LY * LZ; SPACERAND = 0;
  g_dbw2rand = 0; g_debug_level = 0; g_sloppy_precision_flag = 0; g_sloppy_precision = 0;
  g_rgi_C1 = 0.; g_c_sw = 0.; g_use_clover_flag = 0; lowmem_flag = 0;
  DUM_DERI = 4; DUM_MATRIX = 8; NO_OF_SPINORFIELDS = NSF;
#ifdef TM_USE_OMP
  omp_num_threads = nthreads > 0 ? nthreads : 1;
  init_openmp();
#else
  (void)nthreads;
#endif
  if (init_gauge_field(VOLUMEPLUSRAND, 1) != 0) return -2;
  if (init_geometry_indices(VOLUMEPLUSRAND) != 0) return -3;
  /* full-volume sized slots so that D_psi (lexicographic, V sites) can use them too */
  if (init_spinor_field(VOLUMEPLUSRAND, NSF) != 0) return -4;
  geometry();
  g_kappa = 0.16; g_mu = 0.0; X0 = X1 = X2 = X3 = 0.;
  boundary(g_kappa);
#ifdef _USE_HALFSPINOR
  if (init_dirac_halfspinor() != 0) return -5;
#endif
  g_update_gauge_copy = 1;
  ref_initialised = 1;
  return 0;
}

int ref_num_threads(void) {
#ifdef TM_USE_OMP
  return omp_num_threads;
#else
  return 1;
#endif
}

int ref_is_halfspinor(void) {
#ifdef _USE_HALFSPINOR
  return 1;
#else
  return 0;
#endif
}

/* g_mu is 2*kappa*mu (invert_eo.c:255, operator.c:337) */
void ref_set_params(double kappa, double gmu, double x0, double x1, double x2, double x3) {
  g_kappa = kappa; g_mu = gmu; X0 = x0; X1 = x1; X2 = x2; X3 = x3;
  boundary(g_kappa);
}
void ref_set_nd_params(double mubar, double epsbar, double invmaxev) {
  g_mubar = mubar; g_epsbar = epsbar; phmc_invmaxev = invmaxev;
}
void ref_set_debug_level(int l) { g_debug_level = l; }

void ref_start_ranlux(int level, int seed) { start_ranlux(level, seed); }
/* benchmark.c:247-248 */
void ref_random_gauge(int seed) {
  start_ranlux(1, seed);
  random_gauge_field(1, g_gauge_field);
  g_update_gauge_copy = 1;
}
void ref_get_gauge(double *out) { memcpy(out, g_gauge_field[0], (size_t)VOLUME * 4 * sizeof(su3)); }
void ref_set_gauge(const double *in) {
  memcpy(g_gauge_field[0], in, (size_t)VOLUME * 4 * sizeof(su3));
  g_update_gauge_copy = 1;
}
/* benchmark.c:259 */
void ref_random_spinor_eo(double *out) { random_spinor_field_eo((spinor *)out, 1, RN_GAUSS); }
void ref_random_spinor_lexic(double *out) { random_spinor_field_lexic((spinor *)out, 1, RN_GAUSS); }

void ref_get_eo2lexic(int *out) { memcpy(out, g_eo2lexic, (size_t)VOLUME * sizeof(int)); }
void ref_get_lexic2eosub(int *out) { memcpy(out, g_lexic2eosub, (size_t)VOLUME * sizeof(int)); }
void ref_get_hi(int *out) { memcpy(out, g_hi, (size_t)VOLUME * 16 * sizeof(int)); }
void ref_get_iup(int *out) { for (int i = 0; i < VOLUME; i++) for (int m = 0; m < 4; m++) out[4 * i + m] = g_iup[i][m]; }
void ref_get_idn(int *out) { for (int i = 0; i < VOLUME; i++) for (int m = 0; m < 4; m++) out[4 * i + m] = g_idn[i][m]; }
void ref_get_ka(double *out) {
  _Complex double k[4] = {ka0, ka1, ka2, ka3};
  for (int i = 0; i < 4; i++) { out[2 * i] = creal(k[i]); out[2 * i + 1] = cimag(k[i]); }
}

/* ---- operators (a6..a19) ---- */
void ref_Hopping_Matrix(int ieo, double *l, double *k) { Hopping_Matrix(ieo, (spinor *)l, (spinor *)k); }
void ref_tm_times_Hopping_Matrix(int ieo, double *l, double *k, double cre, double cim) {
  tm_times_Hopping_Matrix(ieo, (spinor *)l, (spinor *)k, cre + cim * I);
}
void ref_tm_sub_Hopping_Matrix(int ieo, double *l, double *p, double *k, double cre, double cim) {
  tm_sub_Hopping_Matrix(ieo, (spinor *)l, (spinor *)p, (spinor *)k, cre + cim * I);
}
void ref_H_eo_tm_inv_psi(double *l, double *k, int ieo, double sign) { H_eo_tm_inv_psi((spinor *)l, (spinor *)k, ieo, sign); }
void ref_Qtm_pm_psi(double *l, double *k) { Qtm_pm_psi((spinor *)l, (spinor *)k); }
void ref_Qtm_plus_psi(double *l, double *k) { Qtm_plus_psi((spinor *)l, (spinor *)k); }
void ref_Qtm_minus_psi(double *l, double *k) { Qtm_minus_psi((spinor *)l, (spinor *)k); }
void ref_Mtm_plus_psi(double *l, double *k) { Mtm_plus_psi((spinor *)l, (spinor *)k); }
void ref_Mtm_minus_psi(double *l, double *k) { Mtm_minus_psi((spinor *)l, (spinor *)k); }
void ref_M_full(double *en, double *on, double *e, double *o) { M_full((spinor *)en, (spinor *)on, (spinor *)e, (spinor *)o); }
void ref_Q_full(double *en, double *on, double *e, double *o) { Q_full((spinor *)en, (spinor *)on, (spinor *)e, (spinor *)o); }
void ref_D_psi(double *p, double *q) { D_psi((spinor *)p, (spinor *)q); }
void ref_Q_pm_psi(double *l, double *k) { Q_pm_psi((spinor *)l, (spinor *)k); }
void ref_gamma5(double *l, double *k, int n) { gamma5((spinor *)l, (spinor *)k, n); }
void ref_mul_one_pm_imu_inv(double *l, double sign, int n) { mul_one_pm_imu_inv((spinor *)l, sign, n); }
void ref_assign_mul_one_pm_imu_inv(double *l, double *k, double sign, int n) { assign_mul_one_pm_imu_inv((spinor *)l, (spinor *)k, sign, n); }
void ref_assign_mul_one_pm_imu(double *l, double *k, double sign, int n) { assign_mul_one_pm_imu((spinor *)l, (spinor *)k, sign, n); }
void ref_mul_one_pm_imu_sub_mul_gamma5(double *l, double *k, double *j, double sign) {
  mul_one_pm_imu_sub_mul_gamma5((spinor *)l, (spinor *)k, (spinor *)j, sign);
}
void ref_convert_eo_to_lexic(double *p, double *s, double *r) { convert_eo_to_lexic((spinor *)p, (spinor *)s, (spinor *)r); }
void ref_convert_lexic_to_eo(double *s, double *r, double *p) { convert_lexic_to_eo((spinor *)s, (spinor *)r, (spinor *)p); }

/* ---- BLAS-1 (a20..a25) ---- */
double ref_square_norm(double *p, int n) { return square_norm((spinor *)p, n, 1); }
double ref_scalar_prod_r(double *s, double *r, int n) { return scalar_prod_r((spinor *)s, (spinor *)r, n, 1); }
void ref_assign_add_mul_r(double *p, double *q, double c, int n) { assign_add_mul_r((spinor *)p, (spinor *)q, c, n); }
void ref_assign_mul_add_r(double *r, double c, double *s, int n) { assign_mul_add_r((spinor *)r, c, (spinor *)s, n); }
double ref_assign_mul_add_r_and_square(double *r, double c, double *s, int n) {
  return assign_mul_add_r_and_square((spinor *)r, c, (spinor *)s, n, 1);
}
void ref_diff(double *q, double *r, double *s, int n) { diff((spinor *)q, (spinor *)r, (spinor *)s, n); }
void ref_assign(double *r, double *s, int n) { assign((spinor *)r, (spinor *)s, n); }
void ref_mul_r(double *r, double c, double *s, int n) { mul_r((spinor *)r, c, (spinor *)s, n); }

/* ---- solvers ---- */
int ref_cg_her(double *p, double *q, int max_iter, double eps_sq, int rel_prec) {
  return cg_her((spinor *)p, (spinor *)q, max_iter, eps_sq, rel_prec, VOLUME / 2, &Qtm_pm_psi);
}

/* invert_eo itself, compiled unmodified (invert_eo.c:83-561; lime.h comes from stubs/, the solvers it can dispatch to
 * outside the scoped path are abort stubs in ref_shim.c).  even_odd_flag = 1, no extra masses, NO_EXT_INV. */
int ref_invert_eo(double *even_new, double *odd_new, double *even, double *odd,
                  double precision, int max_iter, int rel_prec, int solver_flag) {
  solver_params_t sp;
  memset(&sp, 0, sizeof(sp));
  return invert_eo((spinor *)even_new, (spinor *)odd_new, (spinor *)even, (spinor *)odd, precision, max_iter, solver_flag,
                   rel_prec, 0, 1, 0, NULL, sp, 0, NO_EXT_INV, 0, NO_COMPRESSION);
}
/* the same with every flag of the scoped branches open: solver_flag CG / MIXEDCG / RGMIXEDCG (delta = solver_params.mcg_delta
 * of the reliable updates), even_odd_flag 0 = the full-lattice branch (invert_eo.c:364-558: cg_her on Q_pm_psi, VOLUME sites) */
int ref_invert_eo_flags(double *even_new, double *odd_new, double *even, double *odd, double precision, int max_iter,
                        int rel_prec, int solver_flag, int even_odd_flag, double delta) {
  solver_params_t sp;
  memset(&sp, 0, sizeof(sp));
  sp.mcg_delta = (float)delta;
  return invert_eo((spinor *)even_new, (spinor *)odd_new, (spinor *)even, (spinor *)odd, precision, max_iter, solver_flag,
                   rel_prec, 0, even_odd_flag, 0, NULL, sp, 0, NO_EXT_INV, 0, NO_COMPRESSION);
}
int ref_invert_eo_cg(double *even_new, double *odd_new, double *even, double *odd,
                     double precision, int max_iter, int rel_prec) {
  return ref_invert_eo(even_new, odd_new, even, odd, precision, max_iter, rel_prec, CG);
}

/* ---- ND doublet (a29, a30) ---- */
void ref_Qtm_pm_ndpsi(double *ls, double *lc, double *ks, double *kc) {
  Qtm_pm_ndpsi((spinor *)ls, (spinor *)lc, (spinor *)ks, (spinor *)kc);
}
void ref_Qtm_ndpsi(double *ls, double *lc, double *ks, double *kc) {
  Qtm_ndpsi((spinor *)ls, (spinor *)lc, (spinor *)ks, (spinor *)kc);
}
void ref_Qtm_dagger_ndpsi(double *ls, double *lc, double *ks, double *kc) {
  Qtm_dagger_ndpsi((spinor *)ls, (spinor *)lc, (spinor *)ks, (spinor *)kc);
}
void ref_M_ee_inv_ndpsi(double *ls, double *lc, double *ks, double *kc, double mu, double eps) {
  M_ee_inv_ndpsi((spinor *)ls, (spinor *)lc, (spinor *)ks, (spinor *)kc, mu, eps);
}
int ref_cg_her_nd(double *ps, double *pc, double *qs, double *qc, int max_iter, double eps_sq, int rel_prec) {
  return cg_her_nd((spinor *)ps, (spinor *)pc, (spinor *)qs, (spinor *)qc, max_iter, eps_sq, rel_prec,
                   VOLUME / 2, &Qtm_pm_ndpsi);
}
/* invert_doublet_eo itself, compiled unmodified (invert_doublet_eo.c:68-187): solver_flag CG or RGMIXEDCG (mcg_delta = delta) */
int ref_invert_doublet_eo(double *ens, double *ons, double *enc, double *onc,
                          double *es, double *os, double *ec, double *oc,
                          double precision, int max_iter, int rel_prec, int solver_flag, double delta) {
  solver_params_t sp;
  memset(&sp, 0, sizeof(sp));
  sp.mcg_delta = (float)delta;
  return invert_doublet_eo((spinor *)ens, (spinor *)ons, (spinor *)enc, (spinor *)onc, (spinor *)es, (spinor *)os,
                           (spinor *)ec, (spinor *)oc, precision, max_iter, solver_flag, rel_prec, sp, NO_EXT_INV, 0,
                           NO_COMPRESSION);
}
int ref_invert_doublet_eo_cg(double *ens, double *ons, double *enc, double *onc,
                             double *es, double *os, double *ec, double *oc,
                             double precision, int max_iter, int rel_prec) {
  return ref_invert_doublet_eo(ens, ons, enc, onc, es, os, ec, oc, precision, max_iter, rel_prec, CG, 0.);
}

/* ---- timing, the benchmark.c:262-327 recipe: nreps x { H(0, sf1, sf0); H(1, sf2, sf1) } ---- */
double ref_bench_hopping(int nreps) {
  double t1, t2;
  random_spinor_field_eo(g_spinor_field[0], 1, RN_GAUSS);
  Hopping_Matrix(0, g_spinor_field[1], g_spinor_field[0]); /* warm-up, builds the gauge copy */
  Hopping_Matrix(1, g_spinor_field[2], g_spinor_field[1]);
  t1 = gettime();
  for (int j = 0; j < nreps; j++) {
    Hopping_Matrix(0, g_spinor_field[1], g_spinor_field[0]);
    Hopping_Matrix(1, g_spinor_field[2], g_spinor_field[1]);
  }
  t2 = gettime();
  return t2 - t1;
}
double ref_bench_D_psi(int nreps) {
  double t1, t2;
  random_spinor_field_lexic(g_spinor_field[0], 1, RN_GAUSS);
  D_psi(g_spinor_field[1], g_spinor_field[0]);
  t1 = gettime();
  for (int j = 0; j < nreps; j++) {
    D_psi(g_spinor_field[1], g_spinor_field[0]);
    D_psi(g_spinor_field[0], g_spinor_field[1]);
  }
  t2 = gettime();
  return t2 - t1;
}
/* nreps applications of Qtm_pm_psi (the CG matrix) */
double ref_bench_Qtm_pm(int nreps) {
  double t1, t2;
  random_spinor_field_eo(g_spinor_field[0], 1, RN_GAUSS);
  Qtm_pm_psi(g_spinor_field[1], g_spinor_field[0]);
  t1 = gettime();
  for (int j = 0; j < nreps; j++) Qtm_pm_psi(g_spinor_field[1], g_spinor_field[0]);
  t2 = gettime();
  return t2 - t1;
}
double ref_gettime(void) { return gettime(); }

/* ---- single precision operator + mixed CG; half-spinor build only:
 *      Hopping_Matrix_32 exits with "only implemented with HALFSPINOR" otherwise
 *      (operator/Hopping_Matrix_32.c:112-114) ---- */
double mixcg_innereps = 5.0e-5;      /* default_input_values.h:193, normally set by read_input.l:2912 */
int mixcg_maxinnersolverit = 5000;   /* default_input_values.h:194 */
static int ref32_up = 0;
int ref_init32(void) {
#ifdef _USE_HALFSPINOR
  if (ref32_up) return 0;
  if (init_gauge_field_32(VOLUMEPLUSRAND, 1) != 0) return -1;
  if (init_spinor_field_32(VOLUMEPLUSRAND / 2, 6) != 0) return -2;
  if (init_dirac_halfspinor32() != 0) return -3;
  ref32_up = 1;
  return 0;
#else
  return -9;
#endif
}
/* lib_wrapper.c:232: convert_32_gauge_field after every gauge change */
void ref_update_gauge32(void) {
  convert_32_gauge_field(g_gauge_field_32, g_gauge_field, VOLUMEPLUSRAND);
  g_update_gauge_copy_32 = 1;
}
void ref_set_mixcg(double innereps, int maxinner) { mixcg_innereps = innereps; mixcg_maxinnersolverit = maxinner; }
void ref_Hopping_Matrix_32(int ieo, float *l, float *k) { Hopping_Matrix_32(ieo, (spinor32 *)l, (spinor32 *)k); }
void ref_Qtm_pm_psi_32(float *l, float *k) { Qtm_pm_psi_32((spinor32 *)l, (spinor32 *)k); }
/* operator/D_psi.h:28 and its caller operator/tm_operators_32.c:141: lexicographic spinor32 fields of VOLUME sites
 * (Q_pm_psi_32 uses g_spinor_field32[0] as a VOLUME-site scratch: fields 0 and 1 of the contiguous slab) */
void ref_D_psi_32(float *p, float *q) { D_psi_32((spinor32 *)p, (spinor32 *)q); }
void ref_Q_pm_psi_32(float *l, float *k) { Q_pm_psi_32((spinor32 *)l, (spinor32 *)k); }
int ref_mixed_cg_her(double *p, double *q, int max_iter, double eps_sq, int rel_prec) {
  solver_params_t sp;
  memset(&sp, 0, sizeof(sp));
  return mixed_cg_her((spinor *)p, (spinor *)q, sp, max_iter, eps_sq, rel_prec, VOLUME / 2, &Qtm_pm_psi, &Qtm_pm_psi_32);
}

/* ---- HMC pieces: deriv_Sb, chronological guess, solve_degenerate and the
 *      DET / DETRATIO monomials, all the unmodified reference (deriv_Sb.c, solver/chrono_guess.c,
 *      solver/monomial_solve.c, monomial/{monomial,det_monomial,detratio_monomial}.c) ---- */
#include "hamiltonian_field.h"
#include "deriv_Sb.h"
#include "init/init_moment_field.h"
#include "monomial/monomial.h"
#include "solver/chrono_guess.h"
#include "solver/monomial_solve.h"

extern int even_odd_flag; /* read_input.h, defined in ref_shim.c */
static hamiltonian_field_t ref_hf;
static int ref_hmc_up = 0;
int ref_hmc_init(void) {
  if (ref_hmc_up) return 0;
  if (init_moment_field(VOLUME, VOLUMEPLUSRAND) != 0) return -1;
  ref_hf.gaugefield = g_gauge_field; ref_hf.momenta = moment; ref_hf.derivative = df0;
  ref_hf.update_gauge_copy = g_update_gauge_copy; ref_hf.traj_counter = 0;
  g_relative_precision_flag = 0;
  ref_hmc_up = 1;
  return 0;
}
void ref_set_relative_precision_flag(int f) { g_relative_precision_flag = f; }
/* df: [VOLUME][4][8] doubles = hf->derivative (su3adj.h:25-27), accumulated in place */
void ref_deriv_Sb(int ieo, double *l, double *k, double *df, double factor) {
  memcpy(df0[0], df, (size_t)VOLUME * 4 * sizeof(su3adj));
  deriv_Sb(ieo, (spinor *)l, (spinor *)k, &ref_hf, factor);
  memcpy(df, df0[0], (size_t)VOLUME * 4 * sizeof(su3adj));
}
/* what read_input.l does for a BeginMonomial DET / DETRATIO block, then init_monomials (hmc_tm.c:300) */
int ref_mnl_add(int type, double kappa, double mu, double kappa2, double mu2, int solver, int maxiter,
                double forceprec, double accprec, int csg_N) {
  int id = add_monomial(type) - 1;
  monomial *m = &monomial_list[id];
  m->type = type; /* read_input.l sets it after add_monomial() */
  m->kappa = kappa; m->mu = mu; m->kappa2 = kappa2; m->mu2 = mu2; m->solver = solver; m->maxiter = maxiter;
  m->forceprec = forceprec; m->accprec = accprec; m->csg_N = csg_N; m->csg_N2 = 0; m->even_odd_flag = 1;
  m->solver_params.mcg_delta = (float)mixcg_innereps;
  return id;
}
int ref_mnl_init(void) {
  if (init_monomials(VOLUMEPLUSRAND / 2, even_odd_flag) != 0) return -1;
  if (init_csg_field(VOLUMEPLUSRAND / 2) != 0) return -2;
  return 0;
}
void ref_mnl_heatbath(int id) { monomial_list[id].hbfunction(id, &ref_hf); }
double ref_mnl_acc(int id) { return monomial_list[id].accfunction(id, &ref_hf); }
void ref_mnl_derivative(int id, double *df) {
  memcpy(df0[0], df, (size_t)VOLUME * 4 * sizeof(su3adj));
  monomial_list[id].derivativefunction(id, &ref_hf);
  memcpy(df, df0[0], (size_t)VOLUME * 4 * sizeof(su3adj));
}
void ref_mnl_get_pf(int id, double *out) { memcpy(out, monomial_list[id].pf, (size_t)(VOLUME / 2) * sizeof(spinor)); }
void ref_mnl_set_pf(int id, const double *in) { memcpy(monomial_list[id].pf, in, (size_t)(VOLUME / 2) * sizeof(spinor)); }
void ref_mnl_info(int id, double *energy0, double *energy1, int *iter0, int *iter1, int *csg_n) {
  monomial *m = &monomial_list[id];
  *energy0 = m->energy0; *energy1 = m->energy1; *iter0 = m->iter0; *iter1 = m->iter1; *csg_n = m->csg_n;
}
/* direct access to the chronological guess for unit parity: history of `n` fields, newest last */
int ref_solve_degenerate(double *p, double *q, int max_iter, double eps_sq, int rel_prec, int solver) {
  solver_params_t sp;
  memset(&sp, 0, sizeof(sp));
  sp.mcg_delta = (float)mixcg_innereps;
  return solve_degenerate((spinor *)p, (spinor *)q, sp, max_iter, eps_sq, rel_prec, VOLUME / 2, &Qtm_pm_psi, solver);
}
/* timing of one derivative call of monomial id (bench.py hmc leg) */
double ref_bench_mnl_derivative(int id, int nreps) {
  double t1 = gettime();
  for (int j = 0; j < nreps; j++) monomial_list[id].derivativefunction(id, &ref_hf);
  return gettime() - t1;
}

/* ---- remaining members of the operator families ---- */
#define REF_UNARY(name) void ref_##name(double *l, double *k) { name((spinor *)l, (spinor *)k); }
REF_UNARY(Qtm_plus_sym_psi) REF_UNARY(Qtm_minus_sym_psi) REF_UNARY(Mtm_plus_sym_psi) REF_UNARY(Mtm_minus_sym_psi)
REF_UNARY(Mtm_plus_sym_dagg_psi) REF_UNARY(Qtm_pm_sym_psi) REF_UNARY(M_minus_psi) REF_UNARY(D_dagg_psi)
REF_UNARY(Q_plus_psi) REF_UNARY(Q_minus_psi)
void ref_Mee_psi(double *l, double *k, double mu) { Mee_psi((spinor *)l, (spinor *)k, mu); }
void ref_Mee_inv_psi(double *l, double *k, double mu) { Mee_inv_psi((spinor *)l, (spinor *)k, mu); }
void ref_mul_one_sub_mul_gamma5(double *l, double *k, double *j) { mul_one_sub_mul_gamma5((spinor *)l, (spinor *)k, (spinor *)j); }
void ref_mul_one_pm_imu_sub_mul(double *l, double *k, double *j, double sign, int n) {
  mul_one_pm_imu_sub_mul((spinor *)l, (spinor *)k, (spinor *)j, sign, n);
}
void ref_M_minus_1_timesC(double *en, double *on, double *e, double *o) {
  M_minus_1_timesC((spinor *)en, (spinor *)on, (spinor *)e, (spinor *)o);
}
void ref_H_eo_tm_ndpsi(double *ls, double *lc, double *ks, double *kc, int ieo) {
  H_eo_tm_ndpsi((spinor *)ls, (spinor *)lc, (spinor *)ks, (spinor *)kc, ieo);
}
void M_oo_sub_g5_ndpsi(spinor *const, spinor *const, spinor *const, spinor *const, spinor *const, spinor *const, const double, const double);
void mul_one_pm_iconst(spinor *const, spinor *const, const double, const int);
void ref_M_oo_sub_g5_ndpsi(double *ls, double *lc, double *ks, double *kc, double *js, double *jc, double mu, double eps) {
  M_oo_sub_g5_ndpsi((spinor *)ls, (spinor *)lc, (spinor *)ks, (spinor *)kc, (spinor *)js, (spinor *)jc, mu, eps);
}
void ref_mul_one_pm_iconst(double *l, double *k, double mu, int sign) { mul_one_pm_iconst((spinor *)l, (spinor *)k, mu, sign); }
#include "solver/rg_mixed_cg_her.h"
int ref_rg_mixed_cg_her(double *p, double *q, int max_iter, double eps_sq, int rel_prec, double delta) {
  solver_params_t sp;
  memset(&sp, 0, sizeof(sp));
  sp.mcg_delta = (float)delta;
  return rg_mixed_cg_her((spinor *)p, (spinor *)q, sp, max_iter, eps_sq, rel_prec, VOLUME / 2, &Qtm_pm_psi, &Qtm_pm_psi_32);
}

/* ---- gauge / propagator files: the reference's unmodified io/ code over the stand-in
 *      LIME layer (stubs/lime.h) ---- */
#include "io/gauge.h"
#include "io/spinor.h"
#include "io/params.h"
#include "io/utils.h"
#include "solver/solver_types.h"
extern int gauge_precision_read_flag; /* read_input.h:69, defined in ref_shim.c */
int ref_write_gauge(const char *filename, int prec, double plaq, int counter) {
  paramsXlfInfo *xlf = construct_paramsXlfInfo(plaq, counter);
  int st = write_gauge_field((char *)filename, prec, xlf);
  free(xlf);
  return st;
}
int ref_read_gauge(const char *filename, int prec) {
  gauge_precision_read_flag = prec;
  int st = read_gauge_field((char *)filename, g_gauge_field);
  g_update_gauge_copy = 1;
  return st;
}
/* what op_write_prop (operator.c:532-605) writes for one flavour with PropInfo.format == 0 */
int ref_write_propagator(const char *filename, double *even, double *odd, int prec, double epssq, int iter) {
  WRITER *writer = NULL;
  spinor *s = (spinor *)even, *r = (spinor *)odd;
  construct_writer(&writer, (char *)filename, 0);
  write_propagator_type(writer, 0);
  paramsInverterInfo *info = construct_paramsInverterInfo(epssq, iter, CG, 1);
  write_spinor_info(writer, 0, info, 0);
  free(info);
  paramsPropagatorFormat *fmt = construct_paramsPropagatorFormat(prec, 1);
  write_propagator_format(writer, fmt);
  free(fmt);
  int st = write_spinor(writer, &s, &r, 1, prec);
  destruct_writer(writer);
  return st;
}
int ref_read_spinor(double *even, double *odd, const char *filename, int position) {
  return read_spinor((spinor *)even, (spinor *)odd, (char *)filename, position);
}

/* ---- plaquette (measure_gauge_action.c:46): what tmLQCD_read_gauge prints after reading a configuration
 *      (wrapper/lib_wrapper.c:232-235) ---- */
#include "measure_gauge_action.h"
double ref_measure_plaquette(void) { return measure_plaquette((const su3 **)g_gauge_field); }

/* ---- single-precision BLAS-1 of the mixed solvers: linalg/..._32.c, operator/tm_operators_32.c:130 ---- */
#include "linalg/square_norm_32.h"
#include "linalg/scalar_prod_r_32.h"
#include "linalg/assign_add_mul_r_32.h"
#include "linalg/assign_mul_add_r_32.h"
#include "linalg/diff_32.h"
#include "linalg/mul_r_32.h"
#include "linalg/assign_mul_add_mul_r_32.h"
#include "operator/tm_operators_32.h"
float ref_square_norm_32(float *p, int n) { return square_norm_32((spinor32 *)p, n, 0); }
float ref_scalar_prod_r_32(float *s, float *r, int n) { return scalar_prod_r_32((spinor32 *)s, (spinor32 *)r, n, 0); }
void ref_assign_add_mul_r_32(float *r, float *s, float c, int n) { assign_add_mul_r_32((spinor32 *)r, (spinor32 *)s, c, n); }
void ref_assign_mul_add_r_32(float *r, float c, float *s, int n) { assign_mul_add_r_32((spinor32 *)r, c, (spinor32 *)s, n); }
void ref_diff_32(float *q, float *r, float *s, int n) { diff_32((spinor32 *)q, (spinor32 *)r, (spinor32 *)s, n); }
void ref_mul_r_32(float *r, float c, float *s, int n) { mul_r_32((spinor32 *)r, c, (spinor32 *)s, n); }
void ref_assign_mul_add_mul_r_32(float *r, float *s, float c1, float c2, int n) { assign_mul_add_mul_r_32((spinor32 *)r, (spinor32 *)s, c1, c2, n); }
void ref_gamma5_32(float *l, float *k, int n) { gamma5_32((spinor32 *)l, (spinor32 *)k, n); }

/* ---- the RGMIXEDCG branch of invert_doublet_eo (invert_doublet_eo.c:145-150): groundwork for the next round ---- */
#include "operator/tm_operators_nd_32.h"
#include "solver/rg_mixed_cg_her_nd.h"
void ref_Qtm_pm_ndpsi_32(float *ls, float *lc, float *ks, float *kc) {
  Qtm_pm_ndpsi_32((spinor32 *)ls, (spinor32 *)lc, (spinor32 *)ks, (spinor32 *)kc);
}
int ref_rg_mixed_cg_her_nd(double *pu, double *pd, double *qu, double *qd, int max_iter, double eps_sq, int rel_prec, double delta) {
  solver_params_t sp;
  memset(&sp, 0, sizeof(sp));
  sp.mcg_delta = (float)delta;
  return rg_mixed_cg_her_nd((spinor *)pu, (spinor *)pd, (spinor *)qu, (spinor *)qd, sp, max_iter, eps_sq, rel_prec, VOLUME / 2,
                            &Qtm_pm_ndpsi, &Qtm_pm_ndpsi_32);
}

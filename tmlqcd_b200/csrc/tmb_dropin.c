/* tmb_dropin.c - host side, plain C99: the reference's own entry points for the even/odd
 * twisted-mass path (include/tmlqcd_b200_dropin.h) implemented on the device-level C ABI
 * (include/tmlqcd_b200.h).  No arithmetic on spinors or links happens in this file: it
 * moves caller-owned host buffers to the device, calls tmb_* and moves results back.
 *
 * Implicit inputs are re-read from the reference's globals at EVERY call (callers flip
 * g_mu around operator calls, tm_operators.c:382-386, and call boundary() with another
 * kappa per monomial, monomial/detratio_monomial.c:57-59); the gauge field is re-uploaded
 * when g_update_gauge_copy is set (the reference's dirty flag: start.c:506,
 * update_gauge.c:109; consumer resets it, update_backward_gauge.c:240).
 */
#include <complex.h>
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/tmlqcd_b200.h"
#include "../../include/tmlqcd_b200_dropin.h"

/* the device layer takes ka_mu exactly as boundary() computed them on the host */
int tmb_set_hopping_phases(const double ka_re_im[8]);

/* ---- the reference's globals (global.h, boundary.c:34-38, phmc.h) ---- */
int T, L, LX, LY, LZ, VOLUME, RAND, VOLUMEPLUSRAND;
int g_update_gauge_copy = 1, g_proc_id = 0, g_debug_level = 0, g_nproc = 1, g_nproc_t = 1;
double g_kappa = 0., g_mu = 0., g_mubar = 0., g_epsbar = 0., phmc_invmaxev = 1.;
double X0 = 0., X1 = 0., X2 = 0., X3 = 0.;
_Complex double ka0, ka1, ka2, ka3, phase_0, phase_1, phase_2, phase_3;
su3 **g_gauge_field = NULL;
double mixcg_innereps = 5.0e-5;    /* read_input.h / default_input_values.h:193 */
int mixcg_maxinnersolverit = 5000; /* default_input_values.h:194 */

static su3 *gauge_slab = NULL;
static int dropin_up = 0;
#define NDEV 12
static void *D[NDEV];
static void *D32[4];

static void die(const char *where) {
  /* like fatal_error() (fatal_error.c): message, then abort the program */
  fprintf(stderr, "tmLQCD-B200 FATAL in %s: %s\n", where, tmb_last_error());
  fflush(stderr);
  exit(1);
}
#define CHK(x) do { if ((x) < 0) die(__func__); } while (0)

static void *dev(int k) {
  if (!D[k]) { D[k] = tmb_field_alloc(); if (!D[k]) die("tmb_field_alloc"); }
  return D[k];
}

int tmb_dropin_init(int t, int lx, int ly, int lz, int device) {
  if (device < 0) { const char *lr = getenv("LOCAL_RANK"); device = lr ? atoi(lr) : 0; }
  if (tmb_init(t, lx, ly, lz, device) < 0) return -1;
  T = t; L = lx; LX = lx; LY = ly; LZ = lz;
  VOLUME = t * lx * ly * lz; RAND = 0; VOLUMEPLUSRAND = VOLUME;
  if (!g_gauge_field) { /* one contiguous slab, g_gauge_field[ix][mu] (init/init_gauge_field.c:51-68) */
    gauge_slab = (su3 *)calloc((size_t)VOLUME * 4 + 1, sizeof(su3));
    g_gauge_field = (su3 **)calloc((size_t)VOLUME, sizeof(su3 *));
    if (!gauge_slab || !g_gauge_field) return -2;
    for (int ix = 0; ix < VOLUME; ix++) g_gauge_field[ix] = gauge_slab + 4 * (size_t)ix;
  }
  g_update_gauge_copy = 1;
  dropin_up = 1;
  return 0;
}
int tmb_dropin_finalize(void) {
  for (int k = 0; k < NDEV; k++) D[k] = NULL; /* freed by tmb_finalize */
  for (int k = 0; k < 4; k++) D32[k] = NULL;
  tmb_finalize();
  free(gauge_slab); free(g_gauge_field); gauge_slab = NULL; g_gauge_field = NULL;
  dropin_up = 0;
  return 0;
}

/* boundary.c:40-55 */
void boundary(const double kappa) {
  const double PI_ = 3.14159265358979;
  double x0 = X0 * PI_ / ((T)*g_nproc_t), x1 = X1 * PI_ / (LX), x2 = X2 * PI_ / (LY), x3 = X3 * PI_ / (LZ);
  ka0 = kappa * cexp(x0 * I); ka1 = kappa * cexp(x1 * I);
  ka2 = kappa * cexp(x2 * I); ka3 = kappa * cexp(x3 * I);
  phase_0 = -ka0; phase_1 = -ka1; phase_2 = -ka2; phase_3 = -ka3;
}

/* push the reference's implicit inputs to the device context */
static void sync_globals(void) {
  if (!dropin_up) { fprintf(stderr, "tmLQCD-B200 FATAL: tmb_dropin_init has not been called\n"); exit(1); }
  const double ka[8] = {creal(ka0), cimag(ka0), creal(ka1), cimag(ka1), creal(ka2), cimag(ka2), creal(ka3), cimag(ka3)};
  CHK(tmb_set_hopping_phases(ka));
  CHK(tmb_set_mu(g_mu));
  CHK(tmb_set_nd(g_mubar, g_epsbar, phmc_invmaxev));
  if (g_update_gauge_copy) {
    CHK(tmb_gauge_upload((const double *)g_gauge_field[0]));
    g_update_gauge_copy = 0;
  }
}
static void up(int k, const spinor *h) { CHK(tmb_field_upload(dev(k), (const double *)h)); }
static void down(spinor *h, int k) { CHK(tmb_field_download((double *)h, dev(k))); }

/* ---------------- operators: one upload per input, one download per output ---------------- */
void Hopping_Matrix(const int ieo, spinor *const l, spinor *const k) {
  sync_globals(); CHK(tmb_Hopping_Matrix_host(ieo, (double *)l, (const double *)k, 0, 1., 0.));
}
void Hopping_Matrix_nocom(const int ieo, spinor *const l, spinor *const k) { Hopping_Matrix(ieo, l, k); }
void tm_times_Hopping_Matrix(const int ieo, spinor *const l, spinor *const k, _Complex double const cf) {
  sync_globals(); CHK(tmb_Hopping_Matrix_host(ieo, (double *)l, (const double *)k, 1, creal(cf), cimag(cf)));
}
void tm_sub_Hopping_Matrix(const int ieo, spinor *const l, spinor *const p, spinor *const k, _Complex double const cf) {
  sync_globals(); up(0, k); up(2, p);
  CHK(tmb_tm_sub_Hopping_Matrix(ieo, dev(1), dev(2), dev(0), creal(cf), cimag(cf))); down(l, 1);
}
void H_eo_tm_inv_psi(spinor *const l, spinor *const k, const int ieo, const double sign) {
  /* tm_operators.c:514-521: z = (1 -+ i mu)/(1+mu^2) */
  const double nrm = 1. / (1. + g_mu * g_mu), sg = sign < 0. ? 1. : -1.;
  sync_globals(); CHK(tmb_Hopping_Matrix_host(ieo, (double *)l, (const double *)k, 1, nrm, sg * nrm * g_mu));
}
void tm_sub_H_eo_gamma5(spinor *const l, spinor *const p, spinor *const k, const int ieo, const double sign) {
  sync_globals(); up(0, k); up(2, p); CHK(tmb_tm_sub_H_eo_gamma5(dev(1), dev(2), dev(0), ieo, sign)); down(l, 1);
}
#define UNARY(name) \
  void name(spinor *const l, spinor *const k) { sync_globals(); up(0, k); CHK(tmb_##name(dev(1), dev(0))); down(l, 1); }
UNARY(Qtm_pm_psi)
UNARY(Qtm_plus_psi)
UNARY(Qtm_minus_psi)
UNARY(Mtm_plus_psi)
UNARY(Mtm_minus_psi)

void M_full(spinor *const En, spinor *const On, spinor *const E, spinor *const O) {
  sync_globals(); up(0, E); up(1, O); CHK(tmb_M_full(dev(2), dev(3), dev(0), dev(1))); down(En, 2); down(On, 3);
}
void Q_full(spinor *const En, spinor *const On, spinor *const E, spinor *const O) {
  sync_globals(); up(0, E); up(1, O); CHK(tmb_Q_full(dev(2), dev(3), dev(0), dev(1))); down(En, 2); down(On, 3);
}

/* D_psi_body.c:266-375; aborts on P == Q exactly like the reference (:267-272) */
void D_psi(spinor *const P, spinor *const Q) {
  if (P == Q) {
    printf("Error in D_psi (operator.c):\n");
    printf("Arguments must be different spinor fields\n");
    printf("Program aborted\n");
    exit(1);
  }
  sync_globals();
  CHK(tmb_field_upload_lexic(dev(0), dev(1), (const double *)Q));
  CHK(tmb_D_psi_eo(dev(2), dev(3), dev(0), dev(1)));
  CHK(tmb_field_download_lexic((double *)P, dev(2), dev(3)));
}
/* tm_operators.c:488 / :463 / :380: full-lattice compositions; the reference flips the global g_mu */
void Q_plus_psi(spinor *const l, spinor *const k) {
  sync_globals();
  CHK(tmb_field_upload_lexic(dev(0), dev(1), (const double *)k));
  CHK(tmb_Q_full(dev(2), dev(3), dev(0), dev(1)));
  CHK(tmb_field_download_lexic((double *)l, dev(2), dev(3)));
}
void Q_minus_psi(spinor *const l, spinor *const k) {
  g_mu = -g_mu; Q_plus_psi(l, k); g_mu = -g_mu;
}
void Q_pm_psi(spinor *const l, spinor *const k) {
  sync_globals();
  CHK(tmb_field_upload_lexic(dev(0), dev(1), (const double *)k));
  CHK(tmb_set_mu(-g_mu));
  CHK(tmb_Q_full(dev(2), dev(3), dev(0), dev(1)));
  CHK(tmb_set_mu(g_mu));
  CHK(tmb_Q_full(dev(0), dev(1), dev(2), dev(3)));
  CHK(tmb_field_download_lexic((double *)l, dev(0), dev(1)));
}

/* ---------------- elementwise + BLAS-1: N is VOLUME/2 (one eo field) or VOLUME ---------------- */
static int nparts(int N, const char *who) {
  if (N == VOLUME / 2) return 1;
  if (N == VOLUME) return 2;
  fprintf(stderr, "tmLQCD-B200 FATAL in %s: N=%d is neither VOLUME/2 nor VOLUME\n", who, N);
  exit(1);
}
#define PART(h, j) ((spinor *)(h) + (size_t)(j) * (VOLUME / 2))

void gamma5(spinor *const l, spinor *const k, const int V) {
  sync_globals();
  for (int j = 0, n = nparts(V, __func__); j < n; j++) { up(0, PART(k, j)); CHK(tmb_gamma5(dev(1), dev(0))); down(PART(l, j), 1); }
}
void assign_mul_one_pm_imu_inv(spinor *const l, spinor *const k, const double sign, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up(0, PART(k, j)); CHK(tmb_assign_mul_one_pm_imu_inv(dev(1), dev(0), sign)); down(PART(l, j), 1); }
}
void mul_one_pm_imu_inv(spinor *const l, const double sign, const int N) { assign_mul_one_pm_imu_inv(l, l, sign, N); }
void assign_mul_one_pm_imu(spinor *const l, spinor *const k, const double sign, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up(0, PART(k, j)); CHK(tmb_assign_mul_one_pm_imu(dev(1), dev(0), sign)); down(PART(l, j), 1); }
}
void mul_one_pm_imu(spinor *const l, const double sign) { assign_mul_one_pm_imu(l, l, sign, VOLUME / 2); }
void mul_one_pm_imu_sub_mul_gamma5(spinor *const l, spinor *const k, spinor *const j, const double sign) {
  sync_globals(); up(0, k); up(1, j); CHK(tmb_mul_one_pm_imu_sub_mul_gamma5(dev(2), dev(0), dev(1), sign)); down(l, 2);
}
double square_norm(const spinor *const P, const int N, const int parallel) {
  (void)parallel; /* the device reduction is always global over ranks */
  sync_globals();
  double acc = 0.;
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { double r; up(0, PART(P, j)); CHK(tmb_square_norm(dev(0), &r)); acc += r; }
  return acc;
}
double scalar_prod_r(const spinor *const S, const spinor *const R, const int N, const int parallel) {
  (void)parallel;
  sync_globals();
  double acc = 0.;
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { double r; up(0, PART(S, j)); up(1, PART(R, j)); CHK(tmb_scalar_prod_r(dev(0), dev(1), &r)); acc += r; }
  return acc;
}
void assign_add_mul_r(spinor *const P, spinor *const Q, const double c, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up(0, PART(P, j)); up(1, PART(Q, j)); CHK(tmb_assign_add_mul_r(dev(0), dev(1), c)); down(PART(P, j), 0); }
}
void assign_mul_add_r(spinor *const R, const double c, const spinor *const S, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up(0, PART(R, j)); up(1, PART(S, j)); CHK(tmb_assign_mul_add_r(dev(0), c, dev(1))); down(PART(R, j), 0); }
}
double assign_mul_add_r_and_square(spinor *const R, const double c, const spinor *const S, const int N, const int parallel) {
  (void)parallel;
  sync_globals();
  double acc = 0.;
  for (int j = 0, n = nparts(N, __func__); j < n; j++) {
    double r; up(0, PART(R, j)); up(1, PART(S, j)); CHK(tmb_assign_mul_add_r_and_square(dev(0), c, dev(1), &r)); down(PART(R, j), 0); acc += r;
  }
  return acc;
}
void diff(spinor *const Q, const spinor *const R, const spinor *const S, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up(0, PART(R, j)); up(1, PART(S, j)); CHK(tmb_diff(dev(2), dev(0), dev(1))); down(PART(Q, j), 2); }
}
void add(spinor *const Q, const spinor *const R, const spinor *const S, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up(0, PART(R, j)); up(1, PART(S, j)); CHK(tmb_add(dev(2), dev(0), dev(1))); down(PART(Q, j), 2); }
}
void mul_r(spinor *const R, const double c, spinor *const S, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up(0, PART(S, j)); CHK(tmb_mul_r(dev(1), c, dev(0))); down(PART(R, j), 1); }
}
/* host-to-host copies / permutations stay on the host: no arithmetic, nothing to accelerate.
 * (assign.c:42; convert_eo_to_lexic.c:35-115 through the device permutation kernels) */
void assign(spinor *const R, spinor *const S, const int N) { memmove(R, S, (size_t)N * sizeof(spinor)); }
void convert_eo_to_lexic(spinor *const P, spinor *const s, spinor *const r) {
  sync_globals(); up(0, s); up(1, r); CHK(tmb_field_download_lexic((double *)P, dev(0), dev(1)));
}
void convert_lexic_to_eo(spinor *const s, spinor *const r, spinor *const P) {
  sync_globals(); CHK(tmb_field_upload_lexic(dev(0), dev(1), (const double *)P)); down(s, 0); down(r, 1);
}

/* ---------------- solvers ---------------- */
/* solver/cg_her.c:62-143.  f == Qtm_pm_psi on VOLUME/2 sites (the case invert_eo and
 * solve_degenerate use; monomial_solve.c:134 selects by function-pointer identity the same
 * way) runs entirely on the device.  Any other f is applied through its own host-pointer
 * entry point with the same recurrence driven from here - correct, but every step crosses
 * PCIe; it exists so that the symbol is a complete replacement. */
int cg_her(spinor *const P, spinor *const Q, const int max_iter, double eps_sq, const int rel_prec, const int N, matrix_mult f) {
  if (f == &Qtm_pm_psi && N == VOLUME / 2) {
    sync_globals();
    up(6, Q); up(7, P);
    int iter = tmb_cg_her(dev(7), dev(6), max_iter, eps_sq, rel_prec);
    if (iter < -1) die(__func__);
    down(P, 7);
    if (g_debug_level > 0 && g_proc_id == 0) {
      int it; double err, sec; tmb_solver_stats(&it, &err, &sec);
      printf("# CG: iter: %d eps_sq: %1.4e t/s: %1.4e\n", it, eps_sq, sec); /* cg_her.c:134 */
    }
    return iter;
  }
  const size_t n = (size_t)N;
  spinor *sf0 = calloc(n + 1, sizeof(spinor)), *sf1 = calloc(n + 1, sizeof(spinor)), *sf2 = calloc(n + 1, sizeof(spinor)), *tmp;
  if (!sf0 || !sf1 || !sf2) { fprintf(stderr, "cg_her: out of memory\n"); exit(1); }
  double squarenorm = square_norm(Q, N, 1), normsq, pro, err, alpha, beta;
  int it;
  f(sf0, P);
  diff(sf1, Q, sf0, N);
  assign(sf2, sf1, N);
  normsq = square_norm(sf1, N, 1);
  for (it = 1; it <= max_iter; it++) {
    f(sf0, sf2);
    pro = scalar_prod_r(sf2, sf0, N, 1);
    alpha = normsq / pro;
    assign_add_mul_r(P, sf2, alpha, N);
    err = assign_mul_add_r_and_square(sf0, -alpha, sf1, N, 1);
    if ((err <= eps_sq && rel_prec == 0) || (err <= eps_sq * squarenorm && rel_prec == 1)) break;
    beta = err / normsq;
    assign_mul_add_r(sf2, beta, sf0, N);
    tmp = sf0; sf0 = sf1; sf1 = tmp;
    normsq = err;
  }
  free(sf0); free(sf1); free(sf2);
  return it > max_iter ? -1 : it;
}

/* solver/mixed_cg_her.c:65: the (f, f32) = (Qtm_pm_psi, Qtm_pm_psi_32) pair runs on the device */
int mixed_cg_her(spinor *const P, spinor *const Q, solver_params_t solver_params, const int max_iter, double eps_sq,
                 const int rel_prec, const int N, matrix_mult f, matrix_mult32 f32) {
  (void)solver_params;
  if (f != &Qtm_pm_psi || f32 != (matrix_mult32)&Qtm_pm_psi_32 || N != VOLUME / 2) {
    fprintf(stderr, "tmLQCD-B200 FATAL in mixed_cg_her: only (Qtm_pm_psi, Qtm_pm_psi_32) on VOLUME/2 sites is implemented\n");
    exit(1);
  }
  sync_globals();
  CHK(tmb_set_mixcg(mixcg_innereps, mixcg_maxinnersolverit));
  up(6, Q);
  int iter = tmb_mixed_cg_her(dev(7), dev(6), max_iter, eps_sq, rel_prec);
  if (iter < -1) die(__func__);
  down(P, 7);
  if (g_debug_level > 0 && g_proc_id == 0) {
    int it; double err, sec; tmb_solver_stats(&it, &err, &sec);
    printf("# mixed CG: iter: %d eps_sq: %1.4e t/s: %1.4e\n", it, eps_sq, sec); /* mixed_cg_her.c:181 */
  }
  return iter;
}

/* operator/Hopping_Matrix_32.c:119, operator/tm_operators_32.c:94 on host spinor32 buffers */
static void *dev32(int k) {
  if (!D32[k]) { D32[k] = tmb_field32_alloc(); if (!D32[k]) die("tmb_field32_alloc"); }
  return D32[k];
}
void Hopping_Matrix_32(const int ieo, spinor32 *const l, spinor32 *const k) {
  sync_globals(); CHK(tmb_field32_upload(dev32(0), (const float *)k));
  CHK(tmb_Hopping_Matrix_32(ieo, dev32(1), dev32(0))); CHK(tmb_field32_download((float *)l, dev32(1)));
}
void Qtm_pm_psi_32(spinor32 *const l, spinor32 *const k) {
  sync_globals(); CHK(tmb_field32_upload(dev32(0), (const float *)k));
  CHK(tmb_Qtm_pm_psi_32(dev32(1), dev32(0))); CHK(tmb_field32_download((float *)l, dev32(1)));
}

/* invert_eo.c:83-561: the even/odd CG branch (:152-157, :252, :268-270, :306-310) */
int invert_eo(spinor *const Even_new, spinor *const Odd_new, spinor *const Even, spinor *const Odd,
              const double precision, const int max_iter, const int solver_flag, const int rel_prec,
              const int sub_evs_flag, const int even_odd_flag, const int no_extra_masses,
              double *const extra_masses, solver_params_t solver_params, const int id,
              const ExternalInverter external_inverter, const SloppyPrecision sloppy,
              const CompressionType compression) {
  (void)sub_evs_flag; (void)no_extra_masses; (void)extra_masses; (void)solver_params; (void)id;
  (void)external_inverter; (void)sloppy;
  if (!even_odd_flag || (solver_flag != TMB_SOLVER_CG && solver_flag != TMB_SOLVER_MIXEDCG)) {
    fprintf(stderr, "tmLQCD-B200 FATAL in invert_eo: only the even/odd CG and MIXEDCG branches (even_odd_flag != 0) "
                    "are implemented on the GPU; got solver_flag=%d even_odd_flag=%d\n", solver_flag, even_odd_flag);
    exit(1);
  }
  if (g_proc_id == 0 && g_debug_level > 0) {
    printf("# Using even/odd preconditioning!\n# Using CG!\n# mu = %.12f, kappa = %.12f\n", g_mu / 2. / g_kappa, g_kappa);
    fflush(stdout);
  }
  sync_globals();
  /* CompressionType as the reference hands it to its external inverters (invert_eo.c:93-101);
   * COMPRESSION_8 has no double-precision-exact reconstruction and is served as COMPRESSION_12 */
  CHK(tmb_set_compression(compression == NO_COMPRESSION ? 18 : 12));
  up(6, Even); up(7, Odd); up(9, Odd_new); /* Odd_new is the CG's initial guess (cg_her.c:84) */
  int iter;
  if (solver_flag == TMB_SOLVER_MIXEDCG) { /* invert_eo.c:225-232; mixed_cg_her zeroes the guess (:108) */
    CHK(tmb_set_mixcg(mixcg_innereps, mixcg_maxinnersolverit));
    iter = tmb_invert_eo_mixed(dev(8), dev(9), dev(6), dev(7), precision, max_iter, rel_prec);
  } else
    iter = tmb_invert_eo(dev(8), dev(9), dev(6), dev(7), precision, max_iter, rel_prec);
  if (iter < -1) die(__func__);
  CHK(tmb_set_compression(18));
  down(Even_new, 8); down(Odd_new, 9);
  return iter;
}

/* ---------------- non-degenerate doublet ---------------- */
void M_ee_inv_ndpsi(spinor *const ls, spinor *const lc, spinor *const ks, spinor *const kc, const double mu, const double eps) {
  sync_globals(); up(0, ks); up(1, kc); CHK(tmb_M_ee_inv_ndpsi(dev(2), dev(3), dev(0), dev(1), mu, eps)); down(ls, 2); down(lc, 3);
}
#define ND_OP(name) \
  void name(spinor *const ls, spinor *const lc, spinor *const ks, spinor *const kc) { \
    sync_globals(); up(0, ks); up(1, kc); CHK(tmb_##name(dev(2), dev(3), dev(0), dev(1))); down(ls, 2); down(lc, 3); }
ND_OP(Qtm_ndpsi)
ND_OP(Qtm_dagger_ndpsi)
ND_OP(Qtm_pm_ndpsi)

int cg_her_nd(spinor *const P_up, spinor *P_dn, spinor *const Q_up, spinor *const Q_dn, const int max_iter,
              double eps_sq, const int rel_prec, const int N, matrix_mult_nd f) {
  if (f != &Qtm_pm_ndpsi || N != VOLUME / 2) {
    fprintf(stderr, "tmLQCD-B200 FATAL in cg_her_nd: only f == Qtm_pm_ndpsi on VOLUME/2 sites is implemented\n");
    exit(1);
  }
  sync_globals();
  up(0, P_up); up(1, P_dn); up(2, Q_up); up(3, Q_dn);
  int iter = tmb_cg_her_nd(dev(0), dev(1), dev(2), dev(3), max_iter, eps_sq, rel_prec);
  if (iter < -1) die(__func__);
  down(P_up, 0); down(P_dn, 1);
  return iter;
}
int invert_doublet_eo(spinor *const Even_new_s, spinor *const Odd_new_s, spinor *const Even_new_c, spinor *const Odd_new_c,
                      spinor *const Even_s, spinor *const Odd_s, spinor *const Even_c, spinor *const Odd_c,
                      const double precision, const int max_iter, const int solver_flag, const int rel_prec,
                      solver_params_t solver_params, const ExternalInverter external_inverter,
                      const SloppyPrecision sloppy, const CompressionType compression) {
  (void)solver_flag; (void)solver_params; (void)external_inverter; (void)sloppy; (void)compression;
  sync_globals();
  up(4, Even_s); up(5, Odd_s); up(6, Even_c); up(7, Odd_c);
  up(1, Odd_new_s); up(3, Odd_new_c); /* initial guess of cg_her_nd */
  int iter = tmb_invert_doublet_eo(dev(0), dev(1), dev(2), dev(3), dev(4), dev(5), dev(6), dev(7), precision, max_iter, rel_prec);
  if (iter < -1) die(__func__);
  down(Even_new_s, 0); down(Odd_new_s, 1); down(Even_new_c, 2); down(Odd_new_c, 3);
  return iter;
}

/* ---------------- include/tmLQCD.h facade (wrapper/lib_wrapper.c:77-370) ---------------- */
#define MAX_OPS 16
static struct { double kappa, mu, eps_sq, reached_prec; int max_iter, rel_prec, iterations; } ops[MAX_OPS];
static int no_operators = 0, facade_up = 0, lat[4] = {0, 0, 0, 0};

int tmLQCD_b200_set_lattice(int t, int lx, int ly, int lz) { lat[0] = t; lat[1] = lx; lat[2] = ly; lat[3] = lz; return 0; }
int tmLQCD_b200_set_theta(double x0, double x1, double x2, double x3) { X0 = x0; X1 = x1; X2 = x2; X3 = x3; return 0; }
int tmLQCD_b200_add_operator(double kappa, double two_kappa_mu, double eps_sq, int max_iter, int rel_prec) {
  if (no_operators >= MAX_OPS) return -1;
  ops[no_operators].kappa = kappa; ops[no_operators].mu = two_kappa_mu; ops[no_operators].eps_sq = eps_sq;
  ops[no_operators].max_iter = max_iter; ops[no_operators].rel_prec = rel_prec;
  ops[no_operators].iterations = 0; ops[no_operators].reached_prec = -1.;
  return no_operators++;
}
int tmLQCD_b200_get_solver_info(int op_id, int *iterations, double *reached_prec) {
  if (op_id < 0 || op_id >= no_operators) return -1;
  if (iterations) *iterations = ops[op_id].iterations;
  if (reached_prec) *reached_prec = ops[op_id].reached_prec;
  return 0;
}

/* Minimal reader for the handful of invert.input keys this path needs (the reference's flex
 * grammar read_input.l is out of scope): T, L, LX, LY, LZ, ThetaT/X/Y/Z and, inside
 * BeginOperator TMWILSON ... EndOperator, kappa, 2KappaMu, SolverPrecision, MaxSolverIterations,
 * UseRelativePrecision.  Keys are case-insensitive, '#' starts a comment. */
static int read_invert_input(const char *fn) {
  FILE *f = fopen(fn, "r");
  if (!f) return -1;
  char line[512], key[128], val[128];
  int in_op = 0;
  double kappa = 0., mu = 0., prec = 1e-14; int maxit = 1000, rel = 0;
  while (fgets(line, sizeof(line), f)) {
    char *h = strchr(line, '#'); if (h) *h = 0;
    for (char *c = line; *c; c++) *c = (char)tolower((unsigned char)*c);
    if (strstr(line, "beginoperator")) { in_op = 1; kappa = g_kappa; mu = 0.; prec = 1e-14; maxit = 1000; rel = 0; continue; }
    if (strstr(line, "endoperator")) { if (in_op) tmLQCD_b200_add_operator(kappa, mu, prec, maxit, rel); in_op = 0; continue; }
    if (sscanf(line, " %127[a-z0-9] = %127s", key, val) != 2) continue;
    if (!strcmp(key, "t")) lat[0] = atoi(val);
    else if (!strcmp(key, "l")) lat[1] = lat[2] = lat[3] = atoi(val);
    else if (!strcmp(key, "lx")) lat[1] = atoi(val);
    else if (!strcmp(key, "ly")) lat[2] = atoi(val);
    else if (!strcmp(key, "lz")) lat[3] = atoi(val);
    else if (!strcmp(key, "thetat")) X0 = atof(val);
    else if (!strcmp(key, "thetax")) X1 = atof(val);
    else if (!strcmp(key, "thetay")) X2 = atof(val);
    else if (!strcmp(key, "thetaz")) X3 = atof(val);
    else if (!strcmp(key, "kappa")) { if (in_op) kappa = atof(val); else g_kappa = atof(val); }
    else if (!strcmp(key, "2kappamu")) { if (in_op) mu = atof(val); else g_mu = atof(val); }
    else if (!strcmp(key, "solverprecision")) prec = atof(val);
    else if (!strcmp(key, "maxsolveriterations")) maxit = atoi(val);
    else if (!strcmp(key, "userelativeprecision")) rel = !strcmp(val, "yes");
  }
  fclose(f);
  return 0;
}

int tmLQCD_invert_init(int argc, char *argv[], const int verbose, const int external_id) {
  (void)argc; (void)argv; (void)external_id;
  g_debug_level = verbose;
  if (lat[0] == 0) read_invert_input("invert.input"); /* lib_wrapper.c:96 reads the same file name */
  if (lat[0] == 0) { fprintf(stderr, "tmLQCD_invert_init: lattice size unknown (no invert.input, no tmLQCD_b200_set_lattice)\n"); return -1; }
  if (tmb_dropin_init(lat[0], lat[1], lat[2], lat[3], -1) != 0) { fprintf(stderr, "tmLQCD_invert_init: %s\n", tmb_last_error()); return -1; }
  facade_up = 1;
  return 0;
}
int tmLQCD_read_gauge(const int nconfig) {
  (void)nconfig;
  if (!facade_up) { fprintf(stderr, "tmLQCD_read_gauge: tmLQCD_inver_init must be called first. Aborting...\n"); return -1; }
  /* ILDG/LIME input (io/gauge_read.c, needs c-lime) is outside this path: the caller fills the
   * array returned by tmLQCD_get_gauge_field_pointer and sets g_update_gauge_copy = 1. */
  fprintf(stderr, "tmLQCD_read_gauge: LIME I/O is not part of the B200 path; fill tmLQCD_get_gauge_field_pointer() instead\n");
  return -1;
}
int tmLQCD_get_gauge_field_pointer(double **gf) {
  if (!facade_up) return -1;
  *gf = (double *)g_gauge_field[0];
  g_update_gauge_copy = 1; /* the caller is about to read or write links */
  return 0;
}
int tmLQCD_get_lat_params(tmLQCD_lat_params *p) {
  if (!facade_up) return -1;
  p->LX = LX; p->LY = LY; p->LZ = LZ; p->T = T; p->nstore = 0; p->nsave = 0; p->no_operators = no_operators;
  return 0;
}
int tmLQCD_get_mpi_params(tmLQCD_mpi_params *p) {
  if (!facade_up) return -1;
  memset(p, 0, sizeof(*p));
  p->nproc = g_nproc; p->nproc_t = g_nproc_t; p->nproc_x = p->nproc_y = p->nproc_z = 1;
  p->proc_id = g_proc_id; p->cart_id = g_proc_id; p->time_rank = g_proc_id; p->omp_num_threads = 1;
  p->proc_coords[0] = g_proc_id;
  return 0;
}
/* lib_wrapper.c:242-279 + op_invert (operator.c:312-391): lexicographic source -> eo, CG on the
 * Schur complement, residual check with M_full, normalisation by 2 kappa, eo -> lexicographic.
 * Everything between the upload of `source` and the download of `propagator` stays in HBM. */
int tmLQCD_invert(double *const propagator, double *const source, const int op_id, const int write_prop) {
  (void)write_prop;
  if (!facade_up) { fprintf(stderr, "tmLQCD_invert: tmLQCD_inver_init must be called first. Aborting...\n"); return -1; }
  if (op_id < 0 || op_id >= no_operators) { fprintf(stderr, "tmLQCD_invert: op_id=%d not in valid range. Aborting...\n", op_id); return -1; }
  g_mu = ops[op_id].mu; g_kappa = ops[op_id].kappa; /* op_set_globals, operator.c:320 */
  boundary(g_kappa);
  sync_globals();
  CHK(tmb_field_upload_lexic(dev(6), dev(7), source));
  CHK(tmb_field_zero(dev(8))); CHK(tmb_field_zero(dev(9)));
  int iter = tmb_invert_eo(dev(8), dev(9), dev(6), dev(7), ops[op_id].eps_sq, ops[op_id].max_iter, ops[op_id].rel_prec);
  if (iter < -1) die(__func__);
  ops[op_id].iterations = iter;
  /* reached_prec = |M x - b|^2 (operator.c:358, :379-384) */
  double n1 = 0., n2 = 0.;
  CHK(tmb_M_full(dev(10), dev(11), dev(8), dev(9)));
  CHK(tmb_diff(dev(10), dev(10), dev(6))); CHK(tmb_diff(dev(11), dev(11), dev(7)));
  CHK(tmb_square_norm(dev(10), &n1)); CHK(tmb_square_norm(dev(11), &n2));
  ops[op_id].reached_prec = n1 + n2;
  if (g_kappa != 0.) { CHK(tmb_mul_r(dev(8), 2. * g_kappa, dev(8))); CHK(tmb_mul_r(dev(9), 2. * g_kappa, dev(9))); }
  CHK(tmb_field_download_lexic(propagator, dev(8), dev(9)));
  if (g_proc_id == 0 && g_debug_level > 0)
    printf("# Inversion done in %d iterations, squared residue = %e!\n", iter, ops[op_id].reached_prec);
  return 0;
}
int tmLQCD_finalise(void) {
  if (!facade_up) return -1;
  tmb_dropin_finalize();
  facade_up = 0; no_operators = 0; lat[0] = 0;
  return 0;
}

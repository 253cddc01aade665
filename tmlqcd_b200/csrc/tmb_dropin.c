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
#include <time.h>
#include "../../include/tmlqcd_b200.h"
#include "../../include/tmlqcd_b200_dropin.h"

/* the device layer takes ka_mu exactly as boundary() computed them on the host */
int tmb_set_hopping_phases(const double ka_re_im[8]);

/* ---- the reference's globals (global.h, boundary.c:34-38, phmc.h) ---- */
int T, L, LX, LY, LZ, VOLUME, RAND, VOLUMEPLUSRAND;
int g_update_gauge_copy = 1, g_proc_id = 0, g_debug_level = 0, g_nproc = 1, g_nproc_t = 1, g_nproc_x = 1, g_nproc_y = 1, g_nproc_z = 1;
double g_kappa = 0., g_mu = 0., g_mubar = 0., g_epsbar = 0., phmc_invmaxev = 1.;
double X0 = 0., X1 = 0., X2 = 0., X3 = 0.;
_Complex double ka0, ka1, ka2, ka3, phase_0, phase_1, phase_2, phase_3;
su3 **g_gauge_field = NULL;
double mixcg_innereps = 5.0e-5;    /* read_input.h / default_input_values.h:193 */
int mixcg_maxinnersolverit = 5000; /* default_input_values.h:194 */

static su3 *gauge_slab = NULL;
static int dropin_up = 0;
#define NDEV 14
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
    /* pinned: the links cross PCIe at every g_update_gauge_copy (each MD step in the HMC) */
    gauge_slab = (su3 *)tmb_host_alloc(((size_t)VOLUME * 4 + 1) * sizeof(su3));
    g_gauge_field = (su3 **)calloc((size_t)VOLUME, sizeof(su3 *));
    if (!gauge_slab || !g_gauge_field) return -2;
    memset(gauge_slab, 0, ((size_t)VOLUME * 4 + 1) * sizeof(su3));
    for (int ix = 0; ix < VOLUME; ix++) g_gauge_field[ix] = gauge_slab + 4 * (size_t)ix;
  }
  g_update_gauge_copy = 1;
  dropin_up = 1;
  return 0;
}
static void hmc_forget(void);
int tmb_dropin_finalize(void) {
  for (int k = 0; k < NDEV; k++) D[k] = NULL; /* freed by tmb_finalize */
  for (int k = 0; k < 4; k++) D32[k] = NULL;
  hmc_forget();
  if (gauge_slab) tmb_host_free(gauge_slab); /* before the context goes away; a caller-owned g_gauge_field is left alone */
  tmb_finalize();
  if (gauge_slab) free(g_gauge_field);
  gauge_slab = NULL; g_gauge_field = NULL;
  dropin_up = 0;
  return 0;
}

/* boundary.c:40-55 */
void boundary(const double kappa) {
  const double PI_ = 3.14159265358979;
  double x0 = X0 * PI_ / ((T)*g_nproc_t), x1 = X1 * PI_ / ((LX)*g_nproc_x), x2 = X2 * PI_ / ((LY)*g_nproc_y), x3 = X3 * PI_ / ((LZ)*g_nproc_z);
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
static double wall(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static void up(int k, const spinor *h) { CHK(tmb_field_upload(dev(k), (const double *)h)); }
static void down(spinor *h, int k) { CHK(tmb_field_download((double *)h, dev(k))); }

/* ---------------- operators: one upload per input, one download per output ---------------- */
void Hopping_Matrix(const int ieo, spinor *const l, spinor *const k) {
  sync_globals(); CHK(tmb_Hopping_Matrix_host(ieo, (double *)l, (const double *)k, 0, 1., 0.));
}
/* operator/Hopping_Matrix_nocom.c: no exchange (benchmark.c:337-373 times it against Hopping_Matrix); one rank: identical */
void Hopping_Matrix_nocom(const int ieo, spinor *const l, spinor *const k) {
  if (tmb_comm_nranks() == 1) { Hopping_Matrix(ieo, l, k); return; }
  sync_globals(); up(0, k); CHK(tmb_Hopping_Matrix_nocom(ieo, dev(1), dev(0))); down(l, 1);
}
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

/* measure_gauge_action.c:46-106.  gf is normally g_gauge_field (every caller in the reference passes it); another
 * field is uploaded for the measurement and the device copy is marked dirty again afterwards. */
double measure_plaquette(const su3 **const gf) {
  double res = 0.;
  if ((su3 **)gf != g_gauge_field) {
    if (!dropin_up) { fprintf(stderr, "tmLQCD-B200 FATAL: tmb_dropin_init has not been called\n"); exit(1); }
    CHK(tmb_gauge_upload((const double *)gf[0]));
    g_update_gauge_copy = 1;
  } else
    sync_globals();
  CHK(tmb_measure_plaquette(&res));
  return res;
}

/* ---------------- solvers ---------------- */
/* solver/cg_her.c:62-143.  f == Qtm_pm_psi on VOLUME/2 sites (the case invert_eo and
 * solve_degenerate use; monomial_solve.c:134 selects by function-pointer identity the same
 * way) runs entirely on the device.  Any other f is applied through its own host-pointer
 * entry point with the same recurrence driven from here - correct, but every step crosses
 * PCIe; it exists so that the symbol is a complete replacement. */
/* cg_her on the full lattice with f == Q_pm_psi (the non-even/odd inversions: invert_eo.c:505-545, the det monomial without
 * even/odd preconditioning): the recurrence of cg_her.c:80-127 on (even, odd) PAIRS of device fields.  A lexicographic field
 * is exactly such a pair on the device (the permutation is part of the transfer), Q_pm_psi = Q_+ Q_- is two M_full-type
 * launches per parity with g_mu flipped in between (tm_operators.c:380-388); only the scalars cross PCIe.
 * Slots: x = (6,7) initial guess and result, b = (4,5) source; work fields r (8,9), p (10,11), A p (2,3), temporaries (0,1). */
static void pair_Q_pm(int o0, int o1, int i0, int i1) {
  CHK(tmb_set_mu(-g_mu));
  CHK(tmb_Q_full(dev(0), dev(1), dev(i0), dev(i1)));
  CHK(tmb_set_mu(g_mu));
  CHK(tmb_Q_full(dev(o0), dev(o1), dev(0), dev(1)));
}
static double pair_norm(int a0, int a1) {
  double r0, r1; CHK(tmb_square_norm(dev(a0), &r0)); CHK(tmb_square_norm(dev(a1), &r1)); return r0 + r1;
}
static int cg_pair_Q_pm(const int max_iter, const double eps_sq, const int rel_prec) {
  int r0 = 8, r1 = 9, q0 = 2, q1 = 3, it; /* cg_her.c's solver_field[1] and [0], swapped every iteration */
  const int p0 = 10, p1 = 11;
  double normsq, pro, pr0, pr1, err, e0, e1, alpha, beta;
  const double t0 = wall(), squarenorm = pair_norm(4, 5);
  pair_Q_pm(q0, q1, 6, 7);
  CHK(tmb_diff(dev(r0), dev(4), dev(q0))); CHK(tmb_diff(dev(r1), dev(5), dev(q1)));
  CHK(tmb_assign(dev(p0), dev(r0))); CHK(tmb_assign(dev(p1), dev(r1)));
  normsq = pair_norm(r0, r1);
  for (it = 1; it <= max_iter; it++) {
    pair_Q_pm(q0, q1, p0, p1);
    CHK(tmb_scalar_prod_r(dev(p0), dev(q0), &pr0)); CHK(tmb_scalar_prod_r(dev(p1), dev(q1), &pr1));
    pro = pr0 + pr1;
    alpha = normsq / pro;
    CHK(tmb_assign_add_mul_r(dev(6), dev(p0), alpha)); CHK(tmb_assign_add_mul_r(dev(7), dev(p1), alpha));
    CHK(tmb_assign_mul_add_r_and_square(dev(q0), -alpha, dev(r0), &e0));
    CHK(tmb_assign_mul_add_r_and_square(dev(q1), -alpha, dev(r1), &e1));
    err = e0 + e1;
    if ((err <= eps_sq && rel_prec == 0) || (err <= eps_sq * squarenorm && rel_prec == 1)) break;
    beta = err / normsq;
    CHK(tmb_assign_mul_add_r(dev(p0), beta, dev(q0))); CHK(tmb_assign_mul_add_r(dev(p1), beta, dev(q1)));
    { int t = q0; q0 = r0; r0 = t; t = q1; q1 = r1; r1 = t; }
    normsq = err;
  }
  if (g_debug_level > 0 && g_proc_id == 0) printf("# CG: iter: %d eps_sq: %1.4e t/s: %1.4e\n", it, eps_sq, wall() - t0); /* cg_her.c:134 */
  return it > max_iter ? -1 : it;
}

int cg_her(spinor *const P, spinor *const Q, const int max_iter, double eps_sq, const int rel_prec, const int N, matrix_mult f) {
  if (f == &Q_pm_psi && N == VOLUME) {
    sync_globals();
    CHK(tmb_field_upload_lexic(dev(4), dev(5), (const double *)Q));
    CHK(tmb_field_upload_lexic(dev(6), dev(7), (const double *)P));
    const int iter = cg_pair_Q_pm(max_iter, eps_sq, rel_prec);
    CHK(tmb_field_download_lexic((double *)P, dev(6), dev(7)));
    return iter;
  }
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
/* operator/D_psi.h:28 (body: operator/D_psi.c, the float instantiation of D_psi_body.c) and its caller Q_pm_psi_32,
 * operator/tm_operators_32.c:141-149: lexicographic spinor32 fields of VOLUME sites */
void D_psi_32(spinor32 *const P, spinor32 *const Q) {
  if (P == Q) { /* D_psi_body.c:267-272 */
    printf("Error in D_psi (operator.c):\n");
    printf("Arguments must be different spinor fields\n");
    printf("Program aborted\n");
    exit(1);
  }
  sync_globals();
  CHK(tmb_field32_upload_lexic(dev32(0), dev32(1), (const float *)Q));
  CHK(tmb_D_psi_eo_32(dev32(2), dev32(3), dev32(0), dev32(1)));
  CHK(tmb_field32_download_lexic((float *)P, dev32(2), dev32(3)));
}
void Q_pm_psi_32(spinor32 *const l, spinor32 *const k) {
  sync_globals();
  CHK(tmb_field32_upload_lexic(dev32(0), dev32(1), (const float *)k));
  CHK(tmb_set_mu(-g_mu));
  CHK(tmb_M_full_32(dev32(2), dev32(3), dev32(0), dev32(1), 1)); /* g5 D(-mu) */
  CHK(tmb_set_mu(g_mu));
  CHK(tmb_M_full_32(dev32(0), dev32(1), dev32(2), dev32(3), 1)); /* g5 D(+mu) */
  CHK(tmb_field32_download_lexic((float *)l, dev32(0), dev32(1)));
}
void Qtm_pm_psi_32(spinor32 *const l, spinor32 *const k) {
  sync_globals(); CHK(tmb_field32_upload(dev32(0), (const float *)k));
  CHK(tmb_Qtm_pm_psi_32(dev32(1), dev32(0))); CHK(tmb_field32_download((float *)l, dev32(1)));
}

/* ---------------- single-precision BLAS-1 of the mixed solvers (the _32.c files of linalg/, operator/tm_operators_32.c:130) ----------------
 * N is VOLUME/2 or VOLUME like the double-precision family; every call moves its operands across PCIe. */
#define PART32(h, j) ((spinor32 *)(h) + (size_t)(j) * (VOLUME / 2))
static void up32(int k, const spinor32 *h) { CHK(tmb_field32_upload(dev32(k), (const float *)h)); }
static void down32(spinor32 *h, int k) { CHK(tmb_field32_download((float *)h, dev32(k))); }
float square_norm_32(const spinor32 *const P, const int N, const int parallel) {
  (void)parallel; sync_globals();
  double acc = 0.;
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { double r; up32(0, PART32(P, j)); CHK(tmb_square_norm_32(dev32(0), &r)); acc += r; }
  return (float)acc;
}
float scalar_prod_r_32(const spinor32 *const S, const spinor32 *const R, const int N, const int parallel) {
  (void)parallel; sync_globals();
  double acc = 0.;
  for (int j = 0, n = nparts(N, __func__); j < n; j++) {
    double r; up32(0, PART32(S, j)); up32(1, PART32(R, j)); CHK(tmb_scalar_prod_r_32(dev32(0), dev32(1), &r)); acc += r;
  }
  return (float)acc;
}
void assign_add_mul_r_32(spinor32 *const R, spinor32 *const S, const float c, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up32(0, PART32(R, j)); up32(1, PART32(S, j)); CHK(tmb_blas32(0, dev32(0), dev32(1), NULL, c, 0.)); down32(PART32(R, j), 0); }
}
void assign_mul_add_r_32(spinor32 *const R, const float c, const spinor32 *const S, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up32(0, PART32(R, j)); up32(1, PART32(S, j)); CHK(tmb_blas32(1, dev32(0), dev32(1), NULL, c, 0.)); down32(PART32(R, j), 0); }
}
void diff_32(spinor32 *const Q, const spinor32 *const R, const spinor32 *const S, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up32(0, PART32(R, j)); up32(1, PART32(S, j)); CHK(tmb_blas32(2, dev32(2), dev32(0), dev32(1), 0., 0.)); down32(PART32(Q, j), 2); }
}
void mul_r_32(spinor32 *const R, const float c, spinor32 *const S, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up32(0, PART32(S, j)); CHK(tmb_blas32(3, dev32(1), dev32(0), NULL, c, 0.)); down32(PART32(R, j), 1); }
}
void assign_mul_add_mul_r_32(spinor32 *const R, spinor32 *const S, const float c1, const float c2, const int N) {
  sync_globals();
  for (int j = 0, n = nparts(N, __func__); j < n; j++) { up32(0, PART32(R, j)); up32(1, PART32(S, j)); CHK(tmb_blas32(4, dev32(0), dev32(1), NULL, c1, c2)); down32(PART32(R, j), 0); }
}
void gamma5_32(spinor32 *const l, spinor32 *const k, const int V) {
  sync_globals();
  for (int j = 0, n = nparts(V, __func__); j < n; j++) { up32(0, PART32(k, j)); CHK(tmb_blas32(5, dev32(1), dev32(0), NULL, 0., 0.)); down32(PART32(l, j), 1); }
}

/* invert_eo without even/odd preconditioning, the device part: source pair in slots (6, 7) (overwritten: it is the CG's initial guess and iterate), solution in (2, 3) */
static int invert_no_eo_dev(const double precision, const int max_iter, const int rel_prec) {
  if (g_proc_id == 0 && g_debug_level > 0) { printf("# Not using even/odd preconditioning!\n# Using CG!\n"); fflush(stdout); }
  CHK(tmb_gamma5(dev(4), dev(6))); CHK(tmb_gamma5(dev(5), dev(7)));
  const int iter = cg_pair_Q_pm(max_iter, precision, rel_prec);
  CHK(tmb_set_mu(-g_mu));
  CHK(tmb_Q_full(dev(2), dev(3), dev(6), dev(7)));
  CHK(tmb_set_mu(g_mu));
  return iter;
}
static int invert_no_eo(spinor *const Even_new, spinor *const Odd_new, spinor *const Even, spinor *const Odd,
                        const double precision, const int max_iter, const int rel_prec) {
  /* invert_eo.c:367, :505-545, :558: convert_eo_to_lexic(DUM_DERI, Even, Odd); gamma5(DUM_DERI+1, DUM_DERI);
   * cg_her(DUM_DERI, DUM_DERI+1, .., VOLUME, &Q_pm_psi) - the source is its own initial guess -;
   * Q_minus_psi(DUM_DERI+1, DUM_DERI); convert_lexic_to_eo(Even_new, Odd_new, DUM_DERI+1).  On the device the lexicographic
   * field IS the (even, odd) pair, so the two permutations fall away. */
  sync_globals();
  up(6, Even); up(7, Odd);
  const int iter = invert_no_eo_dev(precision, max_iter, rel_prec);
  down(Even_new, 2); down(Odd_new, 3);
  return iter;
}
/* invert_eo.c:83-561.  With even/odd preconditioning (:126-318) the CG, MIXEDCG and RGMIXEDCG branches (:250-276, :234-241,
 * :242-249); without it (:364-558) the CG branch (:505-545).  Every other solver_flag terminates with a message. */
int invert_eo(spinor *const Even_new, spinor *const Odd_new, spinor *const Even, spinor *const Odd,
              const double precision, const int max_iter, const int solver_flag, const int rel_prec,
              const int sub_evs_flag, const int even_odd_flag, const int no_extra_masses,
              double *const extra_masses, solver_params_t solver_params, const int id,
              const ExternalInverter external_inverter, const SloppyPrecision sloppy,
              const CompressionType compression) {
  (void)sub_evs_flag; (void)no_extra_masses; (void)extra_masses; (void)id;
  (void)external_inverter; (void)sloppy;
  const int eo_ok = solver_flag == TMB_SOLVER_CG || solver_flag == TMB_SOLVER_MIXEDCG || solver_flag == TMB_SOLVER_RGMIXEDCG;
  if ((even_odd_flag && !eo_ok) || (!even_odd_flag && solver_flag != TMB_SOLVER_CG)) {
    fprintf(stderr, "tmLQCD-B200 FATAL in invert_eo: implemented on the GPU are CG, MIXEDCG and RGMIXEDCG with even/odd "
                    "preconditioning and CG without; got solver_flag=%d even_odd_flag=%d\n", solver_flag, even_odd_flag);
    exit(1);
  }
  if (!even_odd_flag) return invert_no_eo(Even_new, Odd_new, Even, Odd, precision, max_iter, rel_prec);
  if (g_proc_id == 0 && g_debug_level > 0) {
    printf("# Using even/odd preconditioning!\n");
    if (solver_flag == TMB_SOLVER_CG) printf("# Using CG!\n# mu = %.12f, kappa = %.12f\n", g_mu / 2. / g_kappa, g_kappa);
    else printf("# Using Mixed Precision CG!\n"); /* invert_eo.c:228, :236 */
    fflush(stdout);
  }
  const double t0 = wall();
  sync_globals();
  /* CompressionType as the reference hands it to its external inverters (invert_eo.c:93-101);
   * COMPRESSION_8 has no double-precision-exact reconstruction and is served as COMPRESSION_12 */
  CHK(tmb_set_compression(compression == NO_COMPRESSION ? 18 : 12));
  const double t1 = wall();
  up(6, Even); up(7, Odd); up(9, Odd_new); /* Odd_new is the CG's initial guess (cg_her.c:84) */
  const double t2 = wall();
  int iter;
  if (solver_flag == TMB_SOLVER_MIXEDCG) { /* invert_eo.c:234-241; mixed_cg_her zeroes the guess (:108) */
    CHK(tmb_set_mixcg(mixcg_innereps, mixcg_maxinnersolverit));
    iter = tmb_invert_eo_mixed(dev(8), dev(9), dev(6), dev(7), precision, max_iter, rel_prec);
  } else if (solver_flag == TMB_SOLVER_RGMIXEDCG) { /* invert_eo.c:242-249; rg_mixed_cg_her.c:243-245 always starts from a zero guess */
    CHK(tmb_set_mcg_delta((double)solver_params.mcg_delta));
    iter = tmb_invert_eo_rgmixed(dev(8), dev(9), dev(6), dev(7), precision, max_iter, rel_prec);
  } else
    iter = tmb_invert_eo(dev(8), dev(9), dev(6), dev(7), precision, max_iter, rel_prec);
  if (iter < -1) die(__func__);
  const double t3 = wall();
  CHK(tmb_set_compression(18));
  down(Even_new, 8); down(Odd_new, 9);
  if (g_proc_id == 0 && g_debug_level > 1) /* where the time of a host-pointer solve goes */
    printf("# invert_eo (B200): globals/gauge %.3f ms, upload %.3f ms, solve %.3f ms, download %.3f ms\n",
           1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (wall() - t3));
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
/* operator/tm_operators_nd_32.c:215 on host spinor32 buffers; solver/rg_mixed_cg_her_nd.c:182 */
void Qtm_pm_ndpsi_32(spinor32 *const l_strange, spinor32 *const l_charm, spinor32 *const k_strange, spinor32 *const k_charm) {
  sync_globals();
  CHK(tmb_field32_upload(dev32(0), (const float *)k_strange)); CHK(tmb_field32_upload(dev32(1), (const float *)k_charm));
  CHK(tmb_Qtm_pm_ndpsi_32(dev32(2), dev32(3), dev32(0), dev32(1)));
  CHK(tmb_field32_download((float *)l_strange, dev32(2))); CHK(tmb_field32_download((float *)l_charm, dev32(3)));
}
int rg_mixed_cg_her_nd(spinor *const P_up, spinor *const P_dn, spinor *const Q_up, spinor *const Q_dn, solver_params_t solver_params,
                       const int max_iter, const double eps_sq, const int rel_prec, const int N, matrix_mult_nd f, matrix_mult_nd32 f32) {
  if (f != &Qtm_pm_ndpsi || f32 != (matrix_mult_nd32)&Qtm_pm_ndpsi_32 || N != VOLUME / 2) {
    fprintf(stderr, "tmLQCD-B200 FATAL in rg_mixed_cg_her_nd: only (Qtm_pm_ndpsi, Qtm_pm_ndpsi_32) on VOLUME/2 sites is implemented\n");
    exit(1);
  }
  sync_globals();
  CHK(tmb_set_mcg_delta((double)solver_params.mcg_delta));
  up(2, Q_up); up(3, Q_dn);
  int iter = tmb_rg_mixed_cg_her_nd(dev(0), dev(1), dev(2), dev(3), max_iter, eps_sq, rel_prec);
  if (iter < -1) die(__func__);
  down(P_up, 0); down(P_dn, 1);
  return iter;
}
int invert_doublet_eo(spinor *const Even_new_s, spinor *const Odd_new_s, spinor *const Even_new_c, spinor *const Odd_new_c,
                      spinor *const Even_s, spinor *const Odd_s, spinor *const Even_c, spinor *const Odd_c,
                      const double precision, const int max_iter, const int solver_flag, const int rel_prec,
                      solver_params_t solver_params, const ExternalInverter external_inverter,
                      const SloppyPrecision sloppy, const CompressionType compression) {
  (void)external_inverter; (void)sloppy; (void)compression;
  /* invert_doublet_eo.c:143-156: RGMIXEDCG -> rg_mixed_cg_her_nd (float inner loops on Qtm_pm_ndpsi_32, delta =
   * solver_params.mcg_delta), every other flag -> cg_her_nd */
  sync_globals();
  if (solver_flag == TMB_SOLVER_RGMIXEDCG) CHK(tmb_set_mcg_delta((double)solver_params.mcg_delta));
  up(4, Even_s); up(5, Odd_s); up(6, Even_c); up(7, Odd_c);
  up(1, Odd_new_s); up(3, Odd_new_c); /* initial guess of cg_her_nd */
  int iter = tmb_invert_doublet_eo_solver(dev(0), dev(1), dev(2), dev(3), dev(4), dev(5), dev(6), dev(7), precision, max_iter, rel_prec, solver_flag);
  if (iter < -1) die(__func__);
  down(Even_new_s, 0); down(Odd_new_s, 1); down(Even_new_c, 2); down(Odd_new_c, 3);
  return iter;
}

/* ---------------- include/tmLQCD.h facade (wrapper/lib_wrapper.c:77-370) ---------------- */
#define MAX_OPS 16
static struct { double kappa, mu, eps_sq, reached_prec, mcg_delta; int max_iter, rel_prec, iterations, solver, even_odd_flag; } ops[MAX_OPS];
static int no_operators = 0, facade_up = 0, lat[4] = {0, 0, 0, 0};

int tmLQCD_b200_set_lattice(int t, int lx, int ly, int lz) { lat[0] = t; lat[1] = lx; lat[2] = ly; lat[3] = lz; return 0; }
int tmLQCD_b200_set_theta(double x0, double x1, double x2, double x3) { X0 = x0; X1 = x1; X2 = x2; X3 = x3; return 0; }
int tmLQCD_b200_add_operator(double kappa, double two_kappa_mu, double eps_sq, int max_iter, int rel_prec) {
  if (no_operators >= MAX_OPS) return -1;
  ops[no_operators].kappa = kappa; ops[no_operators].mu = two_kappa_mu; ops[no_operators].eps_sq = eps_sq;
  ops[no_operators].max_iter = max_iter; ops[no_operators].rel_prec = rel_prec;
  ops[no_operators].iterations = 0; ops[no_operators].reached_prec = -1.;
  /* init_operators' defaults (operator.c:102-103, :125): CG with even/odd preconditioning, mcg_delta = _default_mixcg_innereps */
  ops[no_operators].solver = TMB_SOLVER_CG; ops[no_operators].even_odd_flag = 1; ops[no_operators].mcg_delta = 5.0e-5;
  return no_operators++;
}
/* the operator's Solver / UseEvenOdd / mcgdelta keys (read_input.l:1108-1139, :967-974, :835-838): what invert_eo implements
 * here - CG, MIXEDCG, RGMIXEDCG with even/odd preconditioning, CG without */
int tmLQCD_b200_set_operator_solver(int op_id, int solver_flag, int even_odd_flag, double mcg_delta) {
  if (op_id < 0 || op_id >= no_operators) return -1;
  const int eo_ok = solver_flag == TMB_SOLVER_CG || solver_flag == TMB_SOLVER_MIXEDCG || solver_flag == TMB_SOLVER_RGMIXEDCG;
  if ((even_odd_flag && !eo_ok) || (!even_odd_flag && solver_flag != TMB_SOLVER_CG)) {
    fprintf(stderr, "tmLQCD_b200_set_operator_solver: solver_flag=%d with even_odd_flag=%d is not implemented "
                    "(CG, MIXEDCG, RGMIXEDCG with even/odd preconditioning; CG without)\n", solver_flag, even_odd_flag);
    return -1;
  }
  ops[op_id].solver = solver_flag; ops[op_id].even_odd_flag = even_odd_flag ? 1 : 0;
  if (mcg_delta > 0.) ops[op_id].mcg_delta = mcg_delta;
  return 0;
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
int tmLQCD_b200_set_io(const char *gauge_file, const char *prop_base, int prec);
static int read_invert_input(const char *fn) {
  FILE *f = fopen(fn, "r");
  if (!f) return -1;
  char line[512], orig[512], key[128], val[128];
  int in_op = 0;
  double kappa = 0., mu = 0., prec = 1e-14, delta = 0.; int maxit = 1000, rel = 0, solver = TMB_SOLVER_CG, eo = 1, bad = 0;
  while (fgets(line, sizeof(line), f)) {
    char *h = strchr(line, '#'); if (h) *h = 0;
    strcpy(orig, line); /* file names keep their case */
    for (char *c = line; *c; c++) *c = (char)tolower((unsigned char)*c);
    if (strstr(line, "beginoperator")) { in_op = 1; kappa = g_kappa; mu = 0.; prec = 1e-14; maxit = 1000; rel = 0; solver = TMB_SOLVER_CG; eo = 1; delta = 0.; continue; }
    if (strstr(line, "endoperator")) {
      if (in_op) {
        const int id = tmLQCD_b200_add_operator(kappa, mu, prec, maxit, rel);
        if (id < 0 || tmLQCD_b200_set_operator_solver(id, solver, eo, delta) != 0) bad = 1;
      }
      in_op = 0; continue;
    }
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
    else if (!strcmp(key, "userelativeprecision") || !strcmp(key, "solverrelativeprecision")) rel = !strcmp(val, "yes"); /* read_input.l:824-833 */
    else if (!strcmp(key, "useevenodd")) eo = !strcmp(val, "yes");
    else if (!strcmp(key, "mcgdelta")) delta = atof(val);
    else if (!strcmp(key, "solver")) { /* read_input.l:1108-1139 */
      if (!strcmp(val, "cg")) solver = TMB_SOLVER_CG;
      else if (!strcmp(val, "mixedcg")) solver = TMB_SOLVER_MIXEDCG;
      else if (!strcmp(val, "rgmixedcg")) solver = TMB_SOLVER_RGMIXEDCG;
      else { fprintf(stderr, "invert.input: Solver = %s is not implemented (cg, mixedcg, rgmixedcg)\n", val); bad = 1; }
    }
    else if (!strcmp(key, "gaugeconfiginputfile") || !strcmp(key, "sourcefilename") || !strcmp(key, "propagatorprecision")) {
      char raw[128] = ""; /* read_input.l:399 (GaugeConfigInputFile), :372 (SourceFilename), :808-816 (PropagatorPrecision) */
      const char *eq = strchr(orig, '=');
      if (eq && sscanf(eq + 1, " %127s", raw) == 1) {
        if (key[0] == 'g') tmLQCD_b200_set_io(raw, NULL, 0);
        else if (key[0] == 's') tmLQCD_b200_set_io(NULL, raw, 0);
        else tmLQCD_b200_set_io(NULL, NULL, atoi(raw));
      }
    }
  }
  fclose(f);
  return bad ? -2 : 0;
}

int tmLQCD_invert_init(int argc, char *argv[], const int verbose, const int external_id) {
  (void)argc; (void)argv; (void)external_id;
  g_debug_level = verbose;
  if (lat[0] == 0 && read_invert_input("invert.input") == -2) { /* lib_wrapper.c:96 reads the same file name */
    no_operators = 0; lat[0] = 0; /* nothing of a refused input file stays behind */
    return -1;
  }
  if (lat[0] == 0) { fprintf(stderr, "tmLQCD_invert_init: lattice size unknown (no invert.input, no tmLQCD_b200_set_lattice)\n"); return -1; }
  if (tmb_dropin_init(lat[0], lat[1], lat[2], lat[3], -1) != 0) { fprintf(stderr, "tmLQCD_invert_init: %s\n", tmb_last_error()); return -1; }
  facade_up = 1;
  return 0;
}
static char gauge_input_filename[500] = "conf"; /* default_input_values.h:91 */
static char prop_basename[400] = "source";      /* default_input_values.h:93 */
static int prop_precision = 32, nstore = 0;     /* default_input_values.h:126, :89 */
int tmLQCD_b200_set_io(const char *gauge_file, const char *prop_base, int prec) {
  if (gauge_file) { strncpy(gauge_input_filename, gauge_file, sizeof(gauge_input_filename) - 1); }
  if (prop_base) { strncpy(prop_basename, prop_base, sizeof(prop_basename) - 1); }
  if (prec == 32 || prec == 64) prop_precision = prec;
  return 0;
}
/* wrapper/lib_wrapper.c:203-239: "<GaugeConfigInputFile>.<nconfig, 4 digits>" through read_gauge_field */
int tmLQCD_read_gauge(const int nconfig) {
  char conf_filename[600];
  if (!facade_up) { fprintf(stderr, "tmLQCD_read_gauge: tmLQCD_inver_init must be called first. Aborting...\n"); return -1; }
  nstore = nconfig;
  sprintf(conf_filename, "%s.%.4d", gauge_input_filename, nconfig);
  if (g_proc_id == 0 && g_debug_level > 0)
    printf("#\n# Trying to read gauge field from file %s.\n", conf_filename);
  int j = read_gauge_field(conf_filename, g_gauge_field);
  if (j != 0) {
    fprintf(stderr, "tmLQCD_read_gauge: Error %d while reading gauge field from %s\n ...\n", j, conf_filename);
    return -1;
  }
  if (g_proc_id == 0 && g_debug_level > 0) printf("# Finished reading gauge field.\n");
  /* lib_wrapper.c:232-235 */
  const double plaquette = measure_plaquette((const su3 **)g_gauge_field) / (6. * VOLUME * g_nproc);
  if (g_proc_id == 0) printf("# The computed plaquette value is %.16e.\n", plaquette);
  return 0;
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
  p->nproc = g_nproc; p->nproc_t = g_nproc_t; p->nproc_x = g_nproc_x; p->nproc_y = g_nproc_y; p->nproc_z = g_nproc_z;
  p->proc_id = g_proc_id; p->cart_id = g_proc_id; p->time_rank = g_proc_id; p->omp_num_threads = 1;
  p->proc_coords[0] = g_proc_id;
  return 0;
}
/* lib_wrapper.c:242-279 + op_invert (operator.c:312-391): lexicographic source -> eo, CG on the
 * Schur complement, residual check with M_full, normalisation by 2 kappa, eo -> lexicographic.
 * Everything between the upload of `source` and the download of `propagator` stays in HBM. */
int tmLQCD_invert(double *const propagator, double *const source, const int op_id, const int write_prop) {
  if (!facade_up) { fprintf(stderr, "tmLQCD_invert: tmLQCD_inver_init must be called first. Aborting...\n"); return -1; }
  if (op_id < 0 || op_id >= no_operators) { fprintf(stderr, "tmLQCD_invert: op_id=%d not in valid range. Aborting...\n", op_id); return -1; }
  g_mu = ops[op_id].mu; g_kappa = ops[op_id].kappa; /* op_set_globals, operator.c:320 */
  boundary(g_kappa);
  sync_globals();
  /* source pair in (be, bo), solution in (se, so) */
  int be = 6, bo = 7, se = 8, so = 9, iter;
  const double eps_sq = ops[op_id].eps_sq; const int max_iter = ops[op_id].max_iter, rel_prec = ops[op_id].rel_prec;
  if (!ops[op_id].even_odd_flag) { /* invert_eo.c:364-558: the source is the CG's initial guess and gets overwritten: keep a copy */
    be = 12; bo = 13; se = 2; so = 3;
    CHK(tmb_field_upload_lexic(dev(be), dev(bo), source));
    CHK(tmb_assign(dev(6), dev(be))); CHK(tmb_assign(dev(7), dev(bo)));
    iter = invert_no_eo_dev(eps_sq, max_iter, rel_prec);
  } else {
    CHK(tmb_field_upload_lexic(dev(be), dev(bo), source));
    CHK(tmb_field_zero(dev(se))); CHK(tmb_field_zero(dev(so)));
    if (ops[op_id].solver == TMB_SOLVER_MIXEDCG) {
      CHK(tmb_set_mixcg(mixcg_innereps, mixcg_maxinnersolverit));
      iter = tmb_invert_eo_mixed(dev(se), dev(so), dev(be), dev(bo), eps_sq, max_iter, rel_prec);
    } else if (ops[op_id].solver == TMB_SOLVER_RGMIXEDCG) {
      CHK(tmb_set_mcg_delta(ops[op_id].mcg_delta));
      iter = tmb_invert_eo_rgmixed(dev(se), dev(so), dev(be), dev(bo), eps_sq, max_iter, rel_prec);
    } else
      iter = tmb_invert_eo(dev(se), dev(so), dev(be), dev(bo), eps_sq, max_iter, rel_prec);
  }
  if (iter < -1) die(__func__);
  ops[op_id].iterations = iter;
  /* reached_prec = |M x - b|^2 (operator.c:358, :379-384) */
  double n1 = 0., n2 = 0.;
  CHK(tmb_M_full(dev(10), dev(11), dev(se), dev(so)));
  CHK(tmb_diff(dev(10), dev(10), dev(be))); CHK(tmb_diff(dev(11), dev(11), dev(bo)));
  CHK(tmb_square_norm(dev(10), &n1)); CHK(tmb_square_norm(dev(11), &n2));
  ops[op_id].reached_prec = n1 + n2;
  if (g_kappa != 0.) { CHK(tmb_mul_r(dev(se), 2. * g_kappa, dev(se))); CHK(tmb_mul_r(dev(so), 2. * g_kappa, dev(so))); }
  CHK(tmb_field_download_lexic(propagator, dev(se), dev(so)));
  if (write_prop) { /* op_write_prop (operator.c:532-605): point-source naming, splitted files */
    char fn[600];
    spinor *pe = (spinor *)malloc((size_t)VOLUME / 2 * sizeof(spinor)), *po = (spinor *)malloc((size_t)VOLUME / 2 * sizeof(spinor));
    if (!pe || !po) { fprintf(stderr, "tmLQCD_invert: out of memory\n"); return -1; }
    down(pe, se); down(po, so);
    sprintf(fn, "%s.%.4d.%.2d.%.2d.inverted", prop_basename, nstore, 0, 0);
    /* the solver's name in inverter-info: io/params_construct_InverterInfo.c:57-79 knows "CG", the mixed solvers are "other" */
    const int st = tmb_write_propagator(fn, pe, po, prop_precision, ops[op_id].reached_prec, iter,
                                        ops[op_id].solver == TMB_SOLVER_CG ? "CG" : "other", 0);
    free(pe); free(po);
    if (st != 0) { fprintf(stderr, "tmLQCD_invert: writing %s failed\n", fn); return -1; }
  }
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

/* ==================================================================================================
 * Remaining members of the operator families
 * ================================================================================================== */
/* tm_operators.c:587 / :723: the mass is an ARGUMENT here, not g_mu */
void Mee_inv_psi(spinor *const l, spinor *const k, const double mu) {
  const double nrm = 1. / (1. + mu * mu);
  sync_globals(); up(0, k); CHK(tmb_diag(dev(1), dev(0), nrm, -nrm * mu)); down(l, 1);
}
void Mee_psi(spinor *const l, spinor *const k, const double mu) {
  sync_globals(); up(0, k); CHK(tmb_diag(dev(1), dev(0), 1., mu)); down(l, 1);
}
/* tm_operators.c:813 (mul_one_pm_imu_sub_mul_body.c), :781 */
void mul_one_pm_imu_sub_mul(spinor *const l, spinor *const k, spinor *const j, const double sign, const int N) {
  sync_globals();
  for (int q = 0, n = nparts(N, __func__); q < n; q++) {
    up(0, PART(k, q)); up(1, PART(j, q)); CHK(tmb_mul_one_pm_imu_sub_mul(dev(2), dev(0), dev(1), sign)); down(PART(l, q), 2);
  }
}
void mul_one_sub_mul_gamma5(spinor *const l, spinor *const k, spinor *const j) {
  sync_globals(); up(0, k); up(1, j); CHK(tmb_diag_sub(dev(2), dev(0), dev(1), 1., 0., 1)); down(l, 2);
}
/* tm_operators.c:145 */
void M_minus_1_timesC(spinor *const En, spinor *const On, spinor *const E, spinor *const O) {
  sync_globals(); up(0, E); up(1, O);
  CHK(tmb_H_eo_tm_inv_psi(dev(2), dev(1), EO, +1.)); CHK(tmb_H_eo_tm_inv_psi(dev(3), dev(0), OE, +1.));
  down(En, 2); down(On, 3);
}
/* the "symmetric" even/odd operators, tm_operators.c:186-310: 1 - (M_oo)^-1 H_oe (M_ee)^-1 H_eo.
 * which: 0 -> l = k - w ; 1 -> l = g5 (k - w), with w = (1 +- i mu g5)^-1 H_oe (1 +- i mu g5)^-1 H_eo k */
static void sym_op(spinor *const l, spinor *const k, double sign, int g5) {
  sync_globals(); up(0, k);
  CHK(tmb_H_eo_tm_inv_psi(dev(1), dev(0), EO, sign));
  CHK(tmb_H_eo_tm_inv_psi(dev(2), dev(1), OE, sign));
  CHK(tmb_diag_sub(dev(3), dev(0), dev(2), 1., 0., g5));
  down(l, 3);
}
void Qtm_plus_sym_psi(spinor *const l, spinor *const k) { sym_op(l, k, +1., 1); }
void Qtm_minus_sym_psi(spinor *const l, spinor *const k) { sym_op(l, k, -1., 1); }
void Mtm_plus_sym_psi(spinor *const l, spinor *const k) { sym_op(l, k, +1., 0); }
void Mtm_minus_sym_psi(spinor *const l, spinor *const k) { sym_op(l, k, -1., 0); }
void Qtm_plus_sym_psi_nocom(spinor *const l, spinor *const k) { sym_op(l, k, +1., 1); }
void Mtm_plus_sym_psi_nocom(spinor *const l, spinor *const k) { sym_op(l, k, +1., 0); }
void Mtm_minus_sym_psi_nocom(spinor *const l, spinor *const k) { sym_op(l, k, -1., 0); }
/* tm_operators.c:312-322 */
void Mtm_plus_sym_dagg_psi(spinor *const l, spinor *const k) {
  sync_globals(); up(0, k);
  CHK(tmb_gamma5(dev(1), dev(0)));
  CHK(tmb_assign_mul_one_pm_imu_inv(dev(1), dev(1), -1.));
  CHK(tmb_H_eo_tm_inv_psi(dev(2), dev(1), EO, -1.));
  CHK(tmb_Hopping_Matrix(OE, dev(3), dev(2)));
  CHK(tmb_gamma5(dev(2), dev(3)));
  CHK(tmb_diff(dev(1), dev(0), dev(2)));
  down(l, 1);
}
/* tm_operators.c:347-364, mirrored LITERALLY: its second half multiplies the scratch field DUM_MATRIX (not
 * the freshly hopped DUM_MATRIX+1) by the inverse, so the function is not Q_+^sym Q_-^sym; the reference's
 * behaviour is kept, not silently fixed ("odd corner cases") */
void Qtm_pm_sym_psi(spinor *const l, spinor *const k) {
  sync_globals(); up(0, k);
  void *L = dev(3), *M0 = dev(1), *M1 = dev(2);
  CHK(tmb_H_eo_tm_inv_psi(M1, dev(0), EO, -1.));
  CHK(tmb_H_eo_tm_inv_psi(M0, M1, OE, -1.));
  CHK(tmb_diag_sub(L, dev(0), M0, 1., 0., 1));          /* l = g5 (k - M0) */
  CHK(tmb_H_eo_tm_inv_psi(L, M0, EO, +1.));             /* Hopping_Matrix(EO, l, M0); mul_one_pm_imu_inv(l, +1) */
  CHK(tmb_Hopping_Matrix(OE, M1, L));
  CHK(tmb_assign_mul_one_pm_imu_inv(M0, M0, +1.));      /* applied to M0, as in the reference */
  CHK(tmb_diag_sub(L, dev(0), M0, 1., 0., 1));
  down(l, 3);
}
/* no halo exchange exists at the host-pointer level of a single rank: the *_nocom forms coincide */
void Qtm_plus_psi_nocom(spinor *const l, spinor *const k) { Qtm_plus_psi(l, k); }
void Mtm_plus_psi_nocom(spinor *const l, spinor *const k) { Mtm_plus_psi(l, k); }
void Qtm_pm_psi_nocom(spinor *const l, spinor *const k) { Qtm_pm_psi(l, k); }
/* tm_operators.c:471, :390: full-lattice; the reference flips the global g_mu around D_psi */
void M_minus_psi(spinor *const l, spinor *const k) {
  sync_globals();
  CHK(tmb_field_upload_lexic(dev(0), dev(1), (const double *)k));
  CHK(tmb_set_mu(-g_mu)); CHK(tmb_M_full(dev(2), dev(3), dev(0), dev(1))); CHK(tmb_set_mu(g_mu));
  CHK(tmb_field_download_lexic((double *)l, dev(2), dev(3)));
}
void D_dagg_psi(spinor *const l, spinor *const k) { /* g5 D(-mu) g5 */
  sync_globals();
  CHK(tmb_field_upload_lexic(dev(0), dev(1), (const double *)k));
  CHK(tmb_gamma5(dev(0), dev(0))); CHK(tmb_gamma5(dev(1), dev(1)));
  CHK(tmb_set_mu(-g_mu)); CHK(tmb_Q_full(dev(2), dev(3), dev(0), dev(1))); CHK(tmb_set_mu(g_mu));
  CHK(tmb_field_download_lexic((double *)l, dev(2), dev(3)));
}
/* start.c:354; linalg/assign_to_32.c:37,:84; linalg/addto_32.c:16: host-to-host, conversion on the device */
void zero_spinor_field(spinor *const k, const int N) { memset(k, 0, sizeof(spinor) * (size_t)N); }
void assign_to_32(spinor32 *const R, spinor *const S, const int N) {
  (void)nparts(N, __func__);
  sync_globals();
  for (int q = 0, n = nparts(N, __func__); q < n; q++) {
    up(0, PART(S, q)); CHK(tmb_assign_to_32(dev32(0), dev(0)));
    CHK(tmb_field32_download((float *)((spinor32 *)R + (size_t)q * (VOLUME / 2)), dev32(0)));
  }
}
void assign_to_64(spinor *const R, spinor32 *const S, const int N) {
  sync_globals();
  for (int q = 0, n = nparts(N, __func__); q < n; q++) {
    CHK(tmb_field32_upload(dev32(0), (const float *)((spinor32 *)S + (size_t)q * (VOLUME / 2))));
    CHK(tmb_assign_to_64(dev(0), dev32(0))); down(PART(R, q), 0);
  }
}
void addto_32(spinor *const Q, const spinor32 *const R, const int N) {
  sync_globals();
  for (int q = 0, n = nparts(N, __func__); q < n; q++) {
    CHK(tmb_field32_upload(dev32(0), (const float *)((const spinor32 *)R + (size_t)q * (VOLUME / 2))));
    CHK(tmb_assign_to_64(dev(1), dev32(0))); up(0, PART(Q, q)); CHK(tmb_add(dev(0), dev(0), dev(1))); down(PART(Q, q), 0);
  }
}
/* solver/solver_field.c:31-71: host scratch for callers that drive their own recurrences */
/* solver/solver_field.c:31-66: one slab of nr fields.  This library owns the allocation (it replaces solver_field.o), so the
 * slab is page-locked host memory: solver fields are what a generic-f solver hands to the per-call operators over and over */
#define MAX_PINNED_SLABS 64
static void *pinned_slabs[MAX_PINNED_SLABS];
int init_solver_field(spinor ***const solver_field, const int V, const int nr) {
  if ((*solver_field = (spinor **)malloc((size_t)(nr + 1) * sizeof(spinor *))) == NULL) return 2;
  const size_t bytes = ((size_t)nr * V + 1) * sizeof(spinor);
  spinor *slab = NULL;
  if (dropin_up) {
    for (int i = 0; i < MAX_PINNED_SLABS && !slab; i++)
      if (!pinned_slabs[i] && (slab = (spinor *)tmb_host_alloc(bytes)) != NULL) { pinned_slabs[i] = slab; memset(slab, 0, bytes); break; }
  }
  if (!slab && (slab = (spinor *)calloc((size_t)nr * V + 1, sizeof(spinor))) == NULL) return 1;
  (*solver_field)[nr] = slab;
  (*solver_field)[0] = (*solver_field)[nr];
  for (int i = 1; i < nr; i++) (*solver_field)[i] = (*solver_field)[i - 1] + V;
  return 0;
}
void finalize_solver(spinor **solver_field, const int nr) {
  int pinned = 0;
  for (int i = 0; i < MAX_PINNED_SLABS; i++)
    if (pinned_slabs[i] && pinned_slabs[i] == (void *)solver_field[nr]) { pinned_slabs[i] = NULL; pinned = 1; }
  if (pinned) tmb_host_free(solver_field[nr]); else free(solver_field[nr]);
  free(solver_field);
}
/* tm_operators_nd.c:508, :698, :599 */
void H_eo_tm_ndpsi(spinor *const ls, spinor *const lc, spinor *const ks, spinor *const kc, const int ieo) {
  sync_globals(); up(0, ks); up(1, kc);
  CHK(tmb_Hopping_Matrix(ieo, dev(2), dev(0))); CHK(tmb_Hopping_Matrix(ieo, dev(3), dev(1)));
  CHK(tmb_M_ee_inv_ndpsi(dev(1), dev(0), dev(2), dev(3), -g_mubar, g_epsbar)); /* (l_charm, l_strange, ...) as :515 */
  down(ls, 0); down(lc, 1);
}
void M_oo_sub_g5_ndpsi(spinor *const ls, spinor *const lc, spinor *const ks, spinor *const kc, spinor *const js,
                       spinor *const jc, const double mu, const double eps) {
  sync_globals(); up(0, ks); up(1, kc); up(2, js); up(3, jc);
  CHK(tmb_M_oo_sub_g5_ndpsi(dev(4), dev(5), dev(0), dev(1), dev(2), dev(3), mu, eps)); down(ls, 4); down(lc, 5);
}
void mul_one_pm_iconst(spinor *const l, spinor *const k, const double mu_, const int sign_) {
  sync_globals(); up(0, k); CHK(tmb_diag(dev(1), dev(0), 1., sign_ < 0 ? -mu_ : mu_)); down(l, 1);
}
/* solver/rg_mixed_cg_her.c:180 */
int rg_mixed_cg_her(spinor *const P, spinor *const Q, solver_params_t solver_params, const int max_iter, double eps_sq,
                    const int rel_prec, const int N, matrix_mult f, matrix_mult32 f32) {
  if (f != &Qtm_pm_psi || f32 != (matrix_mult32)&Qtm_pm_psi_32 || N != VOLUME / 2) {
    fprintf(stderr, "tmLQCD-B200 FATAL in rg_mixed_cg_her: only (Qtm_pm_psi, Qtm_pm_psi_32) on VOLUME/2 sites is implemented\n");
    exit(1);
  }
  sync_globals();
  CHK(tmb_set_mcg_delta((double)solver_params.mcg_delta));
  up(6, Q);
  int iter = tmb_rg_mixed_cg_her(dev(7), dev(6), max_iter, eps_sq, rel_prec);
  if (iter < -1) die(__func__);
  down(P, 7);
  return iter;
}

/* ==================================================================================================
 * HMC side with the reference's names and host pointers
 * ================================================================================================== */
int g_relative_precision_flag = 0; /* global.h:75 */

/* deriv_Sb.c:402: accumulates into hf->derivative (host).  hf->gaugefield must be g_gauge_field. */
void deriv_Sb(const int ieo, spinor *const l, spinor *const k, hamiltonian_field_t *const hf, const double factor) {
  if (hf->gaugefield != g_gauge_field) {
    fprintf(stderr, "tmLQCD-B200 FATAL in deriv_Sb: hf->gaugefield is not g_gauge_field\n"); exit(1);
  }
  sync_globals();
  up(0, l); up(1, k);
  CHK(tmb_derivative_upload((const double *)hf->derivative[0]));
  CHK(tmb_deriv_Sb(ieo, dev(0), dev(1), factor));
  CHK(tmb_derivative_download((double *)hf->derivative[0]));
}

static int op_id(matrix_mult f, const char *who) {
  if (f == &Qtm_pm_psi) return TMB_OP_QTM_PM;
  if (f == &Qtm_plus_psi) return TMB_OP_QTM_PLUS;
  if (f == &Qtm_minus_psi) return TMB_OP_QTM_MINUS;
  fprintf(stderr, "tmLQCD-B200 FATAL in %s: matrix_mult must be Qtm_pm_psi, Qtm_plus_psi or Qtm_minus_psi\n", who);
  exit(1);
}
/* solver/chrono_guess.c:43, :82 on host fields: the history is moved to the device for the call and the
 * (orthogonalised / appended) vectors are written back, as the reference modifies v[] in place.  The
 * monomial entry points below keep the history resident instead. */
#define CSG_MAX 20
static void *csg_dev[CSG_MAX];
static void *csg_slot(int i) {
  if (!csg_dev[i]) { csg_dev[i] = tmb_field_alloc(); if (!csg_dev[i]) die("tmb_field_alloc"); }
  return csg_dev[i];
}
void chrono_add_solution(spinor *const trial, spinor **const v, int index_array[], const int N, int *_n, const int V) {
  if (N <= 0) return;
  if (V != VOLUME / 2 || N > CSG_MAX) { fprintf(stderr, "tmLQCD-B200 FATAL in chrono_add_solution: V != VOLUME/2 or N > %d\n", CSG_MAX); exit(1); }
  sync_globals();
  void *hist[CSG_MAX];
  for (int i = 0; i < N; i++) hist[i] = csg_slot(i);
  up(0, trial);
  CHK(tmb_chrono_add_solution(dev(0), hist, index_array, N, _n));
  const int slot = index_array[(*_n) - 1 < N ? (*_n) - 1 : N - 1];
  CHK(tmb_field_download((double *)v[slot], hist[slot]));
}
int chrono_guess(spinor *const trial, spinor *const phi, spinor **const v, int index_array[], const int N, const int n,
                 const int V, matrix_mult f) {
  if (N <= 0) { zero_spinor_field(trial, V); return 0; }
  if (V != VOLUME / 2 || N > CSG_MAX) { fprintf(stderr, "tmLQCD-B200 FATAL in chrono_guess: V != VOLUME/2 or N > %d\n", CSG_MAX); exit(1); }
  sync_globals();
  void *hist[CSG_MAX];
  for (int i = 0; i < N; i++) hist[i] = csg_slot(i);
  for (int j = 0; j < n; j++) CHK(tmb_field_upload(hist[index_array[j]], (const double *)v[index_array[j]]));
  up(0, phi);
  CHK(tmb_chrono_guess(dev(1), dev(0), hist, index_array, N, n, op_id(f, __func__)));
  for (int j = 0; j < n - 1; j++) CHK(tmb_field_download((double *)v[index_array[j]], hist[index_array[j]])); /* orthogonalised */
  down(trial, 1);
  return 0;
}
/* solver/monomial_solve.c:86 */
int solve_degenerate(spinor *const P, spinor *const Q, solver_params_t solver_params, const int max_iter, double eps_sq,
                     const int rel_prec, const int N, matrix_mult f, int solver_type) {
  if (f == &Q_pm_psi && N == VOLUME && solver_type == TMB_SOLVER_CG) /* monomials without even/odd preconditioning: monomial_solve.c:149 */
    return cg_her(P, Q, max_iter, eps_sq, rel_prec, N, f);
  if (f != &Qtm_pm_psi || N != VOLUME / 2) {
    fprintf(stderr, "tmLQCD-B200 FATAL in solve_degenerate: implemented are f == Qtm_pm_psi on VOLUME/2 sites (CG, MIXEDCG, RGMIXEDCG) "
                    "and f == Q_pm_psi on VOLUME sites (CG)\n");
    exit(1);
  }
  if (solver_type != TMB_SOLVER_CG && solver_type != TMB_SOLVER_MIXEDCG && solver_type != TMB_SOLVER_RGMIXEDCG) {
    if (g_proc_id == 0) printf("Error: solver not allowed for degenerate solve. Aborting...\n"); /* monomial_solve.c:164 */
    exit(2);
  }
  sync_globals();
  CHK(tmb_set_mixcg(mixcg_innereps, mixcg_maxinnersolverit));
  CHK(tmb_set_mcg_delta((double)solver_params.mcg_delta));
  up(6, Q); up(7, P); /* P is the initial guess of the CG branch */
  int iter = tmb_solve_degenerate(dev(7), dev(6), max_iter, eps_sq, rel_prec, solver_type);
  if (iter < -1) die(__func__);
  down(P, 7);
  if (g_debug_level > 0) { /* monomial_solve.c:167-174 */
    double r = 0.;
    CHK(tmb_Qtm_pm_psi(dev(8), dev(7))); CHK(tmb_diff(dev(8), dev(8), dev(6))); CHK(tmb_square_norm(dev(8), &r));
    if (g_proc_id == 0) printf("# solve_degenerate residual check: %e\n", r);
  }
  return iter;
}

/* ---- DET / DETRATIO monomials: the reference's hbfunction / accfunction / derivativefunction signatures
 *      (monomial.h:125-127).  The reference keeps the parameters in monomial_list[id] (monomial.h:53-131), a
 *      struct this library does not bind; the glue on the reference side registers them once
 *      (INTEGRATION.md), after which pf, w_fields and the chronological history stay in HBM. ---- */
#define MAX_MNL 30
static int mnl_map[MAX_MNL];
static int mnl_registered[MAX_MNL];
static tmb_random_spinor_fn rng_fn = NULL;
int tmb_dropin_register_monomial(int id, int type, double kappa, double mu, double kappa2, double mu2, int solver,
                                 int maxiter, double forceprec, double accprec, int csg_N) {
  if (id < 0 || id >= MAX_MNL) return -1;
  int did = tmb_monomial_add(type, kappa, mu, kappa2, mu2, solver, maxiter, forceprec, accprec, csg_N);
  if (did < 0) return -1;
  mnl_map[id] = did; mnl_registered[id] = 1;
  return 0;
}
void tmb_dropin_set_random_spinor_field_eo(tmb_random_spinor_fn fn) { rng_fn = fn; }
int tmb_dropin_monomial_info(int id, double *energy0, double *energy1, int *iter0, int *iter1) {
  if (id < 0 || id >= MAX_MNL || !mnl_registered[id]) return -1;
  return tmb_monomial_info(mnl_map[id], energy0, energy1, iter0, iter1, NULL);
}
static void hmc_forget(void) { /* device objects are gone after tmb_finalize */
  for (int i = 0; i < CSG_MAX; i++) csg_dev[i] = NULL;
  for (int i = 0; i < MAX_MNL; i++) mnl_registered[i] = 0;
}
static int mnl_dev(int id, const char *who) {
  if (id < 0 || id >= MAX_MNL || !mnl_registered[id]) {
    fprintf(stderr, "tmLQCD-B200 FATAL in %s: monomial %d was not registered (tmb_dropin_register_monomial)\n", who, id);
    exit(1);
  }
  return mnl_map[id];
}
static void mnl_heatbath(const int id, hamiltonian_field_t *const hf, const char *who) {
  (void)hf;
  const int did = mnl_dev(id, who);
  if (!rng_fn) { fprintf(stderr, "tmLQCD-B200 FATAL in %s: no random_spinor_field_eo registered (tmb_dropin_set_random_spinor_field_eo)\n", who); exit(1); }
  sync_globals();
  CHK(tmb_set_relative_precision_flag(g_relative_precision_flag));
  spinor *eta = (spinor *)calloc((size_t)VOLUME / 2 + 1, sizeof(spinor));
  if (!eta) { fprintf(stderr, "%s: out of memory\n", who); exit(1); }
  rng_fn(eta, 1 /* rngrepro: reproduce_randomnumber_flag */, 0 /* RN_GAUSS */); /* det_monomial.c:165 */
  up(0, eta);
  free(eta);
  double e0;
  CHK(tmb_monomial_heatbath(did, dev(0), &e0));
}
static double mnl_acc(const int id, hamiltonian_field_t *const hf, const char *who) {
  (void)hf;
  const int did = mnl_dev(id, who);
  double dH = 0.;
  sync_globals();
  CHK(tmb_set_relative_precision_flag(g_relative_precision_flag));
  CHK(tmb_monomial_acc(did, &dH));
  return dH;
}
static void mnl_derivative(const int id, hamiltonian_field_t *const hf, const char *who) {
  const int did = mnl_dev(id, who);
  if (hf->gaugefield != g_gauge_field) { fprintf(stderr, "tmLQCD-B200 FATAL in %s: hf->gaugefield is not g_gauge_field\n", who); exit(1); }
  sync_globals();
  CHK(tmb_set_relative_precision_flag(g_relative_precision_flag));
  CHK(tmb_derivative_upload((const double *)hf->derivative[0]));
  CHK(tmb_monomial_derivative(did));
  CHK(tmb_derivative_download((double *)hf->derivative[0]));
}
void det_heatbath(const int id, hamiltonian_field_t *const hf) { mnl_heatbath(id, hf, __func__); }          /* det_monomial.c:150 */
double det_acc(const int id, hamiltonian_field_t *const hf) { return mnl_acc(id, hf, __func__); }           /* det_monomial.c:202 */
void det_derivative(const int id, hamiltonian_field_t *const hf) { mnl_derivative(id, hf, __func__); }      /* det_monomial.c:47 */
void detratio_heatbath(const int id, hamiltonian_field_t *const hf) { mnl_heatbath(id, hf, __func__); }     /* detratio_monomial.c:199 */
double detratio_acc(const int id, hamiltonian_field_t *const hf) { return mnl_acc(id, hf, __func__); }      /* detratio_monomial.c:266 */
void detratio_derivative(const int id, hamiltonian_field_t *const hf) { mnl_derivative(id, hf, __func__); } /* detratio_monomial.c:49 */

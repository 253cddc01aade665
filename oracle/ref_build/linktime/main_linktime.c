/* main_linktime.c - TEST INFRASTRUCTURE ONLY.
 *
 * A tmLQCD-style main program that proves the link-time replacement described in INTEGRATION.md section A: this file and
 * the UNMODIFIED reference callers listed in the Makefile (invert_eo.c, invert_doublet_eo.c, solver/monomial_solve.c,
 * monomial/{monomial,det_monomial,detratio_monomial}.c, start.c, geometry_eo.c, init/ ...) are compiled from
 * /root/reference and linked against libtmlqcd_b200.so INSTEAD of the reference's operator/, linalg/, solver/cg_her*.o,
 * deriv_Sb.o, chrono_guess.o ... objects.  Like every reference main it defines the globals itself (INIT_GLOBALS before
 * global.h, invert.c:28 / hmc_tm.c:27): the executable's definitions interpose the data globals exported by the library,
 * which is how g_mu / ka0..3 / g_update_gauge_copy / g_gauge_field set HERE reach the device.  Function pointers taken
 * here (&Qtm_pm_psi handed to solve_degenerate by the reference's own det_monomial.c) must compare equal to the
 * library's own: that is the `f == Qtm_pm_psi` dispatch of solver/monomial_solve.c:134.
 *
 * usage: linktime <input.bin> <output.bin>; the test (tests/test_linktime.py) writes the inputs of the golden fixtures
 * and compares the outputs with the unmodified reference's results.
 */
#define INIT_GLOBALS
#ifdef HAVE_CONFIG_H
#include <config.h>
#endif
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "global.h"
#include "boundary.h"
#include "geometry_eo.h"
#include "start.h"
#include "init/init.h"
#include "invert_eo.h"
#include "invert_doublet_eo.h"
#include "monomial/monomial.h"
#include "operator/tm_operators.h"
#include "operator/tm_operators_32.h"
#include "phmc.h"
#include "solver/monomial_solve.h"
#include "solver/solver_types.h"

/* new lifecycle symbols of the library (include/tmlqcd_b200_dropin.h; not included here because it re-declares the
 * reference's types for callers that do NOT have the reference headers) */
extern int tmb_dropin_init(int t, int lx, int ly, int lz, int device);
extern int tmb_dropin_finalize(void);
extern const char *tmb_last_error(void);

/* globals the reference's other mains / the flex parser / phmc.c define (read_input.h, phmc.h) */
double mixcg_innereps = 5.0e-5;
int mixcg_maxinnersolverit = 5000;
double phmc_invmaxev = 1.;
double X0 = 0., X1 = 0., X2 = 0., X3 = 0.; /* read_input.h: the flex parser's theta angles, read by boundary() */
extern int even_odd_flag;

static void rd(FILE *f, void *p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "linktime: short read\n"); exit(3); } }
static void wr(FILE *f, const void *p, size_t n) { if (fwrite(p, 1, n, f) != n) { fprintf(stderr, "linktime: short write\n"); exit(3); } }

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s input.bin output.bin\n", argv[0]); return 2; }
  FILE *in = fopen(argv[1], "rb"), *out = fopen(argv[2], "wb");
  if (!in || !out) { perror("linktime"); return 2; }
  int dims[4]; double par[12];
  rd(in, dims, sizeof(dims));
  rd(in, par, sizeof(par)); /* kappa, gmu, theta[4], mubar, epsbar, invmaxev, kappa2, gmu2, delta */

  /* what tmlqcd_mpi_init (mpi_init.c:321-357) and the mains do for a single process */
  T = dims[0]; L = dims[1]; LX = dims[1]; LY = dims[2]; LZ = dims[3]; T_global = T;
  N_PROC_T = N_PROC_X = N_PROC_Y = N_PROC_Z = 1;
  g_nproc = g_nproc_t = g_nproc_x = g_nproc_y = g_nproc_z = 1;
  g_proc_id = 0; g_cart_id = 0; g_stdio_proc = 0;
  for (int i = 0; i < 4; i++) g_proc_coords[i] = 0;
  VOLUME = T * LX * LY * LZ; RAND = 0; EDGES = 0; VOLUMEPLUSRAND = VOLUME;
  SPACEVOLUME = LX * LY * LZ; SPACERAND = 0;
  g_dbw2rand = 0; g_debug_level = 0; g_sloppy_precision_flag = 0; g_sloppy_precision = 0;
  g_rgi_C1 = 0.; g_c_sw = 0.; g_use_clover_flag = 0; lowmem_flag = 0;
  DUM_DERI = 4; DUM_MATRIX = 8; NO_OF_SPINORFIELDS = 14;
  if (init_gauge_field(VOLUMEPLUSRAND, 0) != 0) return 4;      /* the reference's own g_gauge_field (init/init_gauge_field.c) */
  if (init_geometry_indices(VOLUMEPLUSRAND) != 0) return 4;
  if (init_spinor_field(VOLUMEPLUSRAND, NO_OF_SPINORFIELDS) != 0) return 4; /* g_spinor_field: DUM_DERI.. scratch of invert_eo.c */
  if (init_moment_field(VOLUME, VOLUMEPLUSRAND) != 0) return 4;
  geometry();

  /* the one call INTEGRATION.md asks a main for; the library sees THIS program's T, LX, ..., g_gauge_field */
  if (tmb_dropin_init(T, LX, LY, LZ, -1) != 0) { fprintf(stderr, "tmb_dropin_init: %s\n", tmb_last_error()); return 5; }

  rd(in, g_gauge_field[0], (size_t)VOLUME * 4 * sizeof(su3));
  g_update_gauge_copy = 1;                                      /* the dirty flag, as start.c:506 / io/gauge_read.c:186 set it */
  g_kappa = par[0]; g_mu = par[1]; X0 = par[2]; X1 = par[3]; X2 = par[4]; X3 = par[5];
  g_mubar = par[6]; g_epsbar = par[7]; phmc_invmaxev = par[8];
  boundary(g_kappa);

  const size_t fb = (size_t)(VOLUME / 2) * sizeof(spinor);
  spinor *src[4], *sol[4];
  for (int i = 0; i < 4; i++) { src[i] = calloc(1, fb); sol[i] = calloc(1, fb); rd(in, src[i], fb); }
  solver_params_t sp;
  memset(&sp, 0, sizeof(sp));
  sp.mcg_delta = (float)par[11];
  int it;

  /* (1) unmodified invert_eo.c, CG branch: per-call operators + cg_her(&Qtm_pm_psi) of the library */
  it = invert_eo(sol[0], sol[1], src[0], src[1], 1e-20, 1000, CG, 1, 0, 1, 0, NULL, sp, 0, NO_EXT_INV, 0, NO_COMPRESSION);
  wr(out, &it, sizeof(int)); wr(out, sol[0], fb); wr(out, sol[1], fb);
  /* (1b) the same file's RGMIXEDCG branch (:242-249: the library's rg_mixed_cg_her must recognise the executable's
   *      &Qtm_pm_psi / &Qtm_pm_psi_32) and its branch WITHOUT even/odd preconditioning (:364-558: convert_eo_to_lexic, gamma5,
   *      cg_her(.., VOLUME, &Q_pm_psi) - the library's CG on (even, odd) pairs, found by f == Q_pm_psi -, Q_minus_psi) */
  for (int pass = 0; pass < 2; pass++) {
    memset(sol[0], 0, fb); memset(sol[1], 0, fb);
    it = invert_eo(sol[0], sol[1], src[0], src[1], pass ? 1e-24 : 1e-20, 3000, pass ? CG : RGMIXEDCG, 1, 0, pass ? 0 : 1, 0, NULL, sp, 0,
                   NO_EXT_INV, 0, NO_COMPRESSION);
    wr(out, &it, sizeof(int)); wr(out, sol[0], fb); wr(out, sol[1], fb);
  }

  /* (2) unmodified invert_doublet_eo.c: CG -> the library's cg_her_nd, RGMIXEDCG -> its rg_mixed_cg_her_nd */
  for (int pass = 0; pass < 2; pass++) {
    for (int i = 0; i < 4; i++) memset(sol[i], 0, fb);
    it = invert_doublet_eo(sol[0], sol[1], sol[2], sol[3], src[0], src[1], src[2], src[3], 1e-18, 1000,
                           pass ? RGMIXEDCG : CG, 1, sp, NO_EXT_INV, 0, NO_COMPRESSION);
    wr(out, &it, sizeof(int));
    for (int i = 0; i < 4; i++) wr(out, sol[i], fb);
  }

  /* (3) unmodified solver/monomial_solve.c: the f == Qtm_pm_psi dispatch with every scoped solver */
  const int solvers[3] = {CG, MIXEDCG, RGMIXEDCG};
  for (int k = 0; k < 3; k++) {
    memset(sol[0], 0, fb);
    it = solve_degenerate(sol[0], src[0], sp, 2000, 1e-20, 1, VOLUME / 2, &Qtm_pm_psi, solvers[k]);
    wr(out, &it, sizeof(int)); wr(out, sol[0], fb);
  }

  /* (4) unmodified monomial/det_monomial.c + detratio_monomial.c over the library's operators, chrono_guess, deriv_Sb:
   *     heatbath (noise from the reference's RANLUX: start.c compiled here), three derivatives, acc */
  hamiltonian_field_t hf;
  hf.gaugefield = g_gauge_field; hf.momenta = moment; hf.derivative = df0; hf.update_gauge_copy = 0; hf.traj_counter = 0;
  g_relative_precision_flag = 0;
  const int types[2] = {DET, DETRATIO};
  for (int k = 0; k < 2; k++) {
    const int id = add_monomial(types[k]) - 1;
    monomial *m = &monomial_list[id];
    m->type = types[k]; m->kappa = par[0]; m->mu = par[1]; m->kappa2 = par[9]; m->mu2 = par[10]; m->solver = CG; m->maxiter = 2000;
    m->forceprec = 1e-22; m->accprec = 1e-24; m->csg_N = 2; m->csg_N2 = 0; m->even_odd_flag = 1;
    m->solver_params.mcg_delta = (float)mixcg_innereps;
  }
  if (init_monomials(VOLUMEPLUSRAND / 2, even_odd_flag) != 0 || init_csg_field(VOLUMEPLUSRAND / 2) != 0) return 6;
  for (int id = 0; id < 2; id++) {
    monomial *m = &monomial_list[id];
    start_ranlux(1, 1000 + id + 2); /* the seeds of tests/golden/make_golden_hmc.py for its monomials 2 and 3 (csg_N = 2) */
    m->hbfunction(id, &hf);
    wr(out, &m->energy0, sizeof(double));
    memset(df0[0], 0, (size_t)VOLUME * 4 * sizeof(su3adj));
    for (int call = 0; call < 3; call++) {
      m->derivativefunction(id, &hf);
      wr(out, &m->iter1, sizeof(int));
      wr(out, df0[0], (size_t)VOLUME * 4 * sizeof(su3adj));
    }
    const double dH = m->accfunction(id, &hf);
    wr(out, &dH, sizeof(double)); wr(out, &m->iter0, sizeof(int));
  }
  fclose(in); fclose(out);
  tmb_dropin_finalize();
  printf("linktime: done\n");
  return 0;
}

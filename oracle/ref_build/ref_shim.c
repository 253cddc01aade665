/* Definitions of the globals that the reference's flex-generated input parser
 * (read_input.l, declared in read_input.h:45-137) and monomial.c would provide.
 * flex is not available, so read_input.c cannot be generated; this shim gives the
 * compiled reference objects the symbols they link against.
 * TEST INFRASTRUCTURE ONLY. */
#include <stdio.h>
#include <stdlib.h>

int verbose = 0;
int index_start = 0;
int reproduce_randomnumber_flag = 1;
int even_odd_flag = 1;
int bc_flag = 0;
int usegpu_flag = 0;
int use_preconditioning = 0;
int Nmeas = 1, Nsave = 1;          /* read_input.h: only printed into xlf-info style records */
int gauge_precision_read_flag = 64; /* read_input.h:69, GaugeConfigReadPrecision */
#ifndef TM_USE_OMP
int omp_num_threads = 1;
#endif

void fatal_error(char const *error, char const *function) {
  fprintf(stderr, "FATAL ERROR in reference (%s): %s\n", function, error);
  exit(500);
}

/* Symbols referenced by reference objects from code paths that are NOT on the
 * even/odd twisted-mass path (clover term, deflation blocks, spectral preconditioner,
 * chronological-guess fields).  They only have to exist for the shared object to load;
 * reaching one of them means the oracle left the scoped path, so they abort. */
#define OFF_PATH(name) void name() { fatal_error("off-path symbol " #name " called", "ref_shim"); }
OFF_PATH(assign_mul_one_sw_pm_imu_eps)
OFF_PATH(assign_mul_one_sw_pm_imu_site_lexic)
OFF_PATH(assign_mul_one_sw_pm_imu_site_lexic_32)
OFF_PATH(clover_gamma5_nd)
OFF_PATH(clover_inv_nd)
/* referenced by Qsw_pm_ndpsi_32 in operator/tm_operators_nd_32.c (compiled for Qtm_pm_ndpsi_32) */
OFF_PATH(assign_mul_one_sw_pm_imu_eps_32_orphaned) OFF_PATH(clover_gamma5_nd_32_orphaned) OFF_PATH(clover_inv_nd_32_orphaned)
OFF_PATH(init_blocks_eo_gaugefield)
OFF_PATH(init_blocks_eo_gaugefield_32)
OFF_PATH(init_blocks_gaugefield)
OFF_PATH(init_blocks_gaugefield_32)
OFF_PATH(spinorPrecondition)
double g_prec_sequence_d_dagger_d[3] = {0., 0., 0.};
/* monomial/monomial.c (compiled unmodified for add_monomial / init_monomials /
 * mnl_backup_restore_globals and the DET / DETRATIO wiring) references every other monomial type
 * and solver/monomial_solve.c every other solver; none of them is on the scoped path. */
OFF_PATH(Qsw_full_minus_psi) OFF_PATH(Qsw_full_plus_psi) OFF_PATH(Qsw_full_pm_psi)
OFF_PATH(Qsw_minus_psi) OFF_PATH(Qsw_plus_psi) OFF_PATH(Qsw_pm_psi) OFF_PATH(Qsw_pm_psi_32)
OFF_PATH(copy_32_sw_fields) OFF_PATH(init_swpm)
OFF_PATH(clover_trlog_acc) OFF_PATH(clover_trlog_heatbath)
OFF_PATH(cloverdet_acc) OFF_PATH(cloverdet_derivative) OFF_PATH(cloverdet_heatbath)
OFF_PATH(cloverdetratio_acc) OFF_PATH(cloverdetratio_derivative) OFF_PATH(cloverdetratio_heatbath)
OFF_PATH(cloverdetratio_rwacc) OFF_PATH(clovernd_trlog_acc) OFF_PATH(clovernd_trlog_heatbath)
OFF_PATH(cloverndpoly_acc) OFF_PATH(cloverndpoly_derivative) OFF_PATH(cloverndpoly_heatbath)
OFF_PATH(gauge_EMderivative) OFF_PATH(gauge_acc) OFF_PATH(gauge_derivative) OFF_PATH(gauge_heatbath)
OFF_PATH(init_ndpoly_monomial) OFF_PATH(init_ndrat_monomial) OFF_PATH(nddetratio_acc)
OFF_PATH(ndpoly_acc) OFF_PATH(ndpoly_derivative) OFF_PATH(ndpoly_heatbath)
OFF_PATH(ndrat_acc) OFF_PATH(ndrat_derivative) OFF_PATH(ndrat_heatbath)
OFF_PATH(ndratcor_acc) OFF_PATH(ndratcor_heatbath)
OFF_PATH(poly_acc) OFF_PATH(poly_derivative) OFF_PATH(poly_heatbath)
OFF_PATH(rat_acc) OFF_PATH(rat_derivative) OFF_PATH(rat_heatbath) OFF_PATH(ratcor_acc) OFF_PATH(ratcor_heatbath)
OFF_PATH(bicgstab_complex) OFF_PATH(cg_mms_tm) OFF_PATH(cg_mms_tm_nd) OFF_PATH(mixed_cg_mms_tm_nd)
OFF_PATH(sw_term) OFF_PATH(sw_invert) OFF_PATH(sw_deriv) OFF_PATH(sw_all)
/* invert_eo.c (compiled unmodified) can dispatch to every Krylov solver of solver/; only CG, MIXEDCG and RGMIXEDCG are on
 * the scoped path */
OFF_PATH(bicg_complex) OFF_PATH(bicgstabell) OFF_PATH(cgs_real) OFF_PATH(cr) OFF_PATH(fgmres) OFF_PATH(gcr) OFF_PATH(gmres)
OFF_PATH(gmres_dr) OFF_PATH(incr_eigcg) OFF_PATH(mcr) OFF_PATH(mr) OFF_PATH(pcg_her)
int gmres_m_parameter = 10, gmresdr_nr_ev = 0; /* read_input.h, only passed to the solvers above */

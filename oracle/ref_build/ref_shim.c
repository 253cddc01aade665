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
int no_monomials = 0;
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
OFF_PATH(init_blocks_eo_gaugefield)
OFF_PATH(init_blocks_eo_gaugefield_32)
OFF_PATH(init_blocks_gaugefield)
OFF_PATH(init_blocks_gaugefield_32)
OFF_PATH(spinorPrecondition)
double g_prec_sequence_d_dagger_d[3] = {0., 0., 0.};
/* monomial_list is only touched by init_csg_field (init/init_spinor_field.c:183), never called here */
char monomial_list[1 << 20];

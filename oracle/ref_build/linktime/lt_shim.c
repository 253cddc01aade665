/* lt_shim.c - TEST INFRASTRUCTURE ONLY: symbols the unmodified callers of the link-time test reference from code paths
 * outside the even/odd twisted-mass path (clover / non-eo branches of det_monomial.c, the CGMMS propagator writer of
 * invert_eo.c).  In a full tmLQCD build they come from objects that stay with the reference; here they abort. */
#ifdef HAVE_CONFIG_H
#include <config.h>
#endif
#include <stdlib.h>
void fatal_error(char const *error, char const *function);
/* io/params.h: paramsSourceInfo / paramsPropInfo (about a hundred bytes each); only their addresses are taken by the
 * off-path branches, so opaque storage of ample size does */
char SourceInfo[4096], PropInfo[4096];
#define OFF_PATH(name) void name() { fatal_error("off-path symbol " #name " called", "lt_shim"); }
OFF_PATH(deriv_Sb_D_psi) OFF_PATH(D_psi_prec) OFF_PATH(Q_pm_psi_prec)
OFF_PATH(Qsw_pm_ndpsi) OFF_PATH(Qsw_pm_ndpsi_32) OFF_PATH(Qsw_dagger_ndpsi) OFF_PATH(Msw_ee_inv_ndpsi)
OFF_PATH(write_spinor_info) OFF_PATH(write_spinor) OFF_PATH(write_propagator_format) OFF_PATH(destruct_writer)
OFF_PATH(construct_writer) OFF_PATH(construct_paramsPropagatorFormat) OFF_PATH(construct_paramsInverterInfo)

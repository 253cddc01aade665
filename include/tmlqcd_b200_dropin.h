/* tmlqcd_b200_dropin.h - the reference's own C symbols for the even/odd twisted-mass path,
 * re-implemented on the B200 (implementation: tmlqcd_b200/csrc/tmb_dropin.c on top of the
 * device-level C ABI in tmlqcd_b200.h).
 *
 * Same names, same signatures, same host data layouts, same implicit inputs (the globals
 * g_mu, g_kappa, ka0..3 via boundary(), g_mubar, g_epsbar, phmc_invmaxev, g_gauge_field,
 * g_update_gauge_copy, VOLUME, T, LX, LY, LZ) and the same error behaviour as urbach/tmLQCD:
 * operators are void and abort with a message on fatal conditions, solvers return the
 * iteration count or -1 (solver/cg_her.c:141).  A reference executable links this library
 * INSTEAD of the objects named beside each prototype.
 *
 * Pointer semantics: every spinor* / su3** is caller-owned HOST memory in the reference
 * layout.  Each operator call uploads its inputs, runs on the GPU and downloads its outputs;
 * the solvers (cg_her with f == Qtm_pm_psi, invert_eo, cg_her_nd, invert_doublet_eo) upload
 * once, iterate with all fields and scalars resident in HBM, and download the solution.
 */
#ifndef TMLQCD_B200_DROPIN_H
#define TMLQCD_B200_DROPIN_H
#include <complex.h>
#ifdef __cplusplus
#error "C header (uses C99 _Complex like the reference's su3.h); bind from C"
#endif

/* ---- ABI types: su3.h:40-63 ---- */
typedef struct { _Complex double c00, c01, c02, c10, c11, c12, c20, c21, c22; } su3;
typedef struct { _Complex double c0, c1, c2; } su3_vector;
typedef struct { su3_vector s0, s1, s2, s3; } spinor;
typedef struct { _Complex float c0, c1, c2; } su3_vector32;
typedef struct { su3_vector32 s0, s1, s2, s3; } spinor32; /* su3.h:55-78 */
/* solver/matrix_mult_typedef.h:30, matrix_mult_typedef_nd.h */
typedef void (*matrix_mult)(spinor *const, spinor *const);
typedef void (*matrix_mult_nd)(spinor *const, spinor *const, spinor *const, spinor *const);
typedef void (*matrix_mult32)(void *const, void *const);
typedef void (*matrix_mult_nd32)(void *const, void *const, void *const, void *const);
/* misc_types.h:26-43 */
typedef enum SloppyPrecision_s { SLOPPY_DOUBLE = 0, SLOPPY_SINGLE, SLOPPY_HALF } SloppyPrecision;
typedef enum CompressionType_s { NO_COMPRESSION = 18, COMPRESSION_12 = 12, COMPRESSION_8 = 8 } CompressionType;
typedef enum ExternalInverter_s { NO_EXT_INV = 0, QUDA_INVERTER, QPHIX_INVERTER } ExternalInverter;
/* solver/solver_params.h:39-101 (passed BY VALUE to invert_eo: layout must match) */
typedef enum solution_type_t { TM_SOLUTION_M_MDAG = 0, TM_SOLUTION_M } solution_type_t;
typedef struct {
  int eigcg_nrhs, eigcg_nrhs1, eigcg_nev, eigcg_vmax, eigcg_ldh;
  double eigcg_tolsq1, eigcg_tolsq, eigcg_restolsq;
  int eigcg_rand_guess_opt;
  float mcg_delta;
  int type, max_iter, rel_prec, no_shifts, sdim;
  double squared_solver_prec;
  matrix_mult M_psi;
  matrix_mult32 M_psi32;
  matrix_mult_nd M_ndpsi;
  matrix_mult_nd32 M_ndpsi32;
  double *shifts;
  solution_type_t solution_type;
  CompressionType compression_type;
  SloppyPrecision sloppy_precision;
  ExternalInverter external_inverter;
} solver_params_t;
/* solver/solver_types.h:23-49: TMB_SOLVER_CG / _MIXEDCG / _RGMIXEDCG come from tmlqcd_b200.h */
#define EO 0 /* global.h / operator headers: ieo = 0 -> output on even sites */
#define OE 1

/* ---- globals of global.h / boundary.c / phmc.h that the path reads at call time ---- */
extern int T, L, LX, LY, LZ, VOLUME, RAND, VOLUMEPLUSRAND;
extern int g_update_gauge_copy, g_proc_id, g_debug_level, g_nproc, g_nproc_t, g_nproc_x, g_nproc_y, g_nproc_z; /* global.h:206 */
extern double g_kappa, g_mu, g_mubar, g_epsbar, phmc_invmaxev;
extern double X0, X1, X2, X3;
extern _Complex double ka0, ka1, ka2, ka3, phase_0, phase_1, phase_2, phase_3;
extern su3 **g_gauge_field;
extern double mixcg_innereps;      /* read_input.h: MixCGInnerEps */
extern int mixcg_maxinnersolverit; /* read_input.h: MixCGMaxIter */

/* ---- lifecycle added by the drop-in (the reference allocates in its mains) ----
 * sets T,L,LX,LY,LZ,VOLUME,..., allocates g_gauge_field (init/init_gauge_field.c:41) and
 * brings up the device context.  device < 0: take LOCAL_RANK or 0. */
int tmb_dropin_init(int t, int lx, int ly, int lz, int device);
int tmb_dropin_finalize(void);

/* boundary.c:40 */
void boundary(const double kappa);
/* operator/Hopping_Matrix.c:131, Hopping_Matrix_nocom.c, tm_times_Hopping_Matrix.c:119, tm_sub_Hopping_Matrix.c:122 */
void Hopping_Matrix(const int ieo, spinor *const l, spinor *const k);
void Hopping_Matrix_nocom(const int ieo, spinor *const l, spinor *const k);
void tm_times_Hopping_Matrix(const int ieo, spinor *const l, spinor *const k, _Complex double const cfactor);
void tm_sub_Hopping_Matrix(const int ieo, spinor *const l, spinor *const p, spinor *const k, _Complex double const cfactor);
/* operator/tm_operators.c:338,:172,:216,:245,:289,:508,:528,:117,:130 */
void Qtm_pm_psi(spinor *const l, spinor *const k);
void Qtm_plus_psi(spinor *const l, spinor *const k);
void Qtm_minus_psi(spinor *const l, spinor *const k);
void Mtm_plus_psi(spinor *const l, spinor *const k);
void Mtm_minus_psi(spinor *const l, spinor *const k);
void H_eo_tm_inv_psi(spinor *const l, spinor *const k, const int ieo, const double sign);
void tm_sub_H_eo_gamma5(spinor *const l, spinor *const p, spinor *const k, const int ieo, const double sign);
void M_full(spinor *const Even_new, spinor *const Odd_new, spinor *const Even, spinor *const Odd);
void Q_full(spinor *const Even_new, spinor *const Odd_new, spinor *const Even, spinor *const Odd);
/* operator/mul_one_pm_imu_inv_body.c:1,:44; tm_operators.c:627,:669,:813 */
void mul_one_pm_imu_inv(spinor *const l, const double sign, const int N);
void assign_mul_one_pm_imu_inv(spinor *const l, spinor *const k, const double sign, const int N);
void mul_one_pm_imu(spinor *const l, const double sign);
void assign_mul_one_pm_imu(spinor *const l, spinor *const k, const double sign, const int N);
void mul_one_pm_imu_sub_mul_gamma5(spinor *const l, spinor *const k, spinor *const j, const double sign);
/* operator/D_psi_body.c:266; tm_operators.c:380,:488,:463 */
void D_psi(spinor *const P, spinor *const Q);
void Q_pm_psi(spinor *const l, spinor *const k);
void Q_plus_psi(spinor *const l, spinor *const k);
void Q_minus_psi(spinor *const l, spinor *const k);
/* gamma.c:77 */
void gamma5(spinor *const l, spinor *const k, const int V);
/* linalg/: square_norm.c:253, scalar_prod_r.c:135, assign_add_mul_r.c:346, assign_mul_add_r.c:340,
 * assign_mul_add_r_and_square.c:145, diff.c:270, assign.c:42, mul_r.c:40, add.c,
 * convert_eo_to_lexic.c:35,:77 */
double square_norm(const spinor *const P, const int N, const int parallel);
double scalar_prod_r(const spinor *const S, const spinor *const R, const int N, const int parallel);
void assign_add_mul_r(spinor *const P, spinor *const Q, const double c, const int N);
void assign_mul_add_r(spinor *const R, const double c, const spinor *const S, const int N);
double assign_mul_add_r_and_square(spinor *const R, const double c, const spinor *const S, const int N, const int parallel);
void diff(spinor *const Q, const spinor *const R, const spinor *const S, const int N);
void add(spinor *const Q, const spinor *const R, const spinor *const S, const int N);
void assign(spinor *const R, spinor *const S, const int N);
void mul_r(spinor *const R, const double c, spinor *const S, const int N);
void convert_eo_to_lexic(spinor *const P, spinor *const s, spinor *const r);
void convert_lexic_to_eo(spinor *const s, spinor *const r, spinor *const P);
/* solver/cg_her.c:62: f == Qtm_pm_psi on VOLUME/2 sites and f == Q_pm_psi on VOLUME sites run on device-resident fields;
 * any other f runs the same recurrence with f applied through its host-pointer entry point.
 * invert_eo.c:83: with even/odd preconditioning solver_flag CG (:250-276), MIXEDCG (:234-241), RGMIXEDCG (:242-249; delta =
 * solver_params.mcg_delta); without it (even_odd_flag == 0) CG (:505-545).  Other flags terminate with a message. */
int cg_her(spinor *const P, spinor *const Q, const int max_iter, double eps_sq, const int rel_prec, const int N, matrix_mult f);
int invert_eo(spinor *const Even_new, spinor *const Odd_new, spinor *const Even, spinor *const Odd,
              const double precision, const int max_iter, const int solver_flag, const int rel_prec,
              const int sub_evs_flag, const int even_odd_flag, const int no_extra_masses,
              double *const extra_masses, solver_params_t solver_params, const int id,
              const ExternalInverter external_inverter, const SloppyPrecision sloppy,
              const CompressionType compression);
/* operator/Hopping_Matrix_32.c:119; operator/tm_operators_32.c:94; solver/mixed_cg_her.c:65 */
void Hopping_Matrix_32(const int ieo, spinor32 *const l, spinor32 *const k);
void Qtm_pm_psi_32(spinor32 *const l, spinor32 *const k);
/* operator/D_psi.h:28; operator/tm_operators_32.c:141: lexicographic spinor32 fields of VOLUME sites */
void D_psi_32(spinor32 *const P, spinor32 *const Q);
void Q_pm_psi_32(spinor32 *const l, spinor32 *const k);
int mixed_cg_her(spinor *const P, spinor *const Q, solver_params_t solver_params, const int max_iter, double eps_sq,
                 const int rel_prec, const int N, matrix_mult f, matrix_mult32 f32);
/* linalg/square_norm_32.c:95, scalar_prod_r_32.c:109, assign_add_mul_r_32.c:104, assign_mul_add_r_32.c:81, diff_32.c:39,
 * mul_r_32.c:69, assign_mul_add_mul_r_32.c:37; operator/tm_operators_32.c:130 */
float square_norm_32(const spinor32 *const P, const int N, const int parallel);
float scalar_prod_r_32(const spinor32 *const S, const spinor32 *const R, const int N, const int parallel);
void assign_add_mul_r_32(spinor32 *const R, spinor32 *const S, const float c, const int N);
void assign_mul_add_r_32(spinor32 *const R, const float c, const spinor32 *const S, const int N);
void diff_32(spinor32 *const Q, const spinor32 *const R, const spinor32 *const S, const int N);
void mul_r_32(spinor32 *const R, const float c, spinor32 *const S, const int N);
void assign_mul_add_mul_r_32(spinor32 *const R, spinor32 *const S, const float c1, const float c2, const int N);
void gamma5_32(spinor32 *const l, spinor32 *const k, const int V);
/* operator/tm_operators_nd.c:68,:130,:195,:639; solver/cg_her_nd.c:57; invert_doublet_eo.c:68 */
void M_ee_inv_ndpsi(spinor *const l_s, spinor *const l_c, spinor *const k_s, spinor *const k_c, const double mu, const double eps);
void Qtm_ndpsi(spinor *const l_strange, spinor *const l_charm, spinor *const k_strange, spinor *const k_charm);
void Qtm_dagger_ndpsi(spinor *const l_strange, spinor *const l_charm, spinor *const k_strange, spinor *const k_charm);
void Qtm_pm_ndpsi(spinor *const l_strange, spinor *const l_charm, spinor *const k_strange, spinor *const k_charm);
int cg_her_nd(spinor *const P_up, spinor *P_dn, spinor *const Q_up, spinor *const Q_dn, const int max_iter,
              double eps_sq, const int rel_prec, const int N, matrix_mult_nd f);
/* operator/tm_operators_nd_32.c:215; solver/rg_mixed_cg_her_nd.c:182 (the RGMIXEDCG branch of invert_doublet_eo.c:145-149) */
void Qtm_pm_ndpsi_32(spinor32 *const l_strange, spinor32 *const l_charm, spinor32 *const k_strange, spinor32 *const k_charm);
int rg_mixed_cg_her_nd(spinor *const P_up, spinor *const P_dn, spinor *const Q_up, spinor *const Q_dn, solver_params_t solver_params,
                       const int max_iter, const double eps_sq, const int rel_prec, const int N, matrix_mult_nd f, matrix_mult_nd32 f32);
int invert_doublet_eo(spinor *const Even_new_s, spinor *const Odd_new_s, spinor *const Even_new_c, spinor *const Odd_new_c,
                      spinor *const Even_s, spinor *const Odd_s, spinor *const Even_c, spinor *const Odd_c,
                      const double precision, const int max_iter, const int solver_flag, const int rel_prec,
                      solver_params_t solver_params, const ExternalInverter external_inverter,
                      const SloppyPrecision sloppy, const CompressionType compression);

/* ---- remaining members of the operator families ---- */
/* operator/tm_operators.c:587,:723 (mass as an argument), :813, :781, :145 */
void Mee_inv_psi(spinor *const l, spinor *const k, const double mu);
void Mee_psi(spinor *const l, spinor *const k, const double mu);
void mul_one_pm_imu_sub_mul(spinor *const l, spinor *const k, spinor *const j, const double sign, const int N);
void mul_one_sub_mul_gamma5(spinor *const l, spinor *const k, spinor *const j);
void M_minus_1_timesC(spinor *const Even_new, spinor *const Odd_new, spinor *const Even, spinor *const Odd);
/* operator/tm_operators.c:186,:223,:259,:296,:312,:347 and the _nocom forms :179,:194,:252,:267,:304,:366 */
void Qtm_plus_sym_psi(spinor *const l, spinor *const k);
void Qtm_minus_sym_psi(spinor *const l, spinor *const k);
void Mtm_plus_sym_psi(spinor *const l, spinor *const k);
void Mtm_minus_sym_psi(spinor *const l, spinor *const k);
void Mtm_plus_sym_dagg_psi(spinor *const l, spinor *const k);
void Qtm_pm_sym_psi(spinor *const l, spinor *const k);
void Qtm_plus_sym_psi_nocom(spinor *const l, spinor *const k);
void Mtm_plus_sym_psi_nocom(spinor *const l, spinor *const k);
void Mtm_minus_sym_psi_nocom(spinor *const l, spinor *const k);
void Qtm_plus_psi_nocom(spinor *const l, spinor *const k);
void Mtm_plus_psi_nocom(spinor *const l, spinor *const k);
void Qtm_pm_psi_nocom(spinor *const l, spinor *const k);
/* operator/tm_operators.c:471, :390 */
void M_minus_psi(spinor *const l, spinor *const k);
void D_dagg_psi(spinor *const l, spinor *const k);
/* start.c:354; linalg/assign_to_32.c:37,:84; linalg/addto_32.c:16; solver/solver_field.c:31,:67 */
void zero_spinor_field(spinor *const k, const int N);
void assign_to_32(spinor32 *const R, spinor *const S, const int N);
void assign_to_64(spinor *const R, spinor32 *const S, const int N);
void addto_32(spinor *const Q, const spinor32 *const R, const int N);
int init_solver_field(spinor ***const solver_field, const int V, const int nr);
void finalize_solver(spinor **solver_field, const int nr);
/* operator/tm_operators_nd.c:508, :698, :599 */
void H_eo_tm_ndpsi(spinor *const l_strange, spinor *const l_charm, spinor *const k_strange, spinor *const k_charm, const int ieo);
void M_oo_sub_g5_ndpsi(spinor *const l_s, spinor *const l_c, spinor *const k_s, spinor *const k_c, spinor *const j_s,
                       spinor *const j_c, const double mu, const double eps);
void mul_one_pm_iconst(spinor *const l, spinor *const k, const double mu_, const int sign_);
/* solver/rg_mixed_cg_her.c:180 */
int rg_mixed_cg_her(spinor *const P, spinor *const Q, solver_params_t solver_params, const int max_iter, double eps_sq,
                    const int rel_prec, const int N, matrix_mult f, matrix_mult32 f32);

/* ---- HMC side: su3adj.h:25-27, hamiltonian_field.h:28-34 ---- */
typedef struct { double d1, d2, d3, d4, d5, d6, d7, d8; } su3adj;
typedef struct {
  su3 **gaugefield;
  su3adj **momenta;
  su3adj **derivative;
  int update_gauge_copy;
  int traj_counter;
} hamiltonian_field_t;
extern int g_relative_precision_flag;
/* deriv_Sb.c:402 */
void deriv_Sb(const int ieo, spinor *const l, spinor *const k, hamiltonian_field_t *const hf, const double factor);
/* solver/chrono_guess.c:43, :82 */
void chrono_add_solution(spinor *const trial, spinor **const v, int index_array[], const int N, int *_n, const int V);
int chrono_guess(spinor *const trial, spinor *const phi, spinor **const v, int index_array[], const int N, const int n,
                 const int V, matrix_mult f);
/* solver/monomial_solve.c:86 */
int solve_degenerate(spinor *const P, spinor *const Q, solver_params_t solver_params, const int max_iter, double eps_sq,
                     const int rel_prec, const int N, matrix_mult f, int solver_type);
/* monomial/det_monomial.c:47,:150,:202 and monomial/detratio_monomial.c:49,:199,:266 with the signatures of the
 * function pointers in `monomial` (monomial.h:125-127).  Parameters come from a one-time registration (the
 * reference side copies them out of monomial_list[id]); the heatbath noise comes from the caller's
 * random_spinor_field_eo (start.c:284), registered as a callback so that the RNG stream stays the reference's */
typedef void (*tmb_random_spinor_fn)(spinor *const k, const int repro, const int rn_type);
int tmb_dropin_register_monomial(int id, int type, double kappa, double mu, double kappa2, double mu2, int solver,
                                 int maxiter, double forceprec, double accprec, int csg_N);
void tmb_dropin_set_random_spinor_field_eo(tmb_random_spinor_fn fn);
int tmb_dropin_monomial_info(int id, double *energy0, double *energy1, int *iter0, int *iter1);
void det_heatbath(const int id, hamiltonian_field_t *const hf);
double det_acc(const int id, hamiltonian_field_t *const hf);
void det_derivative(const int id, hamiltonian_field_t *const hf);
void detratio_heatbath(const int id, hamiltonian_field_t *const hf);
double detratio_acc(const int id, hamiltonian_field_t *const hf);
void detratio_derivative(const int id, hamiltonian_field_t *const hf);

/* measure_gauge_action.c:46: sum over sites and planes of Re tr(plaquette)/3, on the device copy of gf */
double measure_plaquette(const su3 **const gf);

/* ---- gauge configurations and propagator files: ILDG / SciDAC records in LIME containers.
 *      Types: io/dml.h (DML_Checksum), io/params.h:78-104 (paramsXlfInfo, paramsGaugeInfo). ---- */
typedef struct { unsigned int suma, sumb; } DML_Checksum;
typedef struct {
  char date[64];
  char package_version[32];
  double beta, c2_rec, epsilonbar, kappa, mu, mubar, plaq;
  int counter;
  long int time;
} paramsXlfInfo;
typedef struct {
  double plaquetteEnergy;
  int gaugeRead;
  DML_Checksum checksum;
  char *xlfInfo;
  char *ildg_data_lfn;
} tmb_gauge_info; /* = paramsGaugeInfo */
extern tmb_gauge_info GaugeInfo;          /* io/gauge_read.c:28 */
extern int gauge_precision_read_flag;     /* read_input.h:69: 64 or 32 */
extern int g_disable_IO_checks;           /* global.h:77 */
extern double g_beta, g_rgi_C1;           /* global.h: only printed into xlf-info */
/* io/params_construct_xlfInfo.c; io/gauge.h:32,:35 */
paramsXlfInfo *construct_paramsXlfInfo(double const plaq, int const counter);
int read_gauge_field(char *filename, su3 **const gf);
int write_gauge_field(char *filename, int prec, paramsXlfInfo const *xlfInfo);
/* io/spinor.h:28 (even/odd pair, position-th scidac-binary-data record of the file) */
int read_spinor(spinor *const s, spinor *const r, char *filename, const int position);
/* the records op_write_prop (operator.c:532-605) writes for one flavour, PropInfo.format == 0 */
int tmb_write_propagator(const char *filename, spinor *const s, spinor *const r, int prec, double epssq, int iter,
                         const char *solver_name, int append);

/* ---- include/tmLQCD.h:37-59, wrapper/lib_wrapper.c:77-370 ---- */
typedef struct { unsigned int LX, LY, LZ, T, nstore, nsave, no_operators; } tmLQCD_lat_params;
typedef struct {
  unsigned int nproc, nproc_t, nproc_x, nproc_y, nproc_z, cart_id, proc_id, time_rank, omp_num_threads;
  unsigned int proc_coords[4];
} tmLQCD_mpi_params;
int tmLQCD_invert_init(int argc, char *argv[], const int verbose, const int external_id);
int tmLQCD_read_gauge(const int nconfig);
int tmLQCD_invert(double *const propagator, double *const source, const int op_id, const int write_prop);
int tmLQCD_finalise(void);
int tmLQCD_get_gauge_field_pointer(double **gf);
int tmLQCD_get_mpi_params(tmLQCD_mpi_params *params);
int tmLQCD_get_lat_params(tmLQCD_lat_params *params);
/* The reference fills its operator list from the flex parser (read_input.l); flex and c-lime are
 * not part of this path, so the same information is given programmatically (or through the
 * small key=value reader used by tmLQCD_invert_init on "invert.input"). */
int tmLQCD_b200_set_lattice(int t, int lx, int ly, int lz);
int tmLQCD_b200_add_operator(double kappa, double two_kappa_mu, double eps_sq, int max_iter, int rel_prec);
int tmLQCD_b200_set_theta(double x0, double x1, double x2, double x3);
int tmLQCD_b200_get_solver_info(int op_id, int *iterations, double *reached_prec);
/* the operator's Solver / UseEvenOdd / mcgdelta keys (read_input.l:1108-1139, :967-974, :835-838; defaults operator.c:102-125:
 * CG, even/odd preconditioning, 5e-5): TMB_SOLVER_CG / _MIXEDCG / _RGMIXEDCG with even_odd_flag != 0, TMB_SOLVER_CG with 0;
 * mcg_delta <= 0 keeps the current value.  -1 for a combination invert_eo does not implement here. */
int tmLQCD_b200_set_operator_solver(int op_id, int solver_flag, int even_odd_flag, double mcg_delta);
/* GaugeConfigInputFile (default "conf", default_input_values.h:91) and the propagator output of
 * tmLQCD_invert(..., write_prop != 0): basename (default "source", :93) and precision (default 32, :126) */
int tmLQCD_b200_set_io(const char *gauge_input_filename, const char *prop_basename, int prop_precision);

#endif /* TMLQCD_B200_DROPIN_H */

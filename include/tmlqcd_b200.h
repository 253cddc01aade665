/* tmlqcd_b200.h - C ABI of the B200-native even/odd twisted-mass Wilson-Dirac path.
 *
 * One process drives one GPU.  All entry points are extern "C", take plain pointers and
 * sizes, and return 0 on success or a negative code (tmb_last_error() gives the text).
 * Two levels:
 *
 *  (1) device level (this file): explicit lifecycle, device-resident eo spinor fields
 *      (opaque device pointers from tmb_field_alloc) and operators on them.  This is what a
 *      caller that wants the fields to stay in HBM binds, e.g. a solver or the HMC.
 *  (2) reference level (tmlqcd_b200_dropin.h): the reference's own symbols
 *      (Hopping_Matrix, Qtm_pm_psi, cg_her, invert_eo, ...) with the reference's signatures
 *      and host pointers, implemented on top of (1).
 *
 * Host layouts are the reference's (su3.h:40-63): spinor = 24 doubles (s0..s3 x c0..c2,
 * re/im), su3 = 18 doubles row-major; eo fields hold VOLUME/2 sites in g_lexic2eosub order;
 * the gauge field is g_gauge_field[ix][mu], ix lexicographic (geometry_eo.c:290).
 *
 * Each function cites the reference interface it replaces (file:line in urbach/tmLQCD).
 */
#ifndef TMLQCD_B200_H
#define TMLQCD_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- lifecycle: replaces tmlqcd_mpi_init (mpi_init.c:321-357), init_gauge_field,
 *      init_spinor_field, geometry() (geometry_eo.c:743) for the device side ---- */
int tmb_init(int T, int LX, int LY, int LZ, int device); /* local extents of this rank; LZ and T even */
int tmb_finalize(void);
int tmb_is_initialized(void);
const char *tmb_last_error(void);
int tmb_volume_half(void); /* VOLUME/2 of this rank */

/* multi-GPU: T is split over `nranks` processes, rank r holds global t in [r*T,(r+1)*T).
 * id128 is a 128-byte ncclUniqueId made by rank 0 and distributed by the caller
 * (torch.distributed / MPI).  Replaces g_cart_grid + xchange/ (mpi_init.c:375, xchange_field.c:583). */
int tmb_comm_unique_id(void *id128);
int tmb_comm_init(const void *id128, int nranks, int rank);
/* the same on an (nt x nz) grid of ranks, rank = ct * nz + cz: T split over nt, Z over nz (north_star "T (then Z)"; the z part of
 * the reference's PARALLELXYZT, mpi_init.c:331-357, xchange/xchange_field.c:583-672).  tmb_init takes the local extents T / nt and
 * LZ / nz (even).  With nz > 1 the hopping kernels run unchanged on the slab and a fix-up over the z-face sites swaps the wrapped
 * z term for the neighbour's (the faces are pushed through peer memory when the arenas are mapped, NCCL
 * otherwise); solvers then use un-fused reductions; the fermion force gets a fix-up of the z links of
 * the last-z sites; the plaquette and the two-flavour float solver are T-split only. */
int tmb_comm_init_grid(const void *id128, int nt, int nz, int rank);
int tmb_comm_grid(int *nt, int *nz);
int tmb_comm_loopback(int on); /* single GPU: exercise the T-split path against itself; 1: halo buffers, 2: peer mode */
int tmb_comm_loopback_z(int on); /* single GPU: exercise the Z-split path (face pack, exchange, fix-up) against itself */
int tmb_comm_peer_mode(void);  /* 1 if the hops read the neighbours' fields directly over NVLink (CUDA IPC), 0: NCCL halos */
int tmb_comm_zpeer_mode(void); /* 1 if the z faces of a Z-split grid are pushed into the z neighbours' memory (CUDA IPC), 0: NCCL send/recv */
int tmb_comm_sequence_counts(unsigned int *t_hops, unsigned int *z_pushes); /* test hook: sequence numbers taken by the peer-mode T hops / the z-face pushes */
int tmb_comm_nranks(void);

/* ---- parameters: the globals the reference operators read at call time ---- */
/* boundary(kappa) with X0..X3 (boundary.c:40-55): ka_mu = kappa*exp(i*theta_mu*pi/L_mu^global) */
int tmb_set_boundary(double kappa, const double theta[4]);
int tmb_set_hopping_phases(const double ka_re_im[8]);              /* ka0..ka3 as (re,im) pairs, boundary.c:47-50 */
int tmb_set_mu(double g_mu);                                       /* g_mu = 2*kappa*mu (invert_eo.c:255) */
int tmb_set_nd(double g_mubar, double g_epsbar, double phmc_invmaxev); /* tm_operators_nd.c */
/* kernel configuration knobs (profiling / tuning; defaults are the measured best) */
/* hop_variant: -1 (default) picks the residency per launch (384 resident threads per SM, or 448 when that saves a
 * nearly empty trailing wave on a short launch, e.g. 16^3x32); 0 forces 384; 1..9 are block/occupancy tuning variants
 * of the plain kernel; 10 forces 448.  cache_hints: 1 = L2 eviction policies on the link / spinor loads, 0 = plain
 * loads, -1 (default) = policies only when links + CG vectors exceed the L2 (see eff_hints() in tmb_capi.cu) */
int tmb_set_tuning(int hop_variant, int cache_hints, int xblock);
/* CTA tile traversal of the hopping kernels (default OFF, TMB_TILE=1 in the environment turns it on): the four warps of a
 * CTA take the same 32-site run at 2 x 2 neighbouring (t, x) instead of 128 consecutive sites, which turns neighbour-spinor
 * requests to L2 into L1 hits (24^3x48: L1 hit rate 18.8 -> 24.3 %, L2 -> SM bytes 720 -> 667 MB per hop).  Memory layout and
 * results are unchanged (bit for bit).  Measured: no gain, 1-5 % slower in every configuration (profiles/r02_tile_ab.jsonl) -
 * the kernel is not bound by the L2 -> SM path - hence off; kept selectable like the other tuning variants.  Lattices whose
 * LY*LZ/2 is not a multiple of 32 or whose T is odd always keep the linear traversal. */
int tmb_set_tile(int on);
/* tuning experiment: with tmb_set_overlap bit 1, the hopping kernels bulk-prefetch into L2 the gauge rows of the CTA `ctas`
 * CTAs ahead of them (0: their own rows) */
int tmb_set_prefetch_distance(int ctas);
/* host-pointer Hopping_Matrix pipeline: explicit chunk sizes in time-slices (n = 0: automatic schedule), and a diagnostic
 * run that returns (kind, first slice, microseconds) rows: kind 0 upload done, 1 kernels of a piece done, 2 download done,
 * 3 end of the call */
int tmb_set_host_chunk_sizes(const int *sizes, int n);
int tmb_host_hop_timeline(int ieo, double *l_host, const double *k_host, double *out, int max_rows);
/* two-flavour hop: 2 = the hopping kernel with two flavour groups of warps per CTA (every precision, compression and
 * communication mode, fused <p, A p>); 0 = both flavours in one thread, 1 = lane-paired flavours (one rank, 18-real links, double
 * only); -1 (default) = 0 where it applies (measured faster there), 2 elsewhere */
int tmb_set_hop2_variant(int v);
/* CompressionType of the reference (misc_types.h:33-37): 18 = full links (default), 12 = two rows streamed,
 * third reconstructed in registers (1152 instead of 1536 B/site); refused unless the field is SU(3) to 1e-13 */
int tmb_set_compression(int nreal);
int tmb_set_host_chunks(int n); /* n equal chunks for the pipelined host-pointer Hopping_Matrix; 0 (default): the automatic schedule (a sixth of the field per chunk, first chunk split 1/4 + 3/4, halving tail) */
int tmb_set_p2p_diag(int bits); /* peer-mode timing diagnostics (results INVALID across ranks); refused unless TMB_P2P_DIAG=1 */
int tmb_set_overlap(int flags); /* unknown bits are refused. bit0: programmatic dependent launch, bit1: L2 bulk prefetch of gauge rows, bit2: no CUDA-graph replay in the CG, bit3: L2 prefetch of the epilogue operands (p, dotw), bit4: the CG takes <p, A p> from the last hop (operand load) instead of |Q- p|^2 from the second, bit5: the CG's x / r update as a separate sweep instead of the last hop's epilogue, bit6 (experiment, measured slower): the CG finishes <p, A p> in a one-CTA kernel on the side stream next to the third hop instead of in the last CTA of the second hop */

/* ---- memory ---- */
void *tmb_field_alloc(void);         /* one eo spinor field, VOLUME/2 sites, device SoA layout */
int tmb_field_free(void *field);
int tmb_field_zero(void *field);
void *tmb_host_alloc(size_t bytes);  /* pinned host memory */
int tmb_host_free(void *p);
int tmb_host_register(void *p, size_t bytes); /* pin caller-owned memory (e.g. the reference's calloc slabs) */
int tmb_host_unregister(void *p);
/* host AoS (reference layout) <-> device field; synchronous on return */
int tmb_field_upload(void *field, const double *host_spinors);
int tmb_field_download(double *host_spinors, const void *field);
/* lexicographic full-volume field <-> (even, odd) pair: convert_lexic_to_eo / convert_eo_to_lexic
 * (linalg/convert_eo_to_lexic.c:35-115) fused into the transfer */
int tmb_field_upload_lexic(void *even, void *odd, const double *host_lexic);
int tmb_field_download_lexic(double *host_lexic, const void *even, const void *odd);
/* g_gauge_field upload; what update_backward_gauge (update_backward_gauge.c:185) + xchange_gauge
 * do for the CPU code whenever g_update_gauge_copy is set */
int tmb_gauge_upload(const double *host_gauge);
int tmb_sync(void);
int tmb_timer_start(void);           /* CUDA events on the library's compute stream */
int tmb_timer_stop(float *ms);

/* ---- operators on device fields ---- */
/* Hopping_Matrix(ieo,l,k): operator/Hopping_Matrix.c:131 */
int tmb_Hopping_Matrix(int ieo, void *l, const void *k);
/* Hopping_Matrix_nocom (operator/Hopping_Matrix_nocom.c): no halo exchange - the comm-off leg of benchmark.c:337-373;
 * with a split T the slab wraps onto itself, on one rank it is Hopping_Matrix */
int tmb_Hopping_Matrix_nocom(int ieo, void *l, const void *k);
/* the same on caller-owned HOST buffers (reference AoS layout), transfers pipelined with the kernel
 * in chunks of time-slices; mode 0: Hopping_Matrix, mode 1: tm_times_Hopping_Matrix with cfactor */
int tmb_Hopping_Matrix_host(int ieo, double *l_host, const double *k_host, int mode, double cf_re, double cf_im);
/* tm_times_Hopping_Matrix(ieo,l,k,cfactor): operator/tm_times_Hopping_Matrix.c:119 */
int tmb_tm_times_Hopping_Matrix(int ieo, void *l, const void *k, double cf_re, double cf_im);
/* tm_sub_Hopping_Matrix(ieo,l,p,k,cfactor): operator/tm_sub_Hopping_Matrix.c:122 */
int tmb_tm_sub_Hopping_Matrix(int ieo, void *l, const void *p, const void *k, double cf_re, double cf_im);
/* H_eo_tm_inv_psi / tm_sub_H_eo_gamma5: operator/tm_operators.c:508, :528 */
int tmb_H_eo_tm_inv_psi(void *l, const void *k, int ieo, double sign);
int tmb_tm_sub_H_eo_gamma5(void *l, const void *p, const void *k, int ieo, double sign);
/* operator/tm_operators.c:338, :172, :216, :245, :289, :117, :130 */
int tmb_Qtm_pm_psi(void *l, const void *k);
int tmb_Qtm_plus_psi(void *l, const void *k);
int tmb_Qtm_minus_psi(void *l, const void *k);
int tmb_Mtm_plus_psi(void *l, const void *k);
int tmb_Mtm_minus_psi(void *l, const void *k);
int tmb_M_full(void *even_new, void *odd_new, const void *even, const void *odd);
int tmb_Q_full(void *even_new, void *odd_new, const void *even, const void *odd);
/* D_psi on an (even,odd) pair = one lexicographic field: operator/D_psi_body.c:266 */
int tmb_D_psi_eo(void *even_new, void *odd_new, const void *even, const void *odd);
/* twisted-mass diagonal: mul_one_pm_imu_inv_body.c:1/:44, tm_operators.c:669, :813; gamma.c:77 */
int tmb_assign_mul_one_pm_imu_inv(void *l, const void *k, double sign);
int tmb_assign_mul_one_pm_imu(void *l, const void *k, double sign);
int tmb_mul_one_pm_imu_sub_mul_gamma5(void *l, const void *k, const void *j, double sign);
int tmb_mul_one_pm_imu_sub_mul(void *l, const void *k, const void *j, double sign);
int tmb_gamma5(void *l, const void *k);
/* the generic forms behind the family above (Mee_psi / Mee_inv_psi take mu as an argument, tm_operators.c:587, :723;
 * mul_one_sub_mul_gamma5, :781): l = (z on s0,s1 | conj z on s2,s3) k  and  l = [g5]((z | conj z) k - j) */
int tmb_diag(void *l, const void *k, double z_re, double z_im);
int tmb_diag_sub(void *l, const void *k, const void *j, double z_re, double z_im, int g5);

/* ---- BLAS-1 (linalg/): results of reductions are global sums over ranks ---- */
int tmb_square_norm(const void *p, double *result);                       /* square_norm.c:253 */
int tmb_scalar_prod_r(const void *s, const void *r, double *result);      /* scalar_prod_r.c:135 */
int tmb_assign_add_mul_r(void *p, const void *q, double c);               /* assign_add_mul_r.c:346 */
int tmb_assign_mul_add_r(void *r, double c, const void *s);               /* assign_mul_add_r.c:340 */
int tmb_assign_mul_add_r_and_square(void *r, double c, const void *s, double *result); /* assign_mul_add_r_and_square.c:145 */
int tmb_diff(void *q, const void *r, const void *s);                      /* diff.c:270 */
int tmb_add(void *q, const void *r, const void *s);                       /* add.c */
int tmb_assign(void *r, const void *s);                                   /* assign.c:42 */
int tmb_mul_r(void *r, double c, const void *s);                          /* mul_r.c:40 */

/* ---- solvers, fields stay resident, scalars stay on the device ---- */
/* cg_her(P,Q,max_iter,eps_sq,rel_prec,VOLUME/2,&Qtm_pm_psi): solver/cg_her.c:62.
 * Returns the iteration count, -1 if not converged (cg_her.c:141), < -1 on a CUDA error. */
int tmb_cg_her(void *P, const void *Q, int max_iter, double eps_sq, int rel_prec);
/* the CG branch of invert_eo: invert_eo.c:152-157, :252, :268-270, :306-310 */
int tmb_invert_eo(void *even_new, void *odd_new, const void *even, const void *odd, double precision, int max_iter,
                  int rel_prec);
/* last solve: iterations, final |r|^2 as seen by the CG, seconds spent in the CG loop */
int tmb_solver_stats(int *iterations, double *final_err, double *seconds);

/* ---- single precision operator + mixed-precision CG ----
 * float fields hold VOLUME/2 spinor32 (su3.h:75-78) in the same SoA layout with float2 elements */
void *tmb_field32_alloc(void);                                   /* free with tmb_field_free */
int tmb_field32_upload(void *field32, const float *host_spinor32);
int tmb_field32_download(float *host_spinor32, const void *field32);
int tmb_assign_to_32(void *field32, const void *field64);        /* linalg/assign_to_32.c */
int tmb_assign_to_64(void *field64, const void *field32);
int tmb_Hopping_Matrix_32(int ieo, void *l32, const void *k32);  /* operator/Hopping_Matrix_32.c:119 */
int tmb_Qtm_pm_psi_32(void *l32, const void *k32);               /* operator/tm_operators_32.c:94 */
/* D_psi_32 on an (even, odd) pair of float fields (operator/D_psi.h:28); M_full_32 with g5 != 0 gives the Q_full rows
 * that Q_pm_psi_32 (tm_operators_32.c:141) is built from */
int tmb_D_psi_eo_32(void *even_new32, void *odd_new32, const void *even32, const void *odd32);
int tmb_M_full_32(void *even_new32, void *odd_new32, const void *even32, const void *odd32, int g5);
int tmb_field32_upload_lexic(void *even32, void *odd32, const float *host_lexic32);
int tmb_field32_download_lexic(float *host_lexic32, const void *even32, const void *odd32);
int tmb_set_mixcg(double innereps, int maxinnersolverit);        /* mixcg_innereps / mixcg_maxinnersolverit, default_input_values.h:193 */
/* mixed_cg_her(P,Q,params,max_iter,eps_sq,rel_prec,VOLUME/2,&Qtm_pm_psi,&Qtm_pm_psi_32): solver/mixed_cg_her.c:65 */
int tmb_mixed_cg_her(void *P, const void *Q, int max_iter, double eps_sq, int rel_prec);
/* invert_eo with solver_flag == MIXEDCG: invert_eo.c:234-241 */
int tmb_invert_eo_mixed(void *even_new, void *odd_new, const void *even, const void *odd, double precision,
                        int max_iter, int rel_prec);
/* invert_eo with solver_flag == RGMIXEDCG: invert_eo.c:242-249 (delta of the reliable updates: tmb_set_mcg_delta) */
int tmb_invert_eo_rgmixed(void *even_new, void *odd_new, const void *even, const void *odd, double precision,
                          int max_iter, int rel_prec);

/* ---- non-degenerate doublet: operator/tm_operators_nd.c:68,:130,:195,:639; cg_her_nd.c:57;
 *      invert_doublet_eo.c:68 ---- */
int tmb_M_ee_inv_ndpsi(void *ls, void *lc, const void *ks, const void *kc, double mu, double eps);
int tmb_M_oo_sub_g5_ndpsi(void *ls, void *lc, const void *ks, const void *kc, const void *js, const void *jc,
                          double mu, double eps); /* tm_operators_nd.c:698 */
int tmb_Qtm_ndpsi(void *ls, void *lc, const void *ks, const void *kc);
int tmb_Qtm_dagger_ndpsi(void *ls, void *lc, const void *ks, const void *kc);
int tmb_Qtm_pm_ndpsi(void *ls, void *lc, const void *ks, const void *kc);
int tmb_cg_her_nd(void *Pup, void *Pdn, const void *Qup, const void *Qdn, int max_iter, double eps_sq, int rel_prec);
int tmb_invert_doublet_eo(void *ens, void *ons, void *enc, void *onc, const void *es, const void *os,
                          const void *ec, const void *oc, double precision, int max_iter, int rel_prec);
/* the same with invert_doublet_eo's solver_flag (invert_doublet_eo.c:145-156): 14 = RGMIXEDCG -> rg_mixed_cg_her_nd, else cg_her_nd */
int tmb_invert_doublet_eo_solver(void *ens, void *ons, void *enc, void *onc, const void *es, const void *os,
                                 const void *ec, const void *oc, double precision, int max_iter, int rel_prec, int solver_flag);
/* Qtm_pm_ndpsi_32 on float fields (operator/tm_operators_nd_32.c:215) and rg_mixed_cg_her_nd (solver/rg_mixed_cg_her_nd.c:182;
 * delta = tmb_set_mcg_delta); counts of the last reliable-update solve: float inner, double inner, outer iterations */
int tmb_Qtm_pm_ndpsi_32(void *ls32, void *lc32, const void *ks32, const void *kc32);
int tmb_rg_mixed_cg_her_nd(void *Pup, void *Pdn, const void *Qup, const void *Qdn, int max_iter, double eps_sq, int rel_prec);
int tmb_solver_stats_rg(int *inner_sp, int *inner_dp, int *outer);

/* rg_mixed_cg_her (solver/rg_mixed_cg_her.c:180), the default mixed solver of solve_degenerate; delta = solver_params.mcg_delta */
int tmb_set_mcg_delta(double delta);
int tmb_rg_mixed_cg_her(void *P, const void *Q, int max_iter, double eps_sq, int rel_prec);

/* ---- HMC side: fermion force, chronological guess, DET / DETRATIO monomials ---- */
/* hf->derivative (hamiltonian_field.h:30): host layout [ix][mu][8] doubles = su3adj d1..d8 (su3adj.h:25-27),
 * ix lexicographic; it lives on the device between the calls below */
int tmb_derivative_zero(void);
int tmb_derivative_upload(const double *host_df);
int tmb_derivative_download(double *host_df);
/* deriv_Sb(ieo, l, k, hf, factor): deriv_Sb.c:402; l has parity ieo, k the other; accumulates into the device
 * derivative field; with a T split the halo exchange of xchange_2fields (deriv_Sb.c:413) is done inside */
int tmb_deriv_Sb(int ieo, const void *l, const void *k, double factor);
/* complex BLAS-1 used by the chronological guess: linalg/scalar_prod_body.c, assign_add_mul.c, assign_diff_mul.c:31, mul.c */
int tmb_scalar_prod(const void *s, const void *r, double *re, double *im);
int tmb_assign_add_mul(void *r, const void *s, double c_re, double c_im);
int tmb_assign_diff_mul(void *r, const void *s, double c_re, double c_im);
int tmb_mul(void *r, double c_re, double c_im, const void *s);
/* solver/solver_types.h:23-49: the solver ids on the scoped path */
#define TMB_SOLVER_CG 1
#define TMB_SOLVER_MIXEDCG 13
#define TMB_SOLVER_RGMIXEDCG 14
/* matrix_mult selector for device-level callers (the reference passes a function pointer) */
enum { TMB_OP_QTM_PM = 0, TMB_OP_QTM_PLUS = 1, TMB_OP_QTM_MINUS = 2 };
/* chrono_add_solution / chrono_guess: solver/chrono_guess.c:43, :82; v = array of N device fields */
int tmb_chrono_add_solution(const void *trial, void *const *v, int *index_array, int N, int *n);
int tmb_chrono_guess(void *trial, const void *phi, void *const *v, const int *index_array, int N, int n, int op);
/* solve_degenerate(P, Q, params, max_iter, eps_sq, rel_prec, VOLUME/2, &Qtm_pm_psi, solver): solver/monomial_solve.c:86;
 * solver_type as solver/solver_types.h:23-49: CG = 1, MIXEDCG = 13, RGMIXEDCG = 14 */
int tmb_solve_degenerate(void *P, const void *Q, int max_iter, double eps_sq, int rel_prec, int solver_type);
/* monomials: type DET = 0, DETRATIO = 1 (monomial.h:27-28); parameters as the BeginMonomial block sets them */
enum { TMB_MNL_DET = 0, TMB_MNL_DETRATIO = 1 };
int tmb_monomial_add(int type, double kappa, double mu, double kappa2, double mu2, int solver, int maxiter,
                     double forceprec, double accprec, int csg_N); /* returns the monomial id */
int tmb_monomial_clear(void);
int tmb_set_relative_precision_flag(int flag);                      /* g_relative_precision_flag */
/* det_heatbath / detratio_heatbath (det_monomial.c:150, detratio_monomial.c:199); `gauss` = device field holding what
 * random_spinor_field_eo(w_fields[0], rngrepro, RN_GAUSS) drew (the RNG stays with the caller) */
int tmb_monomial_heatbath(int id, const void *gauss, double *energy0);
/* det_derivative / detratio_derivative (det_monomial.c:47, detratio_monomial.c:49): adds to the device derivative field */
int tmb_monomial_derivative(int id);
/* det_acc / detratio_acc (det_monomial.c:202, detratio_monomial.c:266): *dH = energy1 - energy0 */
int tmb_monomial_acc(int id, double *dH);
int tmb_monomial_info(int id, double *energy0, double *energy1, int *iter0, int *iter1, int *csg_n);
void *tmb_monomial_pf(int id);     /* the pseudo-fermion field (device) */
void *tmb_monomial_wfield(int k);  /* w_fields[k], k < 6 (device) */

/* number of kernels this library launched since tmb_init (bench.py's gpu_launches) */
/* single-precision BLAS-1 on float fields (the _32.c files of linalg/).  op: 0 R += c1 S1 (assign_add_mul_r_32), 1 R = c1 R + S1
 * (assign_mul_add_r_32), 2 R = S1 - S2 (diff_32), 3 R = c1 S1 (mul_r_32), 4 R = c1 R + c2 S1 (assign_mul_add_mul_r_32),
 * 5 R = gamma5 S1 (gamma5_32) */
int tmb_blas32(int op, void *r, const void *s1, const void *s2, double c1, double c2);
int tmb_square_norm_32(const void *field32, double *result);
int tmb_scalar_prod_r_32(const void *a32, const void *b32, double *result);
/* measure_plaquette (measure_gauge_action.c:46): sum over all sites of all ranks and the 6 planes of Re tr(P)/3;
 * the average plaquette is result / (6 * VOLUME * nranks) */
int tmb_measure_plaquette(double *result);
long long tmb_launch_count(void);
/* measurement aid: sustained device-to-device copy bandwidth (read + write), GB/s, `reps` copies of `bytes` */
int tmb_measure_copy_gbs(size_t bytes, int reps, double *gbs);
/* measurement aid: pinned-memory host link, GB/s per direction: upload alone, download alone, both at once */
int tmb_measure_pcie_gbs(size_t bytes, int reps, double *h2d, double *d2h, double *duplex_each);

#ifdef __cplusplus
}
#endif
#endif /* TMLQCD_B200_H */

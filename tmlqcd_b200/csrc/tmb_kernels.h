/* tmb_kernels.h - internal C++ interface between the C-ABI layer (tmb_capi.cu) and the
 * CUDA kernels (tmb_kernels.cu).  Not installed; the public interface is include/tmlqcd_b200.h.
 * Field pointers are void*: `prec` (0 = double2 elements, 1 = float2 elements) selects the
 * instantiation, reductions always accumulate and land in double. */
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "tmb_geom.h"

/* Device-resident scalar block of the CG (solver/cg_her.c:62-143 keeps these on the host;
 * here they never leave HBM during the iteration, kernels read them directly). */
struct tmb_cg_state {
  double normsq;   /* (r,r) of the previous iteration */
  double pro;      /* (p, A p) */
  double err;      /* (r,r) new */
  double alpha, beta;
  double sqnorm_q; /* |Q|^2, for rel_prec */
  double eps_sq;
  double inner_eps; /* mixed CG: inner stop err <= inner_eps * sqnrm0 (mixed_cg_her.c:141) */
  double sqnrm0;    /* mixed CG: |delta|^2 at the start of the inner solve */
  double tmp[4];   /* landing slots of reductions (all-reduced in place when nranks > 1) */
  int rel_prec;
  int converged;   /* set on device when the stop test fires; later kernels become no-ops */
  int iter;        /* iterations completed */
  int max_iter;    /* mixed CG inner loop: also stop after this many iterations */
  unsigned int ticket[4]; /* last-CTA election of the fused reduction finish, one per slot */
  int fprec;       /* 1: scalars rounded to float like the reference's float inner loops (rg_mixed_cg_her.c:107-111) */
};

enum tmb_fin_op {
  TMB_FIN_STORE = 0,    /* tmp[slot] = sum */
  TMB_FIN_CG_PRO = 1,   /* pro = sum; alpha = normsq/pro              cg_her.c:93-94 */
  TMB_FIN_CG_ERR = 2,   /* err = sum; iter++; stop test; beta; normsq cg_her.c:101-126 */
  TMB_FIN_CG_INIT = 3,  /* normsq = sum                               cg_her.c:88 */
  TMB_FIN_MCG_ERR = 4,  /* inner loop of mixed_cg_her.c:139-150: its four-way stop test */
  TMB_FIN_RG_ERR = 5    /* inner loops of rg_mixed_cg_her.c:75-104 / :107-145: inner_eps = delta, sqnrm0 = rhomax */
};

/* Cross-rank sum inside the reduction finish (peer mode): the CTA that finishes a rank's two-stage reduction stores its
 * partial into EVERY rank's landing array over NVLink (st.release.sys of a sequence word behind it), waits for the
 * nranks words of its own array and adds the values in rank order - the same order on every rank, so all ranks hold
 * bit-identical sums (and take identical stop decisions).  Replaces final_kernel + ncclAllReduce(1 double) +
 * apply_kernel: no extra launches and no NCCL in the CG.  Sequence numbers come from a device counter that only
 * performed reductions advance (kernels that return early after the CG's stop do not), ring of TMB_XR_RING slots. */
#define TMB_XR_MAXR 8
#define TMB_XR_RING 4
struct tmb_xred_table {            /* device memory, written once at tmb_comm_init */
  int nranks, rank;
  double *val[TMB_XR_MAXR];        /* rank q's landing array [RING][MAXR] (peer memory; [rank] = own) */
  unsigned int *seq[TMB_XR_MAXR];  /* rank q's sequence words [RING][MAXR] */
  unsigned int *ctr;               /* this rank's reduction counter */
  int *err;                        /* set to 1 when a wait timed out */
};

struct tmb_hop_launch {
  const void *in; void *out; const void *p; const void *dotw;
  /* nfl == 2 (non-degenerate doublet): the second flavour's fields and the 2x2 flavour mixing of the epilogue
   * (tm_operators_nd.c:639-756): mode 1 out_f = nrm[(1 -+ i mu g5) H in_f + eps H in_f'], mode 2 out_f = scale g5[(1 -+ i mu g5) p_f
   * + eps p_f' - H in_f], flavour 0 with the upper sign; dot == 2 accumulates dot_scale (|out|^2 + |out1|^2) */
  /* mode 4 (CG tail, one field, dot == 2): out is not written; A p = g5((cf|conj cf) p - H in) feeds cg_x += alpha cg_p,
   * cg_r -= alpha A p in registers (alpha = st_fin->alpha) and the reduction is |cg_r|^2 */
  void *cg_x, *cg_r; const void *cg_p;
  int nfl; const void *in1; void *out1; const void *p1; double nd_mu, nd_eps, nd_scale, dot_scale;
  const void *in_up1, *in_dn1; /* peer mode: the neighbours' copies of in1 */
  const void *U; const void *halo_up, *halo_dn, *Uhalo;
  double *partial;          /* fused-dot block partials (DOT) */
  const tmb_cg_state *st;   /* if non-null: kernel exits immediately when st->converged */
  tmb_geom g;
  int par;                  /* parity of the OUTPUT sites = ieo of Hopping_Matrix(ieo,l,k) */
  double2 ka[4]; double2 cf;
  int prec;                 /* 0: double, 1: float */
  int mode;                 /* epilogue 0..3, see tmb_site.cuh */
  int dist;                 /* 1: +-t of the boundary slices from halo buffers */
  int dot;                  /* 1: accumulate Re<dotw, out> into partial[]; 2: accumulate |out|^2 (no operand) */
  int hints;                /* 1: L1/L2 cache-policy loads */
  int variant;              /* 0: production kernel; 1..9: tuning variants of the plain kernel */
  int site0, nsites;        /* contiguous work range ... */
  int split, gap;           /* ... with a hole: i = site0 + w + (w >= split ? gap : 0) */
  int xblock;               /* >0: traverse (t,x) planes x-blocked for L2 locality */
  int tile;                 /* 1: CTA tile traversal 2 x 2 x 32 (tmb_tile_site, tmb_geom.h); whole-lattice launches of 128-site CTAs only */
  int pdl;                  /* launch with programmatic stream serialization (PDL) */
  int prefetch;             /* bulk-prefetch the CTA's gauge rows into L2 before the dependency wait */
  int prefetch_dist;        /* ... the rows of the CTA this many CTAs AHEAD instead (linear traversal only; 0: own rows) */
  int recon12;              /* U / Uhalo hold 12-real compressed links (6 complex per link) */
  /* peer mode (dist == 2): ONE launch per hop.  Its first p2p_copy_ctas CTAs pull the projected boundary
   * time-slices out of the neighbours' copies of `in` (in_up / in_dn: peer memory over NVLink) into halo_up /
   * halo_dn while the other CTAs work through the interior; the CTAs of the two boundary slices come last.
   * flags (own memory, written by the peers): [0] ready from rank+1, [1] ready from rank-1, [2] done from
   * rank+1, [3] done from rank-1; up_flags / dn_flags are the neighbours' arrays seen from here. */
  const void *in_up, *in_dn;
  void *halo_up_w, *halo_dn_w;
  const unsigned int *seq_base; unsigned int seq_off; /* this hop's sequence number is *seq_base + seq_off (a CUDA graph
                             * of a CG chunk carries fixed offsets; its last node advances the base) */
  int p2p_copy_ctas;
  unsigned int *flags, *up_flags, *dn_flags;
  unsigned int *p2p_copied; /* device-local: [0] pull CTAs finished (counter), [1] halo_ready (= seq) */
  int *p2p_err;             /* set to 1 if a flag wait timed out (deadlock guard) */
  int p2p_nohandshake;      /* timing diagnostic: skip the end-of-hop handshake (NOT safe) */
  int p2p_diag;             /* timing diagnostics (results invalid): 4 no pull, 8 no halo path in boundary CTAs, 16 natural slice order */
  /* fused finish of the DOT reduction (fin_op >= 0): the last of fin_total CTAs sums partial_base[0..fin_total) */
  tmb_cg_state *st_fin; const double *partial_base; int fin_op, fin_slot, fin_total;
  const tmb_xred_table *xr; /* non-null: the finish sums over ranks through peer memory */
};

int tmb_hop_grid(const tmb_hop_launch &a);
cudaError_t tmb_launch_hop(const tmb_hop_launch &a, cudaStream_t s);
cudaError_t tmb_launch_hop_nd(const tmb_hop_launch &a, cudaStream_t s); /* nfl == 2 instantiations (tmb_hop2.cu) */

/* reductions: block partials -> one scalar in st->tmp[slot] (+ optional CG bookkeeping) */
cudaError_t tmb_launch_final(const double *partial, int n, tmb_cg_state *st, int slot, int op, int apply,
                             const tmb_xred_table *xr, cudaStream_t s);
cudaError_t tmb_launch_final_hop(const double *partial, int n, tmb_cg_state *st, int slot, int op, int apply,
                             const tmb_xred_table *xr, cudaStream_t s);
cudaError_t tmb_launch_seq_bump(unsigned int *base, unsigned int n, cudaStream_t s); /* *base += n */
cudaError_t tmb_launch_apply(tmb_cg_state *st, int slot, int op, cudaStream_t s);
int tmb_red_grid(size_t n2);
cudaError_t tmb_launch_norm2(int prec, const void *a, size_t n2, double *partial, cudaStream_t s);
cudaError_t tmb_launch_dot(int prec, const void *a, const void *b, size_t n2, double *partial, cudaStream_t s);
cudaError_t tmb_launch_dot_fin(int prec, const void *a, const void *b, size_t n2, double *partial, tmb_cg_state *st, int slot,
                               int op, const tmb_xred_table *xr, cudaStream_t s);
cudaError_t tmb_launch_xpay_norm(double2 *r, double c, const double2 *sv, size_t n2, double *partial, cudaStream_t s);
cudaError_t tmb_launch_cg_update_xr(int prec, void *x, void *r, const void *p, const void *ap, size_t n2,
                                    tmb_cg_state *st, double *partial, int fin_slot, int fin_op, const tmb_xred_table *xr,
                                    cudaStream_t s);
cudaError_t tmb_launch_cg_update_p(int prec, void *p, const void *r, size_t n2, const tmb_cg_state *st, cudaStream_t s);

/* elementwise, n2 = 12*Vh complex elements; `half` = 6*Vh separates spin 0,1 from spin 2,3 */
cudaError_t tmb_launch_axpy(double2 *p, const double2 *q, double c, size_t n2, cudaStream_t s);
cudaError_t tmb_launch_xpay(double2 *r, double c, const double2 *sv, size_t n2, cudaStream_t s);
cudaError_t tmb_launch_lincomb(double2 *q, double a, const double2 *r, double b, const double2 *sv, size_t n2, cudaStream_t s);
cudaError_t tmb_launch_scale(double2 *r, double c, const double2 *sv, size_t n2, cudaStream_t s);
cudaError_t tmb_launch_gamma5(double2 *l, const double2 *k, size_t n2, size_t half, cudaStream_t s);
cudaError_t tmb_launch_diag(double2 *l, const double2 *k, double2 z, size_t n2, size_t half, cudaStream_t s);
/* l = (g5?) ( (z|conj z) k - j ) */
cudaError_t tmb_launch_diag_sub(double2 *l, const double2 *k, const double2 *j, double2 z, int g5, size_t n2,
                                size_t half, cudaStream_t s);
cudaError_t tmb_launch_nd_mee_inv(double2 *ls, double2 *lc, const double2 *ks, const double2 *kc, double mu,
                                  double eps, size_t n2, size_t half, cudaStream_t s);
cudaError_t tmb_launch_nd_moo_sub_g5(double2 *ls, double2 *lc, const double2 *ks, const double2 *kc,
                                     const double2 *js, const double2 *jc, double mu, double eps, size_t n2,
                                     size_t half, cudaStream_t s);
/* complex BLAS-1 (chronological guess): R += c S, R = c S, <S,R> -> partial[0..grid) re, partial[grid..2grid) im */
cudaError_t tmb_launch_caxpy(double2 *r, double2 c, const double2 *sv, size_t n2, cudaStream_t s);
cudaError_t tmb_launch_cscale(double2 *r, double2 c, const double2 *sv, size_t n2, cudaStream_t s);
cudaError_t tmb_launch_cdot(const double2 *a, const double2 *b, size_t n2, double *partial, cudaStream_t s);
/* precision conversion (linalg/assign_to_32.c, addto_32.c): n complex elements */
cudaError_t tmb_launch_to_float(float2 *dst, const double2 *src, size_t n, cudaStream_t s);
cudaError_t tmb_launch_add_from_float(double2 *dst, const float2 *src, size_t n, cudaStream_t s); /* dst += src */

/* layout conversion between the reference's host AoS layouts and the device SoA layout */
cudaError_t tmb_launch_pack_eo(double2 *soa, const double2 *aos, int Vh, cudaStream_t s);
cudaError_t tmb_launch_unpack_eo(double2 *aos, const double2 *soa, int Vh, cudaStream_t s);
cudaError_t tmb_launch_pack_eo_range(double2 *soa, const double2 *aos, int Vh, int i0, int n, cudaStream_t s);
cudaError_t tmb_launch_unpack_eo_range(double2 *aos, const double2 *soa, int Vh, int i0, int n, cudaStream_t s);
cudaError_t tmb_launch_pack_lexic(double2 *even, double2 *odd, const double2 *lex, tmb_geom g, cudaStream_t s);
cudaError_t tmb_launch_unpack_lexic(double2 *lex, const double2 *even, const double2 *odd, tmb_geom g, cudaStream_t s);
cudaError_t tmb_launch_pack_lexic_f(float2 *even, float2 *odd, const float2 *lex, tmb_geom g, cudaStream_t s);
cudaError_t tmb_launch_unpack_lexic_f(float2 *lex, const float2 *even, const float2 *odd, tmb_geom g, cudaStream_t s);
cudaError_t tmb_launch_pack_gauge(double2 *U, const double2 *lex, tmb_geom g, cudaStream_t s);
/* T-face half-spinors: send_up = (1-g0) proj of the last slice, send_dn = (1+g0) proj of the first */
cudaError_t tmb_launch_pack_halo(int prec, void *send_up, void *send_dn, const void *in, tmb_geom g, cudaStream_t s);
/* 12-real compression: dst[l][6][n] = rows 0,1 of src[l][9][n]; su3_defect: max |row2 - conj(row0 x row1)|^2 */
cudaError_t tmb_launch_compress12(double2 *dst, const double2 *src, size_t n, int nlinks, cudaStream_t s);
cudaError_t tmb_launch_su3_defect(const double2 *U, size_t n, int nlinks, double *partial, cudaStream_t s);
/* Uhalo[q][e][j] = U[q][0][e][(T-1)S + j] : what rank+1 needs from this rank */
cudaError_t tmb_launch_pack_gauge_halo(double2 *out, const double2 *U, tmb_geom g, cudaStream_t s);

/* ---- second split direction (Z), see tmb_site.cuh: Sz = T*LX*LY/2 face entries per parity ----
 * pack_zfaces: send_up[6][Sz] = (1 - g3)-type projection (hop direction -z) of this rank's last-z sites of `in` (parity pin), for
 * rank z+1; send_dn[6][Sz] = direction +z projection of its first-z sites, for rank z-1.  pack_gauge_zhalo: U_z of the last-z
 * sites, [2][9][Sz], for rank z+1.  zfix: replaces the wrapped z term of the face sites of `out` by the halo term. */
cudaError_t tmb_launch_pack_zfaces(int prec, void *send_up, void *send_dn, const void *in, tmb_geom g, int pin, cudaStream_t s);
cudaError_t tmb_launch_pack_gauge_zhalo(int prec, void *out, const void *U, tmb_geom g, cudaStream_t s);
/* peer-mode Z exchange: second halo buffers, the two flags of this rank and the hop's sequence number (see tmb_kernels.cu) */
struct tmb_zpeer { const void *hz_up1, *hz_dn1; const unsigned int *flags; const unsigned int *seq_base; unsigned int seq_off; int *err;
  /* every mode: this rank's own projected faces and own last-z links as contiguous arrays (tmb_zfix_side, tmb_site.cuh); may be null */
  const void *own_up, *own_dn, *Uzl; };
cudaError_t tmb_launch_pack_zfaces_push(int prec, void *up0, void *up1, void *dn0, void *dn1, void *own_up, void *own_dn, const void *in,
                                        tmb_geom g, int pin, const unsigned int *seq_base, unsigned int seq_off, unsigned int *flag_up,
                                        unsigned int *flag_dn, cudaStream_t s);
cudaError_t tmb_launch_zfix(int prec, int mode, void *out, const void *in, const void *U, const void *hz_up, const void *hz_dn, const void *Uzh,
                            tmb_geom g, int par, double2 ka3, double2 cf, const tmb_cg_state *st, const tmb_zpeer *w, cudaStream_t s);

/* ---- fermion force (tmb_force.cu): deriv_Sb.c:402-649 as a gather over link owners ---- */
struct tmb_deriv_launch {
  const void *l, *k, *U; double *df;
  const void *halo_k, *halo_l;
  tmb_geom g;
  int ieo;                 /* parity of l (the field that gets the gamma5) */
  int dist;                /* 1: +t remote of slice T-1 from the halo buffers */
  int t0, nt;              /* time-slices [t0, t0+nt) of both parities */
  double2 ka[4]; double c; /* c = 2*factor */
};
cudaError_t tmb_launch_deriv(const tmb_deriv_launch &a, cudaStream_t s);
cudaError_t tmb_launch_deriv_zfix(const tmb_deriv_launch &a, const double2 *halo_k, const double2 *halo_l, cudaStream_t s); /* Z split */
/* first time-slice of k and of l, (1+g0)-projected: out[0][6][S] from k, out[1][6][S] from l */
cudaError_t tmb_launch_pack_deriv_halo(double2 *out, const double2 *k, const double2 *l, tmb_geom g, cudaStream_t s);
/* hf->derivative host layout [ix][mu][8] (init/init_moment_field.c:62-80) <-> device [2][4][8][Vh]; mode 0: set, 1: add */
cudaError_t tmb_launch_pack_deriv(double *dev, const double *lex, tmb_geom g, cudaStream_t s);
cudaError_t tmb_launch_unpack_deriv(double *lex, const double *dev, tmb_geom g, int add, cudaStream_t s);

/* single-precision BLAS-1 (the _32.c files of linalg/), op codes in tmb_kernels.cu */
cudaError_t tmb_launch_blas32(int op, float2 *r, const float2 *sv, const float2 *s2, float c1, float c2, size_t n2, size_t half, cudaStream_t s);
/* ---- plaquette (tmb_force.cu): measure_gauge_action.c:46-106; partial[] gets one sum per CTA (grid from tmb_plaq_grid) ---- */
int tmb_plaq_grid(const tmb_geom &g);
cudaError_t tmb_launch_plaquette(const double2 *U, const double2 *Uup, tmb_geom g, int dist, double *partial, cudaStream_t s);
/* spatial links of the first time-slice, [2][3][9][S]: what the rank below needs for its last slice's t-x, t-y, t-z plaquettes */
cudaError_t tmb_launch_pack_gauge_first_slice(double2 *out, const double2 *U, tmb_geom g, cudaStream_t s);

/* ---- two-flavour hopping (non-degenerate doublet): one gauge stream for both flavours ----
 * mode 0: (out0, out1) = (H in0, H in1)
 * mode 1: (out0, out1) = M_ee_inv_nd(H in0, H in1; mu, eps) with the flavour roles as tmb_launch_nd_mee_inv:
 *         out0 = "ls" from (ks = H in0, kc = H in1)
 * mode 2: (out0, out1) = g5( M_oo(p0, p1; mu, eps) - (H in0, H in1) ), the M_oo_sub_g5_ndpsi epilogue (:698-756),
 *         then scaled by `scale` */
struct tmb_hop2_launch {
  const void *in0, *in1; void *out0, *out1; const void *p0, *p1; const void *U;
  tmb_geom g; int par; double2 ka[4];
  int mode; double mu, eps, scale; int hints;
  int variant; /* 0: one thread carries both flavours (hop2_kernel, default), 1: lane-paired flavours (hop2p_kernel), 3: as 0 with 3 CTAs per SM (168 registers, spills) */
  int prec;    /* 1: float fields and float links (hop2_kernel only) */
  int tile;    /* 1: CTA tile traversal (tmb_tile_site); hop2_kernel only */
  int prefetch, prefetch_dist; /* hop2_kernel: L2 bulk prefetch of the gauge rows of the CTA prefetch_dist CTAs ahead */
  /* hop2_kernel, mode 2 only: dot == 2 accumulates dot_scale (|out0|^2 + |out1|^2) into partial[]; fin_op >= 0: fused finish */
  int dot; double dot_scale; double *partial; const tmb_cg_state *st; tmb_cg_state *st_fin; int fin_op, fin_slot; const tmb_xred_table *xr;
};
cudaError_t tmb_launch_hop2(const tmb_hop2_launch &a, cudaStream_t s);
int tmb_hop2_grid(const tmb_hop2_launch &a);
int tmb_hop_block(const tmb_hop_launch &a);

/* tmb_kernels.cu - hand-written CUDA kernels (sm_100a) for tmLQCD's even/odd twisted-mass path.
 *
 * K1  hop_kernel        Hopping_Matrix / tm_times_ / tm_sub_ (operator/Hopping_Matrix.c:131,
 *                       tm_times_Hopping_Matrix.c:119, tm_sub_Hopping_Matrix.c:122) with the
 *                       epilogue and, for the CG, the <p,Ap> reduction fused in; double and,
 *                       for the mixed CG, single precision (operator/Hopping_Matrix_32.c:99)
 * K3  elementwise       linalg/ BLAS-1 + twisted-mass diagonal (tm_operators.c, tm_operators_nd.c)
 * K4  reductions        square_norm / scalar_prod_r / assign_mul_add_r_and_square, two-stage:
 *                       warp shuffle + block partial, then one CTA finishing and doing the CG
 *                       scalar bookkeeping on the device.
 * Bandwidth-bound stencil: no tensor cores.  One thread per output site, vector loads that are
 * contiguous across the warp (SoA), gauge streamed with L1::no_allocate + L2 evict-first.
 */
#include "tmb_kernels.h"
#include "tmb_site.cuh"

#include "tmb_hop.cuh"

int tmb_hop_block(const tmb_hop_launch &a) { /* threads per CTA of the kernel tmb_launch_hop will pick */
  return a.prec ? TMB_HOP_BLOCK_F : hop_variant_block(a.variant);
}
int tmb_hop_grid(const tmb_hop_launch &a) {
  const int b = (a.prec ? TMB_HOP_BLOCK_F : hop_variant_block(a.variant)) / (a.nfl == 2 ? 2 : 1); /* sites per CTA */
  return (a.nsites + b - 1) / b + (a.dist == 2 ? a.p2p_copy_ctas : 0);
}

/* HINTS is a configuration mask: bit 0 cache-policy loads, bit 1 12-real links (see tmb_site.cuh) */
template <int DIST, int HINTS>
static cudaError_t hop_mode(const tmb_hop_launch &a, cudaStream_t s);
template <int DIST, int CFG>
static cudaError_t hop_mode_f(const tmb_hop_launch &a, cudaStream_t s);
template <int HINTS>
static cudaError_t hop_dist(const tmb_hop_launch &a, cudaStream_t s) {
  if (a.dist == 2) return hop_mode<2, HINTS>(a, s);
  return a.dist ? hop_mode<1, HINTS>(a, s) : hop_mode<0, HINTS>(a, s);
}
template <int CFG>
static cudaError_t hop_dist_f(const tmb_hop_launch &a, cudaStream_t s) {
  if (a.dist == 2) return hop_mode_f<2, CFG>(a, s);
  return a.dist ? hop_mode_f<1, CFG>(a, s) : hop_mode_f<0, CFG>(a, s);
}
/* Variant 10: 64 threads x 7 CTAs per SM = 448 resident threads (144 registers) instead of 384 (168 registers), for
 * ALL epilogues of the plain double-precision kernel.  Slightly more spill traffic per site, but a different wave
 * count: 16^3x32 (65536 sites per parity) is 1.15 waves of 148 x 384 threads - a nearly empty second wave - and
 * 0.99 waves of 148 x 448.  Chosen per lattice by tmb_capi.cu (see hop_residency()). */
template <int HINTS>
static cudaError_t hop_mode_448(const tmb_hop_launch &a, cudaStream_t s) {
  if (a.dot) {
    if (a.mode == 4 && a.dot == 2) return hop_go<double2, 4, 0, 2, HINTS, 64, 7>(a, s); /* CG tail */
    if (a.mode != 2) return cudaErrorInvalidValue;
    if (a.dot == 2) return hop_go<double2, 2, 0, 2, HINTS, 64, 7>(a, s);
    return hop_go<double2, 2, 0, 1, HINTS, 64, 7>(a, s);
  }
  switch (a.mode) {
    case 0: return hop_go<double2, 0, 0, 0, HINTS, 64, 7>(a, s);
    case 1: return hop_go<double2, 1, 0, 0, HINTS, 64, 7>(a, s);
    case 2: return hop_go<double2, 2, 0, 0, HINTS, 64, 7>(a, s);
    case 3: return hop_go<double2, 3, 0, 0, HINTS, 64, 7>(a, s);
  }
  return cudaErrorInvalidValue;
}
template <int DIST, int HINTS>
static cudaError_t hop_mode(const tmb_hop_launch &a, cudaStream_t s) {
  if (a.dot) {
    if (a.mode == 4 && a.dot == 2) return hop_go<double2, 4, DIST, 2, HINTS, TMB_HOP_BLOCK, TMB_HOP_MINB>(a, s); /* CG tail */
    if (a.mode != 2) return cudaErrorInvalidValue;
    if (a.dot == 2) return hop_go<double2, 2, DIST, 2, HINTS, TMB_HOP_BLOCK, TMB_HOP_MINB>(a, s);
    return hop_go<double2, 2, DIST, 1, HINTS, TMB_HOP_BLOCK, TMB_HOP_MINB>(a, s);
  }
  switch (a.mode) {
    case 0: return hop_go<double2, 0, DIST, 0, HINTS, TMB_HOP_BLOCK, TMB_HOP_MINB>(a, s);
    case 1: return hop_go<double2, 1, DIST, 0, HINTS, TMB_HOP_BLOCK, TMB_HOP_MINB>(a, s);
    case 2: return hop_go<double2, 2, DIST, 0, HINTS, TMB_HOP_BLOCK, TMB_HOP_MINB>(a, s);
    case 3: return hop_go<double2, 3, DIST, 0, HINTS, TMB_HOP_BLOCK, TMB_HOP_MINB>(a, s);
  }
  return cudaErrorInvalidValue;
}
/* single precision: cache-policy loads always on */
template <int DIST, int CFG>
static cudaError_t hop_mode_f(const tmb_hop_launch &a, cudaStream_t s) {
  if (a.dot) {
    if (a.mode == 4 && a.dot == 2) return hop_go<float2, 4, DIST, 2, CFG, TMB_HOP_BLOCK_F, TMB_HOP_MINB_F>(a, s); /* CG tail */
    if (a.mode != 2) return cudaErrorInvalidValue;
    if (a.dot == 2) return hop_go<float2, 2, DIST, 2, CFG, TMB_HOP_BLOCK_F, TMB_HOP_MINB_F>(a, s);
    return hop_go<float2, 2, DIST, 1, CFG, TMB_HOP_BLOCK_F, TMB_HOP_MINB_F>(a, s);
  }
  switch (a.mode) {
    case 0: return hop_go<float2, 0, DIST, 0, CFG, TMB_HOP_BLOCK_F, TMB_HOP_MINB_F>(a, s);
    case 1: return hop_go<float2, 1, DIST, 0, CFG, TMB_HOP_BLOCK_F, TMB_HOP_MINB_F>(a, s);
    case 2: return hop_go<float2, 2, DIST, 0, CFG, TMB_HOP_BLOCK_F, TMB_HOP_MINB_F>(a, s);
    case 3: return hop_go<float2, 3, DIST, 0, CFG, TMB_HOP_BLOCK_F, TMB_HOP_MINB_F>(a, s); /* D_psi_32 rows */
  }
  return cudaErrorInvalidValue;
}

/* tuning variants of the plain Hopping_Matrix kernel (MODE 0, no halo): block x min-blocks/SM */
template <int HINTS>
static cudaError_t hop_tune(const tmb_hop_launch &a, int variant, cudaStream_t s) {
  switch (variant) {
    case 1: return hop_go<double2, 0, 0, 0, HINTS, 64, 4>(a, s);
    case 2: return hop_go<double2, 0, 0, 0, HINTS, 64, 6>(a, s);
    case 3: return hop_go<double2, 0, 0, 0, HINTS, 128, 2>(a, s);
    case 4: return hop_go<double2, 0, 0, 0, HINTS, 128, 3>(a, s);
    case 5: return hop_go<double2, 0, 0, 0, HINTS, 128, 4>(a, s);
    case 6: return hop_go<double2, 0, 0, 0, HINTS, 256, 1>(a, s);
    case 7: return hop_go<double2, 0, 0, 0, HINTS, 256, 2>(a, s);
    case 8: return hop_go<double2, 0, 0, 0, HINTS, 96, 4>(a, s);
    case 9: return hop_go<double2, 0, 0, 0, HINTS, 192, 2>(a, s);
  }
  return cudaErrorInvalidValue;
}

cudaError_t tmb_launch_hop(const tmb_hop_launch &a, cudaStream_t s) {
  if (a.nfl == 2) return tmb_launch_hop_nd(a, s); /* tmb_hop2.cu */
  if (a.prec) {
    if (a.recon12) return hop_dist_f<3>(a, s);
    return hop_dist_f<1>(a, s);
  }
  if (a.recon12) return hop_dist<3>(a, s);
  const int variant = a.variant;
  if (variant == 10) {
    if (a.dist) return cudaErrorInvalidValue;
    return a.hints ? hop_mode_448<1>(a, s) : hop_mode_448<0>(a, s);
  }
  if (variant > 0) {
    if (a.mode != 0 || a.dist || a.dot) return cudaErrorInvalidValue;
    return a.hints ? hop_tune<1>(a, variant, s) : hop_tune<0>(a, variant, s);
  }
  return a.hints ? hop_dist<1>(a, s) : hop_dist<0>(a, s);
}

/* ------------------------------------------------------------------ K4: reductions */
#define RED_BLOCK 256
int tmb_red_grid(size_t n2) {
  size_t need = (n2 + RED_BLOCK - 1) / RED_BLOCK;
  size_t cap = (size_t)TMB_SMS * 8;
  return (int)(need < cap ? (need ? need : 1) : cap);
}

template <class V2> struct RedNorm2 {
  const V2 *a;
  __device__ double operator()(size_t k) const { const V2 v = a[k]; return (double)v.x * v.x + (double)v.y * v.y; }
};
template <class V2> struct RedDot { /* Re <a,b> = sum a.re*b.re + a.im*b.im   (linalg/scalar_prod_r.c:159-163) */
  const V2 *a, *b;
  __device__ double operator()(size_t k) const { const V2 v = a[k], w = b[k]; return (double)v.x * w.x + (double)v.y * w.y; }
};
struct RedXpayNorm { /* R = c R + S, |R|^2   (linalg/assign_mul_add_r_and_square.c:145) */
  double2 *r; const double2 *sv; double c;
  __device__ double operator()(size_t k) const {
    double2 v = r[k]; const double2 w = sv[k];
    v.x = c * v.x + w.x; v.y = c * v.y + w.y; r[k] = v;
    return v.x * v.x + v.y * v.y;
  }
};
template <class V2> struct RedCgXR { /* x += alpha p ; r -= alpha Ap ; |r|^2      (cg_her.c:95-101) */
  V2 *x, *r; const V2 *p, *ap; const tmb_cg_state *st;
  __device__ double operator()(size_t k) const {
    typedef typename tmb_real<V2>::type R;
    const R al = (R)st->alpha;
    V2 xv = x[k]; const V2 pv = p[k];
    xv.x += al * pv.x; xv.y += al * pv.y; x[k] = xv;
    V2 rv = r[k]; const V2 av = ap[k];
    rv.x = rv.x - al * av.x; rv.y = rv.y - al * av.y; r[k] = rv;
    return (double)rv.x * rv.x + (double)rv.y * rv.y;
  }
};

template <class F>
__global__ void __launch_bounds__(RED_BLOCK) red_kernel(F f, size_t n2, double *partial, const tmb_cg_state *st,
                                                         tmb_cg_state *st_fin, int fin_slot, int fin_op, const tmb_xred_table *xr) {
  if (st != nullptr && st->converged) return;
  double acc = 0.;
  for (size_t k = (size_t)blockIdx.x * RED_BLOCK + threadIdx.x; k < n2; k += (size_t)gridDim.x * RED_BLOCK)
    acc += f(k);
  const double s = block_sum<RED_BLOCK>(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (fin_op >= 0) finish_last_block<RED_BLOCK>(partial, (int)gridDim.x, st_fin, fin_slot, fin_op, xr);
}

/* one CTA: sums the block partials in a fixed order (deterministic), then the CG bookkeeping */
template <int B>
__global__ void __launch_bounds__(B) final_kernel(const double *partial, int n, tmb_cg_state *st, int slot,
                                                  int op, int apply, const tmb_xred_table *xr) {
  if (op != TMB_FIN_STORE && st->converged) return;
  double acc = 0.;
  for (int k = threadIdx.x; k < n; k += B) acc += partial[k];
  double s = block_sum<B>(acc);
  if (threadIdx.x == 0) {
    if (xr != nullptr) s = xred_sum(xr, s);
    st->tmp[slot] = s;
    if (apply) cg_apply(st, slot, op);
  }
}
__global__ void apply_kernel(tmb_cg_state *st, int slot, int op) {
  if (st->converged) return;
  cg_apply(st, slot, op);
}

cudaError_t tmb_launch_final(const double *partial, int n, tmb_cg_state *st, int slot, int op, int apply,
                             const tmb_xred_table *xr, cudaStream_t s) {
  final_kernel<RED_BLOCK><<<1, RED_BLOCK, 0, s>>>(partial, n, st, slot, op, apply, xr);
  return cudaGetLastError();
}
/* the same with the summation tree of the hopping kernels' own finish (finish_last_block<128>): bit-identical sums
 * whether a hop's partials are finished inside the kernel or next to the following one */
cudaError_t tmb_launch_final_hop(const double *partial, int n, tmb_cg_state *st, int slot, int op, int apply,
                                 const tmb_xred_table *xr, cudaStream_t s) {
  final_kernel<TMB_HOP_BLOCK><<<1, TMB_HOP_BLOCK, 0, s>>>(partial, n, st, slot, op, apply, xr);
  return cudaGetLastError();
}
__global__ void seq_bump_kernel(unsigned int *base, unsigned int n) { *base += n; }
cudaError_t tmb_launch_seq_bump(unsigned int *base, unsigned int n, cudaStream_t s) {
  seq_bump_kernel<<<1, 1, 0, s>>>(base, n);
  return cudaGetLastError();
}
cudaError_t tmb_launch_apply(tmb_cg_state *st, int slot, int op, cudaStream_t s) {
  apply_kernel<<<1, 1, 0, s>>>(st, slot, op);
  return cudaGetLastError();
}
#define RED_LAUNCH(f, n2, partial, st, s) \
  do { red_kernel<<<tmb_red_grid(n2), RED_BLOCK, 0, s>>>(f, n2, partial, st, nullptr, 0, -1, nullptr); return cudaGetLastError(); } while (0)
#define RED_LAUNCH_FIN(f, n2, partial, st, stf, slot, op, xr, s) \
  do { red_kernel<<<tmb_red_grid(n2), RED_BLOCK, 0, s>>>(f, n2, partial, st, stf, slot, op, xr); return cudaGetLastError(); } while (0)

cudaError_t tmb_launch_norm2(int prec, const void *a, size_t n2, double *partial, cudaStream_t s) {
  if (prec) { RedNorm2<float2> f = {(const float2 *)a}; RED_LAUNCH(f, n2, partial, nullptr, s); }
  RedNorm2<double2> f = {(const double2 *)a}; RED_LAUNCH(f, n2, partial, nullptr, s);
}
cudaError_t tmb_launch_dot(int prec, const void *a, const void *b, size_t n2, double *partial, cudaStream_t s) {
  if (prec) { RedDot<float2> f = {(const float2 *)a, (const float2 *)b}; RED_LAUNCH(f, n2, partial, nullptr, s); }
  RedDot<double2> f = {(const double2 *)a, (const double2 *)b}; RED_LAUNCH(f, n2, partial, nullptr, s);
}
/* <a,b> with the fused finish + CG bookkeeping (no all-reduce in between: single rank) */
cudaError_t tmb_launch_dot_fin(int prec, const void *a, const void *b, size_t n2, double *partial, tmb_cg_state *st, int slot, int op,
                               const tmb_xred_table *xr, cudaStream_t s) {
  if (prec) { RedDot<float2> f = {(const float2 *)a, (const float2 *)b}; RED_LAUNCH_FIN(f, n2, partial, st, st, slot, op, xr, s); }
  RedDot<double2> f = {(const double2 *)a, (const double2 *)b}; RED_LAUNCH_FIN(f, n2, partial, st, st, slot, op, xr, s);
}
cudaError_t tmb_launch_xpay_norm(double2 *r, double c, const double2 *sv, size_t n2, double *partial, cudaStream_t s) {
  RedXpayNorm f = {r, sv, c}; RED_LAUNCH(f, n2, partial, nullptr, s);
}
cudaError_t tmb_launch_cg_update_xr(int prec, void *x, void *r, const void *p, const void *ap, size_t n2,
                                    tmb_cg_state *st, double *partial, int fin_slot, int fin_op, const tmb_xred_table *xr,
                                    cudaStream_t s) {
  if (prec) { RedCgXR<float2> f = {(float2 *)x, (float2 *)r, (const float2 *)p, (const float2 *)ap, st}; RED_LAUNCH_FIN(f, n2, partial, st, st, fin_slot, fin_op, xr, s); }
  RedCgXR<double2> f = {(double2 *)x, (double2 *)r, (const double2 *)p, (const double2 *)ap, st}; RED_LAUNCH_FIN(f, n2, partial, st, st, fin_slot, fin_op, xr, s);
}

/* ------------------------------------------------------------------ K3: elementwise */
template <class F>
__global__ void __launch_bounds__(256) ew_kernel(F f, size_t n2, const tmb_cg_state *st) {
  if (st != nullptr && st->converged) return;
  for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < n2; k += (size_t)gridDim.x * 256) f(k);
}
static int ew_grid(size_t n2) {
  size_t need = (n2 + 255) / 256, cap = (size_t)TMB_SMS * 16;
  return (int)(need < cap ? (need ? need : 1) : cap);
}
#define EW_LAUNCH(f, n2, st, s) \
  do { ew_kernel<<<ew_grid(n2), 256, 0, s>>>(f, n2, st); return cudaGetLastError(); } while (0)

struct EwAxpy { double2 *p; const double2 *q; double c; /* P += c Q  (linalg/assign_add_mul_r.c:346) */
  __host__ __device__ void operator()(size_t k) const { double2 v = p[k]; const double2 w = q[k]; v.x += c * w.x; v.y += c * w.y; p[k] = v; } };
struct EwXpay { double2 *r; const double2 *sv; double c; /* R = c R + S  (linalg/assign_mul_add_r.c:340) */
  __host__ __device__ void operator()(size_t k) const { double2 v = r[k]; const double2 w = sv[k]; v.x = c * v.x + w.x; v.y = c * v.y + w.y; r[k] = v; } };
struct EwLin { double2 *q; const double2 *r, *sv; double a, b; /* Q = a R + b S  (diff.c:270, add.c) */
  __host__ __device__ void operator()(size_t k) const { const double2 v = r[k], w = sv[k]; q[k] = make_double2(a * v.x + b * w.x, a * v.y + b * w.y); } };
struct EwScale { double2 *r; const double2 *sv; double c; /* R = c S  (linalg/mul_r.c:40) */
  __host__ __device__ void operator()(size_t k) const { const double2 w = sv[k]; r[k] = make_double2(c * w.x, c * w.y); } };
struct EwG5 { double2 *l; const double2 *kk; size_t half; /* gamma.c:77-98 */
  __host__ __device__ void operator()(size_t k) const { const double2 w = kk[k]; l[k] = (k < half) ? w : make_double2(-w.x, -w.y); } };
struct EwDiag { double2 *l; const double2 *kk; double2 z; size_t half; /* (z | conj z) k : tm_operators.c:669, mul_one_pm_imu_inv_body.c */
  __host__ __device__ void operator()(size_t k) const { l[k] = c_mul((k < half) ? z : c_conj(z), kk[k]); } };
struct EwDiagSub { double2 *l; const double2 *kk, *jj; double2 z; int g5; size_t half; /* tm_operators.c:813-857 */
  __host__ __device__ void operator()(size_t k) const {
    const bool up = k < half;
    const double2 zk = c_mul(up ? z : c_conj(z), kk[k]); const double2 j = jj[k];
    l[k] = (up || !g5) ? c_sub(zk, j) : c_sub(j, zk);
  } };
template <class V2> struct EwCgP { V2 *p; const V2 *r; const tmb_cg_state *st; /* p = beta p + r  (cg_her.c:122) */
  __host__ __device__ void operator()(size_t k) const {
    typedef typename tmb_real<V2>::type R;
    const R b = (R)st->beta; V2 v = p[k]; const V2 w = r[k]; v.x = b * v.x + w.x; v.y = b * v.y + w.y; p[k] = v; } };
/* tm_operators_nd.c:639-695 */
struct EwNdMeeInv { double2 *ls, *lc; const double2 *ks, *kc; double mu, eps, nrm; size_t half;
  __host__ __device__ void operator()(size_t k) const {
    const double2 zs = make_double2(1., (k < half) ? -mu : mu), zc = c_conj(zs);
    const double2 s = ks[k], c = kc[k];
    double2 a = c_mul(zs, s); a.x += eps * c.x; a.y += eps * c.y;
    double2 b = c_mul(zc, c); b.x += eps * s.x; b.y += eps * s.y;
    ls[k] = make_double2(nrm * a.x, nrm * a.y); lc[k] = make_double2(nrm * b.x, nrm * b.y);
  } };
/* tm_operators_nd.c:698-756 */
struct EwNdMooSubG5 { double2 *ls, *lc; const double2 *ks, *kc, *js, *jc; double mu, eps; size_t half;
  __host__ __device__ void operator()(size_t k) const {
    const bool up = k < half;
    const double2 zs = make_double2(1., up ? -mu : mu), zc = c_conj(zs);
    const double2 s = ks[k], c = kc[k], ts = js[k], tc = jc[k];
    double2 a = c_mul(zs, s); a.x += eps * c.x; a.y += eps * c.y;
    double2 b = c_mul(zc, c); b.x += eps * s.x; b.y += eps * s.y;
    ls[k] = up ? c_sub(a, ts) : c_sub(ts, a); lc[k] = up ? c_sub(b, tc) : c_sub(tc, b);
  } };
/* linalg/assign_to_32.c, linalg/addto_32.c */
struct EwToFloat { float2 *d; const double2 *sv;
  __host__ __device__ void operator()(size_t k) const { const double2 v = sv[k]; d[k] = make_float2((float)v.x, (float)v.y); } };
struct EwAddFromFloat { double2 *d; const float2 *sv;
  __host__ __device__ void operator()(size_t k) const { double2 v = d[k]; const float2 w = sv[k]; v.x += (double)w.x; v.y += (double)w.y; d[k] = v; } };

/* single-precision BLAS-1 of the mixed solvers (the _32.c files of linalg/): op 0 R += c1 S (assign_add_mul_r_32.c:42), 1 R = c1 R + S
 * (assign_mul_add_r_32.c:18), 2 R = S - S2 (diff_32.c:39), 3 R = c1 S (mul_r_32.c:69), 4 R = c1 R + c2 S
 * (assign_mul_add_mul_r_32.c:37), 5 R = gamma5 S (tm_operators_32.c:130); arithmetic in float like the reference */
struct EwBlas32 { float2 *r; const float2 *sv, *s2; float c1, c2; int op; size_t half;
  __host__ __device__ void operator()(size_t k) const {
    float2 v;
    if (op == 0) { v = r[k]; const float2 w = sv[k]; v.x += c1 * w.x; v.y += c1 * w.y; }
    else if (op == 1) { v = r[k]; const float2 w = sv[k]; v.x = c1 * v.x + w.x; v.y = c1 * v.y + w.y; }
    else if (op == 2) { const float2 a = sv[k], b = s2[k]; v = make_float2(a.x - b.x, a.y - b.y); }
    else if (op == 3) { const float2 w = sv[k]; v = make_float2(c1 * w.x, c1 * w.y); }
    else if (op == 4) { v = r[k]; const float2 w = sv[k]; v.x = c1 * v.x + c2 * w.x; v.y = c1 * v.y + c2 * w.y; }
    else { const float2 w = sv[k]; v = (k >= half) ? make_float2(-w.x, -w.y) : w; }
    r[k] = v;
  } };
cudaError_t tmb_launch_blas32(int op, float2 *r, const float2 *sv, const float2 *s2, float c1, float c2, size_t n2, size_t half, cudaStream_t s) {
  EwBlas32 f = {r, sv, s2, c1, c2, op, half}; EW_LAUNCH(f, n2, nullptr, s);
}
cudaError_t tmb_launch_axpy(double2 *p, const double2 *q, double c, size_t n2, cudaStream_t s) { EwAxpy f = {p, q, c}; EW_LAUNCH(f, n2, nullptr, s); }
cudaError_t tmb_launch_xpay(double2 *r, double c, const double2 *sv, size_t n2, cudaStream_t s) { EwXpay f = {r, sv, c}; EW_LAUNCH(f, n2, nullptr, s); }
cudaError_t tmb_launch_lincomb(double2 *q, double a, const double2 *r, double b, const double2 *sv, size_t n2, cudaStream_t s) { EwLin f = {q, r, sv, a, b}; EW_LAUNCH(f, n2, nullptr, s); }
cudaError_t tmb_launch_scale(double2 *r, double c, const double2 *sv, size_t n2, cudaStream_t s) { EwScale f = {r, sv, c}; EW_LAUNCH(f, n2, nullptr, s); }
cudaError_t tmb_launch_gamma5(double2 *l, const double2 *k, size_t n2, size_t half, cudaStream_t s) { EwG5 f = {l, k, half}; EW_LAUNCH(f, n2, nullptr, s); }
cudaError_t tmb_launch_diag(double2 *l, const double2 *k, double2 z, size_t n2, size_t half, cudaStream_t s) { EwDiag f = {l, k, z, half}; EW_LAUNCH(f, n2, nullptr, s); }
cudaError_t tmb_launch_diag_sub(double2 *l, const double2 *k, const double2 *j, double2 z, int g5, size_t n2, size_t half, cudaStream_t s) { EwDiagSub f = {l, k, j, z, g5, half}; EW_LAUNCH(f, n2, nullptr, s); }
cudaError_t tmb_launch_cg_update_p(int prec, void *p, const void *r, size_t n2, const tmb_cg_state *st, cudaStream_t s) {
  if (prec) { EwCgP<float2> f = {(float2 *)p, (const float2 *)r, st}; EW_LAUNCH(f, n2, st, s); }
  EwCgP<double2> f = {(double2 *)p, (const double2 *)r, st}; EW_LAUNCH(f, n2, st, s);
}
cudaError_t tmb_launch_nd_mee_inv(double2 *ls, double2 *lc, const double2 *ks, const double2 *kc, double mu, double eps, size_t n2, size_t half, cudaStream_t s) {
  EwNdMeeInv f = {ls, lc, ks, kc, mu, eps, 1. / (1. + mu * mu - eps * eps), half}; EW_LAUNCH(f, n2, nullptr, s); }
cudaError_t tmb_launch_nd_moo_sub_g5(double2 *ls, double2 *lc, const double2 *ks, const double2 *kc, const double2 *js, const double2 *jc, double mu, double eps, size_t n2, size_t half, cudaStream_t s) {
  EwNdMooSubG5 f = {ls, lc, ks, kc, js, jc, mu, eps, half}; EW_LAUNCH(f, n2, nullptr, s); }
cudaError_t tmb_launch_to_float(float2 *dst, const double2 *src, size_t n, cudaStream_t s) { EwToFloat f = {dst, src}; EW_LAUNCH(f, n, nullptr, s); }
cudaError_t tmb_launch_add_from_float(double2 *dst, const float2 *src, size_t n, cudaStream_t s) { EwAddFromFloat f = {dst, src}; EW_LAUNCH(f, n, nullptr, s); }

/* complex BLAS-1 of the chronological guess (solver/chrono_guess.c): linalg/assign_diff_mul.c:31,
 * linalg/mul.c, linalg/assign_add_mul.c */
struct EwCAxpy { double2 *r; const double2 *sv; double2 c; /* R += c S */
  __host__ __device__ void operator()(size_t k) const { double2 v = r[k]; c_mad(v, c, sv[k]); r[k] = v; } };
struct EwCScale { double2 *r; const double2 *sv; double2 c; /* R = c S */
  __host__ __device__ void operator()(size_t k) const { r[k] = c_mul(c, sv[k]); } };
cudaError_t tmb_launch_caxpy(double2 *r, double2 c, const double2 *sv, size_t n2, cudaStream_t s) { EwCAxpy f = {r, sv, c}; EW_LAUNCH(f, n2, nullptr, s); }
cudaError_t tmb_launch_cscale(double2 *r, double2 c, const double2 *sv, size_t n2, cudaStream_t s) { EwCScale f = {r, sv, c}; EW_LAUNCH(f, n2, nullptr, s); }

/* <S,R> = sum conj(S) R (linalg/scalar_prod_body.c): real parts to partial[b], imaginary to partial[grid + b] */
__global__ void __launch_bounds__(RED_BLOCK) cdot_kernel(const double2 *a, const double2 *b, size_t n2, double *partial) {
  double re = 0., im = 0.;
  for (size_t k = (size_t)blockIdx.x * RED_BLOCK + threadIdx.x; k < n2; k += (size_t)gridDim.x * RED_BLOCK) {
    const double2 v = a[k], w = b[k];
    re += v.x * w.x + v.y * w.y;
    im += v.x * w.y - v.y * w.x;
  }
  const double sr = block_sum<RED_BLOCK>(re);
  __syncthreads();
  const double si = block_sum<RED_BLOCK>(im);
  if (threadIdx.x == 0) { partial[blockIdx.x] = sr; partial[gridDim.x + blockIdx.x] = si; }
}
cudaError_t tmb_launch_cdot(const double2 *a, const double2 *b, size_t n2, double *partial, cudaStream_t s) {
  cdot_kernel<<<tmb_red_grid(n2), RED_BLOCK, 0, s>>>(a, b, n2, partial);
  return cudaGetLastError();
}

/* ------------------------------------------------------------------ layout conversion */
/* host AoS spinor (su3.h:60-63: 12 complex per site, site-major)  <->  device SoA [12][Vh] */
struct EwPackEo { double2 *soa; const double2 *aos; int Vh;
  __host__ __device__ void operator()(size_t k) const { const int c = (int)(k / Vh), i = (int)(k - (size_t)c * Vh); soa[k] = aos[(size_t)i * 12 + c]; } };
struct EwUnpackEo { double2 *aos; const double2 *soa; int Vh;
  __host__ __device__ void operator()(size_t k) const { const int c = (int)(k / Vh), i = (int)(k - (size_t)c * Vh); aos[(size_t)i * 12 + c] = soa[k]; } };
/* the same for a contiguous site range [i0, i0+n): used by the pipelined host-pointer operators */
struct EwPackEoRange { double2 *soa; const double2 *aos; int Vh, i0, n;
  __host__ __device__ void operator()(size_t k) const { const int c = (int)(k / n), i = i0 + (int)(k - (size_t)c * n); soa[(size_t)c * Vh + i] = aos[(size_t)i * 12 + c]; } };
struct EwUnpackEoRange { double2 *aos; const double2 *soa; int Vh, i0, n;
  __host__ __device__ void operator()(size_t k) const { const int c = (int)(k / n), i = i0 + (int)(k - (size_t)c * n); aos[(size_t)i * 12 + c] = soa[(size_t)c * Vh + i]; } };
/* lexicographic host field of V sites <-> (even, odd) device fields (linalg/convert_eo_to_lexic.c:35-115) */
template <class V2> struct EwPackLex { V2 *even, *odd; const V2 *lex; tmb_geom g;
  __host__ __device__ void operator()(size_t k) const {
    const size_t per = (size_t)12 * g.Vh; const int par = (int)(k / per); const size_t kk = k - par * per;
    const int c = (int)(kk / g.Vh), i = (int)(kk - (size_t)c * g.Vh);
    const int ix = tmb_eo_to_lexic(g, par, i);
    (par ? odd : even)[kk] = lex[(size_t)ix * 12 + c];
  } };
template <class V2> struct EwUnpackLex { V2 *lex; const V2 *even, *odd; tmb_geom g;
  __host__ __device__ void operator()(size_t k) const {
    const size_t per = (size_t)12 * g.Vh; const int par = (int)(k / per); const size_t kk = k - par * per;
    const int c = (int)(kk / g.Vh), i = (int)(kk - (size_t)c * g.Vh);
    const int ix = tmb_eo_to_lexic(g, par, i);
    lex[(size_t)ix * 12 + c] = (par ? odd : even)[kk];
  } };
/* host gauge g_gauge_field[ix][mu] (init/init_gauge_field.c:51-68, su3 = 9 complex row-major)
 * -> U[((q*4+mu)*9+e)*Vh + i] */
struct EwPackGauge { double2 *U; const double2 *lex; tmb_geom g;
  __host__ __device__ void operator()(size_t k) const {
    const int i = (int)(k % g.Vh); const int row = (int)(k / g.Vh); /* row = (q*4+mu)*9+e */
    const int e = row % 9, qm = row / 9, mu = qm & 3, q = qm >> 2;
    const int ix = tmb_eo_to_lexic(g, q, i);
    U[k] = lex[((size_t)ix * 4 + mu) * 9 + e];
  } };
template <class V2> struct EwPackHalo { V2 *up, *dn; const V2 *in; tmb_geom g;
  __host__ __device__ void operator()(size_t k) const { /* k in [0, 6*S) */
    const int c = (int)(k / g.S), j = (int)(k - (size_t)c * g.S);
    const size_t last = (size_t)(g.T - 1) * g.S + j;
    const V2 a = in[(size_t)c * g.Vh + last], b = in[(size_t)(c + 6) * g.Vh + last];
    up[k] = c_sub(a, b);                 /* (1-g0): s0-s2, s1-s3 of the last slice -> rank+1 */
    const V2 a0 = in[(size_t)c * g.Vh + j], b0 = in[(size_t)(c + 6) * g.Vh + j];
    dn[k] = c_add(a0, b0);               /* (1+g0): s0+s2, s1+s3 of the first slice -> rank-1 */
  } };
struct EwPackGaugeHalo { double2 *out; const double2 *U; tmb_geom g;
  __host__ __device__ void operator()(size_t k) const { /* k in [0, 2*9*S) : out[(q*9+e)*S + j] */
    const int j = (int)(k % g.S); const int qe = (int)(k / g.S); const int e = qe % 9, q = qe / 9;
    out[k] = U[(size_t)((q * 4 + 0) * 9 + e) * g.Vh + (size_t)(g.T - 1) * g.S + j];
  } };

cudaError_t tmb_launch_pack_eo(double2 *soa, const double2 *aos, int Vh, cudaStream_t s) { EwPackEo f = {soa, aos, Vh}; EW_LAUNCH(f, (size_t)12 * Vh, nullptr, s); }
cudaError_t tmb_launch_unpack_eo(double2 *aos, const double2 *soa, int Vh, cudaStream_t s) { EwUnpackEo f = {aos, soa, Vh}; EW_LAUNCH(f, (size_t)12 * Vh, nullptr, s); }
cudaError_t tmb_launch_pack_eo_range(double2 *soa, const double2 *aos, int Vh, int i0, int n, cudaStream_t s) { EwPackEoRange f = {soa, aos, Vh, i0, n}; EW_LAUNCH(f, (size_t)12 * n, nullptr, s); }
cudaError_t tmb_launch_unpack_eo_range(double2 *aos, const double2 *soa, int Vh, int i0, int n, cudaStream_t s) { EwUnpackEoRange f = {aos, soa, Vh, i0, n}; EW_LAUNCH(f, (size_t)12 * n, nullptr, s); }
cudaError_t tmb_launch_pack_lexic(double2 *even, double2 *odd, const double2 *lex, tmb_geom g, cudaStream_t s) { EwPackLex<double2> f = {even, odd, lex, g}; EW_LAUNCH(f, (size_t)24 * g.Vh, nullptr, s); }
cudaError_t tmb_launch_unpack_lexic(double2 *lex, const double2 *even, const double2 *odd, tmb_geom g, cudaStream_t s) { EwUnpackLex<double2> f = {lex, even, odd, g}; EW_LAUNCH(f, (size_t)24 * g.Vh, nullptr, s); }
cudaError_t tmb_launch_pack_lexic_f(float2 *even, float2 *odd, const float2 *lex, tmb_geom g, cudaStream_t s) { EwPackLex<float2> f = {even, odd, lex, g}; EW_LAUNCH(f, (size_t)24 * g.Vh, nullptr, s); }
cudaError_t tmb_launch_unpack_lexic_f(float2 *lex, const float2 *even, const float2 *odd, tmb_geom g, cudaStream_t s) { EwUnpackLex<float2> f = {lex, even, odd, g}; EW_LAUNCH(f, (size_t)24 * g.Vh, nullptr, s); }
cudaError_t tmb_launch_pack_gauge(double2 *U, const double2 *lex, tmb_geom g, cudaStream_t s) { EwPackGauge f = {U, lex, g}; EW_LAUNCH(f, (size_t)72 * g.Vh, nullptr, s); }
cudaError_t tmb_launch_pack_halo(int prec, void *up, void *dn, const void *in, tmb_geom g, cudaStream_t s) {
  if (prec) { EwPackHalo<float2> f = {(float2 *)up, (float2 *)dn, (const float2 *)in, g}; EW_LAUNCH(f, (size_t)6 * g.S, nullptr, s); }
  EwPackHalo<double2> f = {(double2 *)up, (double2 *)dn, (const double2 *)in, g}; EW_LAUNCH(f, (size_t)6 * g.S, nullptr, s);
}
/* 12-real copies: rows 0,1 of every link; n = sites per link row (Vh for the bulk, S for the halo), nl links rows */
struct EwCompress12 { double2 *dst; const double2 *src; size_t n;
  __host__ __device__ void operator()(size_t k) const { /* k in [0, nl*6*n) */
    const size_t i = k % n, row = k / n, e = row % 6, l = row / 6;
    dst[k] = src[(l * 9 + e) * n + i];
  } };
cudaError_t tmb_launch_compress12(double2 *dst, const double2 *src, size_t n, int nlinks, cudaStream_t s) {
  EwCompress12 f = {dst, src, n}; EW_LAUNCH(f, (size_t)nlinks * 6 * n, nullptr, s);
}
/* max over links of |row2 - conj(row0 x row1)|^2: how far the field is from what compression assumes */
struct RedSu3Defect { const double2 *U; size_t n;
  __device__ double operator()(size_t k) const { /* k in [0, nl*n) */
    const size_t i = k % n, l = k / n;
    double2 u[9];
    for (int e = 0; e < 9; e++) u[e] = U[(l * 9 + e) * n + i];
    double2 w[9];
    for (int e = 0; e < 6; e++) w[e] = u[e];
    tmb_reconstruct_row2(w);
    double d = 0.;
    for (int e = 6; e < 9; e++) { const double dx = w[e].x - u[e].x, dy = w[e].y - u[e].y; d += dx * dx + dy * dy; }
    return d;
  } };
template <class F>
__global__ void __launch_bounds__(RED_BLOCK) max_kernel(F f, size_t n2, double *partial) {
  double acc = 0.;
  for (size_t k = (size_t)blockIdx.x * RED_BLOCK + threadIdx.x; k < n2; k += (size_t)gridDim.x * RED_BLOCK) {
    const double v = f(k); acc = v > acc ? v : acc;
  }
  __shared__ double sh[RED_BLOCK];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = RED_BLOCK / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] = sh[threadIdx.x + o] > sh[threadIdx.x] ? sh[threadIdx.x + o] : sh[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
/* partial[b] = max defect seen by block b; the caller takes the max over tmb_red_grid(nl*n) entries */
cudaError_t tmb_launch_su3_defect(const double2 *U, size_t n, int nlinks, double *partial, cudaStream_t s) {
  RedSu3Defect f = {U, n};
  max_kernel<<<tmb_red_grid((size_t)nlinks * n), RED_BLOCK, 0, s>>>(f, (size_t)nlinks * n, partial);
  return cudaGetLastError();
}
/* ---- second split direction (Z): face pack, gauge z-halo, fix-up of the face sites (tmb_site.cuh) ---- */
cudaError_t tmb_launch_pack_zfaces(int prec, void *up, void *dn, const void *in, tmb_geom g, int pin, cudaStream_t s) {
  const size_t n = (size_t)6 * (g.T * g.LX * g.LY / 2);
  if (prec) { EwPackZFaces<float2> f = {(float2 *)up, (float2 *)dn, (const float2 *)in, g, pin}; EW_LAUNCH(f, n, nullptr, s); }
  EwPackZFaces<double2> f = {(double2 *)up, (double2 *)dn, (const double2 *)in, g, pin}; EW_LAUNCH(f, n, nullptr, s);
}
/* Z faces through PEER MEMORY: the projected faces are written straight into the z neighbours' halo buffers (remote stores
 * over NVLink), double-buffered on the parity of the hop's sequence number; a one-thread kernel behind the pack kernel (a
 * kernel boundary: the stores have been performed) raises the neighbours' flags, and the neighbours' fix-up kernels wait
 * for theirs.  No NCCL call, so solver chunks on a Z-split grid can be captured as CUDA graphs like the T-split ones. */
template <class V2>
__global__ void __launch_bounds__(256) pack_zfaces_push_kernel(V2 *up0, V2 *up1, V2 *dn0, V2 *dn1, V2 *own_up, V2 *own_dn, const V2 *in,
                                                               tmb_geom g, int pin, const unsigned int *seq_base, unsigned int seq_off) {
  const unsigned int seq = *seq_base + seq_off;
  EwPackZFaces<V2> f = {(seq & 1u) ? up1 : up0, (seq & 1u) ? dn1 : dn0, in, g, pin, own_up, own_dn}; /* remote + local copy */
  const size_t n = (size_t)6 * (g.T * g.LX * g.LY / 2);
  for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (size_t)gridDim.x * 256) f(k);
}
__global__ void zflag_kernel(unsigned int *fa, unsigned int *fb, const unsigned int *seq_base, unsigned int seq_off) {
  const unsigned int seq = *seq_base + seq_off;
  __threadfence_system();
  st_release_sys(fa, seq);
  st_release_sys(fb, seq);
}
cudaError_t tmb_launch_pack_zfaces_push(int prec, void *up0, void *up1, void *dn0, void *dn1, void *own_up, void *own_dn, const void *in,
                                        tmb_geom g, int pin, const unsigned int *seq_base, unsigned int seq_off, unsigned int *flag_up,
                                        unsigned int *flag_dn, cudaStream_t s) {
  const size_t n = (size_t)6 * (g.T * g.LX * g.LY / 2);
  const size_t need = (n + 255) / 256; const int grid = (int)(need < (size_t)148 * 16 ? (need ? need : 1) : (size_t)148 * 16);
  if (prec) pack_zfaces_push_kernel<float2><<<grid, 256, 0, s>>>((float2 *)up0, (float2 *)up1, (float2 *)dn0, (float2 *)dn1, (float2 *)own_up, (float2 *)own_dn, (const float2 *)in, g, pin, seq_base, seq_off);
  else pack_zfaces_push_kernel<double2><<<grid, 256, 0, s>>>((double2 *)up0, (double2 *)up1, (double2 *)dn0, (double2 *)dn1, (double2 *)own_up, (double2 *)own_dn, (const double2 *)in, g, pin, seq_base, seq_off);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  zflag_kernel<<<1, 1, 0, s>>>(flag_up, flag_dn, seq_base, seq_off);
  return cudaGetLastError();
}
cudaError_t tmb_launch_pack_gauge_zhalo(int prec, void *out, const void *U, tmb_geom g, cudaStream_t s) {
  const size_t n = (size_t)18 * (g.T * g.LX * g.LY / 2);
  if (prec) { EwPackGaugeZHalo<float2> f = {(float2 *)out, (const float2 *)U, g}; EW_LAUNCH(f, n, nullptr, s); }
  EwPackGaugeZHalo<double2> f = {(double2 *)out, (const double2 *)U, g}; EW_LAUNCH(f, n, nullptr, s);
}
/* peer mode (w.flags != nullptr): the halo buffers are the ones of parity seq & 1, filled by the z neighbours' pack kernels;
 * thread 0 of every CTA waits for the two flags to reach seq */
template <int MODE, class V2>
__global__ void __launch_bounds__(64) zfix_kernel(V2 *out, const V2 *in, const V2 *U, const V2 *hz_up, const V2 *hz_dn, const V2 *Uzh,
                                                   tmb_geom g, int par, double2 ka3, double2 cf, const tmb_cg_state *st, const tmb_zpeer w) {
  if (st != nullptr && st->converged) return;
  if (w.flags != nullptr) {
    const unsigned int seq = *w.seq_base + w.seq_off;
    if (threadIdx.x == 0) { wait_flag(w.flags + 0, seq, w.err); wait_flag(w.flags + 1, seq, w.err); }
    __syncthreads();
    if (seq & 1u) { hz_up = (const V2 *)w.hz_up1; hz_dn = (const V2 *)w.hz_dn1; }
  }
  /* one thread per (face site, side): 2 Sz threads in CTAs of 64 - the face accesses are strided (one element per row of
   * Lzh), latency-bound, so the fix-up wants as many independent threads as it can get (one thread per site PAIR in CTAs
   * of 128: 52 us at 24x48x48x24, a third of the hop) */
  const int Sz = g.T * g.LX * g.LY / 2;
  const int q = blockIdx.x * 64 + threadIdx.x;
  if (q >= 2 * Sz) return;
  const int side = q >= Sz ? 1 : 0, j = q - side * Sz;
  tmb_zfix_side<MODE>(out, in, U, hz_up, hz_dn, Uzh, g, par, j, side, cvt2<V2>(ka3), cvt2<V2>(cf), (const V2 *)w.own_up, (const V2 *)w.own_dn, (const V2 *)w.Uzl);
}
template <class V2>
static cudaError_t zfix_go(int mode, V2 *out, const V2 *in, const V2 *U, const V2 *hu, const V2 *hd, const V2 *Uzh, tmb_geom g, int par,
                           double2 ka3, double2 cf, const tmb_cg_state *st, const tmb_zpeer &w, cudaStream_t s) {
  const int grid = (2 * (g.T * g.LX * g.LY / 2) + 63) / 64;
  switch (mode) {
    case 0: zfix_kernel<0, V2><<<grid, 64, 0, s>>>(out, in, U, hu, hd, Uzh, g, par, ka3, cf, st, w); break;
    case 1: zfix_kernel<1, V2><<<grid, 64, 0, s>>>(out, in, U, hu, hd, Uzh, g, par, ka3, cf, st, w); break;
    case 2: zfix_kernel<2, V2><<<grid, 64, 0, s>>>(out, in, U, hu, hd, Uzh, g, par, ka3, cf, st, w); break;
    case 3: zfix_kernel<3, V2><<<grid, 64, 0, s>>>(out, in, U, hu, hd, Uzh, g, par, ka3, cf, st, w); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
cudaError_t tmb_launch_zfix(int prec, int mode, void *out, const void *in, const void *U, const void *hz_up, const void *hz_dn, const void *Uzh,
                            tmb_geom g, int par, double2 ka3, double2 cf, const tmb_cg_state *st, const tmb_zpeer *w, cudaStream_t s) {
  tmb_zpeer none; memset(&none, 0, sizeof(none));
  const tmb_zpeer &ww = w ? *w : none;
  if (prec) return zfix_go<float2>(mode, (float2 *)out, (const float2 *)in, (const float2 *)U, (const float2 *)hz_up, (const float2 *)hz_dn, (const float2 *)Uzh, g, par, ka3, cf, st, ww, s);
  return zfix_go<double2>(mode, (double2 *)out, (const double2 *)in, (const double2 *)U, (const double2 *)hz_up, (const double2 *)hz_dn, (const double2 *)Uzh, g, par, ka3, cf, st, ww, s);
}
cudaError_t tmb_launch_pack_gauge_halo(double2 *out, const double2 *U, tmb_geom g, cudaStream_t s) { EwPackGaugeHalo f = {out, U, g}; EW_LAUNCH(f, (size_t)18 * g.S, nullptr, s); }

/* tmb_site.cuh - per-site arithmetic of the even/odd Wilson hopping term, written once as
 * __host__ __device__ code: the CUDA kernels in tmb_kernels.cu call it on the device; the
 * CPU-only logic tests compile the very same functions for the host (tests/emul/) to check
 * layout, neighbour arithmetic and spin algebra without a GPU.  The shipped library has no
 * host compute path: nothing in tmlqcd_b200/csrc calls these functions from host code.
 *
 * What is computed (reference: operator/hopping.h:574-694 generic-C macros, index walk
 * operator/hopping_body_dbl.c:64-181):
 *   r(x) = sum_mu [ ka_mu U_mu(x) (1+g_mu)-proj k(x+mu) + conj(ka_mu) U_mu(x-mu)^+ (1-g_mu)-proj k(x-mu) ]
 *
 * Everything is templated on the complex element type V2: double2 for the double-precision
 * path, float2 for the single-precision operator of the mixed CG (operator/Hopping_Matrix_32.c,
 * operator/tm_operators_32.c).
 *
 * Device layout (B200-first, not the reference's AoS):
 *   spinor field  f[c*Vh + i]                 c = 3*spin+colour (12), i = eo-sub site, V2 = (re,im)
 *   gauge field   U[((q*4+mu)*9 + e)*Vh + i]  q = parity of the site that owns the forward link,
 *                                             e = 3*row+col; each link stored ONCE (4V links);
 *                                             the backward hop gathers U[1-p][mu][.][nb] at the
 *                                             same neighbour index as the neighbour spinor.
 * Every load of a warp is 32 consecutive V2: 512 (double) or 256 (float) contiguous bytes.
 */
#pragma once
#include <cuda_runtime.h>
#include "tmb_geom.h"

template <class V2> struct tmb_real;
template <> struct tmb_real<double2> { typedef double type; };
template <> struct tmb_real<float2> { typedef float type; };
template <class V2> TMB_HD V2 mk2(typename tmb_real<V2>::type x, typename tmb_real<V2>::type y) {
  V2 v; v.x = x; v.y = y; return v;
}
template <class V2> TMB_HD V2 cvt2(double2 a) {
  return mk2<V2>((typename tmb_real<V2>::type)a.x, (typename tmb_real<V2>::type)a.y);
}

/* ---------------- loads with cache policy ---------------- */
#if defined(__CUDACC__)
/* gauge links: read exactly once per hop -> do not allocate in L1, evict-first in L2 so the
 * 126 MB L2 keeps the input spinor (8-fold reuse) resident across time-slices */
__device__ __forceinline__ double2 tmb_ld_stream(const double2 *p, unsigned long long pol) {
  double2 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
      : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 tmb_ld_stream(const float2 *p, unsigned long long pol) {
  float2 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
      : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
/* gauge links of the two-flavour kernel: read by the two flavour groups of a CTA within a short time -> allocate in L1
 * (the second group's request should hit there), still evict-first in L2 */
__device__ __forceinline__ double2 tmb_ld_stream_l1(const double2 *p, unsigned long long pol) {
  double2 v;
  asm("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 tmb_ld_stream_l1(const float2 *p, unsigned long long pol) {
  float2 v;
  asm("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
/* neighbour spinors: reused by 8 output sites -> keep in L1/L2 */
__device__ __forceinline__ double2 tmb_ld_reuse(const double2 *p, unsigned long long pol) {
  double2 v;
  asm("ld.global.nc.L1::evict_last.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
      : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 tmb_ld_reuse(const float2 *p, unsigned long long pol) {
  float2 v;
  asm("ld.global.nc.L1::evict_last.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
      : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void tmb_st_stream(double2 *p, double2 v, unsigned long long pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;"
               :: "l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void tmb_st_stream(float2 *p, float2 v, unsigned long long pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;"
               :: "l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
/* halo buffers of the peer mode: written early, read late -> ask L2 to keep them */
__device__ __forceinline__ void tmb_st_keep(double2 *p, double2 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" :: "l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void tmb_st_keep(float2 *p, float2 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" :: "l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned long long tmb_policy_evict_first() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned long long tmb_policy_evict_last() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
#endif

struct tmb_policies { unsigned long long stream, reuse; };

template <int HINTS, class V2>
TMB_HD V2 tmb_load_gauge(const V2 *p, const tmb_policies &pol) {
#if defined(__CUDA_ARCH__)
  if (HINTS & 4) return tmb_ld_stream_l1(p, pol.stream);
  if (HINTS) return tmb_ld_stream(p, pol.stream);
  return __ldg(p);
#else
  (void)pol; return *p;
#endif
}
template <int HINTS, class V2>
TMB_HD V2 tmb_load_spinor(const V2 *p, const tmb_policies &pol) {
#if defined(__CUDA_ARCH__)
  if (HINTS) return tmb_ld_reuse(p, pol.reuse);
  return __ldg(p);
#else
  (void)pol; return *p;
#endif
}
template <int HINTS, class V2>
TMB_HD void tmb_store_out(V2 *p, V2 v, const tmb_policies &pol) {
#if defined(__CUDA_ARCH__)
  if (HINTS) { tmb_st_stream(p, v, pol.stream); return; }
  *p = v;
#else
  (void)pol; *p = v;
#endif
}

/* ---------------- complex helpers on V2 ---------------- */
template <class V2> TMB_HD V2 c_add(V2 a, V2 b) { return mk2<V2>(a.x + b.x, a.y + b.y); }
template <class V2> TMB_HD V2 c_sub(V2 a, V2 b) { return mk2<V2>(a.x - b.x, a.y - b.y); }
template <class V2> TMB_HD V2 c_mul(V2 a, V2 b) { return mk2<V2>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <class V2> TMB_HD V2 c_conj(V2 a) { return mk2<V2>(a.x, -a.y); }
template <class V2> TMB_HD void c_mad(V2 &acc, V2 a, V2 b) { /* acc += a*b */
  acc.x += a.x * b.x; acc.x -= a.y * b.y; acc.y += a.x * b.y; acc.y += a.y * b.x;
}
template <class V2> TMB_HD void c_madc(V2 &acc, V2 a, V2 b) { /* acc += conj(a)*b */
  acc.x += a.x * b.x; acc.x += a.y * b.y; acc.y += a.x * b.y; acc.y -= a.y * b.x;
}
/* x + coef*y, coef code: 0:+1  1:-1  2:+i  3:-i */
template <int C, class V2> TMB_HD V2 c_comb(V2 x, V2 y) {
  if (C == 0) return mk2<V2>(x.x + y.x, x.y + y.y);
  if (C == 1) return mk2<V2>(x.x - y.x, x.y - y.y);
  if (C == 2) return mk2<V2>(x.x - y.y, x.y + y.x);
  return mk2<V2>(x.x + y.y, x.y - y.x);
}
/* the conjugate coefficient: 0->0, 1->1, 2->3, 3->2 */
template <int C> struct conj_code { static const int v = (C < 2) ? C : (5 - C); };

/* ---------------- projector table, hopping.h:578-672 ----------------
 * direction D = 2*mu + (0 forward | 1 backward):  a = s0 + CA*s[PA],  b = s1 + CB*s[PB]
 * reconstruction:  r0 += phi_a, r[PA] += conj(CA) phi_a, r1 += phi_b, r[PB] += conj(CB) phi_b */
template <int D> struct hop_tab;
template <> struct hop_tab<0> { static const int PA = 2, CA = 0, PB = 3, CB = 0; };
template <> struct hop_tab<1> { static const int PA = 2, CA = 1, PB = 3, CB = 1; };
template <> struct hop_tab<2> { static const int PA = 3, CA = 2, PB = 2, CB = 2; };
template <> struct hop_tab<3> { static const int PA = 3, CA = 3, PB = 2, CB = 3; };
template <> struct hop_tab<4> { static const int PA = 3, CA = 0, PB = 2, CB = 1; };
template <> struct hop_tab<5> { static const int PA = 3, CA = 1, PB = 2, CB = 0; };
template <> struct hop_tab<6> { static const int PA = 2, CA = 2, PB = 3, CB = 3; };
template <> struct hop_tab<7> { static const int PA = 2, CA = 3, PB = 3, CB = 2; };

/* half-spinor of direction D from a full neighbour spinor at SoA index n */
template <int D, int HINTS, class V2>
TMB_HD void tmb_project(V2 a[3], V2 b[3], const V2 *__restrict__ in, int Vh, int n, const tmb_policies &pol) {
  typedef hop_tab<D> Tb;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    V2 s0 = tmb_load_spinor<HINTS>(in + (size_t)(0 + c) * Vh + n, pol);
    V2 s1 = tmb_load_spinor<HINTS>(in + (size_t)(3 + c) * Vh + n, pol);
    V2 sa = tmb_load_spinor<HINTS>(in + (size_t)(3 * Tb::PA + c) * Vh + n, pol);
    V2 sb = tmb_load_spinor<HINTS>(in + (size_t)(3 * Tb::PB + c) * Vh + n, pol);
    a[c] = c_comb<Tb::CA>(s0, sa);
    b[c] = c_comb<Tb::CB>(s1, sb);
  }
}

/* phi = c * (U a) for forward, c * (U^dagger a) for backward (su3.h:308-316), then scatter */
template <int D, class V2>
TMB_HD void tmb_link_accumulate(V2 r[12], const V2 u[9], const V2 a[3], const V2 b[3], V2 ka) {
  typedef hop_tab<D> Tb;
  const int BWD = D & 1;
  const V2 c = BWD ? c_conj(ka) : ka;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    V2 xa = mk2<V2>(0, 0), xb = mk2<V2>(0, 0);
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (BWD) { c_madc(xa, u[3 * j + i], a[j]); c_madc(xb, u[3 * j + i], b[j]); }
      else     { c_mad(xa, u[3 * i + j], a[j]);  c_mad(xb, u[3 * i + j], b[j]); }
    }
    xa = c_mul(c, xa); xb = c_mul(c, xb);
    r[0 + i] = c_add(r[0 + i], xa);
    r[3 + i] = c_add(r[3 + i], xb);
    r[3 * Tb::PA + i] = c_comb<conj_code<Tb::CA>::v>(r[3 * Tb::PA + i], xa);
    r[3 * Tb::PB + i] = c_comb<conj_code<Tb::CB>::v>(r[3 * Tb::PB + i], xb);
  }
}

template <class V2> struct tmb_hop_fields {
  const V2 *in;      /* k : field of the opposite parity, 12*Vh elements */
  const V2 *U;       /* gauge, [2][4][9][Vh]  (12-real compression: [2][4][6][Vh]) */
  const V2 *halo_up; /* dist_t: [6][S] (1+g0)-projected first slice of rank+1 */
  const V2 *halo_dn; /* dist_t: [6][S] (1-g0)-projected last slice of rank-1 */
  const V2 *Uhalo;   /* dist_t: [2][9|6][S] U_0 of rank-1's last slice, by owner parity */
};

/* Peer mode: element k in [0, 12 S) of the two halo buffers, PULLED out of the neighbours' copies of the input
 * field (peer memory over NVLink, same layout as the local field): halo_up <- (1+g0) projection of rank+1's
 * first time-slice, halo_dn <- (1-g0) projection of rank-1's last one (what pack_halo + send/recv deliver in
 * the NCCL path).  k < 6 S: halo_up, else halo_dn. */
template <class V2>
TMB_HD void tmb_pull_halo_element(V2 *halo_up, V2 *halo_dn, const V2 *in_up, const V2 *in_dn, const tmb_geom &g, size_t k) {
  const size_t half = (size_t)6 * g.S;
  const bool up = k < half;
  const size_t kk = up ? k : k - half;
  const int c = (int)(kk / g.S), j = (int)(kk - (size_t)c * g.S);
  if (up) halo_up[kk] = c_add(in_up[(size_t)c * g.Vh + j], in_up[(size_t)(c + 6) * g.Vh + j]);
  else {
    const size_t last = (size_t)(g.T - 1) * g.S + j;
    halo_dn[kk] = c_sub(in_dn[(size_t)c * g.Vh + last], in_dn[(size_t)(c + 6) * g.Vh + last]);
  }
}

/* 12-real gauge compression (the reference's CompressionType, misc_types.h:33-37, offered to its
 * external inverters): only the first two rows of an SU(3) link are stored and streamed, the third
 * is conj(row0 x row1).  Exact to rounding for unitary links; refused at upload otherwise. */
template <class V2> TMB_HD void tmb_reconstruct_row2(V2 u[9]) {
  u[6] = c_conj(c_sub(c_mul(u[1], u[5]), c_mul(u[2], u[4])));
  u[7] = c_conj(c_sub(c_mul(u[2], u[3]), c_mul(u[0], u[5])));
  u[8] = c_conj(c_sub(c_mul(u[0], u[4]), c_mul(u[1], u[3])));
}

/* CFG bit 0: cache-policy loads; bit 1: 12-real links (6 stored complex numbers instead of 9) */
template <int D, int CFG, class V2>
TMB_HD void tmb_hop_dir(V2 r[12], const tmb_hop_fields<V2> &f, const tmb_geom &g, int par, int i, int n, V2 ka,
                        const tmb_policies &pol) {
  const int mu = D >> 1, BWD = D & 1, HINTS = CFG & 1, NE = (CFG & 2) ? 6 : 9;
  V2 a[3], b[3], u[9];
  /* forward link lives at the output site (parity par), backward link at the neighbour (parity 1-par) */
  const V2 *ub = f.U + (size_t)(((BWD ? 1 - par : par) * 4 + mu) * NE) * g.Vh + (BWD ? n : i);
#pragma unroll
  for (int e = 0; e < NE; e++) u[e] = tmb_load_gauge<CFG & 5>(ub + (size_t)e * g.Vh, pol);
  tmb_project<D, HINTS>(a, b, f.in, g.Vh, n, pol);
  if (NE == 6) tmb_reconstruct_row2(u);
  tmb_link_accumulate<D>(r, u, a, b, ka);
}

/* halo variants for the distributed T direction: the half-spinor arrives already projected */
template <int D, int CFG, class V2>
TMB_HD void tmb_hop_dir_halo(V2 r[12], const tmb_hop_fields<V2> &f, const tmb_geom &g, int par, int i, int j, V2 ka,
                             const tmb_policies &pol) {
  const int HINTS = CFG & 1, NE = (CFG & 2) ? 6 : 9;
  V2 a[3], b[3], u[9];
  if (D == 0) { /* +t at t == T-1: local forward link, half-spinor from rank+1 */
    const V2 *ub = f.U + (size_t)((par * 4 + 0) * NE) * g.Vh + i;
#pragma unroll
    for (int e = 0; e < NE; e++) u[e] = tmb_load_gauge<CFG & 5>(ub + (size_t)e * g.Vh, pol);
#pragma unroll
    for (int c = 0; c < 3; c++) { a[c] = f.halo_up[(size_t)c * g.S + j]; b[c] = f.halo_up[(size_t)(3 + c) * g.S + j]; }
  } else {      /* -t at t == 0: link and half-spinor from rank-1 */
    const V2 *ub = f.Uhalo + (size_t)((1 - par) * NE) * g.S + j;
#pragma unroll
    for (int e = 0; e < NE; e++) u[e] = ub[(size_t)e * g.S];
#pragma unroll
    for (int c = 0; c < 3; c++) { a[c] = f.halo_dn[(size_t)c * g.S + j]; b[c] = f.halo_dn[(size_t)(3 + c) * g.S + j]; }
  }
  if (NE == 6) tmb_reconstruct_row2(u);
  tmb_link_accumulate<D>(r, u, a, b, ka);
}

/* The full 8-direction sum for output site i of parity par.  ka[mu] = kappa*exp(i theta_mu pi/L_mu)
 * exactly as boundary.c:40-55 computes them on the host. */
template <int DIST, int HINTS, class V2>
TMB_HD void tmb_hop_site(V2 r[12], const tmb_hop_fields<V2> &f, const tmb_geom &g, int par, int i, const V2 ka[4],
                         const tmb_policies &pol) {
  int nb[8];
  const int t = tmb_neighbours(g, par, i, nb);
#pragma unroll
  for (int c = 0; c < 12; c++) r[c] = mk2<V2>(0, 0);
  if (DIST && t == g.T - 1) tmb_hop_dir_halo<0, HINTS>(r, f, g, par, i, i - t * g.S, ka[0], pol);
  else                      tmb_hop_dir<0, HINTS>(r, f, g, par, i, nb[0], ka[0], pol);
  if (DIST && t == 0)       tmb_hop_dir_halo<1, HINTS>(r, f, g, par, i, i, ka[0], pol);
  else                      tmb_hop_dir<1, HINTS>(r, f, g, par, i, nb[1], ka[0], pol);
  tmb_hop_dir<2, HINTS>(r, f, g, par, i, nb[2], ka[1], pol);
  tmb_hop_dir<3, HINTS>(r, f, g, par, i, nb[3], ka[1], pol);
  tmb_hop_dir<4, HINTS>(r, f, g, par, i, nb[4], ka[2], pol);
  tmb_hop_dir<5, HINTS>(r, f, g, par, i, nb[5], ka[2], pol);
  tmb_hop_dir<6, HINTS>(r, f, g, par, i, nb[6], ka[3], pol);
  tmb_hop_dir<7, HINTS>(r, f, g, par, i, nb[7], ka[3], pol);
}

/* Two right-hand sides at once (the two flavours of the non-degenerate doublet, operator/tm_operators_nd.c:
 * every Hopping_Matrix there is applied to a strange and a charm field with the same links): each link is
 * loaded ONCE and applied to both projected spinors, so the 1152 B/site gauge stream is shared:
 * 1920 B per site-pair instead of 2 x 1536. */
template <int D, int CFG, class V2>
TMB_HD void tmb_hop_dir2(V2 r0[12], V2 r1[12], const V2 *in0, const V2 *in1, const V2 *U, const tmb_geom &g, int par, int i,
                         int n, V2 ka, const tmb_policies &pol) {
  const int mu = D >> 1, BWD = D & 1, HINTS = CFG & 1, NE = (CFG & 2) ? 6 : 9;
  V2 a[3], b[3], u[9];
  const V2 *ub = U + (size_t)(((BWD ? 1 - par : par) * 4 + mu) * NE) * g.Vh + (BWD ? n : i);
#pragma unroll
  for (int e = 0; e < NE; e++) u[e] = tmb_load_gauge<HINTS>(ub + (size_t)e * g.Vh, pol);
  if (NE == 6) tmb_reconstruct_row2(u);
  tmb_project<D, HINTS>(a, b, in0, g.Vh, n, pol);
  tmb_link_accumulate<D>(r0, u, a, b, ka);
  tmb_project<D, HINTS>(a, b, in1, g.Vh, n, pol);
  tmb_link_accumulate<D>(r1, u, a, b, ka);
}
template <int CFG, class V2>
TMB_HD void tmb_hop_site2(V2 r0[12], V2 r1[12], const V2 *in0, const V2 *in1, const V2 *U, const tmb_geom &g, int par, int i,
                          const V2 ka[4], const tmb_policies &pol) {
  int nb[8];
  tmb_neighbours(g, par, i, nb);
#pragma unroll
  for (int c = 0; c < 12; c++) { r0[c] = mk2<V2>(0, 0); r1[c] = mk2<V2>(0, 0); }
  tmb_hop_dir2<0, CFG>(r0, r1, in0, in1, U, g, par, i, nb[0], ka[0], pol);
  tmb_hop_dir2<1, CFG>(r0, r1, in0, in1, U, g, par, i, nb[1], ka[0], pol);
  tmb_hop_dir2<2, CFG>(r0, r1, in0, in1, U, g, par, i, nb[2], ka[1], pol);
  tmb_hop_dir2<3, CFG>(r0, r1, in0, in1, U, g, par, i, nb[3], ka[1], pol);
  tmb_hop_dir2<4, CFG>(r0, r1, in0, in1, U, g, par, i, nb[4], ka[2], pol);
  tmb_hop_dir2<5, CFG>(r0, r1, in0, in1, U, g, par, i, nb[5], ka[2], pol);
  tmb_hop_dir2<6, CFG>(r0, r1, in0, in1, U, g, par, i, nb[6], ka[3], pol);
  tmb_hop_dir2<7, CFG>(r0, r1, in0, in1, U, g, par, i, nb[7], ka[3], pol);
}
/* M_ee_inv_ndpsi in registers (tm_operators_nd.c:639-695): (ls, lc) = nrm [ (1 -+ i mu) ks + eps kc , (1 +- i mu) kc + eps ks ],
 * upper sign on s0,s1 (c < 6), lower on s2,s3 */
template <class V2>
TMB_HD void tmb_nd_mee_inv_regs(V2 &ls, V2 &lc, V2 ks, V2 kc, int c, double mu, double eps, double nrm) {
  typedef typename tmb_real<V2>::type R;
  const V2 zs = mk2<V2>((R)1, (R)((c < 6) ? -mu : mu)), zc = c_conj(zs);
  V2 a = c_mul(zs, ks); a.x += (R)eps * kc.x; a.y += (R)eps * kc.y;
  V2 b = c_mul(zc, kc); b.x += (R)eps * ks.x; b.y += (R)eps * ks.y;
  ls = mk2<V2>((R)nrm * a.x, (R)nrm * a.y); lc = mk2<V2>((R)nrm * b.x, (R)nrm * b.y);
}

/* ====================================================================================
 * Second split direction (Z).  Ranks form an (nt x nz) grid (the reference's PARALLELXT.. family, mpi_init.c:331-357, here
 * T then Z); the local z extent is LZ / nz (even, so parities need no offset).  Per row (t, x, y) of the eo layout exactly
 * ONE site of each parity sits on a z face: with r = (t + x + y + parity) & 1, the site zh = Lzh-1 of a row with r = 1 has its
 * +z neighbour on rank z+1, the site zh = 0 of a row with r = 0 has its -z neighbour on rank z-1.  Rows come in (y even, y odd)
 * pairs with opposite r, so face buffers are indexed by j = row >> 1, Sz = T*LX*LY/2 entries.
 *
 * The hopping kernels are NOT touched: they run on the slab as if it were periodic in z (the face sites pick up the slab's
 * own opposite face: the WRONG neighbour), and a fix-up over the face sites replaces that one term by the right one from
 * the halo:  out += L( ka U (h_halo - h_wrap) )  for +z,  out += L( conj(ka) [Uhalo^+ h_halo - Uwrap^+ h_wrap] )  for -z,
 * with L the linear map of the epilogue (MODE 0: +, 1: (cf|conj cf), 2: -g5, 3: -).  Faces are 1/Lzh of the sites.
 *   face buffers  hz[c * Sz + j], c < 6: the projected half-spinor (a0..a2, b0..b2) of direction +z (D = 6, from rank z+1's
 *                 first z) or -z (D = 7, from rank z-1's last z)
 *   Uz halo       Uzh[(q * 9 + e) * Sz + j]: U_z of rank z-1's last-z sites of parity q (for the -z hop of parity 1-q) */
template <class V2>
TMB_HD void tmb_zface_row(const tmb_geom &g, int par, int j, int want_r, int *row, int *t) {
  /* the row of pair j whose r = (t + x + y + par) & 1 equals want_r */
  const int r0 = 2 * j; /* y even */
  const int y = r0 % g.LY; const int tx = r0 / g.LY; const int x = tx % g.LX; const int tt = tx / g.LX;
  (void)y;
  const int re = (tt + x + par) & 1; /* r of the y-even row */
  *row = re == want_r ? r0 : r0 + 1; *t = tt;
}
/* what rank z-1 needs for its +z hops (send_dn: D = 6 projection of this rank's z = 0 sites) and what rank z+1 needs for its
 * -z hops (send_up: D = 7 projection of this rank's z = LZ-1 sites); `in` has parity pin, k in [0, 6 * Sz) */
template <class V2> struct EwPackZFaces { V2 *up, *dn; const V2 *in; tmb_geom g; int pin; V2 *up2 = nullptr, *dn2 = nullptr; /* second copies (peer push: remote + local) */
  __host__ __device__ void operator()(size_t k) const {
    const int Sz = g.T * g.LX * g.LY / 2;
    const int c = (int)(k / Sz), j = (int)(k - (size_t)c * Sz);
    int row, t;
    /* z = 0 site of the input parity: zh = 0 in a row with r_in = 0 */
    tmb_zface_row<V2>(g, pin, j, 0, &row, &t);
    {
      const size_t i = (size_t)row * g.Lzh;
      const int cc = c % 3, hi = c / 3; /* hi 0: a = s0 + i s2, hi 1: b = s1 - i s3 */
      const V2 x = in[(size_t)(hi ? 3 + cc : cc) * g.Vh + i], y = in[(size_t)(hi ? 9 + cc : 6 + cc) * g.Vh + i];
      const V2 v = hi ? c_comb<3>(x, y) : c_comb<2>(x, y);
      dn[k] = v; if (dn2 != nullptr) dn2[k] = v;
    }
    tmb_zface_row<V2>(g, pin, j, 1, &row, &t);
    {
      const size_t i = (size_t)row * g.Lzh + (g.Lzh - 1);
      const int cc = c % 3, hi = c / 3; /* hi 0: a = s0 - i s2, hi 1: b = s1 + i s3 */
      const V2 x = in[(size_t)(hi ? 3 + cc : cc) * g.Vh + i], y = in[(size_t)(hi ? 9 + cc : 6 + cc) * g.Vh + i];
      const V2 v = hi ? c_comb<2>(x, y) : c_comb<3>(x, y);
      up[k] = v; if (up2 != nullptr) up2[k] = v;
    }
  } };
/* U_z of this rank's last-z sites, both owner parities: out[(q * 9 + e) * Sz + j] (what rank z+1 needs), k in [0, 18 * Sz) */
template <class V2> struct EwPackGaugeZHalo { V2 *out; const V2 *U; tmb_geom g;
  __host__ __device__ void operator()(size_t k) const {
    const int Sz = g.T * g.LX * g.LY / 2;
    const int j = (int)(k % Sz); const int qe = (int)(k / Sz); const int e = qe % 9, q = qe / 9;
    int row, t;
    tmb_zface_row<V2>(g, q, j, 1, &row, &t); /* the site of parity q at zh = Lzh-1 has z = LZ-1 iff its row has r = 1 */
    out[k] = U[(size_t)((q * 4 + 3) * 9 + e) * g.Vh + (size_t)row * g.Lzh + (g.Lzh - 1)];
  } };
/* the fix-up of one face site pair j of the OUTPUT parity par; U is the full 18-real link field */
/* own_up / own_dn: this rank's OWN projected faces, as packed for the neighbours ([6][Sz], the same row enumeration): the
 * wrapped term the kernel used for the +z hop of the last-z sites is the D = 6 projection of the slab's first-z sites =
 * own_dn, for the -z hop of the first-z sites it is own_up.  Uzl: this rank's own U_z of the last-z sites ([2][9][Sz], what
 * it sent to rank z+1).  With these every operand of the fix-up except `out` itself is read contiguously; nullptr (the CPU
 * emulation, older callers) = gather them from `in` and `U` with the stride of a z row. */
template <int MODE, class V2>
TMB_HD void tmb_zfix_side(V2 *out, const V2 *in, const V2 *U, const V2 *hz_up, const V2 *hz_dn, const V2 *Uzh, const tmb_geom &g,
                          int par, int j, int side, V2 ka3, V2 cf, const V2 *own_up = nullptr, const V2 *own_dn = nullptr,
                          const V2 *Uzl = nullptr) {
  const int Sz = g.T * g.LX * g.LY / 2;
  int row, t;
  tmb_zface_row<V2>(g, par, j, side ? 0 : 1, &row, &t); /* side 0: the +z face site (r = 1, zh = Lzh-1), 1: the -z face site (r = 0, zh = 0) */
  const size_t i = (size_t)row * g.Lzh + (side ? 0 : g.Lzh - 1);
  const size_t iw = (size_t)row * g.Lzh + (side ? g.Lzh - 1 : 0); /* the slab's own opposite face: what the kernel used */
  V2 d[12];
#pragma unroll
  for (int c = 0; c < 12; c++) d[c] = mk2<V2>(0, 0);
  V2 ah[3], bh[3], aw[3], bw[3], u[9];
  const V2 *hz = side ? hz_dn : hz_up;
#pragma unroll
  for (int c = 0; c < 3; c++) { ah[c] = hz[(size_t)c * Sz + j]; bh[c] = hz[(size_t)(3 + c) * Sz + j]; }
  tmb_policies pol = {0, 0};
  const V2 *own = side ? own_up : own_dn;
  if (own != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; c++) { aw[c] = own[(size_t)c * Sz + j]; bw[c] = own[(size_t)(3 + c) * Sz + j]; }
  } else if (!side) tmb_project<6, 0>(aw, bw, in, g.Vh, (int)iw, pol);
  else tmb_project<7, 0>(aw, bw, in, g.Vh, (int)iw, pol);
  if (!side) { /* +z: same local link for both terms */
#pragma unroll
    for (int c = 0; c < 3; c++) { ah[c] = c_sub(ah[c], aw[c]); bh[c] = c_sub(bh[c], bw[c]); }
#pragma unroll
    for (int e = 0; e < 9; e++) u[e] = Uzl != nullptr ? Uzl[(size_t)(par * 9 + e) * Sz + j] : U[(size_t)((par * 4 + 3) * 9 + e) * g.Vh + i];
    tmb_link_accumulate<6>(d, u, ah, bh, ka3);
  } else {     /* -z: the right link comes with the halo, the wrong one is the local link at the wrapped neighbour */
#pragma unroll
    for (int e = 0; e < 9; e++) u[e] = Uzh[(size_t)((1 - par) * 9 + e) * Sz + j];
    tmb_link_accumulate<7>(d, u, ah, bh, ka3);
#pragma unroll
    for (int c = 0; c < 3; c++) { aw[c] = mk2<V2>(-aw[c].x, -aw[c].y); bw[c] = mk2<V2>(-bw[c].x, -bw[c].y); }
#pragma unroll
    for (int e = 0; e < 9; e++) u[e] = Uzl != nullptr ? Uzl[(size_t)((1 - par) * 9 + e) * Sz + j] : U[(size_t)(((1 - par) * 4 + 3) * 9 + e) * g.Vh + iw];
    tmb_link_accumulate<7>(d, u, aw, bw, ka3);
  }
#pragma unroll
  for (int c = 0; c < 12; c++) {
    V2 o = out[(size_t)c * g.Vh + i], x = d[c];
    if (MODE == 1) x = c_mul((c < 6) ? cf : c_conj(cf), x);
    else if (MODE == 2) x = (c < 6) ? mk2<V2>(-x.x, -x.y) : x;
    else if (MODE == 3) x = mk2<V2>(-x.x, -x.y);
    out[(size_t)c * g.Vh + i] = c_add(o, x);
  }
}
template <int MODE, class V2>
TMB_HD void tmb_zfix_pair(V2 *out, const V2 *in, const V2 *U, const V2 *hz_up, const V2 *hz_dn, const V2 *Uzh, const tmb_geom &g,
                          int par, int j, V2 ka3, V2 cf) {
  tmb_zfix_side<MODE>(out, in, U, hz_up, hz_dn, Uzh, g, par, j, 0, ka3, cf);
  tmb_zfix_side<MODE>(out, in, U, hz_up, hz_dn, Uzh, g, par, j, 1, ka3, cf);
}

/* Epilogues (hopping.h:674-694):
 *   MODE 0  l = r                                   _store_res                 Hopping_Matrix
 *   MODE 1  l = (cf on s0,s1 | conj(cf) on s2,s3) r _hop_mul_g5_cmplx_and_store tm_times_Hopping_Matrix
 *   MODE 2  l = g5( (cf|conj cf) p - r )            _g5_cmplx_sub_hop_and_g5store tm_sub_Hopping_Matrix
 *   MODE 3  l = (cf|conj cf) p - r                  (M_full / D_psi rows: tm_operators.c:117-128)
 */
template <int MODE, class V2>
TMB_HD V2 tmb_epilogue(int c, V2 r, V2 p, V2 cf) {
  if (MODE == 0) return r;
  const V2 f = (c < 6) ? cf : c_conj(cf);
  if (MODE == 1) return c_mul(f, r);
  const V2 zp = c_mul(f, p);
  if (MODE == 2) return (c < 6) ? c_sub(zp, r) : c_sub(r, zp);
  return c_sub(zp, r);
}

/* ====================================================================================
 * Fermion force: deriv_Sb(ieo, l, k, hf, factor), deriv_Sb.c:402-649 (generic C branch).
 *
 * The reference loops over the sites x of parity ieo and SCATTERS two terms per direction:
 *   forward   hf->derivative[x][mu]    += 2 factor tr_lambda( ka_mu U_mu(x)    v^+ ),  v = phi(x) psi(x+mu)^+
 *   backward  hf->derivative[x-mu][mu] += 2 factor tr_lambda( ka_mu U_mu(x-mu) v^+ ),  v = psi(x-mu) phi(x)^+
 * with phi = P_d(g5 l), psi = P_d(k), P_d the hopping projector of direction d (deriv_Sb.c:470-640).
 * Every link (z, mu) of the lattice receives exactly ONE of the two terms per call: the forward one
 * if z has parity ieo, the backward one (seen from x = z+mu) otherwise.  So the device version is a
 * GATHER over link owners z of either parity, with no atomics:
 *   v = P_d(local(z)) P_d(remote(z+mu))^+,   d = 2mu (forward type) or 2mu+1 (backward type),
 *   forward type : local = g5 l(z), remote = k(z+mu);   backward type: local = k(z), remote = g5 l(z+mu)
 *   df[z][mu] += 2 factor tr_lambda( ka_mu U_mu(z) v^+ )      (su3.h:605-614, :706-715; su3adj.h:164-172)
 * Momentum-derivative field on the device: df[((q*4 + mu)*8 + a)*Vh + i], same (q, mu, i) indexing
 * as the gauge field, a = 0..7 the su3adj components d1..d8 (su3adj.h:25-27).
 * ==================================================================================== */
template <int D, class V2>
TMB_HD void tmb_project_regs(V2 a[3], V2 b[3], const V2 s[12]) {
  typedef hop_tab<D> Tb;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    a[c] = c_comb<Tb::CA>(s[c], s[3 * Tb::PA + c]);
    b[c] = c_comb<Tb::CB>(s[3 + c], s[3 * Tb::PB + c]);
  }
}

/* r_a += c * Re tr( w lambda_a )-style projection, exactly _trace_lambda_mul_add_assign (su3adj.h:164-172) */
TMB_HD void tmb_trace_lambda_add(double r[8], double c, const double2 w[9]) {
  r[0] += c * (-w[3].y - w[1].y);
  r[1] += c * (+w[3].x - w[1].x);
  r[2] += c * (-w[0].y + w[4].y);
  r[3] += c * (-w[6].y - w[2].y);
  r[4] += c * (+w[6].x - w[2].x);
  r[5] += c * (-w[7].y - w[5].y);
  r[6] += c * (+w[7].x - w[5].x);
  r[7] += c * ((-w[0].y - w[4].y + 2.0 * w[8].y) * 0.577350269189625);
}

/* one link: la/lb = projected local half-spinor, ra/rb = projected remote half-spinor */
TMB_HD void tmb_deriv_link(double r[8], const double2 la[3], const double2 lb[3], const double2 ra[3],
                           const double2 rb[3], const double2 u[9], double2 ka, double c) {
  double2 v[9], w[9];
  /* _vector_tensor_vector_add: v_ij = la_i conj(ra_j) + lb_i conj(rb_j) */
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double2 t = make_double2(0., 0.);
      c_madc(t, ra[j], la[i]);
      c_madc(t, rb[j], lb[i]);
      v[3 * i + j] = t;
    }
  /* _su3_times_su3d: w_ij = sum_k u_ik conj(v_jk);  then _complex_times_su3 with ka */
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double2 t = make_double2(0., 0.);
#pragma unroll
      for (int k = 0; k < 3; k++) c_madc(t, v[3 * j + k], u[3 * i + k]);
      w[3 * i + j] = c_mul(ka, t);
    }
  tmb_trace_lambda_add(r, c, w);
}

struct tmb_deriv_fields {
  const double2 *l, *k;   /* l: field of parity ieo (gets the g5), k: field of parity 1-ieo */
  const double2 *U;       /* [2][4][9][Vh] */
  double *df;             /* [2][4][8][Vh] */
  const double2 *halo_k;  /* dist_t: [6][S] (1+g0)-projected first slice of rank+1's k */
  const double2 *halo_l;  /* dist_t: [6][S] (1+g0)-projected first slice of rank+1's l (= P_1(g5 l)) */
};

template <int MU, int FWD, int DIST>
TMB_HD void tmb_deriv_dir(const tmb_deriv_fields &f, const tmb_geom &g, int q, int i, int n, int t,
                          const double2 loc[12], double2 ka, double c) {
  const int D = 2 * MU + (FWD ? 0 : 1);
  double2 la[3], lb[3], ra[3], rb[3], u[9];
  tmb_project_regs<D>(la, lb, loc);
  if (DIST && MU == 0 && t == g.T - 1) {
    const double2 *h = FWD ? f.halo_k : f.halo_l;
    const int j = i - t * g.S;
#pragma unroll
    for (int cidx = 0; cidx < 3; cidx++) { ra[cidx] = h[(size_t)cidx * g.S + j]; rb[cidx] = h[(size_t)(3 + cidx) * g.S + j]; }
  } else {
    const double2 *rf = FWD ? f.k : f.l;
    double2 rem[12];
#pragma unroll
    for (int cidx = 0; cidx < 12; cidx++) {
      double2 s = rf[(size_t)cidx * g.Vh + n];
      if (!FWD && cidx >= 6) s = make_double2(-s.x, -s.y); /* g5 on the l-derived spinor */
      rem[cidx] = s;
    }
    tmb_project_regs<D>(ra, rb, rem);
  }
  const double2 *ub = f.U + (size_t)((q * 4 + MU) * 9) * g.Vh + i;
#pragma unroll
  for (int e = 0; e < 9; e++) u[e] = ub[(size_t)e * g.Vh];
  double *d = f.df + (size_t)((q * 4 + MU) * 8) * g.Vh + i;
  double r[8];
#pragma unroll
  for (int a = 0; a < 8; a++) r[a] = d[(size_t)a * g.Vh];
  tmb_deriv_link(r, la, lb, ra, rb, u, ka, c);
#pragma unroll
  for (int a = 0; a < 8; a++) d[(size_t)a * g.Vh] = r[a];
}

/* all four links owned by site i of parity q; FWD = (q == ieo) */
template <int FWD, int DIST>
TMB_HD void tmb_deriv_site(const tmb_deriv_fields &f, const tmb_geom &g, int q, int i, const double2 ka[4], double c) {
  int nb[8];
  const int t = tmb_neighbours(g, q, i, nb);
  const double2 *lf = FWD ? f.l : f.k;
  double2 loc[12];
#pragma unroll
  for (int cidx = 0; cidx < 12; cidx++) {
    double2 s = lf[(size_t)cidx * g.Vh + i];
    if (FWD && cidx >= 6) s = make_double2(-s.x, -s.y);
    loc[cidx] = s;
  }
  tmb_deriv_dir<0, FWD, DIST>(f, g, q, i, nb[0], t, loc, ka[0], c);
  tmb_deriv_dir<1, FWD, DIST>(f, g, q, i, nb[2], t, loc, ka[1], c);
  tmb_deriv_dir<2, FWD, DIST>(f, g, q, i, nb[4], t, loc, ka[2], c);
  tmb_deriv_dir<3, FWD, DIST>(f, g, q, i, nb[6], t, loc, ka[3], c);
}

/* Z split: deriv_kernel runs on the slab as if it were periodic in z, so the z links owned by the LAST-z sites took their
 * remote half-spinor from the slab's own first-z sites.  The term is linear in the remote half-spinor, hence
 *   df[x][3] += link( P(local(x)), h_halo - h_wrapped )
 * with h_halo from rank z+1's first-z sites.  Forward type (x of parity ieo) needs P_6(k(x+z)), backward type P_7(g5 l(x+z)) -
 * and P_7 after g5 is the P_6 formula on l itself (s0 + i s2, s1 - i s3), so both halos are the `dn` faces of
 * EwPackZFaces (D = 6 projection of the first-z sites) of k and of l.  halo_k / halo_l: [6][Sz], row enumeration of the
 * receiving side (tmb_zface_row(g, q, j, 1)).  j in [0, Sz), q = parity of the link owner. */
TMB_HD void tmb_deriv_zfix(const tmb_deriv_fields &f, const tmb_geom &g, int ieo, int q, int j, const double2 *halo_k,
                           const double2 *halo_l, double2 ka3, double c) {
  const int Sz = g.T * g.LX * g.LY / 2;
  const bool fwd = q == ieo;
  int row, t;
  tmb_zface_row<double2>(g, q, j, 1, &row, &t);
  const size_t i = (size_t)row * g.Lzh + (g.Lzh - 1), iw = (size_t)row * g.Lzh;
  const double2 *lf = fwd ? f.l : f.k, *rf = fwd ? f.k : f.l, *h = fwd ? halo_k : halo_l;
  double2 loc[12], rem[12], la[3], lb[3], ra[3], rb[3], u[9];
#pragma unroll
  for (int cidx = 0; cidx < 12; cidx++) {
    double2 s = lf[(size_t)cidx * g.Vh + i], w = rf[(size_t)cidx * g.Vh + iw];
    if (cidx >= 6) { if (fwd) s = make_double2(-s.x, -s.y); else w = make_double2(-w.x, -w.y); } /* g5 on the l-derived spinor */
    loc[cidx] = s; rem[cidx] = w;
  }
  if (fwd) { tmb_project_regs<6>(la, lb, loc); tmb_project_regs<6>(ra, rb, rem); }
  else     { tmb_project_regs<7>(la, lb, loc); tmb_project_regs<7>(ra, rb, rem); }
#pragma unroll
  for (int cidx = 0; cidx < 3; cidx++) {
    ra[cidx] = c_sub(h[(size_t)cidx * Sz + j], ra[cidx]);
    rb[cidx] = c_sub(h[(size_t)(3 + cidx) * Sz + j], rb[cidx]);
  }
  const double2 *ub = f.U + (size_t)((q * 4 + 3) * 9) * g.Vh + i;
#pragma unroll
  for (int e = 0; e < 9; e++) u[e] = ub[(size_t)e * g.Vh];
  double *d = f.df + (size_t)((q * 4 + 3) * 8) * g.Vh + i;
  double r[8];
#pragma unroll
  for (int a = 0; a < 8; a++) r[a] = d[(size_t)a * g.Vh];
  tmb_deriv_link(r, la, lb, ra, rb, u, ka3, c);
#pragma unroll
  for (int a = 0; a < 8; a++) d[(size_t)a * g.Vh] = r[a];
}

/* ====================================================================================
 * Plaquette: measure_plaquette, measure_gauge_action.c:46-106 - what tmLQCD_read_gauge prints after reading a
 * configuration (wrapper/lib_wrapper.c:232-235) and every main stores as plaquette_energy.
 *   P(x) = sum_{mu1 < mu2} Re tr( U_mu1(x) U_mu2(x+mu1) [ U_mu2(x) U_mu1(x+mu2) ]^dagger )
 * One thread per site of either parity: 4 own links + 12 links of the forward neighbours, read from the same
 * [parity][mu][9][Vh] field the hopping kernel uses (compulsory traffic 576 B/site, neighbours through L2).
 * Uup (T split only): the spatial links U_1..3 of rank+1's first time-slice, [2][3][9][S] by owner parity.
 * ==================================================================================== */
TMB_HD void tmb_load_link(double2 u[9], const double2 *U, const tmb_geom &g, int q, int mu, int i) {
  const double2 *ub = U + (size_t)((q * 4 + mu) * 9) * g.Vh + i;
#pragma unroll
  for (int e = 0; e < 9; e++) u[e] = ub[(size_t)e * g.Vh];
}
TMB_HD void tmb_su3_times_su3(double2 r[9], const double2 a[9], const double2 b[9]) { /* su3.h:583-592 */
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double2 x = c_mul(a[3 * i], b[j]);
      c_mad(x, a[3 * i + 1], b[3 + j]);
      c_mad(x, a[3 * i + 2], b[6 + j]);
      r[3 * i + j] = x;
    }
}
template <int DIST>
TMB_HD double tmb_plaq_site(const double2 *U, const double2 *Uup, const tmb_geom &g, int q, int i) {
  int nb[8];
  const int t = tmb_neighbours(g, q, i, nb);
  double2 own[4][9];
#pragma unroll
  for (int mu = 0; mu < 4; mu++) tmb_load_link(own[mu], U, g, q, mu, i);
  double sum = 0.;
#pragma unroll
  for (int mu1 = 0; mu1 < 3; mu1++)
#pragma unroll
    for (int mu2 = mu1 + 1; mu2 < 4; mu2++) {
      double2 w[9], w2[9], pr1[9], pr2[9];
      if (DIST && mu1 == 0 && t == g.T - 1) { /* U_mu2(x + t) lives on the rank above */
        const double2 *ub = Uup + (size_t)(((1 - q) * 3 + (mu2 - 1)) * 9) * g.S + (i - t * g.S);
#pragma unroll
        for (int e = 0; e < 9; e++) w[e] = ub[(size_t)e * g.S];
      } else tmb_load_link(w, U, g, 1 - q, mu2, nb[2 * mu1]);
      tmb_load_link(w2, U, g, 1 - q, mu1, nb[2 * mu2]); /* mu2 >= 1: a spatial shift, always local */
      tmb_su3_times_su3(pr1, own[mu1], w);
      tmb_su3_times_su3(pr2, own[mu2], w2);
#pragma unroll
      for (int e = 0; e < 9; e++) { sum += pr1[e].x * pr2[e].x; sum += pr1[e].y * pr2[e].y; } /* su3.h:656-665 */
    }
  return sum;
}

/* tmb_emul.cu - TEST-ONLY host emulation of the device site functions.
 *
 * This container has no GPU.  To check the device code's layout conversion, neighbour
 * arithmetic, spin projection tables, epilogues and the T-slab halo logic on the CPU, this
 * file includes the product kernel source and calls its __host__ __device__ site functions
 * and functors from plain host loops.  It is built into tests/emul/libtmb_emul.so by
 * tests/emul/build.sh and loaded only by tests/ (-m "not gpu").  It is NOT part of the
 * product library and is not a CPU fallback: tmlqcd_b200/csrc never references it.
 */
#include "../../tmlqcd_b200/csrc/tmb_kernels.cu"
#include "../../tmlqcd_b200/csrc/tmb_force.cu"
/* the two-flavour instantiations (tmb_hop2.cu) are not needed by the host emulation: nothing here launches a kernel */
cudaError_t tmb_launch_hop_nd(const tmb_hop_launch &, cudaStream_t) { return cudaErrorNotSupported; }

template <int MODE, int DIST>
static void hop_host(double2 *out, const tmb_hop_fields<double2> &f, const double2 *p, const tmb_geom &g, int par,
                     const double2 ka[4], double2 cf, int site0, int nsites, int split, int gap) {
  tmb_policies pol = {0, 0};
  for (int w = 0; w < nsites; w++) {
    const int i = site0 + w + (w >= split ? gap : 0);
    double2 r[12];
    tmb_hop_site<DIST, 0>(r, f, g, par, i, ka, pol);
    for (int c = 0; c < 12; c++) {
      double2 pc = make_double2(0., 0.);
      if (MODE >= 2) pc = p[(size_t)c * g.Vh + i];
      out[(size_t)c * g.Vh + i] = tmb_epilogue<MODE>(c, r[c], pc, cf);
    }
  }
}

extern "C" {

void emul_pack_eo(double *soa, const double *aos, int Vh) {
  EwPackEo f = {(double2 *)soa, (const double2 *)aos, Vh};
  for (size_t k = 0; k < (size_t)12 * Vh; k++) f(k);
}
void emul_unpack_eo(double *aos, const double *soa, int Vh) {
  EwUnpackEo f = {(double2 *)aos, (const double2 *)soa, Vh};
  for (size_t k = 0; k < (size_t)12 * Vh; k++) f(k);
}
void emul_pack_lexic(double *even, double *odd, const double *lex, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  EwPackLex<double2> f = {(double2 *)even, (double2 *)odd, (const double2 *)lex, g};
  for (size_t k = 0; k < (size_t)24 * g.Vh; k++) f(k);
}
void emul_unpack_lexic(double *lex, const double *even, const double *odd, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  EwUnpackLex<double2> f = {(double2 *)lex, (const double2 *)even, (const double2 *)odd, g};
  for (size_t k = 0; k < (size_t)24 * g.Vh; k++) f(k);
}
void emul_pack_gauge(double *U, const double *lex, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  EwPackGauge f = {(double2 *)U, (const double2 *)lex, g};
  for (size_t k = 0; k < (size_t)72 * g.Vh; k++) f(k);
}
void emul_pack_halo(double *up, double *dn, const double *in, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 1);
  EwPackHalo<double2> f = {(double2 *)up, (double2 *)dn, (const double2 *)in, g};
  for (size_t k = 0; k < (size_t)6 * g.S; k++) f(k);
}
void emul_pack_gauge_halo(double *out, const double *U, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 1);
  EwPackGaugeHalo f = {(double2 *)out, (const double2 *)U, g};
  for (size_t k = 0; k < (size_t)18 * g.S; k++) f(k);
}
/* neighbour table by the closed forms of tmb_geom.h: nb[8*i+d], for comparison with g_hi */
void emul_neighbours(int *nb, int par, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  for (int i = 0; i < g.Vh; i++) tmb_neighbours(g, par, i, nb + 8 * i);
}
void emul_eo2lexic(int *out, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  for (int par = 0; par < 2; par++)
    for (int i = 0; i < g.Vh; i++) out[par * g.Vh + i] = tmb_eo_to_lexic(g, par, i);
}

/* the full hopping term as the product's hop() composes it: one launch when dist == 0,
 * interior + boundary ranges when dist == 1 (same site0/nsites/split/gap arithmetic) */
int emul_hop(int par, double *out, const double *in, const double *p, const double *U, const double *halo_up,
             const double *halo_dn, const double *Uhalo, int T, int LX, int LY, int LZ, const double *ka8,
             double cre, double cim, int mode, int dist) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, dist);
  tmb_hop_fields<double2> f;
  f.in = (const double2 *)in; f.U = (const double2 *)U;
  f.halo_up = (const double2 *)halo_up; f.halo_dn = (const double2 *)halo_dn; f.Uhalo = (const double2 *)Uhalo;
  double2 ka[4];
  for (int m = 0; m < 4; m++) ka[m] = make_double2(ka8[2 * m], ka8[2 * m + 1]);
  const double2 cf = make_double2(cre, cim);
  double2 *o = (double2 *)out; const double2 *pp = (const double2 *)p;
#define GO(M, D, s0, n, sp, gp) hop_host<M, D>(o, f, pp, g, par, ka, cf, s0, n, sp, gp)
#define MODES(D, s0, n, sp, gp) \
  switch (mode) { case 0: GO(0, D, s0, n, sp, gp); break; case 1: GO(1, D, s0, n, sp, gp); break; \
                  case 2: GO(2, D, s0, n, sp, gp); break; case 3: GO(3, D, s0, n, sp, gp); break; default: return -1; }
  if (!dist) { MODES(0, 0, g.Vh, g.Vh, 0); }
  else {
    if (g.Vh > 2 * g.S) { MODES(0, g.S, g.Vh - 2 * g.S, g.Vh - 2 * g.S, 0); }
    MODES(1, 0, 2 * g.S, g.S, g.Vh - 2 * g.S);
  }
  return 0;
}

/* 12-real compressed links: same composition with CFG bit 1 set */
int emul_hop12(int par, double *out, const double *in, const double *U12, const double *halo_up, const double *halo_dn,
               const double *Uhalo12, int T, int LX, int LY, int LZ, const double *ka8, int dist) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, dist);
  tmb_hop_fields<double2> f;
  f.in = (const double2 *)in; f.U = (const double2 *)U12;
  f.halo_up = (const double2 *)halo_up; f.halo_dn = (const double2 *)halo_dn; f.Uhalo = (const double2 *)Uhalo12;
  double2 ka[4];
  for (int m = 0; m < 4; m++) ka[m] = make_double2(ka8[2 * m], ka8[2 * m + 1]);
  tmb_policies pol = {0, 0};
  double2 *o = (double2 *)out;
  for (int i = 0; i < g.Vh; i++) {
    double2 r[12];
    if (dist) tmb_hop_site<1, 2>(r, f, g, par, i, ka, pol); else tmb_hop_site<0, 2>(r, f, g, par, i, ka, pol);
    for (int c = 0; c < 12; c++) o[(size_t)c * g.Vh + i] = r[c];
  }
  return 0;
}
void emul_compress12(double *dst, const double *src, long n, int nlinks) {
  EwCompress12 f = {(double2 *)dst, (const double2 *)src, (size_t)n};
  for (size_t k = 0; k < (size_t)nlinks * 6 * n; k++) f(k);
}
/* second split direction (Z): face pack, gauge z-halo and the fix-up of the face sites, as the device composes them */
void emul_pack_zfaces(double *up, double *dn, const double *in, int T, int LX, int LY, int LZ, int pin) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  EwPackZFaces<double2> f = {(double2 *)up, (double2 *)dn, (const double2 *)in, g, pin};
  for (size_t k = 0; k < (size_t)6 * (T * LX * LY / 2); k++) f(k);
}
void emul_pack_gauge_zhalo(double *out, const double *U, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  EwPackGaugeZHalo<double2> f = {(double2 *)out, (const double2 *)U, g};
  for (size_t k = 0; k < (size_t)18 * (T * LX * LY / 2); k++) f(k);
}
void emul_zfix(int mode, double *out, const double *in, const double *U, const double *hz_up, const double *hz_dn, const double *Uzh,
               int T, int LX, int LY, int LZ, int par, const double *ka8, double cre, double cim) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  const double2 ka3 = make_double2(ka8[6], ka8[7]), cf = make_double2(cre, cim);
  for (int j = 0; j < T * LX * LY / 2; j++) {
    if (mode == 0) tmb_zfix_pair<0>((double2 *)out, (const double2 *)in, (const double2 *)U, (const double2 *)hz_up, (const double2 *)hz_dn, (const double2 *)Uzh, g, par, j, ka3, cf);
    else if (mode == 1) tmb_zfix_pair<1>((double2 *)out, (const double2 *)in, (const double2 *)U, (const double2 *)hz_up, (const double2 *)hz_dn, (const double2 *)Uzh, g, par, j, ka3, cf);
    else if (mode == 2) tmb_zfix_pair<2>((double2 *)out, (const double2 *)in, (const double2 *)U, (const double2 *)hz_up, (const double2 *)hz_dn, (const double2 *)Uzh, g, par, j, ka3, cf);
    else tmb_zfix_pair<3>((double2 *)out, (const double2 *)in, (const double2 *)U, (const double2 *)hz_up, (const double2 *)hz_dn, (const double2 *)Uzh, g, par, j, ka3, cf);
  }
}
/* single precision instantiation of the same site code */
int emul_hop_f(int par, float *out, const float *in, const float *U, int T, int LX, int LY, int LZ, const double *ka8) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  tmb_hop_fields<float2> f;
  f.in = (const float2 *)in; f.U = (const float2 *)U; f.halo_up = f.halo_dn = f.Uhalo = nullptr;
  float2 ka[4];
  for (int m = 0; m < 4; m++) ka[m] = make_float2((float)ka8[2 * m], (float)ka8[2 * m + 1]);
  tmb_policies pol = {0, 0};
  float2 *o = (float2 *)out;
  for (int i = 0; i < g.Vh; i++) {
    float2 r[12];
    tmb_hop_site<0, 0>(r, f, g, par, i, ka, pol);
    for (int c = 0; c < 12; c++) o[(size_t)c * g.Vh + i] = r[c];
  }
  return 0;
}

/* x-blocked traversal: the permutation of work indices used by hop_kernel must be a bijection */
void emul_xblock_perm(int *out, int T, int LX, int LY, int LZ, int XB) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  const int P = g.LY * g.Lzh;
  for (int ww = 0; ww < g.Vh; ww++) {
    const int plane = ww / P, off = ww - plane * P;
    const int per = g.T * XB;
    const int xb = plane / per, rem = plane - xb * per;
    const int t = rem / XB, xi = rem - t * XB;
    out[ww] = (t * g.LX + xb * XB + xi) * P + off;
  }
}

/* chunk schedule of the host-pointer pipeline (tmb_geom.h) */
int emul_host_chunk_schedule(int nt, int small, int *sizes) { return tmb_host_chunk_schedule(nt, small, sizes); }

/* CTA tile traversal (tmb_tile_site / tmb_tile_t of tmb_geom.h, used by hop_kernel and hop2_kernel); returns tmb_tile_ok */
int emul_tile_perm(int *site, int *tslice, int T, int LX, int LY, int LZ, int tshift) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  if (!tmb_tile_ok(g)) return 0;
  for (int w = 0; w < g.Vh; w++) { site[w] = tmb_tile_site(g, w, tshift); tslice[w] = tmb_tile_t(g, w, tshift); }
  return 1;
}

/* peer mode: the halo buffers as the copy CTAs of hop_kernel pull them out of the neighbours' fields */
void emul_pull_halo(double *halo_up, double *halo_dn, const double *in_up, const double *in_dn, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 1);
  for (size_t k = 0; k < (size_t)12 * g.S; k++)
    tmb_pull_halo_element((double2 *)halo_up, (double2 *)halo_dn, (const double2 *)in_up, (const double2 *)in_dn, g, k);
}

/* elementwise functors */
void emul_diag(double *l, const double *k, double zre, double zim, int Vh) {
  EwDiag f = {(double2 *)l, (const double2 *)k, make_double2(zre, zim), (size_t)6 * Vh};
  for (size_t i = 0; i < (size_t)12 * Vh; i++) f(i);
}
void emul_diag_sub(double *l, const double *k, const double *j, double zre, double zim, int g5, int Vh) {
  EwDiagSub f = {(double2 *)l, (const double2 *)k, (const double2 *)j, make_double2(zre, zim), g5, (size_t)6 * Vh};
  for (size_t i = 0; i < (size_t)12 * Vh; i++) f(i);
}
void emul_gamma5(double *l, const double *k, int Vh) {
  EwG5 f = {(double2 *)l, (const double2 *)k, (size_t)6 * Vh};
  for (size_t i = 0; i < (size_t)12 * Vh; i++) f(i);
}
void emul_nd_mee_inv(double *ls, double *lc, const double *ks, const double *kc, double mu, double eps, int Vh) {
  EwNdMeeInv f = {(double2 *)ls, (double2 *)lc, (const double2 *)ks, (const double2 *)kc, mu, eps,
                  1. / (1. + mu * mu - eps * eps), (size_t)6 * Vh};
  for (size_t i = 0; i < (size_t)12 * Vh; i++) f(i);
}
void emul_nd_moo_sub_g5(double *ls, double *lc, const double *ks, const double *kc, const double *js,
                        const double *jc, double mu, double eps, int Vh) {
  EwNdMooSubG5 f = {(double2 *)ls, (double2 *)lc, (const double2 *)ks, (const double2 *)kc,
                    (const double2 *)js, (const double2 *)jc, mu, eps, (size_t)6 * Vh};
  for (size_t i = 0; i < (size_t)12 * Vh; i++) f(i);
}
/* fermion force: the gather over link owners exactly as deriv_kernel composes it (interior slices with
 * DIST = 0, last slice with DIST = 1 reading the (1+g0)-projected halo of the upper neighbour) */
void emul_pack_deriv(double *dev, const double *lex, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  for (size_t x = 0; x < (size_t)64 * g.Vh; x++) {
    const int i = (int)(x % g.Vh), row = (int)(x / g.Vh), a = row & 7, mu = (row >> 3) & 3, q = row >> 5;
    dev[x] = lex[((size_t)tmb_eo_to_lexic(g, q, i) * 4 + mu) * 8 + a];
  }
}
void emul_unpack_deriv(double *lex, const double *dev, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  for (size_t x = 0; x < (size_t)64 * g.Vh; x++) {
    const int i = (int)(x % g.Vh), row = (int)(x / g.Vh), a = row & 7, mu = (row >> 3) & 3, q = row >> 5;
    lex[((size_t)tmb_eo_to_lexic(g, q, i) * 4 + mu) * 8 + a] = dev[x];
  }
}
void emul_pack_deriv_halo(double *out, const double *k, const double *l, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 1);
  const double2 *K = (const double2 *)k, *L = (const double2 *)l; double2 *o = (double2 *)out;
  for (size_t x = 0; x < (size_t)12 * g.S; x++) {
    const int which = (int)(x / ((size_t)6 * g.S)); const size_t y = x - (size_t)which * 6 * g.S;
    const int c = (int)(y / g.S), j = (int)(y - (size_t)c * g.S);
    const double2 *f = which ? L : K;
    o[x] = c_add(f[(size_t)c * g.Vh + j], f[(size_t)(c + 6) * g.Vh + j]);
  }
}
void emul_deriv(int ieo, const double *l, const double *k, const double *U, double *df, const double *halo, int T, int LX,
                int LY, int LZ, const double *ka8, double factor, int dist) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, dist);
  tmb_deriv_fields f;
  f.l = (const double2 *)l; f.k = (const double2 *)k; f.U = (const double2 *)U; f.df = df;
  f.halo_k = (const double2 *)halo; f.halo_l = (const double2 *)halo + (size_t)6 * g.S;
  double2 ka[4];
  for (int m = 0; m < 4; m++) ka[m] = make_double2(ka8[2 * m], ka8[2 * m + 1]);
  for (int q = 0; q < 2; q++)
    for (int i = 0; i < g.Vh; i++) {
      const bool last = dist && i >= (g.T - 1) * g.S;
      if (q == ieo) { if (last) tmb_deriv_site<1, 1>(f, g, q, i, ka, 2. * factor); else tmb_deriv_site<1, 0>(f, g, q, i, ka, 2. * factor); }
      else          { if (last) tmb_deriv_site<0, 1>(f, g, q, i, ka, 2. * factor); else tmb_deriv_site<0, 0>(f, g, q, i, ka, 2. * factor); }
    }
}
/* Z split of the fermion force: the fix-up of the z links owned by the last-z sites (tmb_deriv_zfix); halo = [k dn-face | l dn-face]
 * of the slab above, each [6][Sz] as emul_pack_zfaces writes its `dn` output */
void emul_deriv_zfix(int ieo, const double *l, const double *k, const double *U, double *df, const double *halo_k, const double *halo_l,
                     int T, int LX, int LY, int LZ, const double *ka8, double factor) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  tmb_deriv_fields f;
  f.l = (const double2 *)l; f.k = (const double2 *)k; f.U = (const double2 *)U; f.df = df; f.halo_k = nullptr; f.halo_l = nullptr;
  const int Sz = T * LX * LY / 2;
  for (int q = 0; q < 2; q++)
    for (int j = 0; j < Sz; j++)
      tmb_deriv_zfix(f, g, ieo, q, j, (const double2 *)halo_k, (const double2 *)halo_l, make_double2(ka8[6], ka8[7]), 2. * factor);
}
/* two-flavour hopping term with the epilogues of hop2_kernel (tmb_force.cu), applied site by site on the host:
 * mode 0 plain, 1 M_ee_inv_ndpsi of the two results, 2 scale * g5( M_oo(p0, p1) - H ) */
void emul_hop2(int par, double *out0, double *out1, const double *in0, const double *in1, const double *p0, const double *p1,
               const double *U, int T, int LX, int LY, int LZ, const double *ka8, int mode, double mu, double eps, double scale) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 0);
  tmb_policies pol = {0, 0};
  double2 ka[4];
  for (int m = 0; m < 4; m++) ka[m] = make_double2(ka8[2 * m], ka8[2 * m + 1]);
  const size_t Vh = g.Vh;
  double2 *o0 = (double2 *)out0, *o1 = (double2 *)out1;
  const double2 *q0 = (const double2 *)p0, *q1 = (const double2 *)p1;
  for (int i = 0; i < g.Vh; i++) {
    double2 r0[12], r1[12];
    tmb_hop_site2<0>(r0, r1, (const double2 *)in0, (const double2 *)in1, (const double2 *)U, g, par, i, ka, pol);
    for (int c = 0; c < 12; c++) {
      if (mode == 1) {
        double2 ls, lc;
        tmb_nd_mee_inv_regs(ls, lc, r0[c], r1[c], c, mu, eps, 1. / (1. + mu * mu - eps * eps));
        r0[c] = ls; r1[c] = lc;
      } else if (mode == 2) { /* the arithmetic of hop2_kernel<2> */
        const bool up = c < 6;
        const double2 zs = make_double2(1., up ? -mu : mu), zc = c_conj(zs);
        const double2 a = q0[c * Vh + i], b = q1[c * Vh + i];
        double2 x = c_mul(zs, a); x.x += eps * b.x; x.y += eps * b.y;
        double2 y = c_mul(zc, b); y.x += eps * a.x; y.y += eps * a.y;
        const double2 d0 = up ? c_sub(x, r0[c]) : c_sub(r0[c], x), d1 = up ? c_sub(y, r1[c]) : c_sub(r1[c], y);
        r0[c] = make_double2(scale * d0.x, scale * d0.y); r1[c] = make_double2(scale * d1.x, scale * d1.y);
      }
      o0[c * Vh + i] = r0[c]; o1[c * Vh + i] = r1[c];
    }
  }
}
/* single-precision BLAS-1 functor of the product (tmb_kernels.cu: EwBlas32), n2 = 12 * sites complex numbers */
void emul_blas32(int op, float *r, const float *s1, const float *s2, float c1, float c2, long n2) {
  EwBlas32 f = {(float2 *)r, (const float2 *)s1, (const float2 *)s2, c1, c2, op, (size_t)n2 / 2};
  for (size_t k = 0; k < (size_t)n2; k++) f(k);
}
/* plaquette sum by the device site function; dist: the +t links of the last slice come from `up` = [2][3][9][S]
 * (emul_pack_gauge_first_slice of the rank above; of the same field for a periodic single rank) */
void emul_pack_gauge_first_slice(double *out, const double *U, int T, int LX, int LY, int LZ) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, 1);
  const double2 *u = (const double2 *)U; double2 *o = (double2 *)out;
  for (size_t x = 0; x < (size_t)54 * g.S; x++) { /* the index arithmetic of pack_gauge_first_slice_kernel */
    const int j = (int)(x % g.S); const int row = (int)(x / g.S);
    const int e = row % 9, m = (row / 9) % 3, q = row / 27;
    o[x] = u[(size_t)((q * 4 + (m + 1)) * 9 + e) * g.Vh + j];
  }
}
double emul_plaquette(const double *U, const double *up, int T, int LX, int LY, int LZ, int dist) {
  tmb_geom g = tmb_make_geom(T, LX, LY, LZ, dist);
  double s = 0.;
  for (int q = 0; q < 2; q++)
    for (int i = 0; i < g.Vh; i++)
      s += dist ? tmb_plaq_site<1>((const double2 *)U, (const double2 *)up, g, q, i)
                : tmb_plaq_site<0>((const double2 *)U, nullptr, g, q, i);
  return s / 3.0;
}
} /* extern "C" */

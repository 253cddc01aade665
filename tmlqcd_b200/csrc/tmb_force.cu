/* tmb_force.cu - fermion force kernel (sm_100a): deriv_Sb(ieo, l, k, hf, factor), deriv_Sb.c:402-649.
 *
 * The reference scatters into hf->derivative with omp atomics (su3adj.h:140-156); here every link
 * owner site gathers its one contribution per call (see tmb_site.cuh), so there are no atomics and
 * the result is deterministic.  One thread per (parity, site): 12 local + 4 x 12 remote spinor
 * components, 4 x 9 link elements, 4 x 8 derivative components read-modify-written, all loads
 * contiguous across the warp.  Bandwidth-bound like the hopping kernel (1.25 flop/B): algorithmic
 * bytes per link-owner site 192 (spinor, perfect reuse) + 4*144 (links) + 2*4*64 (df r/w) = 1280.
 */
#include "tmb_hop.cuh" /* block_sum, finish_last_block for the two-flavour kernel's fused norm */

template <int DIST>
__global__ void __launch_bounds__(128, 3) deriv_kernel(const tmb_deriv_launch a) {
  /* The two parities of the same 128-site range are neighbours in the grid (q is the fastest block index): the CTA of parity
   * ieo reads l locally and k at the +mu neighbours, its sibling k locally and l at the neighbours, so both fields come out of
   * L2 for one of the two and DRAM sees every spinor once.  With the parities as two halves of the grid (round 1) each field
   * crossed DRAM twice: 14 % traffic above the algorithmic 1280 B per link-owner site. */
  const int q = blockIdx.x & 1;                   /* parity of the link owner */
  const int w = (blockIdx.x >> 1) * 128 + threadIdx.x;
  if (w >= a.nt * a.g.S) return;
  const int i = a.t0 * a.g.S + w;
  tmb_deriv_fields f;
  f.l = (const double2 *)a.l; f.k = (const double2 *)a.k; f.U = (const double2 *)a.U; f.df = a.df;
  f.halo_k = (const double2 *)a.halo_k; f.halo_l = (const double2 *)a.halo_l;
  if (q == a.ieo) tmb_deriv_site<1, DIST>(f, a.g, q, i, a.ka, a.c); /* block-uniform branch */
  else            tmb_deriv_site<0, DIST>(f, a.g, q, i, a.ka, a.c);
}

/* Z split: fix-up of the z links owned by the last-z sites of both parities (tmb_deriv_zfix, tmb_site.cuh); one thread per
 * (parity, face site).  halo_k / halo_l: the `dn` faces of the z-face pack of k and of l, from rank z+1. */
__global__ void __launch_bounds__(64) deriv_zfix_kernel(const tmb_deriv_launch a, const double2 *halo_k, const double2 *halo_l) {
  const int Sz = a.g.T * a.g.LX * a.g.LY / 2;
  const int w = blockIdx.x * 64 + threadIdx.x;
  if (w >= 2 * Sz) return;
  const int q = w >= Sz ? 1 : 0, j = w - q * Sz;
  tmb_deriv_fields f;
  f.l = (const double2 *)a.l; f.k = (const double2 *)a.k; f.U = (const double2 *)a.U; f.df = a.df;
  f.halo_k = nullptr; f.halo_l = nullptr;
  tmb_deriv_zfix(f, a.g, a.ieo, q, j, halo_k, halo_l, a.ka[3], a.c);
}
cudaError_t tmb_launch_deriv_zfix(const tmb_deriv_launch &a, const double2 *halo_k, const double2 *halo_l, cudaStream_t s) {
  const int n = 2 * (a.g.T * a.g.LX * a.g.LY / 2);
  deriv_zfix_kernel<<<(n + 63) / 64, 64, 0, s>>>(a, halo_k, halo_l);
  return cudaGetLastError();
}

cudaError_t tmb_launch_deriv(const tmb_deriv_launch &a, cudaStream_t s) {
  const int n = a.nt * a.g.S;
  if (n <= 0) return cudaSuccess;
  const int grid = 2 * ((n + 127) / 128);
  if (a.dist) deriv_kernel<1><<<grid, 128, 0, s>>>(a);
  else deriv_kernel<0><<<grid, 128, 0, s>>>(a);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) pack_deriv_halo_kernel(double2 *out, const double2 *k, const double2 *l, tmb_geom g) {
  const size_t n = (size_t)12 * g.S;
  for (size_t x = (size_t)blockIdx.x * 256 + threadIdx.x; x < n; x += (size_t)gridDim.x * 256) {
    const int which = (int)(x / ((size_t)6 * g.S)); const size_t y = x - (size_t)which * 6 * g.S;
    const int c = (int)(y / g.S), j = (int)(y - (size_t)c * g.S);
    const double2 *f = which ? l : k;
    out[x] = c_add(f[(size_t)c * g.Vh + j], f[(size_t)(c + 6) * g.Vh + j]);
  }
}
cudaError_t tmb_launch_pack_deriv_halo(double2 *out, const double2 *k, const double2 *l, tmb_geom g, cudaStream_t s) {
  const size_t n = (size_t)12 * g.S;
  pack_deriv_halo_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(out, k, l, g);
  return cudaGetLastError();
}

/* host [ix][mu][8] <-> device [q][mu][8][Vh] */
__global__ void __launch_bounds__(256) pack_deriv_kernel(double *dev, const double *lex, tmb_geom g) {
  const size_t n = (size_t)64 * g.Vh;
  for (size_t x = (size_t)blockIdx.x * 256 + threadIdx.x; x < n; x += (size_t)gridDim.x * 256) {
    const int i = (int)(x % g.Vh); const int row = (int)(x / g.Vh);
    const int a = row & 7, mu = (row >> 3) & 3, q = row >> 5;
    const int ix = tmb_eo_to_lexic(g, q, i);
    dev[x] = lex[((size_t)ix * 4 + mu) * 8 + a];
  }
}
__global__ void __launch_bounds__(256) unpack_deriv_kernel(double *lex, const double *dev, tmb_geom g, int add) {
  const size_t n = (size_t)64 * g.Vh;
  for (size_t x = (size_t)blockIdx.x * 256 + threadIdx.x; x < n; x += (size_t)gridDim.x * 256) {
    const int i = (int)(x % g.Vh); const int row = (int)(x / g.Vh);
    const int a = row & 7, mu = (row >> 3) & 3, q = row >> 5;
    const int ix = tmb_eo_to_lexic(g, q, i);
    double *o = lex + ((size_t)ix * 4 + mu) * 8 + a;
    *o = add ? *o + dev[x] : dev[x];
  }
}
static int lin_grid(size_t n) { size_t need = (n + 255) / 256, cap = 148 * 16; return (int)(need < cap ? (need ? need : 1) : cap); }
cudaError_t tmb_launch_pack_deriv(double *dev, const double *lex, tmb_geom g, cudaStream_t s) {
  pack_deriv_kernel<<<lin_grid((size_t)64 * g.Vh), 256, 0, s>>>(dev, lex, g);
  return cudaGetLastError();
}
cudaError_t tmb_launch_unpack_deriv(double *lex, const double *dev, tmb_geom g, int add, cudaStream_t s) {
  unpack_deriv_kernel<<<lin_grid((size_t)64 * g.Vh), 256, 0, s>>>(lex, dev, g, add);
  return cudaGetLastError();
}

/* ------------------------------------------------------------------ plaquette (measure_gauge_action.c:46-106)
 * grid.y = parity; one partial sum per CTA, summed in index order by the caller's final kernel (deterministic). */
template <int DIST>
__global__ void __launch_bounds__(128) plaq_kernel(const double2 *U, const double2 *Uup, tmb_geom g, double *partial) {
  const int q = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
  double v = (i < g.Vh) ? tmb_plaq_site<DIST>(U, Uup, g, q, i) : 0.;
  __shared__ double sh[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = (sh[0] + sh[1]) + (sh[2] + sh[3]);
}
int tmb_plaq_grid(const tmb_geom &g) { return 2 * ((g.Vh + 127) / 128); }
cudaError_t tmb_launch_plaquette(const double2 *U, const double2 *Uup, tmb_geom g, int dist, double *partial, cudaStream_t s) {
  dim3 grid((g.Vh + 127) / 128, 2);
  if (dist) plaq_kernel<1><<<grid, 128, 0, s>>>(U, Uup, g, partial);
  else plaq_kernel<0><<<grid, 128, 0, s>>>(U, Uup, g, partial);
  return cudaGetLastError();
}
__global__ void __launch_bounds__(256) pack_gauge_first_slice_kernel(double2 *out, const double2 *U, tmb_geom g) {
  const size_t n = (size_t)54 * g.S; /* [2][3][9][S] */
  for (size_t x = (size_t)blockIdx.x * 256 + threadIdx.x; x < n; x += (size_t)gridDim.x * 256) {
    const int j = (int)(x % g.S); const int row = (int)(x / g.S);
    const int e = row % 9, m = (row / 9) % 3, q = row / 27;
    out[x] = U[(size_t)((q * 4 + (m + 1)) * 9 + e) * g.Vh + j]; /* slice t = 0: eo index j */
  }
}
cudaError_t tmb_launch_pack_gauge_first_slice(double2 *out, const double2 *U, tmb_geom g, cudaStream_t s) {
  pack_gauge_first_slice_kernel<<<lin_grid((size_t)54 * g.S), 256, 0, s>>>(out, U, g);
  return cudaGetLastError();
}

/* ------------------------------------------------------------------ two-flavour hopping kernel
 * (lives in this translation unit to keep tmb_kernels.cu's compile time down) */
/* One thread carries BOTH flavours of a site: every link is loaded once into registers and applied to the two projected
 * spinors.  V2 = double2: 254 registers, 2 CTAs of 128 per SM; V2 = float2 (Qtm_pm_ndpsi_32, the operator of
 * rg_mixed_cg_her_nd's inner loops): half the registers per value, 4 CTAs per SM.  DOT == 2: the partial sums of
 * dot_scale (|out0|^2 + |out1|^2) and, for the CTA that takes the last ticket, the reduction finish with the CG
 * bookkeeping (tmb_hop.cuh) - the <p, A p> of cg_her_nd out of the second launch of Qtm_pm_ndpsi. */
template <class V2, int MODE, int DOT, int HINTS, int MINB>
__global__ void __launch_bounds__(128, MINB) hop2_kernel(const tmb_hop2_launch a) {
  typedef typename tmb_real<V2>::type R;
  if (a.prefetch && threadIdx.x < 72) { /* L2 bulk prefetch of the 8 x 9 gauge rows of the CTA prefetch_dist CTAs ahead (linear traversal) */
    const int first = (blockIdx.x + a.prefetch_dist) * 128;
    int n = a.g.Vh - first; n = n > 128 ? 128 : n;
    if (n > 0) {
      const int d = threadIdx.x / 9, e = threadIdx.x - 9 * d, mu = d >> 1, bwd = d & 1;
      int j0 = first;
      if (bwd) {
        const int shift = mu == 0 ? a.g.S : (mu == 1 ? a.g.LY * a.g.Lzh : (mu == 2 ? a.g.Lzh : 0));
        j0 = first - shift; if (j0 < 0) j0 += a.g.Vh;
      }
      if (j0 > a.g.Vh - n) j0 = a.g.Vh - n;
      const V2 *src = (const V2 *)a.U + (size_t)(((bwd ? 1 - a.par : a.par) * 4 + mu) * 9 + e) * a.g.Vh + j0;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((int)(n * sizeof(V2))));
    }
  }
  if (a.st != nullptr && a.st->converged) return;
  const int w = blockIdx.x * 128 + threadIdx.x;
  const int i = a.tile ? tmb_tile_site(a.g, w, 0) : w; /* tiled: Vh is a multiple of 128 */
  double dsum = 0.;
  if (w < a.g.Vh) {
    tmb_policies pol;
    pol.stream = tmb_policy_evict_first();
    pol.reuse = tmb_policy_evict_last();
    V2 ka[4];
#pragma unroll
    for (int m = 0; m < 4; m++) ka[m] = cvt2<V2>(a.ka[m]);
    V2 r0[12], r1[12];
    tmb_hop_site2<HINTS>(r0, r1, (const V2 *)a.in0, (const V2 *)a.in1, (const V2 *)a.U, a.g, a.par, i, ka, pol);
    V2 *o0 = (V2 *)a.out0, *o1 = (V2 *)a.out1;
    const size_t Vh = a.g.Vh;
    const R mu = (R)a.mu, eps = (R)a.eps, scale = (R)a.scale;
    if (MODE == 1) {
      const double nrm = 1. / (1. + a.mu * a.mu - a.eps * a.eps);
#pragma unroll
      for (int c = 0; c < 12; c++) {
        V2 ls, lc;
        tmb_nd_mee_inv_regs(ls, lc, r0[c], r1[c], c, a.mu, a.eps, nrm);
        r0[c] = ls; r1[c] = lc;
      }
    } else if (MODE == 2) {
      const V2 *p0 = (const V2 *)a.p0, *p1 = (const V2 *)a.p1;
      V2 q0[12], q1[12]; /* all operand loads before the first store: out may alias p (see hop_kernel) */
#pragma unroll
      for (int c = 0; c < 12; c++) { q0[c] = p0[c * Vh + i]; q1[c] = p1[c * Vh + i]; }
#pragma unroll
      for (int c = 0; c < 12; c++) {
        const bool up = c < 6;
        const V2 zs = mk2<V2>((R)1, up ? -mu : mu), zc = c_conj(zs);
        V2 x = c_mul(zs, q0[c]); x.x += eps * q1[c].x; x.y += eps * q1[c].y;
        V2 y = c_mul(zc, q1[c]); y.x += eps * q0[c].x; y.y += eps * q0[c].y;
        const V2 d0 = up ? c_sub(x, r0[c]) : c_sub(r0[c], x), d1 = up ? c_sub(y, r1[c]) : c_sub(r1[c], y);
        r0[c] = mk2<V2>(scale * d0.x, scale * d0.y); r1[c] = mk2<V2>(scale * d1.x, scale * d1.y);
      }
    }
    if (DOT == 2) {
#pragma unroll
      for (int c = 0; c < 12; c++) {
        dsum += (double)r0[c].x * (double)r0[c].x; dsum += (double)r0[c].y * (double)r0[c].y;
        dsum += (double)r1[c].x * (double)r1[c].x; dsum += (double)r1[c].y * (double)r1[c].y;
      }
      dsum *= a.dot_scale;
    }
#pragma unroll
    for (int c = 0; c < 12; c++) {
      tmb_store_out<HINTS & 1>(o0 + c * Vh + i, r0[c], pol);
      tmb_store_out<HINTS & 1>(o1 + c * Vh + i, r1[c], pol);
    }
  }
  if (DOT) {
    const double sblk = block_sum<128>(dsum);
    if (threadIdx.x == 0) a.partial[blockIdx.x] = sblk;
    if (a.fin_op >= 0) finish_last_block<128>(a.partial, (int)gridDim.x, a.st_fin, a.fin_slot, a.fin_op, a.xr);
  }
}
/* Lane-paired variant (tmb_set_hop2_variant(1); measured SLOWER than hop2_kernel at 32^3x64: 450 vs 413 us per
 * launch, so it is not the default - kept selectable for re-tuning): lanes 2j and 2j+1 of a warp work on the SAME site, one flavour each.  Both lanes
 * issue the gauge loads with the same address, so the coalescer fetches every link once per pair (a warp instruction
 * covers 16 sites x 16 B = two full 128-B lines) - the gauge stream is still shared, 1920 B per site pair - but each
 * thread carries ONE 12-component accumulator: 168 registers and 12 warps/SM like hop_kernel instead of 254 and 8.
 * The 2x2 flavour mixing of the epilogues needs the partner's value: one __shfl_xor_sync per real number. */
__device__ __forceinline__ double2 tmb_partner(double2 v) {
  return make_double2(__shfl_xor_sync(0xffffffffu, v.x, 1), __shfl_xor_sync(0xffffffffu, v.y, 1));
}
template <int MODE, int HINTS>
__global__ void __launch_bounds__(128, 3) hop2p_kernel(const tmb_hop2_launch a) {
  const int gid = blockIdx.x * 128 + threadIdx.x;
  const int fl = gid & 1;
  const bool live = (gid >> 1) < a.g.Vh;
  const int i = live ? (gid >> 1) : a.g.Vh - 1; /* no early exit: the shuffles below want whole warps */
  tmb_policies pol;
  pol.stream = tmb_policy_evict_first();
  pol.reuse = tmb_policy_evict_last();
  tmb_hop_fields<double2> f;
  f.in = (const double2 *)(fl ? a.in1 : a.in0); f.U = (const double2 *)a.U;
  f.halo_up = nullptr; f.halo_dn = nullptr; f.Uhalo = nullptr;
  double2 r[12];
  tmb_hop_site<0, HINTS>(r, f, a.g, a.par, i, a.ka, pol);
  const size_t Vh = a.g.Vh;
  /* flavour 0 (strange) carries 1 -+ i mu, flavour 1 (charm) the conjugate (tm_operators_nd.c:639-756) */
  const double smu = fl ? -a.mu : a.mu;
  if (MODE == 1) {
    const double nrm = 1. / (1. + a.mu * a.mu - a.eps * a.eps);
#pragma unroll
    for (int c = 0; c < 12; c++) {
      const double2 o = tmb_partner(r[c]);
      const double2 z = make_double2(1., (c < 6) ? -smu : smu);
      double2 x = c_mul(z, r[c]); x.x += a.eps * o.x; x.y += a.eps * o.y;
      r[c] = make_double2(nrm * x.x, nrm * x.y);
    }
  } else if (MODE == 2) {
    const double2 *p = (const double2 *)(fl ? a.p1 : a.p0);
    double2 q[12]; /* all operand loads before the first store: out may alias p (see hop_kernel) */
#pragma unroll
    for (int c = 0; c < 12; c++) q[c] = p[c * Vh + i];
#pragma unroll
    for (int c = 0; c < 12; c++) {
      const bool up = c < 6;
      const double2 o = tmb_partner(q[c]);
      const double2 z = make_double2(1., up ? -smu : smu);
      double2 x = c_mul(z, q[c]); x.x += a.eps * o.x; x.y += a.eps * o.y;
      const double2 d = up ? c_sub(x, r[c]) : c_sub(r[c], x);
      r[c] = make_double2(a.scale * d.x, a.scale * d.y);
    }
  }
  if (!live) return;
  double2 *out = (double2 *)(fl ? a.out1 : a.out0);
#pragma unroll
  for (int c = 0; c < 12; c++) tmb_store_out<HINTS & 1>(out + c * Vh + i, r[c], pol);
}
template <class V2, int HINTS, int MINB>
static cudaError_t hop2_go1(const tmb_hop2_launch &a, cudaStream_t s) {
  const int grid = (a.g.Vh + 127) / 128;
  if (a.dot) {
    if (a.mode != 2 || a.dot != 2) return cudaErrorInvalidValue;
    hop2_kernel<V2, 2, 2, HINTS, MINB><<<grid, 128, 0, s>>>(a);
    return cudaGetLastError();
  }
  switch (a.mode) {
    case 0: hop2_kernel<V2, 0, 0, HINTS, MINB><<<grid, 128, 0, s>>>(a); break;
    case 1: hop2_kernel<V2, 1, 0, HINTS, MINB><<<grid, 128, 0, s>>>(a); break;
    case 2: hop2_kernel<V2, 2, 0, HINTS, MINB><<<grid, 128, 0, s>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
template <int HINTS>
static cudaError_t hop2_go(const tmb_hop2_launch &a, cudaStream_t s) {
  if (a.variant == 1) { /* lane-paired */
    if (a.prec || a.dot) return cudaErrorInvalidValue;
    const int gridp = (int)(((size_t)2 * a.g.Vh + 127) / 128);
    switch (a.mode) {
      case 0: hop2p_kernel<0, HINTS><<<gridp, 128, 0, s>>>(a); break;
      case 1: hop2p_kernel<1, HINTS><<<gridp, 128, 0, s>>>(a); break;
      case 2: hop2p_kernel<2, HINTS><<<gridp, 128, 0, s>>>(a); break;
      default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
  }
  if (a.variant == 3) return hop2_go1<double2, HINTS, 3>(a, s); /* 168 registers, ~1 KB of spills per thread */
  return hop2_go1<double2, HINTS, 2>(a, s);
}
int tmb_hop2_grid(const tmb_hop2_launch &a) { return (a.g.Vh + 127) / 128; }
cudaError_t tmb_launch_hop2(const tmb_hop2_launch &a, cudaStream_t s) {
  if (a.prec) return hop2_go1<float2, 1, 4>(a, s); /* single precision: cache-policy loads always on */
  return a.hints ? hop2_go<1>(a, s) : hop2_go<0>(a, s);
}

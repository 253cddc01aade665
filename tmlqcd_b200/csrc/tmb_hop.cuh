/* tmb_hop.cuh - the hopping kernel template (K1) with everything it pulls in: block reduction, the fused reduction
 * finish with its cross-rank sum, the peer-mode flags.  Included by tmb_kernels.cu (one field per launch) and by
 * tmb_hop2.cu (the two flavours of the non-degenerate doublet in one launch), so that the two sets of instantiations
 * compile side by side. */
#pragma once
#include "tmb_kernels.h"
#include "tmb_site.cuh"
#define TMB_SMS 148

/* ------------------------------------------------------------------ block reduction */
template <int BLOCK>
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  v = 0.;
  if (wid == 0) {
    v = (lane < BLOCK / 32) ? sh[lane] : 0.;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v; /* valid in thread 0 */
}

static __device__ void cg_apply(tmb_cg_state *st, int slot, int op) {
  const double sum = st->tmp[slot];
  if (op == TMB_FIN_CG_PRO) {
    st->pro = sum;
    st->alpha = st->fprec ? (double)((float)st->normsq / (float)sum) : st->normsq / sum;
  } else if (op == TMB_FIN_CG_ERR) {
    st->err = sum;
    st->iter += 1;
    const double thr = st->rel_prec ? st->eps_sq * st->sqnorm_q : st->eps_sq;
    if (sum <= thr) {
      st->converged = 1;
    } else {
      st->beta = sum / st->normsq;
      st->normsq = sum;
    }
  } else if (op == TMB_FIN_CG_INIT) {
    st->normsq = sum;
  } else if (op == TMB_FIN_MCG_ERR) {
    /* solver/mixed_cg_her.c:139-150: j counts the iterations that did NOT break */
    st->err = sum;
    const double thr = st->rel_prec ? st->eps_sq * st->sqnorm_q : st->eps_sq;
    if (sum <= st->inner_eps * st->sqnrm0 || st->iter == st->max_iter || 1.3 * sum <= thr) {
      st->converged = 1;
    } else {
      st->beta = sum / st->normsq;
      st->normsq = sum;
      st->iter += 1;
    }
  } else if (op == TMB_FIN_RG_ERR) {
    /* rg_mixed_cg_her.c:118-145 (float) / :75-104 (double): ++j; ...; rho = |r|^2; beta = rho / *rho1; *rho1 = rho;
     * if (1.3 rho < eps_sq) break; if (rho > rhomax) rhomax = rho; while (rho > delta*rhomax && j+iter <= max_iter) */
    const double rho = st->fprec ? (double)(float)sum : sum;
    st->err = rho;
    st->iter += 1;
    st->beta = st->fprec ? (double)((float)rho / (float)st->normsq) : rho / st->normsq;
    st->normsq = rho;
    const double eps = st->fprec ? (double)(float)st->eps_sq : st->eps_sq;
    if (1.3 * rho < eps) {
      st->converged = 1;
    } else {
      if (rho > st->sqnrm0) st->sqnrm0 = rho;
      const double lim = st->fprec ? (double)((float)st->inner_eps * (float)st->sqnrm0) : st->inner_eps * st->sqnrm0;
      if (!(rho > lim && st->iter <= st->max_iter)) st->converged = 1;
    }
  }
}

/* ------------------------------------------------------------------ cross-GPU flags (peer mode) */
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
/* spin until *p has reached seq (wrap-safe); gives up after ~4 s and raises *err instead of hanging the GPU */
__device__ __forceinline__ void wait_flag(const unsigned int *p, unsigned int seq, int *err) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(p) - seq) < 0) {
    if (clock64() - t0 > 8000000000LL) { *err = 1; break; }
    __nanosleep(100);
  }
}

/* Sum of one double over the ranks through peer memory (see tmb_xred_table); one thread per rank calls it.
 * Slot reuse is safe with a ring of 2 already: a rank can only complete reduction n+1 after every rank has contributed
 * to n+1, i.e. after every rank has finished reading n. */
static __device__ double xred_sum(const tmb_xred_table *t, double mine) {
  const int n = t->nranks, me = t->rank;
  const unsigned int seq = *t->ctr + 1u;
  *t->ctr = seq;
  const int k = (int)(seq & (TMB_XR_RING - 1)) * TMB_XR_MAXR;
  for (int q = 0; q < n; q++) *(volatile double *)(t->val[q] + k + me) = mine;
  __threadfence_system();
  for (int q = 0; q < n; q++) st_release_sys(t->seq[q] + k + me, seq);
  double s = 0.;
  for (int q = 0; q < n; q++) {
    wait_flag(t->seq[me] + k + q, seq, t->err);
    s += *(volatile double *)(t->val[me] + k + q);
  }
  return s;
}

/* Fused finish of a two-stage reduction: the CTA that takes the last ticket sums all block partials
 * in index order (same order whichever CTA is last -> deterministic) and does the CG bookkeeping,
 * which saves the separate one-CTA launch per reduction.  Used when no all-reduce sits in between. */
template <int BLOCK>
__device__ __forceinline__ void finish_last_block(const double *partial, int total, tmb_cg_state *st, int slot, int op,
                                                  const tmb_xred_table *xr) {
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    __threadfence(); /* this CTA's partial is visible before the ticket is taken */
    const unsigned t = atomicAdd(&st->ticket[slot], 1u);
    is_last = (t == (unsigned)(total - 1));
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double acc = 0.;
    for (int k = threadIdx.x; k < total; k += BLOCK) acc += __ldcg(partial + k);
    double s = block_sum<BLOCK>(acc);
    if (threadIdx.x == 0) {
      if (xr != nullptr) s = xred_sum(xr, s);
      st->tmp[slot] = s;
      if (op != TMB_FIN_STORE) cg_apply(st, slot, op);
      st->ticket[slot] = 0;
    }
  }
}

/* ------------------------------------------------------------------ K1: hopping
 * NFL = 1: one field per launch (Hopping_Matrix and its epilogues).
 * NFL = 2: the two flavours of the non-degenerate doublet in ONE launch (every Hopping_Matrix of tm_operators_nd.c is
 * applied to a strange and a charm field with the same links).  A CTA of BLOCK threads works on BLOCK/2 sites: its
 * first half of warps takes flavour 0, the second half flavour 1 of the SAME sites, so every warp-wide load still
 * moves 32 consecutive elements (512 B in double) and each thread carries one 12-component accumulator (the registers
 * and residency of the one-field kernel).  Both flavour groups issue the same gauge addresses at about the same time
 * (links allocate in L1 here, unlike in the one-field kernel): the second request is served by L1 / L2 and DRAM streams every
 * link once for both flavours (1920 B per site pair instead of 2 x 1536; ncu: 515 + 93 MB per MODE 1 launch at 24^3x48,
 * the same as the one-thread kernel).  Staging the eight links of a site in shared memory instead (cp.async, four directions
 * per flavour thread, one barrier) was measured SLOWER (32^3x64 Qtm_pm_ndpsi 1.92 ms against 1.82 ms: the barrier keeps the
 * spinor loads behind the link latency) and is not kept.  The 2x2 flavour mixing of M_ee_inv_ndpsi / M_oo_sub_g5_ndpsi (tm_operators_nd.c:639-756) is the
 * epilogue; the partner flavour's value comes through shared memory.
 *   NFL = 2, MODE 1: out_f = nrm [ (1 -+ i mu g5) H in_f + eps H in_f' ]                    (mu sign flips with f)
 *   NFL = 2, MODE 2: out_f = scale g5 [ (1 -+ i mu g5) p_f + eps p_f' - H in_f ]
 *   NFL = 2, DOT 2 : partial sums of dot_scale |out_0|^2 + |out_1|^2 */
template <class V2, int MODE, int DIST, int DOT, int HINTS, int BLOCK, int MINB, int NFL = 1>
__global__ void __launch_bounds__(BLOCK, MINB) hop_kernel(const tmb_hop_launch a) {
  constexpr int SITES = BLOCK / NFL; /* sites per CTA */
  static_assert(NFL == 1 || (NFL == 2 && SITES % 32 == 0), "flavour groups are whole warps");
  /* Programmatic dependent launch: let the next kernel of the stream start filling SMs while this
   * grid drains, and do everything that does not depend on the previous kernel before the wait -
   * optionally an L2 bulk prefetch of this CTA's gauge rows.  Both instructions are no-ops when
   * the launch carries no PDL attribute.  (Measured: neither helps this kernel, see DESIGN.md.) */
  asm volatile("griddepcontrol.launch_dependents;");
  const int NE = (HINTS & 2) ? 6 : 9; /* stored complex numbers per link (12-real compression: 6) */
  if ((a.prefetch & 1) && threadIdx.x < 8 * NE) {
    const int first = (blockIdx.x + a.prefetch_dist) * SITES;
    int n = a.nsites - first; n = n > SITES ? SITES : n;
    if (n > 0) {
      const int i0 = a.site0 + first + (first >= a.split ? a.gap : 0);
      const int d = threadIdx.x / NE, e = threadIdx.x - NE * d, mu = d >> 1, bwd = d & 1;
      int j0 = i0;
      if (bwd) {
        const int shift = mu == 0 ? a.g.S : (mu == 1 ? a.g.LY * a.g.Lzh : (mu == 2 ? a.g.Lzh : 0));
        j0 = i0 - shift; if (j0 < 0) j0 += a.g.Vh;
      }
      if (j0 > a.g.Vh - n) j0 = a.g.Vh - n;
      const V2 *src = (const V2 *)a.U + (size_t)(((bwd ? 1 - a.par : a.par) * 4 + mu) * NE + e) * a.g.Vh + j0;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((int)(n * sizeof(V2))));
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (a.st != nullptr && a.st->converged) return; /* CG already stopped: uniform early exit (same on every rank) */
  const int fl = NFL == 2 ? (int)threadIdx.x / SITES : 0;           /* flavour of this thread (warp-uniform) */
  const int lt = NFL == 2 ? (int)threadIdx.x - fl * SITES : (int)threadIdx.x;
  int w = blockIdx.x * SITES + lt;
  bool worker = true, bcta = false;
  if (DIST == 2) {
    const unsigned int seq = *a.seq_base + a.seq_off; /* the base only moves between launches (last node of a CG graph) */
    /* Peer mode: ONE launch does the whole hop and its halo exchange.
     * (1) Block 0 tells both neighbours that this rank's input field is complete (everything before this
     *     kernel in the stream has finished).
     * (2) The first p2p_copy_ctas CTAs wait for the neighbours' ready flags and PULL the two boundary
     *     time-slices of the neighbours' fields over NVLink, projecting them to half-spinors on the way
     *     (192 B read remotely, 96 B written locally per site), with 12 loads in flight per thread.
     * (3) Block 0 then acts as the closer: when all pulls have landed it publishes halo_ready = seq for the
     *     boundary CTAs, tells the neighbours that this rank no longer reads their memory (only the pull CTAs
     *     ever do) and waits for the same from them.  The kernel cannot complete before that, so whatever
     *     follows in the stream may overwrite the input field.  No election, no per-CTA atomics.
     * (4) All other CTAs do the stencil in the ROTATED slice order T/2, .., T-1, 0, .., T/2-1: consecutive slices
     *     stay adjacent in time (the +-t neighbour slices are L2 hits, exactly one pair is cut), the two
     *     boundary slices T-1 and 0 sit in the middle of the launch - half a hop after the pull started, and
     *     not in the tail - and only their CTAs wait for halo_ready. */
    const int Gc = a.p2p_copy_ctas;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      __threadfence_system();
      st_release_sys(a.up_flags + 1, seq);
      st_release_sys(a.dn_flags + 0, seq);
    }
    if ((int)blockIdx.x < Gc) {
      worker = false;
      if (threadIdx.x == 0) { wait_flag(a.flags + 0, seq, a.p2p_err); wait_flag(a.flags + 1, seq, a.p2p_err); }
      __syncthreads();
      /* the halo buffers hold NFL faces back to back: flavour f at offset f * 6 S */
      const size_t per = (size_t)12 * a.g.S, n = NFL * per, stride = (size_t)Gc * BLOCK;
      const size_t half = (size_t)6 * a.g.S;
      V2 *hu = (V2 *)a.halo_up_w, *hd = (V2 *)a.halo_dn_w;
      const unsigned long long keep = tmb_policy_evict_last();
      for (size_t k0 = (size_t)blockIdx.x * BLOCK + threadIdx.x; k0 < n && !(a.p2p_diag & 4); k0 += 6 * stride) {
        V2 x[6], y[6]; /* 12 remote loads in flight before the first store */
#pragma unroll
        for (int u = 0; u < 6; u++) {
          const size_t kq = k0 + u * stride;
          if (kq < n) {
            const int pf = (NFL == 2 && kq >= per) ? 1 : 0; const size_t k = kq - pf * per;
            const bool up = k < half; const size_t kk = up ? k : k - half;
            const int c = (int)(kk / a.g.S), j = (int)(kk - (size_t)c * a.g.S);
            const V2 *src = (const V2 *)(up ? (pf ? a.in_up1 : a.in_up) : (pf ? a.in_dn1 : a.in_dn));
            const size_t site = up ? (size_t)j : (size_t)(a.g.T - 1) * a.g.S + j;
            x[u] = src[(size_t)c * a.g.Vh + site]; y[u] = src[(size_t)(c + 6) * a.g.Vh + site];
          }
        }
#pragma unroll
        for (int u = 0; u < 6; u++) { /* halo buffers: keep them in L2 until the boundary CTAs come */
          const size_t kq = k0 + u * stride;
          if (kq < n) {
            const int pf = (NFL == 2 && kq >= per) ? 1 : 0; const size_t k = kq - pf * per;
            if (k < half) tmb_st_keep(hu + pf * half + k, c_add(x[u], y[u]), keep);
            else tmb_st_keep(hd + pf * half + (k - half), c_sub(x[u], y[u]), keep);
          }
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.p2p_copied, 1u);
        if (blockIdx.x == 0) { /* the closer */
          const long long t0 = clock64();
          while (*(volatile unsigned int *)a.p2p_copied < (unsigned int)Gc) {
            if (clock64() - t0 > 8000000000LL) { *a.p2p_err = 1; break; }
            __nanosleep(200);
          }
          *a.p2p_copied = 0;
          __threadfence();
          *(volatile unsigned int *)(a.p2p_copied + 1) = seq; /* halo_ready */
          __threadfence_system();
          st_release_sys(a.up_flags + 3, seq);
          st_release_sys(a.dn_flags + 2, seq);
          if (!a.p2p_nohandshake) { wait_flag(a.flags + 2, seq, a.p2p_err); wait_flag(a.flags + 3, seq, a.p2p_err); }
        }
      }
    } else {
      const int wb = (int)blockIdx.x - Gc;
      w = wb * SITES + lt;
      const int S = a.g.S, Vh = a.g.Vh;
      const int s0 = (a.g.T + 1) / 2;                               /* first slice of the rotated order */
      const int b0 = (a.g.T - 1 - s0) * S, b1 = b0 + 2 * S;        /* work range of slices T-1 and 0 */
      (void)Vh;
      bool touches;
      if (a.tile) { /* tiled traversal: the CTA holds the slice pair (tl, tl + 1 mod T) */
        const int tl = tmb_tile_t(a.g, wb * SITES, (a.p2p_diag & 16) ? 0 : ((a.g.T >> 1) | 1));
        touches = tl == 0 || tl >= a.g.T - 2;
      } else touches = wb * SITES + SITES > b0 && wb * SITES < b1;
      if (touches) { /* block-uniform: this CTA touches slice T-1 or slice 0 */
        bcta = !(a.p2p_diag & 8);
        if (threadIdx.x == 0) {
          const long long t0 = clock64();
          while ((int)(*(volatile unsigned int *)(a.p2p_copied + 1) - seq) < 0) {
            if (clock64() - t0 > 8000000000LL) { *a.p2p_err = 1; break; }
            __nanosleep(100);
          }
          __threadfence();
        }
        __syncthreads();
      }
    }
  }
  double dsum = 0.;
  const bool active = worker && w < a.nsites;
  int i = 0;
  tmb_policies pol;
  pol.stream = tmb_policy_evict_first();
  pol.reuse = tmb_policy_evict_last();
  V2 r[12];
  if (active) {
    int ww = w;
    if (a.tile) { /* 2 x 2 x 32 CTA tiles; peer mode: slices rotated by an odd shift, so that T-1 and 0 share a tile layer */
      ww = tmb_tile_site(a.g, w, (DIST == 2 && !(a.p2p_diag & 16)) ? ((a.g.T >> 1) | 1) : 0);
    } else if (DIST == 2 && !(a.p2p_diag & 16)) { /* rotated slice order s0, .., T-1, 0, .., s0-1 */
      ww = w + ((a.g.T + 1) / 2) * a.g.S;
      if (ww >= a.g.Vh) ww -= a.g.Vh;
    }
    if (a.xblock > 0) { /* x-blocked traversal of the (t,x) planes, memory layout unchanged */
      const int P = a.g.LY * a.g.Lzh, XB = a.xblock;
      const int plane = ww / P, off = ww - plane * P;
      const int per = a.g.T * XB;
      const int xb = plane / per, rem = plane - xb * per;
      const int t = rem / XB, xi = rem - t * XB;
      ww = (t * a.g.LX + xb * XB + xi) * P + off;
    }
    i = a.site0 + ww + (ww >= a.split ? a.gap : 0);
    tmb_hop_fields<V2> f;
    f.in = (const V2 *)(fl ? a.in1 : a.in); f.U = (const V2 *)a.U;
    f.halo_up = (const V2 *)a.halo_up + (size_t)fl * 6 * a.g.S; f.halo_dn = (const V2 *)a.halo_dn + (size_t)fl * 6 * a.g.S;
    f.Uhalo = (const V2 *)a.Uhalo;
    V2 ka[4];
#pragma unroll
    for (int m = 0; m < 4; m++) ka[m] = cvt2<V2>(a.ka[m]);
    /* optional (tmb_set_overlap bit 3): ask L2 for the epilogue operands of this site now, so that the batch of
     * loads after the 8 directions finds them there instead of paying a DRAM round trip with all registers live */
    if (NFL == 1 && (MODE >= 2 || DOT) && (a.prefetch & 2) && (threadIdx.x & 7) == 0) {
#pragma unroll
      for (int c = 0; c < 12; c++) {
        if (MODE >= 2) asm volatile("prefetch.global.L2 [%0];" ::"l"((const V2 *)a.p + (size_t)c * a.g.Vh + i));
        if (DOT == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"((const V2 *)a.dotw + (size_t)c * a.g.Vh + i));
      }
    }
    /* peer mode: interior CTAs run the branch-free site code (all 8 directions' loads can be batched), only
     * the CTAs of the two boundary slices take the variant with the per-site halo branches */
    if (DIST == 1 || (DIST == 2 && bcta)) tmb_hop_site<1, HINTS>(r, f, a.g, a.par, i, ka, pol);
    else tmb_hop_site<0, HINTS>(r, f, a.g, a.par, i, ka, pol);
  }
  if constexpr (NFL == 1 && MODE == 4) {
    /* CG tail (the last hop of Qtm_pm_psi inside cg_her): A p = g5((cf | conj cf) w0 - H w1) stays in registers and goes
     * straight into  x += alpha p ; r -= alpha A p ; |r|^2  (cg_her.c:95-101).  alpha = normsq / <p, A p> is known by now
     * because <p, A p> = |Q- p|^2 came out of the SECOND hop.  Saves writing A p and reading it back (384 B per site) and
     * the separate sweep's launch: 2496 B per site for this launch + nothing, instead of 1728 + 1152. */
    if (active) {
      typedef typename tmb_real<V2>::type R;
      const V2 cf = cvt2<V2>(a.cf);
      const R al = (R)a.st_fin->alpha;
      const V2 *__restrict__ w0 = (const V2 *)a.p, *__restrict__ pv = (const V2 *)a.cg_p;
      V2 *__restrict__ xv = (V2 *)a.cg_x, *__restrict__ rv = (V2 *)a.cg_r;
#pragma unroll
      for (int h = 0; h < 2; h++) { /* two batches of six components: 24 operand loads in flight, then their stores */
        V2 w[6], pp[6], xx[6], rr[6];
#pragma unroll
        for (int u = 0; u < 6; u++) {
          const size_t k = (size_t)(6 * h + u) * a.g.Vh + i;
          w[u] = w0[k]; pp[u] = pv[k]; xx[u] = xv[k]; rr[u] = rv[k];
        }
#pragma unroll
        for (int u = 0; u < 6; u++) {
          const int c = 6 * h + u;
          const size_t k = (size_t)c * a.g.Vh + i;
          const V2 ap = tmb_epilogue<2>(c, r[c], w[u], cf);
          xx[u].x += al * pp[u].x; xx[u].y += al * pp[u].y;
          rr[u].x = rr[u].x - al * ap.x; rr[u].y = rr[u].y - al * ap.y;
          dsum += (double)rr[u].x * (double)rr[u].x; dsum += (double)rr[u].y * (double)rr[u].y;
          xv[k] = xx[u]; rv[k] = rr[u];
        }
      }
    }
  } else if constexpr (NFL == 1) {
    if (active) {
      const V2 cf = cvt2<V2>(a.cf);
      const V2 *pp = (const V2 *)a.p, *dw_ = (const V2 *)a.dotw;
      V2 *out = (V2 *)a.out;
      /* All epilogue operands are loaded as one batch BEFORE the first store: `out` may alias `p`
       * (Qtm_minus_psi(l, l), invert_eo.c:270), so the compiler must not move a load across a store,
       * and interleaving them serialises 12 DRAM round trips per thread (measured: 138 us instead of
       * 80 us per launch at 24^3x48, profiles/r01_cg_launches_before_epilogue_fix.csv). */
      /* DOT == 1: Re <dotw, out>.  DOT == 2: the squared norm of the OUTPUT (no operand; a compile-time choice - as a
       * run-time branch around the operand loads it cost the whole CG 8 %).  The CG uses it on the second hop of
       * Qtm_pm_psi: <p, Q+ Q- p> = |Q- p|^2 because Q+ is the adjoint of Q- (gamma5-hermiticity). */
      V2 pc[12], dw[12];
#pragma unroll
      for (int c = 0; c < 12; c++) {
        if (MODE >= 2) pc[c] = pp[(size_t)c * a.g.Vh + i];
        if (DOT == 1) dw[c] = dw_[(size_t)c * a.g.Vh + i];
      }
      V2 o[12];
#pragma unroll
      for (int c = 0; c < 12; c++) {
        o[c] = tmb_epilogue<MODE>(c, r[c], MODE >= 2 ? pc[c] : mk2<V2>(0, 0), cf);
        if (DOT == 1) {
          dsum += (double)dw[c].x * (double)o[c].x;
          dsum += (double)dw[c].y * (double)o[c].y;
        } else if (DOT == 2) {
          dsum += (double)o[c].x * (double)o[c].x;
          dsum += (double)o[c].y * (double)o[c].y;
        }
      }
#pragma unroll
      for (int c = 0; c < 12; c++) tmb_store_out<HINTS & 1>(out + (size_t)c * a.g.Vh + i, o[c], pol);
    }
  } else if (worker) { /* NFL == 2; `worker` is block-uniform (pull CTAs never come here), so the barrier below is safe */
    typedef typename tmb_real<V2>::type R;
    __shared__ V2 xch[12 * BLOCK]; /* what the partner flavour needs: H in_f (MODE 1) or the operand p_f (MODE 2) */
    const R mu = (R)(fl ? -a.nd_mu : a.nd_mu), eps = (R)a.nd_eps;
    V2 own[12];
    if (MODE == 2 && active) { /* operand loads as one batch before any store: out may alias p */
      const V2 *pp = (const V2 *)(fl ? a.p1 : a.p);
#pragma unroll
      for (int c = 0; c < 12; c++) own[c] = pp[(size_t)c * a.g.Vh + i];
    }
    if (MODE >= 1) {
      if (active) {
#pragma unroll
        for (int c = 0; c < 12; c++) xch[c * BLOCK + threadIdx.x] = MODE == 1 ? r[c] : own[c];
      }
      __syncthreads();
    }
    if (active) {
      V2 o[12];
      const R nrm = (R)(1. / (1. + a.nd_mu * a.nd_mu - a.nd_eps * a.nd_eps)), scale = (R)a.nd_scale;
#pragma unroll
      for (int c = 0; c < 12; c++) {
        if (MODE == 0) { o[c] = r[c]; continue; }
        const V2 other = xch[c * BLOCK + (threadIdx.x ^ SITES)];
        const V2 z = mk2<V2>((R)1, (c < 6) ? -mu : mu);
        V2 x = c_mul(z, MODE == 1 ? r[c] : own[c]);
        x.x += eps * other.x; x.y += eps * other.y;
        if (MODE == 1) o[c] = mk2<V2>(nrm * x.x, nrm * x.y);
        else {
          const V2 d = (c < 6) ? c_sub(x, r[c]) : c_sub(r[c], x);
          o[c] = mk2<V2>(scale * d.x, scale * d.y);
        }
        if (DOT == 2) {
          dsum += (double)o[c].x * (double)o[c].x;
          dsum += (double)o[c].y * (double)o[c].y;
        }
      }
      V2 *out = (V2 *)(fl ? a.out1 : a.out);
#pragma unroll
      for (int c = 0; c < 12; c++) tmb_store_out<HINTS & 1>(out + (size_t)c * a.g.Vh + i, o[c], pol);
    }
    if (DOT) dsum *= a.dot_scale;
  }
  if (DOT) {
    const double s = block_sum<BLOCK>(dsum);
    if (threadIdx.x == 0) a.partial[blockIdx.x] = s;
    if (a.fin_op >= 0) finish_last_block<BLOCK>(a.partial_base, a.fin_total, a.st_fin, a.fin_slot, a.fin_op, a.xr);
  }
}

/* production configuration: chosen from the sweep in profiles/  */
#ifndef TMB_HOP_BLOCK
#define TMB_HOP_BLOCK 128
#endif
#ifndef TMB_HOP_MINB
#define TMB_HOP_MINB 3
#endif
/* single precision: half the registers per value -> more CTAs per SM */
#define TMB_HOP_BLOCK_F 128
#define TMB_HOP_MINB_F 4

static inline int hop_variant_block(int variant) {
  static const int b[11] = {TMB_HOP_BLOCK, 64, 64, 128, 128, 128, 256, 256, 96, 192, 64};
  return (variant >= 0 && variant < 11) ? b[variant] : TMB_HOP_BLOCK;
}
template <class V2, int MODE, int DIST, int DOT, int HINTS, int BLOCK, int MINB, int NFL = 1>
static cudaError_t hop_go(const tmb_hop_launch &a, cudaStream_t s) {
  const int grid = (a.nsites + BLOCK / NFL - 1) / (BLOCK / NFL) + (DIST == 2 ? a.p2p_copy_ctas : 0);
  if (grid <= 0) return cudaSuccess;
  if (a.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(BLOCK); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, hop_kernel<V2, MODE, DIST, DOT, HINTS, BLOCK, MINB, NFL>, a);
  }
  hop_kernel<V2, MODE, DIST, DOT, HINTS, BLOCK, MINB, NFL><<<grid, BLOCK, 0, s>>>(a);
  return cudaGetLastError();
}


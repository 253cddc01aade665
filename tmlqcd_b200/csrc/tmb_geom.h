/* tmb_geom.h - even/odd lattice geometry of one GPU's slab, shared by host and device code.
 *
 * Rebuilt from the rules of the reference's geometry_eo.c (not from its tables):
 *   lexicographic index  ix = ((t*LX + x)*LY + y)*LZ + z          geometry_eo.c:290
 *   parity               even iff (t+x+y+z) % 2 == 0              geometry_eo.c:807-814
 *   eo-sub index         rank of ix among the sites of its own parity in increasing ix
 *                                                                 geometry_eo.c:869-884
 * With LZ even every (t,x,y) row holds LZ/2 sites of each parity in increasing z, hence
 *   i = ((t*LX + x)*LY + y)*Lzh + zh,   z = 2*zh + ((t+x+y+parity) & 1).
 * tests/test_geometry.py checks these closed forms against the reference's
 * g_eo2lexic / g_lexic2eosub / g_hi tables.
 */
#ifndef TMB_GEOM_H
#define TMB_GEOM_H

#if defined(__CUDACC__)
#define TMB_HD __host__ __device__ __forceinline__
#else
#define TMB_HD static inline
#endif

typedef struct {
  int T, LX, LY, LZ; /* local extents of this rank's slab */
  int Lzh;           /* LZ/2 */
  int S;             /* LX*LY*Lzh: sites of one parity in one time-slice (a T-face) */
  int Vh;            /* T*S: sites of one parity */
  int dist_t;        /* 1: T is split over ranks, +-t neighbours of the boundary slices come from halo buffers */
} tmb_geom;

TMB_HD tmb_geom tmb_make_geom(int T, int LX, int LY, int LZ, int dist_t) {
  tmb_geom g;
  g.T = T; g.LX = LX; g.LY = LY; g.LZ = LZ; g.Lzh = LZ / 2;
  g.S = LX * LY * g.Lzh; g.Vh = T * g.S; g.dist_t = dist_t;
  return g;
}

/* eo-sub index -> coordinates; `par` is the parity of the site (0 even, 1 odd) */
TMB_HD void tmb_eo_coords(const tmb_geom g, int par, int i, int *t, int *x, int *y, int *z) {
  unsigned u = (unsigned)i;
  unsigned zh = u % (unsigned)g.Lzh; u /= (unsigned)g.Lzh;
  unsigned yy = u % (unsigned)g.LY;  u /= (unsigned)g.LY;
  unsigned xx = u % (unsigned)g.LX;  unsigned tt = u / (unsigned)g.LX;
  *t = (int)tt; *x = (int)xx; *y = (int)yy;
  *z = (int)(2u * zh + ((tt + xx + yy + (unsigned)par) & 1u));
}

TMB_HD int tmb_eo_to_lexic(const tmb_geom g, int par, int i) {
  int t, x, y, z;
  tmb_eo_coords(g, par, i, &t, &x, &y, &z);
  return ((t * g.LX + x) * g.LY + y) * g.LZ + z;
}

/* CTA tile traversal of the hopping kernels (memory layout UNCHANGED): work index w -> eo-sub index such that the four
 * warps of a 128-thread CTA take the SAME run of 32 consecutive (y,z) sites at (t0,x0), (t0,x0+1), (t0+1,x0), (t0+1,x0+1)
 * instead of 128 consecutive sites of one (t,x) plane.  Every warp-wide load still covers 32 consecutive elements; what
 * changes is how many of the 8 neighbour spinors of a site are ALSO neighbours (or the sites themselves) of other threads
 * of the CTA, i.e. L1 hits instead of L2 requests.  Distinct input spinors a CTA touches, per output site:
 *   128 consecutive sites (LZ/2 = 12):  1 + 2/10.7 (y) + 2 (x) + 2 (t)       = 5.2   (measured: L1 hit rate 35 % of 8)
 *   2 x 2 x 32:                         1 + 2/2.7 (y) + 2/2 (x) + 2/2 (t)    = 3.75
 * MEASURED (B200, scripts/tile_ab.py, profiles/r02_tile_ab.jsonl, profiles/r02_tile_ncu_metrics.csv): the L1 hit rate goes
 * from 18.8 to 24.3 % and the L2 -> SM traffic from 720 to 667 MB per hop at 24^3x48, but the kernel gets 1-5 % SLOWER in
 * every configuration (double, float, 12-real links, peer mode, two flavours; burst and power-capped): the L2 -> SM path is
 * not what bounds it.  Off by default (tmb_set_tile); kept as a selectable variant.  tshift rotates the time-slices (peer
 * mode: the boundary slices T-1 and 0 form ONE pair in the middle of the launch when tshift is odd). */
TMB_HD int tmb_tile_ok(const tmb_geom g) {
  return ((g.LY * g.Lzh) % 32 == 0) && (g.T % 2 == 0) && (g.LX % 2 == 0);
}
TMB_HD int tmb_tile_t(const tmb_geom g, int w, int tshift) { /* time-slice of work index w */
  const int R = (g.LY * g.Lzh) >> 5, q = (w >> 5) & 3, c = (w >> 7) / R;
  int t = 2 * (c / (g.LX >> 1)) + (q >> 1) + tshift;
  return t >= g.T ? t - g.T : t;
}
TMB_HD int tmb_tile_site(const tmb_geom g, int w, int tshift) {
  const int P = g.LY * g.Lzh, R = P >> 5;
  const int lane = w & 31, q = (w >> 5) & 3, b = w >> 7;
  const int c = b / R, r = b - c * R;
  const int hx = g.LX >> 1, bt = c / hx, bx = c - bt * hx;
  int t = 2 * bt + (q >> 1) + tshift;
  if (t >= g.T) t -= g.T;
  return (t * g.LX + 2 * bx + (q & 1)) * P + (r << 5) + lane;
}

/* Neighbour eo-sub indices (in the field of the OPPOSITE parity) of site i of parity `par`,
 * order +t,-t,+x,-x,+y,-y,+z,-z as g_hi[16*icx + 2d+1] (geometry_eo.c:1470-1536).
 * Periodic wrap inside the slab; with dist_t the caller replaces nb[0] for t==T-1 and nb[1]
 * for t==0 by halo reads.  Returns t. */
TMB_HD int tmb_neighbours(const tmb_geom g, int par, int i, int nb[8]) {
  unsigned u = (unsigned)i;
  unsigned zh = u % (unsigned)g.Lzh; u /= (unsigned)g.Lzh;
  unsigned y = u % (unsigned)g.LY;   u /= (unsigned)g.LY;
  unsigned x = u % (unsigned)g.LX;   unsigned t = u / (unsigned)g.LX;
  const int r = (int)((t + x + y + (unsigned)par) & 1u); /* z = 2*zh + r */
  const int sx = g.LY * g.Lzh;
  nb[0] = ((int)t + 1 < g.T) ? i + g.S : i - (g.T - 1) * g.S;
  nb[1] = (t > 0) ? i - g.S : i + (g.T - 1) * g.S;
  nb[2] = ((int)x + 1 < g.LX) ? i + sx : i - (g.LX - 1) * sx;
  nb[3] = (x > 0) ? i - sx : i + (g.LX - 1) * sx;
  nb[4] = ((int)y + 1 < g.LY) ? i + g.Lzh : i - (g.LY - 1) * g.Lzh;
  nb[5] = (y > 0) ? i - g.Lzh : i + (g.LY - 1) * g.Lzh;
  nb[6] = r ? (((int)zh + 1 < g.Lzh) ? i + 1 : i - (g.Lzh - 1)) : i;
  nb[7] = r ? i : ((zh > 0) ? i - 1 : i + (g.Lzh - 1));
  return (int)t;
}

/* Chunk schedule of the pipelined host-pointer Hopping_Matrix (tmb_capi.cu, host_hop_enqueue): `nt` time-slices to move, chunks
 * never smaller than `small` slices (about 1 MB).  Few LARGE chunks (a sixth of the field): once both directions are busy a
 * copy gets slower the smaller it is (a 10.6 MB upload takes ~255 us next to a running download, a 2.6 MB one ~105 us).  The
 * first big chunk goes up as a quarter and three quarters, so that the first download starts early, and the tail halves,
 * so that little is left to come down when the upload ends (24^3x48: 2,6,8,8,8,8,4,2,1,1 slices, 1.66 ms per call against
 * 1.70 for 8,8,8,8,8,4,2,2 and 1.80 for 12 equal chunks; profiles/r02_pipe_diag.log).  Returns the number of chunks
 * (<= 8 + log2(nt)); sizes[] gets the sizes in upload order of the slices. */
TMB_HD int tmb_host_chunk_schedule(int nt, int small, int *sizes) {
  int n = 0, rem = nt;
  if (small < 1) small = 1;
  int big = (nt + 5) / 6;
  if (big < small) big = small;
  if (big >= 4 * small && 4 * rem > 7 * big) {
    const int q = big / 4;
    sizes[n++] = q; sizes[n++] = big - q; rem -= big;
  }
  while (4 * rem > 7 * big) { sizes[n++] = big; rem -= big; }
  while (rem > 0) {
    const int sz = rem <= small ? rem : ((rem + 1) / 2 > small ? (rem + 1) / 2 : small);
    sizes[n++] = sz; rem -= sz;
  }
  return n;
}

#endif /* TMB_GEOM_H */

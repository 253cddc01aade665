/* tmb_hop2.cu - the two-flavour instantiations of the hopping kernel (tmb_hop.cuh, NFL = 2): every Hopping_Matrix pair of
 * operator/tm_operators_nd.c (Qtm_ndpsi :68, Qtm_dagger_ndpsi :130, Qtm_pm_ndpsi :195) as ONE launch with the flavour
 * mixing of M_ee_inv_ndpsi (:639) / M_oo_sub_g5_ndpsi (:698) and the phmc_invmaxev scaling in the epilogue, in double and
 * - for rg_mixed_cg_her_nd's inner loops (operator/tm_operators_nd_32.c) - in single precision, on one rank, with halo
 * buffers (NCCL) and in peer mode.  Its own translation unit so that it compiles beside tmb_kernels.cu. */
#include "tmb_hop.cuh"

template <class V2, int DIST, int CFG, int BLOCK, int MINB>
static cudaError_t nd_mode(const tmb_hop_launch &a, cudaStream_t s) {
  if (a.dot) {
    if (a.mode != 2 || a.dot != 2) return cudaErrorInvalidValue;
    return hop_go<V2, 2, DIST, 2, CFG, BLOCK, MINB, 2>(a, s);
  }
  switch (a.mode) {
    case 1: return hop_go<V2, 1, DIST, 0, CFG, BLOCK, MINB, 2>(a, s);
    case 2: return hop_go<V2, 2, DIST, 0, CFG, BLOCK, MINB, 2>(a, s);
  }
  return cudaErrorInvalidValue;
}
template <class V2, int CFG, int BLOCK, int MINB>
static cudaError_t nd_dist(const tmb_hop_launch &a, cudaStream_t s) {
  if (a.dist == 2) return nd_mode<V2, 2, CFG, BLOCK, MINB>(a, s);
  return a.dist ? nd_mode<V2, 1, CFG, BLOCK, MINB>(a, s) : nd_mode<V2, 0, CFG, BLOCK, MINB>(a, s);
}
cudaError_t tmb_launch_hop_nd(const tmb_hop_launch &a, cudaStream_t s) {
  if (a.prec) { /* single precision: cache-policy loads always on */
    if (a.recon12) return nd_dist<float2, 3, TMB_HOP_BLOCK_F, TMB_HOP_MINB_F>(a, s);
    return nd_dist<float2, 1, TMB_HOP_BLOCK_F, TMB_HOP_MINB_F>(a, s);
  }
  if (a.recon12) return nd_dist<double2, 3, TMB_HOP_BLOCK, TMB_HOP_MINB>(a, s);
  if (a.hints == 5) return nd_dist<double2, 5, TMB_HOP_BLOCK, TMB_HOP_MINB>(a, s); /* policies + gauge links through L1 */
  return a.hints ? nd_dist<double2, 1, TMB_HOP_BLOCK, TMB_HOP_MINB>(a, s) : nd_dist<double2, 0, TMB_HOP_BLOCK, TMB_HOP_MINB>(a, s);
}

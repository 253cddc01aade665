/* tmb_capi.cu - the C-ABI layer (include/tmlqcd_b200.h) over the kernels in tmb_kernels.cu.
 *
 * Holds the one-GPU context: streams, the device gauge field, halo buffers, the NCCL
 * communicator for the T-split, the scratch fields that play the role of the reference's
 * g_spinor_field[DUM_MATRIX..] and the device-resident CG.  No CPU compute path exists here:
 * every operator launches CUDA kernels and fails loudly when CUDA is unavailable.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <chrono>
#include <string>
#include <vector>
#include "../../include/tmlqcd_b200.h"
#include "tmb_kernels.h"

/* ------------------------------------------------------------------ NCCL, bound at run time
 * (the process usually already has torch's libnccl.so.2 loaded; dlopen gives us that copy) */
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_SUM = 0, NCCL_MIN = 3, NCCL_FLOAT64 = 8, NCCL_UINT8 = 1 };
struct NcclApi {
  void *handle;
  int (*GetUniqueId)(ncclUniqueId *);
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  int (*CommDestroy)(ncclComm_t);
  int (*GroupStart)(void);
  int (*GroupEnd)(void);
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(int);
};

#define NSCRATCH 14
#define MAXCHUNK 64
#define TMB_MAXCSG 20 /* chrono_guess.c:91: max_N = 20 */
#define TMB_MAXMNL 30 /* monomial.h:51: max_no_monomials */

/* one DET / DETRATIO monomial (the fields of the reference's `monomial` struct this path uses, monomial.h:53-131) */
struct Monomial {
  int type = 0, solver = 1, maxiter = 5000, csg_N = 0, csg_n = 0, iter0 = 0, iter1 = 0;
  double kappa = 0., mu = 0., kappa2 = 0., mu2 = 0., forceprec = 1e-7, accprec = 1e-16, forcefactor = 1.;
  double energy0 = 0., energy1 = 0.;
  double2 *pf = nullptr;
  double2 *csg[TMB_MAXCSG] = {nullptr};
  int idx[TMB_MAXCSG] = {0};
};

struct Ctx {
  bool init = false;
  int device = 0;
  tmb_geom g;
  int nranks = 1, rank = 0;
  bool dist = false, loopback = false;
  /* rank grid (nt x nz), rank = ct * nz + cz: T is split over nt, Z over nz (the reference's PARALLELT / the z part of
   * PARALLELXYZT, mpi_init.c:321-357).  `dist` means "T is split (or looped back)"; zsplit "Z is split (or looped back)" */
  int nt = 1, nz = 1, ct = 0, cz = 0; bool zsplit = false, loop_z = false;
  double2 *zsend_up = nullptr, *zsend_dn = nullptr, *zhalo_up = nullptr, *zhalo_dn = nullptr, *Uzh = nullptr, *Uzl = nullptr; float2 *Uzh32 = nullptr, *Uzl32 = nullptr;
  cudaEvent_t ev_z = nullptr;
  /* Z faces through peer memory (push): this rank's double-buffered halo buffers, where its own faces go on the z
   * neighbours, this rank's two flags and the neighbours' (ARENA_ZFLAGS) */
  bool zpeer = false; char *zloop_buf = nullptr;
  double2 *zd_send = nullptr, *zd_recv = nullptr; /* fermion force on a split Z: first-z faces of k and l (send: + scratch) */
  void *zloc_up[2] = {nullptr, nullptr}, *zloc_dn[2] = {nullptr, nullptr}, *zdst_up[2] = {nullptr, nullptr}, *zdst_dn[2] = {nullptr, nullptr};
  unsigned int *zflags = nullptr, *zflag_at_up = nullptr, *zflag_at_dn = nullptr;
  cudaStream_t s_main = nullptr, s_comm = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_up[MAXCHUNK] = {nullptr}, ev_done[MAXCHUNK] = {nullptr};
  cudaEvent_t ev_side[2] = {nullptr, nullptr}; /* fork / join of the CG's <p,Ap> finish on the side stream */
  int host_sched[MAXCHUNK] = {0}, host_sched_n = 0; /* explicit chunk sizes of the host-pointer pipeline (tmb_set_host_chunk_sizes) */
  cudaEvent_t ev_in = nullptr, ev_halo = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_chk[2] = {nullptr, nullptr};
  double2 *U = nullptr, *Uhalo = nullptr;
  double2 *send_up = nullptr, *send_dn = nullptr, *halo_up = nullptr, *halo_dn = nullptr;
  double2 *gauge_raw = nullptr; int gauge_uploads = 0; /* AoS staging of tmb_gauge_upload, kept once the links change repeatedly */
  double2 *stage = nullptr; /* AoS staging, 32*Vh double2 (one lexicographic spinor field or one derivative field) */
  double *partial = nullptr; int npartial = 0;
  tmb_cg_state *st = nullptr;      /* device */
  tmb_cg_state *st_host = nullptr; /* pinned, 2 slots */
  double2 ka[4];
  double kappa = 0., mu = 0., mubar = 0., epsbar = 0., invmaxev = 1.;
  double2 *scratch[NSCRATCH] = {nullptr};
  float2 *scratch32[NSCRATCH] = {nullptr};
  float2 *U32 = nullptr, *Uhalo32 = nullptr; bool gauge32_valid = false;
  int compression = 18; /* 18: full links; 12: two rows stored, third reconstructed (CompressionType, misc_types.h:33) */
  double2 *U12 = nullptr, *Uhalo12 = nullptr; float2 *U12f = nullptr, *Uhalo12f = nullptr; bool c12_valid = false, c12f_valid = false;
  double mixcg_innereps = 5.0e-5; int mixcg_maxinnersolverit = 5000; /* default_input_values.h:193-194 */
  int hop_variant = -1, hints = -1, xblock = 0, tile = 0, pdl = 0, prefetch = 0, prefetch_dist = 0, cg_graph = 1, hop2_variant = -1, cg_selfnorm = 1, cg_tail = 1, cg_side = 0;
  NcclApi nccl = {};
  ncclComm_t comm = nullptr;
  std::vector<void *> fields; std::vector<size_t> field_bytes;
  long long launches = 0;
  int last_iters = 0; double last_err = 0., last_seconds = 0.;
  int last_inner_sp = 0, last_inner_dp = 0, last_outer = 0;
  bool gauge_loaded = false;
  /* peer mode: symmetric arena for spinor fields + flags, neighbours' arenas mapped through CUDA IPC */
  bool p2p = false; int loopback_mode = 0;
  char *arena = nullptr; size_t arena_bytes = 0, arena_used = 0;
  std::vector<std::pair<size_t, size_t>> arena_free; /* (offset, bytes) */
  char *up_base = nullptr, *dn_base = nullptr;
  unsigned int *flags = nullptr, *up_flags = nullptr, *dn_flags = nullptr, *p2p_ticket = nullptr;
  int host_chunks = 0; /* equal chunks of the pipelined host-pointer hop; 0: the automatic schedule (tmb_host_chunk_schedule) */
  unsigned long long param_gen = 1; /* bumped by every setter whose value is baked into a cached host-hop graph */
  int *p2p_err = nullptr; bool arena_warned = false; int p2p_diag = 0, p2p_copy_ctas = 64;
  /* sequence numbers of the peer-mode hops are *seq_dev + hop_off: the host counts offsets, the device base only moves
   * at the end of a replayed CG graph (whose kernels carry fixed offsets); seq_dev[1] is the reduction counter of xred_sum */
  unsigned int *seq_dev = nullptr; unsigned int hop_off = 0;
  unsigned int zhop_off = 0; /* the same for the z-face pushes (base seq_dev[2]): their own count, its parity picks the halo buffer */
  char *peer_base[TMB_XR_MAXR] = {nullptr};   /* every rank's arena ([rank] = own) when nranks <= TMB_XR_MAXR */
  tmb_xred_table *xr_tab = nullptr; bool xred = false; /* cross-rank sums inside the reduction finish (no NCCL in the CG) */
  /* HMC side (tmb_capi_hmc.inc) */
  double2 phase[4] = {{1., 0.}, {1., 0.}, {1., 0.}, {1., 0.}}; /* exp(i theta_mu pi / L_mu): ka_mu / kappa */
  double *df = nullptr;                       /* hf->derivative on the device, [2][4][8][Vh] */
  double2 *dhalo_send = nullptr, *dhalo_recv = nullptr;
  Monomial mnl[TMB_MAXMNL]; int nmnl = 0;
  double2 *w[6] = {nullptr};                  /* w_fields (monomial.c:57) */
  double2 *nd[8] = {nullptr};                 /* two-flavour CG vectors of tmb_cg_her_nd (0..4), temporaries of tmb_invert_doublet_eo (5, 6), xhigh of tmb_rg_mixed_cg_her_nd (7), [2][12][Vh] each */
  float2 *nd32[4] = {nullptr};                /* float two-flavour vectors of tmb_rg_mixed_cg_her_nd's inner loops */
  int rel_prec_flag = 0;                      /* g_relative_precision_flag */
  double mcg_delta = 5.0e-5;                  /* solver_params.mcg_delta = _default_mixcg_innereps (monomial.c:106) */
};
static Ctx C;
static void hgraphs_clear(); static void pinned_clear();
static int ensure_gauge12(int prec);
extern "C" int tmb_rg_mixed_cg_her(void *P, const void *Q, int max_iter, double eps_sq, int rel_prec);
static std::string g_err;

static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  g_err = buf;
  return code;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(-100, "%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)
#define KL(x) do { cudaError_t e_ = (x); C.launches++; if (e_ != cudaSuccess) return fail(-101, "%s:%d kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)
#define NC(x) do { int e_ = (x); if (e_ != 0) return fail(-102, "%s:%d NCCL error: %s", __FILE__, __LINE__, C.nccl.GetErrorString ? C.nccl.GetErrorString(e_) : "?"); } while (0)
#define NEED_INIT() do { if (!C.init) return fail(-1, "tmb_init has not been called"); } while (0)
#define TRY(x) do { int r_ = (x); if (r_ < 0) return r_; } while (0)

static inline size_t N2() { return (size_t)12 * C.g.Vh; }
static inline size_t HALF() { return (size_t)6 * C.g.Vh; }
static inline size_t FIELD_BYTES() { return N2() * sizeof(double2); }
static inline double2 *F(void *p) { return (double2 *)p; }
static inline const double2 *F(const void *p) { return (const double2 *)p; }

/* ------------------------------------------------------------------ symmetric field memory
 * With more than one rank every spinor field comes out of ONE arena per rank, allocated in the same order
 * on all ranks (SPMD), so that "the same field on the neighbouring rank" is the neighbour's arena base plus
 * this rank's offset - what the peer-mode hopping kernel dereferences over NVLink. */
#define ARENA_RESERVED 4096 /* start of the arena: hop flags at 0, landing arrays of the cross-rank sums at 1024 (values) and 2048 (sequence words) */
#define ARENA_ZFLAGS 512 /* two words: [0] this rank's halo_dn has been filled (by rank z-1), [1] its halo_up (by rank z+1) */
#define ARENA_XR_VAL 1024
#define ARENA_XR_SEQ 2048
static inline bool in_arena(const void *p) {
  return C.arena && (const char *)p >= C.arena && (const char *)p < C.arena + C.arena_bytes;
}
static void *sym_malloc(size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (C.arena) {
    for (size_t i = 0; i < C.arena_free.size(); i++)
      if (C.arena_free[i].second == bytes) { void *p = C.arena + C.arena_free[i].first; C.arena_free.erase(C.arena_free.begin() + i); return p; }
    if (C.arena_used + bytes <= C.arena_bytes) { void *p = C.arena + C.arena_used; C.arena_used += bytes; return p; }
    if (!C.arena_warned) {
      fprintf(stderr, "tmlqcd_b200: field arena of %zu MB exhausted (set TMB_ARENA_MB); further fields use NCCL halos\n", C.arena_bytes >> 20);
      C.arena_warned = true;
    }
  }
  void *p = nullptr;
  return cudaMalloc(&p, bytes) == cudaSuccess ? p : nullptr;
}
static void sym_free(void *p) {
  if (!p) return;
  if (in_arena(p)) return; /* arena blocks are recycled by size through arena_release() or die with the arena */
  cudaFree(p);
}
static void arena_release(void *p, size_t bytes) {
  if (in_arena(p)) C.arena_free.push_back(std::make_pair((size_t)((char *)p - C.arena), (bytes + 255) & ~(size_t)255));
  else cudaFree(p);
}

extern "C" const char *tmb_last_error(void) { return g_err.c_str(); }
extern "C" int tmb_is_initialized(void) { return C.init ? 1 : 0; }
extern "C" int tmb_volume_half(void) { return C.init ? C.g.Vh : 0; }
extern "C" long long tmb_launch_count(void) { return C.launches; }
extern "C" int tmb_comm_nranks(void) { return C.nranks; }
extern "C" int tmb_comm_grid(int *nt, int *nz) { if (nt) *nt = C.nt; if (nz) *nz = C.nz; return 0; }
static inline int t_up() { return ((C.ct + 1) % C.nt) * C.nz + C.cz; }
static inline int t_dn() { return ((C.ct + C.nt - 1) % C.nt) * C.nz + C.cz; }
static inline int z_up() { return C.ct * C.nz + (C.cz + 1) % C.nz; }
static inline int z_dn() { return C.ct * C.nz + (C.cz + C.nz - 1) % C.nz; }
static inline size_t SZ() { return (size_t)C.g.T * C.g.LX * C.g.LY / 2; } /* sites of one parity on a z face */

extern "C" int tmb_init(int T, int LX, int LY, int LZ, int device) {
  if (C.init) {
    if (C.g.T == T && C.g.LX == LX && C.g.LY == LY && C.g.LZ == LZ) return 0;
    return fail(-2, "tmb_init: already initialised with %dx%dx%dx%d; call tmb_finalize first", C.g.T, C.g.LX, C.g.LY, C.g.LZ);
  }
  /* even/odd preconditioning on a periodic lattice needs every extent even (a wrap-around
   * neighbour of an odd extent would have the SAME parity) */
  if (T < 2 || LX < 2 || LY < 2 || LZ < 2 || ((T | LX | LY | LZ) & 1))
    return fail(-3, "tmb_init: all local extents must be even and >= 2 (got T=%d LX=%d LY=%d LZ=%d)", T, LX, LY, LZ);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(-4, "tmb_init: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
  CU(cudaSetDevice(device));
  C.device = device;
  C.g = tmb_make_geom(T, LX, LY, LZ, 0);
  /* the comm stream (halo pack + NCCL send/recv) gets the highest priority: its few CTAs must be
   * scheduled ahead of the thousands of pending CTAs of the interior kernel, otherwise the exchange
   * only starts when the interior kernel has drained and nothing overlaps */
  int prio_lo = 0, prio_hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CU(cudaStreamCreateWithPriority(&C.s_main, cudaStreamNonBlocking, prio_lo));
  CU(cudaStreamCreateWithPriority(&C.s_comm, cudaStreamNonBlocking, prio_hi));
  CU(cudaStreamCreateWithFlags(&C.s_h2d, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&C.s_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < MAXCHUNK; i++) {
    CU(cudaEventCreateWithFlags(&C.ev_up[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&C.ev_done[i], cudaEventDisableTiming));
  }
  CU(cudaEventCreateWithFlags(&C.ev_in, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&C.ev_halo, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&C.ev_z, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&C.ev_side[0], cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&C.ev_side[1], cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&C.ev_chk[0], cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&C.ev_chk[1], cudaEventDisableTiming));
  CU(cudaEventCreate(&C.ev_t0));
  CU(cudaEventCreate(&C.ev_t1));
  CU(cudaMalloc(&C.U, (size_t)72 * C.g.Vh * sizeof(double2)));
  CU(cudaMalloc(&C.stage, (size_t)32 * C.g.Vh * sizeof(double2))); /* one lexicographic spinor field (24 Vh) or one derivative field (32 Vh) */
  C.npartial = C.g.Vh / 64 + 4096;
  CU(cudaMalloc(&C.partial, (size_t)C.npartial * sizeof(double)));
  CU(cudaMalloc(&C.st, sizeof(tmb_cg_state)));
  CU(cudaMemset(C.st, 0, sizeof(tmb_cg_state)));
  CU(cudaHostAlloc(&C.st_host, 2 * sizeof(tmb_cg_state), cudaHostAllocDefault));
  const size_t hb = (size_t)2 * 6 * C.g.S * sizeof(double2); /* two faces: the two flavours of the doublet share an exchange */
  CU(cudaMalloc(&C.send_up, hb)); CU(cudaMalloc(&C.send_dn, hb));
  CU(cudaMalloc(&C.halo_up, hb)); CU(cudaMalloc(&C.halo_dn, hb));
  CU(cudaMalloc(&C.Uhalo, (size_t)18 * C.g.S * sizeof(double2)));
  C.kappa = 0.; C.mu = 0.;
  for (int m = 0; m < 4; m++) C.ka[m] = make_double2(0., 0.);
  C.nranks = 1; C.rank = 0; C.dist = false; C.loopback = false; C.gauge_loaded = false;
  C.launches = 0; C.hop_variant = -1; C.hints = -1; C.xblock = 0; C.tile = getenv("TMB_TILE") ? atoi(getenv("TMB_TILE")) : 0; C.pdl = 0; C.prefetch = 0; C.prefetch_dist = 0; C.compression = 18; C.cg_graph = 1; C.hop2_variant = -1; C.cg_selfnorm = 1; C.cg_tail = 1; C.cg_side = 0;
  C.init = true;
  return 0;
}

extern "C" int tmb_finalize(void) {
  if (!C.init) return 0;
  cudaDeviceSynchronize();
  hgraphs_clear(); pinned_clear();
  for (void *p : C.fields) sym_free(p);
  C.fields.clear(); C.field_bytes.clear();
  for (int i = 0; i < NSCRATCH; i++) { sym_free(C.scratch[i]); C.scratch[i] = nullptr; }
  for (int i = 0; i < NSCRATCH; i++) { sym_free(C.scratch32[i]); C.scratch32[i] = nullptr; }
  if (C.U32) cudaFree(C.U32); if (C.Uhalo32) cudaFree(C.Uhalo32);
  if (C.U12) cudaFree(C.U12); if (C.Uhalo12) cudaFree(C.Uhalo12); if (C.U12f) cudaFree(C.U12f); if (C.Uhalo12f) cudaFree(C.Uhalo12f);
  if (C.df) cudaFree(C.df); if (C.dhalo_send) cudaFree(C.dhalo_send); if (C.dhalo_recv) cudaFree(C.dhalo_recv);
  for (int k = 0; k < C.nmnl; k++) { sym_free(C.mnl[k].pf); for (int i = 0; i < TMB_MAXCSG; i++) sym_free(C.mnl[k].csg[i]); }
  for (int i = 0; i < 6; i++) sym_free(C.w[i]);
  for (int i = 0; i < 8; i++) { sym_free(C.nd[i]); C.nd[i] = nullptr; }
  for (int i = 0; i < 4; i++) { sym_free(C.nd32[i]); C.nd32[i] = nullptr; }
  {
    bool closed_up = false, closed_dn = false;
    for (int r = 0; r < TMB_XR_MAXR; r++)
      if (C.peer_base[r] && C.peer_base[r] != C.arena) {
        if (C.peer_base[r] == C.up_base) closed_up = true;
        if (C.peer_base[r] == C.dn_base) closed_dn = true;
        cudaIpcCloseMemHandle(C.peer_base[r]);
      }
    if (C.up_base && C.up_base != C.arena && !closed_up) cudaIpcCloseMemHandle(C.up_base);
    if (C.dn_base && C.dn_base != C.arena && C.dn_base != C.up_base && !closed_dn) cudaIpcCloseMemHandle(C.dn_base);
  }
  /* nobody may free an arena a peer still has mapped: all ranks pass here after unmapping and before freeing */
  if (C.comm && C.nranks > 1 && C.st) {
    C.nccl.AllReduce(&C.st->tmp[0], &C.st->tmp[0], 1, NCCL_FLOAT64, NCCL_SUM, C.comm, C.s_main);
    cudaStreamSynchronize(C.s_main);
  }
  if (C.comm && C.nccl.CommDestroy) { C.nccl.CommDestroy(C.comm); C.comm = nullptr; }
  if (C.seq_dev) cudaFree(C.seq_dev); if (C.xr_tab) cudaFree(C.xr_tab);
  if (C.arena) cudaFree(C.arena); else if (C.flags) cudaFree(C.flags);
  if (C.p2p_ticket) cudaFree(C.p2p_ticket); if (C.p2p_err) cudaFree(C.p2p_err);
  if (C.gauge_raw) cudaFree(C.gauge_raw);
  cudaFree(C.U); cudaFree(C.Uhalo); cudaFree(C.stage); cudaFree(C.partial); cudaFree(C.st);
  cudaFree(C.send_up); cudaFree(C.send_dn); cudaFree(C.halo_up); cudaFree(C.halo_dn);
  cudaFreeHost(C.st_host);
  if (C.zsend_up) cudaFree(C.zsend_up); if (C.zsend_dn) cudaFree(C.zsend_dn); if (C.zhalo_up) cudaFree(C.zhalo_up);
  if (C.zhalo_dn) cudaFree(C.zhalo_dn); if (C.Uzh) cudaFree(C.Uzh); if (C.Uzh32) cudaFree(C.Uzh32); if (C.Uzl) cudaFree(C.Uzl); if (C.Uzl32) cudaFree(C.Uzl32);
  cudaEventDestroy(C.ev_z);
  if (C.zloop_buf) cudaFree(C.zloop_buf);
  if (C.zd_send) cudaFree(C.zd_send); if (C.zd_recv) cudaFree(C.zd_recv);
  cudaEventDestroy(C.ev_in); cudaEventDestroy(C.ev_halo); cudaEventDestroy(C.ev_t0); cudaEventDestroy(C.ev_t1);
  cudaEventDestroy(C.ev_chk[0]); cudaEventDestroy(C.ev_chk[1]);
  for (int i = 0; i < MAXCHUNK; i++) { cudaEventDestroy(C.ev_up[i]); cudaEventDestroy(C.ev_done[i]); }
  cudaStreamDestroy(C.s_main); cudaStreamDestroy(C.s_comm); cudaStreamDestroy(C.s_h2d); cudaStreamDestroy(C.s_d2h);
  C = Ctx();
  return 0;
}

/* ------------------------------------------------------------------ communicator */
static int load_nccl() {
  if (C.nccl.handle) return 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(-110, "cannot dlopen libnccl.so.2: %s", dlerror());
  C.nccl.handle = h;
#define SYM(field, name) *(void **)(&C.nccl.field) = dlsym(h, name); if (!C.nccl.field) return fail(-111, "NCCL symbol %s missing", name)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
  SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
  SYM(AllReduce, "ncclAllReduce"); SYM(AllGather, "ncclAllGather"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return 0;
}
extern "C" int tmb_comm_unique_id(void *id128) {
  TRY(load_nccl());
  ncclUniqueId id;
  NC(C.nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return 0;
}
/* Peer mode set-up: one arena per rank for all spinor fields, its CUDA-IPC handle all-gathered over NCCL,
 * the two T-neighbours' arenas mapped into this process.  Any failure leaves the NCCL halo path in place. */
static int p2p_small_buffers() {
  if (!C.p2p_ticket) { CU(cudaMalloc(&C.p2p_ticket, 2 * sizeof(unsigned int))); CU(cudaMemset(C.p2p_ticket, 0, 2 * sizeof(unsigned int))); }
  const char *cc = getenv("TMB_P2P_COPY_CTAS");
  C.p2p_copy_ctas = cc ? atoi(cc) : 64;
  if (C.p2p_copy_ctas < 1) C.p2p_copy_ctas = 1;
  if (C.p2p_copy_ctas > 2 * 148) C.p2p_copy_ctas = 2 * 148; /* each pull CTA owns a partial slot and delays the stencil CTAs behind it */
  if (!C.p2p_err) { CU(cudaMalloc(&C.p2p_err, sizeof(int))); CU(cudaMemset(C.p2p_err, 0, sizeof(int))); }
  if (!C.seq_dev) CU(cudaMalloc(&C.seq_dev, 4 * sizeof(unsigned int)));
  CU(cudaMemset(C.seq_dev, 0, 4 * sizeof(unsigned int)));
  C.hop_off = 0; C.zhop_off = 0;
  return 0;
}
/* the table xred_sum() reads: every rank's landing arrays as seen from this rank */
static int setup_xred(char *const *bases, int nranks, int rank) {
  C.xred = false;
  const char *env = getenv("TMB_XRED");
  if ((env && atoi(env) == 0) || nranks > TMB_XR_MAXR) return 0;
  tmb_xred_table h;
  memset(&h, 0, sizeof(h));
  h.nranks = nranks; h.rank = rank; h.ctr = C.seq_dev + 1; h.err = C.p2p_err;
  for (int q = 0; q < nranks; q++) { h.val[q] = (double *)(bases[q] + ARENA_XR_VAL); h.seq[q] = (unsigned int *)(bases[q] + ARENA_XR_SEQ); }
  if (!C.xr_tab) CU(cudaMalloc(&C.xr_tab, sizeof(h)));
  CU(cudaMemcpy(C.xr_tab, &h, sizeof(h), cudaMemcpyHostToDevice));
  C.xred = true;
  return 0;
}
static int setup_p2p() {
  const char *env = getenv("TMB_P2P");
  if (env && atoi(env) == 0) return 0;
  if (!C.fields.empty()) { fprintf(stderr, "tmlqcd_b200: fields were allocated before tmb_comm_init; NCCL halos only\n"); return 0; }
  size_t freeb = 0, totalb = 0;
  CU(cudaMemGetInfo(&freeb, &totalb));
  size_t want = (size_t)48 * FIELD_BYTES() + ((size_t)64 << 20);
  const char *mb = getenv("TMB_ARENA_MB");
  if (mb) want = (size_t)atoll(mb) << 20;
  if (want > freeb / 10 * 7) want = freeb / 10 * 7;
  want &= ~(size_t)((2 << 20) - 1);
  { /* every rank must run out of arena at the same allocation (peer mode or NCCL halos is decided per field, and the
     * two sides of a hop must decide alike): agree on the smallest size any rank can afford */
    double *dmin = nullptr; CU(cudaMalloc(&dmin, sizeof(double)));
    double w = (double)want;
    CU(cudaMemcpy(dmin, &w, sizeof(double), cudaMemcpyHostToDevice));
    NC(C.nccl.AllReduce(dmin, dmin, 1, NCCL_FLOAT64, NCCL_MIN, C.comm, C.s_main));
    CU(cudaStreamSynchronize(C.s_main));
    CU(cudaMemcpy(&w, dmin, sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dmin);
    want = (size_t)w;
  }
  if (cudaMalloc(&C.arena, want) != cudaSuccess) { cudaGetLastError(); C.arena = nullptr; fprintf(stderr, "tmlqcd_b200: arena allocation failed; NCCL halos only\n"); return 0; }
  C.arena_bytes = want; C.arena_used = ARENA_RESERVED; C.arena_free.clear();
  CU(cudaMemset(C.arena, 0, ARENA_RESERVED));
  cudaIpcMemHandle_t mine;
  std::vector<cudaIpcMemHandle_t> all(C.nranks);
  bool ok = cudaIpcGetMemHandle(&mine, C.arena) == cudaSuccess;
  char *dsend = nullptr, *drecv = nullptr;
  CU(cudaMalloc(&dsend, sizeof(mine) + 8)); CU(cudaMalloc(&drecv, (sizeof(mine) + 8) * C.nranks));
  struct Msg { cudaIpcMemHandle_t h; long long ok; } msg;
  msg.h = mine; msg.ok = ok ? 1 : 0;
  static_assert(sizeof(Msg) == sizeof(cudaIpcMemHandle_t) + 8, "packing");
  CU(cudaMemcpy(dsend, &msg, sizeof(msg), cudaMemcpyHostToDevice));
  NC(C.nccl.AllGather(dsend, drecv, sizeof(msg), NCCL_UINT8, C.comm, C.s_main));
  CU(cudaStreamSynchronize(C.s_main));
  std::vector<Msg> msgs(C.nranks);
  CU(cudaMemcpy(msgs.data(), drecv, sizeof(Msg) * C.nranks, cudaMemcpyDeviceToHost));
  cudaFree(dsend); cudaFree(drecv);
  for (int r = 0; r < C.nranks; r++) ok = ok && msgs[r].ok;
  const int up = t_up(), dn = t_dn(); /* the hops read the T neighbours' fields */
  void *pu = nullptr, *pd = nullptr;
  /* all ranks' arenas when the cross-rank sums can use them (<= TMB_XR_MAXR ranks), the two T-neighbours' otherwise */
  const bool allp = C.nranks <= TMB_XR_MAXR;
  for (int r = 0; r < C.nranks && ok; r++) {
    if (r == C.rank) { if (allp) C.peer_base[r] = C.arena; continue; }
    if (!allp && r != up && r != dn) continue;
    void *pp = nullptr;
    if (cudaIpcOpenMemHandle(&pp, msgs[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
    if (allp) C.peer_base[r] = (char *)pp;
    if (r == up) pu = pp;
    if (r == dn) pd = pp;
  }
  /* everybody must agree, otherwise one rank would wait for flags nobody writes */
  int *dflag = nullptr; CU(cudaMalloc(&dflag, sizeof(double)));
  double mineok = ok ? 0. : 1., sum = 0.;
  CU(cudaMemcpy(dflag, &mineok, sizeof(double), cudaMemcpyHostToDevice));
  NC(C.nccl.AllReduce(dflag, dflag, 1, NCCL_FLOAT64, NCCL_SUM, C.comm, C.s_main));
  CU(cudaStreamSynchronize(C.s_main));
  CU(cudaMemcpy(&sum, dflag, sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(dflag);
  if (sum != 0.) { fprintf(stderr, "tmlqcd_b200: CUDA IPC peer mapping unavailable; NCCL halos only\n"); return 0; }
  C.up_base = (char *)pu; C.dn_base = (char *)pd;
  C.flags = (unsigned int *)C.arena; C.up_flags = (unsigned int *)C.up_base; C.dn_flags = (unsigned int *)C.dn_base;
  TRY(p2p_small_buffers());
  C.p2p = true;
  if (allp) TRY(setup_xred(C.peer_base, C.nranks, C.rank));
  /* Z split: the z faces go through peer memory too when every arena is mapped.  The four halo buffers (2 sides x 2
   * sequence parities) sit at the same offset of every rank's arena, in front of the fields. */
  C.zpeer = false;
  const char *ez = getenv("TMB_ZPEER");
  if (C.zsplit && allp && !(ez && atoi(ez) == 0)) {
    const size_t fb = ((size_t)6 * SZ() * sizeof(double2) + 255) & ~(size_t)255;
    if (C.arena_used + 4 * fb <= C.arena_bytes / 2) { /* the same decision on every rank: the arena size is an agreed one */
      const size_t off = C.arena_used;
      C.arena_used += 4 * fb;
      char *mine_b = C.arena + off, *up_b = C.peer_base[z_up()] + off, *dn_b = C.peer_base[z_dn()] + off;
      for (int b = 0; b < 2; b++) {
        C.zloc_dn[b] = mine_b + (size_t)b * fb; C.zloc_up[b] = mine_b + (size_t)(2 + b) * fb;
        C.zdst_up[b] = up_b + (size_t)b * fb;        /* my last-z face -> rank z+1's halo_dn */
        C.zdst_dn[b] = dn_b + (size_t)(2 + b) * fb;  /* my first-z face -> rank z-1's halo_up */
      }
      C.zflags = (unsigned int *)(C.arena + ARENA_ZFLAGS);
      C.zflag_at_up = (unsigned int *)(C.peer_base[z_up()] + ARENA_ZFLAGS) + 0;
      C.zflag_at_dn = (unsigned int *)(C.peer_base[z_dn()] + ARENA_ZFLAGS) + 1;
      C.zpeer = true;
    }
  }
  return 0;
}
static int z_buffers() {
  const size_t fb = (size_t)6 * SZ() * sizeof(double2);
  if (!C.zsend_up) { CU(cudaMalloc(&C.zsend_up, fb)); CU(cudaMalloc(&C.zsend_dn, fb)); CU(cudaMalloc(&C.zhalo_up, fb)); CU(cudaMalloc(&C.zhalo_dn, fb)); }
  if (!C.Uzh) CU(cudaMalloc(&C.Uzh, (size_t)18 * SZ() * sizeof(double2)));
  if (!C.Uzl) CU(cudaMalloc(&C.Uzl, (size_t)18 * SZ() * sizeof(double2)));
  return 0;
}
extern "C" int tmb_comm_init_grid(const void *id128, int nt, int nz, int rank);
extern "C" int tmb_comm_init(const void *id128, int nranks, int rank) { return tmb_comm_init_grid(id128, nranks, 1, rank); }
/* rank grid nt x nz, rank = ct * nz + cz; tmb_init's extents are the LOCAL ones (T / nt, LZ / nz, both even) */
extern "C" int tmb_comm_init_grid(const void *id128, int nt, int nz, int rank) {
  NEED_INIT();
  const int nranks = nt * nz;
  if (nt < 1 || nz < 1 || rank < 0 || rank >= nranks) return fail(-5, "tmb_comm_init: bad rank %d of %d x %d", rank, nt, nz);
  if (nranks == 1) { C.nranks = 1; C.rank = 0; C.nt = C.nz = 1; C.ct = C.cz = 0; C.dist = C.loopback; C.g.dist_t = C.dist; C.zsplit = C.loop_z; return 0; }
  TRY(load_nccl());
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NC(C.nccl.CommInitRank(&C.comm, nranks, id, rank));
  C.nranks = nranks; C.rank = rank; C.nt = nt; C.nz = nz; C.ct = rank / nz; C.cz = rank % nz;
  C.dist = nt > 1; C.g.dist_t = C.dist ? 1 : 0; C.zsplit = nz > 1; C.param_gen++;
  if (C.zsplit) TRY(z_buffers());
  /* a gauge field uploaded before this call has no exchanged U_0 halo (nor float / 12-real copies of one): make the
   * next hop fail with "call tmb_gauge_upload first" instead of reading uninitialised Uhalo */
  C.gauge_loaded = false; C.gauge32_valid = false; C.c12_valid = false; C.c12f_valid = false;
  TRY(setup_p2p());
  return 0;
}
extern "C" int tmb_comm_peer_mode(void) { return C.p2p ? 1 : 0; }
extern "C" int tmb_comm_zpeer_mode(void) { return C.zsplit && C.zpeer ? 1 : 0; }
/* test hook: how many sequence numbers the peer-mode T hops and the z-face pushes have taken so far (they count separately:
 * the parity of the push count picks the z halo buffer, and a shared counter would give every push of a T x Z grid the same
 * parity) */
extern "C" int tmb_comm_sequence_counts(unsigned int *t_hops, unsigned int *z_pushes) {
  NEED_INIT();
  if (t_hops) *t_hops = C.hop_off;
  if (z_pushes) *z_pushes = C.zhop_off;
  return 0;
}
/* single GPU: exercise the Z-split path (face pack, exchange with itself, fix-up) */
extern "C" int tmb_comm_loopback_z(int on) {
  NEED_INIT();
  if (C.nranks > 1) return fail(-6, "tmb_comm_loopback_z: only for a single rank");
  C.loop_z = on != 0; C.zsplit = C.loop_z; C.gauge_loaded = false; C.param_gen++;
  if (C.zsplit) TRY(z_buffers());
  C.zpeer = false;
  if (on == 2) { /* the peer-memory push of the z faces against itself: local buffers and flags stand in for the neighbours' */
    const size_t fb = ((size_t)6 * SZ() * sizeof(double2) + 255) & ~(size_t)255;
    if (!C.zloop_buf) CU(cudaMalloc(&C.zloop_buf, 4 * fb));
    if (!C.flags) { CU(cudaMalloc(&C.flags, ARENA_RESERVED)); CU(cudaMemset(C.flags, 0, ARENA_RESERVED)); }
    TRY(p2p_small_buffers());
    CU(cudaMemset((char *)C.flags + ARENA_ZFLAGS, 0, 8));
    for (int b = 0; b < 2; b++) {
      C.zloc_dn[b] = C.zloop_buf + (size_t)b * fb; C.zloc_up[b] = C.zloop_buf + (size_t)(2 + b) * fb;
      C.zdst_up[b] = C.zloc_dn[b]; C.zdst_dn[b] = C.zloc_up[b];
    }
    C.zflags = (unsigned int *)((char *)C.flags + ARENA_ZFLAGS);
    C.zflag_at_up = C.zflags + 0; C.zflag_at_dn = C.zflags + 1;
    C.zpeer = true;
  }
  return 0;
}
/* on = 1: halo buffers (pack / copy / boundary launch); on = 2: peer mode against itself (one launch, flags) */
extern "C" int tmb_comm_loopback(int on) {
  NEED_INIT();
  if (C.nranks > 1) return fail(-6, "tmb_comm_loopback: only for a single rank");
  C.loopback = on != 0; C.loopback_mode = on; C.dist = C.loopback; C.g.dist_t = C.dist ? 1 : 0; C.param_gen++;
  C.gauge_loaded = false; /* Uhalo must be rebuilt */
  C.p2p = false; C.xred = false;
  if (on == 2) {
    if (!C.flags) { CU(cudaMalloc(&C.flags, ARENA_RESERVED)); }
    CU(cudaMemset(C.flags, 0, ARENA_RESERVED));
    C.up_flags = C.dn_flags = C.flags;
    TRY(p2p_small_buffers());
    C.p2p = true;
    char *self[1] = {(char *)C.flags}; /* the sums run through the landing arrays too, over one rank */
    TRY(setup_xred(self, 1, 0));
  }
  return 0;
}
/* reductions finish inside the producing kernel (no separate launch, no NCCL) on one rank and, over several ranks, when the
 * cross-rank sum can go through peer memory */
static inline bool fuse_fin() { return C.nranks == 1 || C.xred; }
static inline const tmb_xred_table *xr_tab() { return C.xred ? C.xr_tab : nullptr; }
static int allreduce_slot(int slot) {
  if (C.nranks > 1) NC(C.nccl.AllReduce(&C.st->tmp[slot], &C.st->tmp[slot], 1, NCCL_FLOAT64, NCCL_SUM, C.comm, C.s_main));
  return 0;
}

/* ------------------------------------------------------------------ parameters */
extern "C" int tmb_set_boundary(double kappa, const double theta[4]) {
  NEED_INIT();
  /* boundary.c:40-55, incl. its PI_ literal; global extents: T is distributed over nt ranks, Z over nz */
  const double PI_ = 3.14159265358979;
  const double ext[4] = {(double)C.g.T * C.nt, (double)C.g.LX, (double)C.g.LY, (double)C.g.LZ * C.nz};
  C.kappa = kappa; C.param_gen++;
  for (int m = 0; m < 4; m++) {
    const double x = (theta ? theta[m] : 0.) * PI_ / ext[m];
    C.ka[m] = make_double2(kappa * cos(x), kappa * sin(x));
    C.phase[m] = make_double2(cos(x), sin(x));
  }
  return 0;
}
/* ka0..ka3 exactly as the caller's boundary() computed them (re,im pairs) */
extern "C" int tmb_set_hopping_phases(const double ka_re_im[8]) {
  NEED_INIT();
  bool same = true;
  for (int m = 0; m < 4; m++) same = same && C.ka[m].x == ka_re_im[2 * m] && C.ka[m].y == ka_re_im[2 * m + 1];
  if (same) return 0; /* the drop-in layer pushes the globals at every call: nothing changed, cached graphs stay valid */
  C.param_gen++;
  for (int m = 0; m < 4; m++) C.ka[m] = make_double2(ka_re_im[2 * m], ka_re_im[2 * m + 1]);
  C.kappa = sqrt(C.ka[0].x * C.ka[0].x + C.ka[0].y * C.ka[0].y);
  for (int m = 0; m < 4; m++) {
    const double a = sqrt(C.ka[m].x * C.ka[m].x + C.ka[m].y * C.ka[m].y);
    C.phase[m] = a > 0. ? make_double2(C.ka[m].x / a, C.ka[m].y / a) : make_double2(1., 0.);
  }
  return 0;
}
extern "C" int tmb_set_mu(double g_mu) { NEED_INIT(); if (C.mu != g_mu) { C.mu = g_mu; C.param_gen++; } return 0; }
extern "C" int tmb_set_nd(double mubar, double epsbar, double invmaxev) {
  NEED_INIT(); C.mubar = mubar; C.epsbar = epsbar; C.invmaxev = invmaxev; return 0;
}
extern "C" int tmb_set_hop2_variant(int v) { NEED_INIT(); if (v < -1 || v > 3) return fail(-7, "hop2 variant must be -1 (automatic), 0, 1, 2 or 3"); C.hop2_variant = v; return 0; }
/* with tmb_set_overlap bit 1 (L2 bulk prefetch of gauge rows): prefetch the rows of the CTA `ctas` CTAs ahead (one wave:
 * 148 x resident CTAs per SM) instead of the CTA's own */
extern "C" int tmb_set_prefetch_distance(int ctas) { NEED_INIT(); if (ctas < 0) return fail(-7, "prefetch distance must be >= 0"); C.prefetch_dist = ctas; C.param_gen++; return 0; }
extern "C" int tmb_set_tile(int on) { /* CTA tile traversal of the hopping kernels (default off: measured, no gain; TMB_TILE=1) */
  if (!C.init) return fail(-1, "not initialised");
  C.tile = on ? 1 : 0; C.param_gen++;
  return 0;
}
extern "C" int tmb_set_tuning(int hop_variant, int cache_hints, int xblock) {
  NEED_INIT();
  if (hop_variant < -1 || hop_variant > 10) return fail(-7, "hop_variant must be -1 (automatic) .. 10");
  if (xblock > 0 && C.g.LX % xblock) return fail(-7, "xblock must divide LX");
  C.hop_variant = hop_variant; C.hints = cache_hints < 0 ? -1 : (cache_hints ? 1 : 0); C.xblock = xblock; C.param_gen++;
  return 0;
}
/* bit 0: programmatic dependent launch of the hopping kernels, bit 1: L2 bulk prefetch of gauge rows */
/* number of time-slice chunks of the pipelined host-pointer Hopping_Matrix (1..64) */
extern "C" int tmb_set_host_chunks(int n) { NEED_INIT(); if (n < 0 || n > MAXCHUNK) return fail(-7, "host chunks must be in [0, %d] (0: automatic schedule)", MAXCHUNK); C.host_chunks = n; C.param_gen++; return 0; }
/* explicit chunk sizes (time-slices, in order) for the pipelined host-pointer Hopping_Matrix; n = 0 returns to the automatic
 * schedule.  Used when the sizes add up to the number of pipelined slices, ignored otherwise. */
extern "C" int tmb_set_host_chunk_sizes(const int *sizes, int n) {
  NEED_INIT();
  if (n < 0 || n > MAXCHUNK - 2) return fail(-7, "at most %d chunks", MAXCHUNK - 2);
  if (n > 0 && sizes == nullptr) return fail(-7, "tmb_set_host_chunk_sizes: null size array");
  for (int i = 0; i < n; i++) if (sizes[i] <= 0) return fail(-7, "chunk sizes must be positive");
  for (int i = 0; i < n; i++) C.host_sched[i] = sizes[i];
  C.host_sched_n = n; C.param_gen++;
  return 0;
}
extern "C" int tmb_set_overlap(int flags) {
  NEED_INIT();
  if (flags & ~127) return fail(-7, "tmb_set_overlap: unknown bits in 0x%x (bits 0..6 are defined)", flags);
  C.pdl = flags & 1; C.prefetch = ((flags >> 1) & 1) | ((flags & 8) ? 2 : 0); C.cg_graph = (flags & 4) ? 0 : 1; C.cg_selfnorm = (flags & 16) ? 0 : 1;
  C.cg_tail = (flags & 32) ? 0 : 1;
  C.cg_side = (flags & 64) ? 1 : 0;
  C.param_gen++;
  return 0;
}
/* Peer-mode TIMING diagnostics, kept apart from the tuning options above because every one of them gives WRONG
 * results across ranks: 1 boundary slices read from the LOCAL field, 2 no end-of-hop handshake, 4 no pull,
 * 8 no halo path in the boundary CTAs, 16 natural slice order.  Refused unless TMB_P2P_DIAG=1 is set in the environment. */
extern "C" int tmb_set_p2p_diag(int bits) {
  NEED_INIT();
  if (bits & ~31) return fail(-7, "tmb_set_p2p_diag: unknown bits in 0x%x", bits);
  const char *e = getenv("TMB_P2P_DIAG");
  if (bits && !(e && atoi(e) == 1)) return fail(-7, "tmb_set_p2p_diag: results become invalid; set TMB_P2P_DIAG=1 to allow it");
  C.p2p_diag = bits;
  return 0;
}

/* ------------------------------------------------------------------ memory */
extern "C" void *tmb_field_alloc(void) {
  if (!C.init) { fail(-1, "tmb_init has not been called"); return nullptr; }
  void *p = sym_malloc(FIELD_BYTES());
  if (!p) { fail(-100, "cudaMalloc of a spinor field failed"); return nullptr; }
  cudaMemsetAsync(p, 0, FIELD_BYTES(), C.s_main);
  C.fields.push_back(p);
  C.field_bytes.push_back(FIELD_BYTES());
  return p;
}
extern "C" int tmb_field_free(void *field) {
  NEED_INIT();
  for (size_t i = 0; i < C.fields.size(); i++)
    if (C.fields[i] == field) {
      CU(cudaStreamSynchronize(C.s_main));
      arena_release(field, C.field_bytes[i]);
      C.fields.erase(C.fields.begin() + i); C.field_bytes.erase(C.field_bytes.begin() + i);
      return 0;
    }
  return fail(-8, "tmb_field_free: unknown field");
}
extern "C" int tmb_field_zero(void *field) { NEED_INIT(); CU(cudaMemsetAsync(field, 0, FIELD_BYTES(), C.s_main)); return 0; }
extern "C" void *tmb_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { fail(-100, "cudaHostAlloc failed"); return nullptr; }
  return p;
}
extern "C" int tmb_host_free(void *p) { CU(cudaFreeHost(p)); return 0; }
extern "C" int tmb_host_register(void *p, size_t bytes) { CU(cudaHostRegister(p, bytes, cudaHostRegisterDefault)); hgraphs_clear(); return 0; }
extern "C" int tmb_host_unregister(void *p) { hgraphs_clear(); CU(cudaHostUnregister(p)); return 0; }
extern "C" int tmb_sync(void) {
  NEED_INIT(); CU(cudaStreamSynchronize(C.s_comm)); CU(cudaStreamSynchronize(C.s_main));
  if (C.p2p_err) {
    int e = 0;
    CU(cudaMemcpy(&e, C.p2p_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) { cudaMemset(C.p2p_err, 0, sizeof(int)); return fail(-120, "peer-mode hop: a neighbour's flag did not arrive within the time-out (ranks out of step?)"); }
  }
  return 0;
}
extern "C" int tmb_timer_start(void) { NEED_INIT(); CU(cudaEventRecord(C.ev_t0, C.s_main)); return 0; }
extern "C" int tmb_timer_stop(float *ms) {
  NEED_INIT();
  CU(cudaEventRecord(C.ev_t1, C.s_main));
  CU(cudaEventSynchronize(C.ev_t1));
  CU(cudaEventElapsedTime(ms, C.ev_t0, C.ev_t1));
  return 0;
}

/* Measurement aid for bench.py's roofline: device-to-device copy bandwidth (read + write bytes) of `bytes` per copy,
 * sustained over `reps` back-to-back copies on the compute stream - the same quantity MEASURED_PEAKS.json quotes as a
 * burst figure, taken in the same process and thermal state as the kernels it is compared with. */
extern "C" int tmb_measure_copy_gbs(size_t bytes, int reps, double *gbs) {
  NEED_INIT();
  if (bytes < 1024 || reps < 1 || !gbs) return fail(-7, "tmb_measure_copy_gbs: bad arguments");
  void *a = nullptr, *b = nullptr;
  CU(cudaMalloc(&a, bytes));
  if (cudaMalloc(&b, bytes) != cudaSuccess) { cudaFree(a); return fail(-100, "tmb_measure_copy_gbs: out of memory"); }
  cudaMemsetAsync(a, 1, bytes, C.s_main);
  for (int i = 0; i < 3; i++) cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, C.s_main);
  cudaEventRecord(C.ev_t0, C.s_main);
  for (int i = 0; i < reps; i++) cudaMemcpyAsync((i & 1) ? a : b, (i & 1) ? b : a, bytes, cudaMemcpyDeviceToDevice, C.s_main);
  cudaEventRecord(C.ev_t1, C.s_main);
  cudaError_t e = cudaEventSynchronize(C.ev_t1);
  float ms = 0.f;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, C.ev_t0, C.ev_t1);
  cudaFree(a); cudaFree(b);
  if (e != cudaSuccess) return fail(-100, "tmb_measure_copy_gbs: %s", cudaGetErrorString(e));
  *gbs = 2. * (double)bytes * reps / (ms * 1e-3) / 1e9;
  return 0;
}

/* Measurement aid for bench.py's e2e leg: what this box's host link gives with pinned memory - host-to-device alone,
 * device-to-host alone, and both directions at once (GB/s per direction) - `reps` copies of `bytes` on the two copy streams the
 * pipelined host-pointer operators use.  The e2e figure is quoted as a fraction of the last one. */
extern "C" int tmb_measure_pcie_gbs(size_t bytes, int reps, double *h2d, double *d2h, double *duplex_each) {
  NEED_INIT();
  if (bytes < 4096 || reps < 1) return fail(-7, "tmb_measure_pcie_gbs: bad arguments");
  void *ha = nullptr, *hb = nullptr, *da = nullptr, *db = nullptr;
  cudaEvent_t e0, e1, e2;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1)); CU(cudaEventCreate(&e2));
  if (cudaHostAlloc(&ha, bytes, cudaHostAllocDefault) != cudaSuccess || cudaHostAlloc(&hb, bytes, cudaHostAllocDefault) != cudaSuccess ||
      cudaMalloc(&da, bytes) != cudaSuccess || cudaMalloc(&db, bytes) != cudaSuccess) {
    cudaFreeHost(ha); cudaFreeHost(hb); cudaFree(da); cudaFree(db);
    return fail(-100, "tmb_measure_pcie_gbs: out of memory");
  }
  memset(ha, 1, bytes); memset(hb, 2, bytes);
  float ms = 0.f;
  auto run = [&](bool up, bool down, double *each) -> int {
    CU(cudaDeviceSynchronize());
    if (up) CU(cudaMemcpyAsync(da, ha, bytes, cudaMemcpyHostToDevice, C.s_h2d));   /* warm-up */
    if (down) CU(cudaMemcpyAsync(hb, db, bytes, cudaMemcpyDeviceToHost, C.s_d2h));
    CU(cudaDeviceSynchronize());
    CU(cudaEventRecord(e0, C.s_main));
    CU(cudaStreamWaitEvent(C.s_h2d, e0, 0)); CU(cudaStreamWaitEvent(C.s_d2h, e0, 0));
    for (int i = 0; i < reps; i++) {
      if (up) CU(cudaMemcpyAsync(da, ha, bytes, cudaMemcpyHostToDevice, C.s_h2d));
      if (down) CU(cudaMemcpyAsync(hb, db, bytes, cudaMemcpyDeviceToHost, C.s_d2h));
    }
    CU(cudaEventRecord(e1, C.s_h2d)); CU(cudaEventRecord(e2, C.s_d2h));
    CU(cudaStreamWaitEvent(C.s_main, e1, 0)); CU(cudaStreamWaitEvent(C.s_main, e2, 0));
    CU(cudaEventRecord(e1, C.s_main));
    CU(cudaEventSynchronize(e1));
    CU(cudaEventElapsedTime(&ms, e0, e1));
    *each = (double)bytes * reps / (ms * 1e-3) / 1e9;
    return 0;
  };
  double a = 0., b = 0., c = 0.;
  int rc = run(true, false, &a);
  if (rc >= 0) rc = run(false, true, &b);
  if (rc >= 0) rc = run(true, true, &c);
  cudaFreeHost(ha); cudaFreeHost(hb); cudaFree(da); cudaFree(db);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  if (rc < 0) return rc;
  if (h2d) *h2d = a; if (d2h) *d2h = b; if (duplex_each) *duplex_each = c;
  return 0;
}

static void ensure_pinned(const void *p, size_t bytes);
extern "C" int tmb_field_upload(void *field, const double *host) {
  NEED_INIT();
  ensure_pinned(host, FIELD_BYTES());
  CU(cudaMemcpyAsync(C.stage, host, FIELD_BYTES(), cudaMemcpyHostToDevice, C.s_main));
  KL(tmb_launch_pack_eo(F(field), C.stage, C.g.Vh, C.s_main));
  CU(cudaStreamSynchronize(C.s_main));
  return 0;
}
extern "C" int tmb_field_download(double *host, const void *field) {
  NEED_INIT();
  ensure_pinned(host, FIELD_BYTES());
  KL(tmb_launch_unpack_eo(C.stage, F(field), C.g.Vh, C.s_main));
  CU(cudaMemcpyAsync(host, C.stage, FIELD_BYTES(), cudaMemcpyDeviceToHost, C.s_main));
  CU(cudaStreamSynchronize(C.s_main));
  return 0;
}
extern "C" int tmb_field_upload_lexic(void *even, void *odd, const double *host) {
  NEED_INIT();
  ensure_pinned(host, 2 * FIELD_BYTES());
  CU(cudaMemcpyAsync(C.stage, host, 2 * FIELD_BYTES(), cudaMemcpyHostToDevice, C.s_main));
  KL(tmb_launch_pack_lexic(F(even), F(odd), C.stage, C.g, C.s_main));
  CU(cudaStreamSynchronize(C.s_main));
  return 0;
}
extern "C" int tmb_field_download_lexic(double *host, const void *even, const void *odd) {
  NEED_INIT();
  ensure_pinned(host, 2 * FIELD_BYTES());
  KL(tmb_launch_unpack_lexic(C.stage, F(even), F(odd), C.g, C.s_main));
  CU(cudaMemcpyAsync(host, C.stage, 2 * FIELD_BYTES(), cudaMemcpyDeviceToHost, C.s_main));
  CU(cudaStreamSynchronize(C.s_main));
  return 0;
}

/* exchange of the two T-face buffers: send_up -> rank+1's halo_dn, send_dn -> rank-1's halo_up */
/* the same for the two Z faces: send_up -> rank z+1's halo_dn, send_dn -> rank z-1's halo_up */
static int exchange_zfaces(const void *sup, const void *sdn, void *hup, void *hdn, size_t bytes, cudaStream_t s) {
  if (C.nz == 1) { /* loopback */
    CU(cudaMemcpyAsync(hdn, sup, bytes, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(hup, sdn, bytes, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  const int up = z_up(), dn = z_dn();
  NC(C.nccl.GroupStart());
  NC(C.nccl.Send(sup, bytes, NCCL_UINT8, up, C.comm, s));
  NC(C.nccl.Recv(hdn, bytes, NCCL_UINT8, dn, C.comm, s));
  NC(C.nccl.Send(sdn, bytes, NCCL_UINT8, dn, C.comm, s));
  NC(C.nccl.Recv(hup, bytes, NCCL_UINT8, up, C.comm, s));
  NC(C.nccl.GroupEnd());
  return 0;
}
static int exchange_faces(const void *sup, const void *sdn, void *hup, void *hdn, size_t bytes, cudaStream_t s) {
  if (C.nt == 1) { /* loopback: this rank is its own neighbour in T */
    CU(cudaMemcpyAsync(hdn, sup, bytes, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(hup, sdn, bytes, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  const int up = t_up(), dn = t_dn();
  NC(C.nccl.GroupStart());
  NC(C.nccl.Send(sup, bytes, NCCL_UINT8, up, C.comm, s));
  NC(C.nccl.Recv(hdn, bytes, NCCL_UINT8, dn, C.comm, s));
  NC(C.nccl.Send(sdn, bytes, NCCL_UINT8, dn, C.comm, s));
  NC(C.nccl.Recv(hup, bytes, NCCL_UINT8, up, C.comm, s));
  NC(C.nccl.GroupEnd());
  return 0;
}

extern "C" int tmb_gauge_upload(const double *host_gauge) {
  NEED_INIT();
  const size_t bytes = (size_t)72 * C.g.Vh * sizeof(double2); /* V*4 links * 9 complex */
  /* AoS staging copy: allocated per upload for a one-off configuration (inversions), kept from the second
   * upload on (HMC: the links change every MD step, update_gauge.c:109 sets the dirty flag each time) */
  double2 *raw = C.gauge_raw;
  if (!raw) CU(cudaMalloc(&raw, bytes));
  CU(cudaMemcpyAsync(raw, host_gauge, bytes, cudaMemcpyHostToDevice, C.s_main));
  KL(tmb_launch_pack_gauge(C.U, raw, C.g, C.s_main));
  CU(cudaStreamSynchronize(C.s_main));
  if (++C.gauge_uploads >= 2) C.gauge_raw = raw; else CU(cudaFree(raw));
  if (C.dist) { /* one-off gauge halo: U_0 of rank-1's last time-slice (xchange_gauge in the reference) */
    double2 *tmp = nullptr;
    const size_t n = (size_t)18 * C.g.S;
    CU(cudaMalloc(&tmp, n * sizeof(double2)));
    KL(tmb_launch_pack_gauge_halo(tmp, C.U, C.g, C.s_main));
    if (C.nt == 1) {
      CU(cudaMemcpyAsync(C.Uhalo, tmp, n * sizeof(double2), cudaMemcpyDeviceToDevice, C.s_main));
    } else {
      const int up = t_up(), dn = t_dn();
      NC(C.nccl.GroupStart());
      NC(C.nccl.Send(tmp, 2 * n, NCCL_FLOAT64, up, C.comm, C.s_main));
      NC(C.nccl.Recv(C.Uhalo, 2 * n, NCCL_FLOAT64, dn, C.comm, C.s_main));
      NC(C.nccl.GroupEnd());
    }
    CU(cudaStreamSynchronize(C.s_main));
    CU(cudaFree(tmp));
  }
  if (C.zsplit) { /* U_z of rank z-1's last-z sites, for the -z hops of the z = 0 face */
    double2 *tmp = C.Uzl; /* kept: the fix-up reads this rank's own last-z links from it (contiguous) */
    const size_t n = (size_t)18 * SZ();
    KL(tmb_launch_pack_gauge_zhalo(0, tmp, C.U, C.g, C.s_main));
    if (C.nz == 1) {
      CU(cudaMemcpyAsync(C.Uzh, tmp, n * sizeof(double2), cudaMemcpyDeviceToDevice, C.s_main));
    } else {
      NC(C.nccl.GroupStart());
      NC(C.nccl.Send(tmp, 2 * n, NCCL_FLOAT64, z_up(), C.comm, C.s_main));
      NC(C.nccl.Recv(C.Uzh, 2 * n, NCCL_FLOAT64, z_dn(), C.comm, C.s_main));
      NC(C.nccl.GroupEnd());
    }
    CU(cudaStreamSynchronize(C.s_main));
  }
  C.gauge_loaded = true;
  C.gauge32_valid = false; C.c12_valid = false; C.c12f_valid = false;
  if (C.compression == 12) TRY(ensure_gauge12(0));
  return 0;
}

static double2 *scratch(int k) {
  if (!C.scratch[k]) {
    if (!(C.scratch[k] = (double2 *)sym_malloc(FIELD_BYTES()))) { fail(-100, "cudaMalloc of a scratch field failed"); return nullptr; }
    cudaMemsetAsync(C.scratch[k], 0, FIELD_BYTES(), C.s_main);
  }
  return C.scratch[k];
}
#define SCR(var, k) double2 *var = scratch(k); if (!var) return -100

/* ------------------------------------------------------------------ the hopping term */
struct HopOpt {
  int mode = 0; double2 cf = {1., 0.}; const void *p = nullptr;
  const void *dotw = nullptr; const tmb_cg_state *st = nullptr;
  bool selfnorm = false;      /* fused reduction of |out|^2 instead of <dotw, out> */
  int *npartial = nullptr;
  int site0 = 0, nsites = -1; /* sub-range of output sites (single rank only); -1: all */
  int prec = 0;               /* 0: double fields, 1: float fields + float gauge copy */
  int fin_op = -1, fin_slot = 0; /* fused finish of the dot reduction (single rank) */
  bool nocom = false;         /* Hopping_Matrix_nocom: no halo exchange, the slab wraps onto itself in T */
  bool boundary_only = false; /* split T: exchange the faces and compute time-slices 0 and T-1 only (the interior was done in sub-ranges) */
  /* two flavours in one launch (hop_kernel NFL = 2): second flavour's fields, flavour mixing of the epilogue */
  void *cg_x = nullptr, *cg_r = nullptr; const void *cg_p = nullptr; /* mode 4: the CG's x / r update in the epilogue */
  int nfl = 1; const void *in1 = nullptr; void *out1 = nullptr; const void *p1 = nullptr;
  double nd_mu = 0., nd_eps = 0., nd_scale = 1., dot_scale = 1.;
};
static int ensure_gauge32();
static int ensure_gauge12(int prec);
/* Wave quantisation: a launch of n sites occupies n / (148 x R) waves of resident threads, R = 384 (168 registers)
 * or 448 (144 registers, variant 10).  The fuller residency wins when it saves a nearly empty trailing wave on a
 * launch of only a few waves (16^3x32: 1.15 -> 0.99); on long launches the extra spill traffic costs more than the
 * tail (24^3x48: 5.84 vs 5.004 waves - the 448 variant would ADD a wave there). */
/* L2 cache policies of the hop kernels (gauge links evict-first / no L1 allocation, neighbour spinors evict-last):
 * right when the links stream through L2 once per hop (24^3x48: 84.6 vs 92.3 us), wrong when gauge field + the six CG
 * vectors (2304 B per site of one parity) fit in the 126 MB L2 - then evict-first throws away links the next hop
 * would have hit (16^4: CG 77.8 -> 60.1 us/iteration without the policies).  -1 = decide by the working set. */
static int eff_hints() {
  if (C.hints >= 0) return C.hints;
  return 2304. * C.g.Vh > 100e6 ? 1 : 0;
}
static bool hop_residency_448(int nsites) {
  const double w384 = nsites / (148. * 384.), w448 = nsites / (148. * 448.);
  if (w384 > 4.) return false;
  return ceil(w448) < ceil(w384);
}
/* CTA tile traversal (tmb_tile_site, tmb_geom.h): whole-lattice launches of the 128-site CTAs of the one-field kernels */
static int hop_tile_ok(const tmb_hop_launch &a, bool whole) {
  return C.tile && whole && a.nfl != 2 && a.xblock == 0 && tmb_tile_ok(a.g) && tmb_hop_block(a) == 128;
}
static int hop(int ieo, void *out, const void *in, const HopOpt &o) {
  if (!C.gauge_loaded) return fail(-9, "no gauge field on the device: call tmb_gauge_upload first");
  if (C.kappa == 0.) return fail(-9, "hopping parameter not set: call tmb_set_boundary first");
  tmb_hop_launch a;
  memset(&a, 0, sizeof(a));
  a.prec = o.prec;
  a.recon12 = C.compression == 12;
  a.in = in; a.out = out; a.p = o.p; a.dotw = o.dotw;
  a.cg_x = o.cg_x; a.cg_r = o.cg_r; a.cg_p = o.cg_p;
  a.nfl = o.nfl; a.in1 = o.in1; a.out1 = o.out1; a.p1 = o.p1;
  a.nd_mu = o.nd_mu; a.nd_eps = o.nd_eps; a.nd_scale = o.nd_scale; a.dot_scale = o.dot_scale;
  a.halo_up = C.halo_up; a.halo_dn = C.halo_dn;
  if (C.zsplit && o.prec) TRY(ensure_gauge32()); /* the z fix-up reads the full float links */
  if (a.recon12) {
    TRY(ensure_gauge12(o.prec));
    a.U = o.prec ? (const void *)C.U12f : (const void *)C.U12;
    a.Uhalo = o.prec ? (const void *)C.Uhalo12f : (const void *)C.Uhalo12;
  } else {
    if (o.prec) TRY(ensure_gauge32());
    a.U = o.prec ? (const void *)C.U32 : (const void *)C.U;
    a.Uhalo = o.prec ? (const void *)C.Uhalo32 : (const void *)C.Uhalo;
  }
  a.partial = C.partial; a.st = o.st; a.g = C.g; a.par = ieo ? 1 : 0;
  for (int m = 0; m < 4; m++) a.ka[m] = C.ka[m];
  a.cf = o.cf; a.mode = o.mode; a.dot = o.selfnorm ? 2 : (o.dotw ? 1 : 0); a.hints = eff_hints();
  a.pdl = C.pdl; a.prefetch = C.prefetch; a.prefetch_dist = 0;
  /* the two flavour groups of the NFL = 2 kernel read the same links: let them allocate in L1 (32^3x64: 1.86 -> 1.81 ms) */
  if (o.nfl == 2 && a.hints == 1 && !o.prec && !a.recon12) a.hints = 5;
  a.st_fin = C.st; a.partial_base = C.partial; a.fin_op = (a.dot && fuse_fin()) ? o.fin_op : -1; a.fin_slot = o.fin_slot;
  a.xr = a.fin_op >= 0 ? xr_tab() : nullptr;
  /* Z split: the faces of `in` go to the z neighbours while the kernels below run on the slab as if it were periodic in z;
   * the fix-up at the end replaces the wrapped z term of the face sites by the halo term (tmb_site.cuh) */
  const bool zs = C.zsplit && !o.nocom;
  unsigned int zseq = 0;
  if (zs) {
    if (a.dot || o.mode == 4 || o.nfl == 2 || o.nsites >= 0 || o.boundary_only)
      return fail(-12, "fused reductions, the CG tail, the two-flavour kernel and site sub-ranges are not available with a split Z direction");
    const size_t fb = (size_t)6 * SZ() * (o.prec ? sizeof(float2) : sizeof(double2));
    CU(cudaEventRecord(C.ev_in, C.s_main));
    CU(cudaStreamWaitEvent(C.s_comm, C.ev_in, 0));
    if (C.zpeer) { /* faces straight into the z neighbours' halo buffers of parity seq & 1, then their flags */
      zseq = ++C.zhop_off;
      KL(tmb_launch_pack_zfaces_push(o.prec, C.zdst_up[0], C.zdst_up[1], C.zdst_dn[0], C.zdst_dn[1], C.zsend_up, C.zsend_dn, in, C.g, 1 - a.par, C.seq_dev + 2, zseq,
                                     C.zflag_at_up, C.zflag_at_dn, C.s_comm));
      C.launches++; /* the flag kernel */
    } else {
      KL(tmb_launch_pack_zfaces(o.prec, C.zsend_up, C.zsend_dn, in, C.g, 1 - a.par, C.s_comm));
      TRY(exchange_zfaces(C.zsend_up, C.zsend_dn, C.zhalo_up, C.zhalo_dn, fb, C.s_comm));
    }
    CU(cudaEventRecord(C.ev_z, C.s_comm));
  }
  int np = 0;
  if (!C.dist || o.nocom) {
    a.dist = 0; a.site0 = o.site0; a.nsites = o.nsites < 0 ? C.g.Vh : o.nsites; a.split = a.nsites; a.gap = 0;
    /* tuning variants exist for the plain Hopping_Matrix kernel only */
    a.variant = (o.mode == 0 && !a.dot && !a.recon12 && !o.prec) ? C.hop_variant : 0;
    if (C.hop_variant == 10 || C.hop_variant == -1) /* 448-thread residency: every epilogue of the plain double kernel */
      a.variant = (!a.recon12 && !o.prec && (C.hop_variant == 10 || hop_residency_448(a.nsites))) ? 10 : 0;
    if (o.nfl == 2) a.variant = 0;
    a.xblock = o.nsites < 0 ? C.xblock : 0;
    a.tile = hop_tile_ok(a, o.nsites < 0);
    if (!a.tile && !a.xblock) a.prefetch_dist = C.prefetch_dist;
    np = tmb_hop_grid(a);
    if (np > C.npartial) return fail(-10, "partial buffer too small (%d > %d)", np, C.npartial);
    a.fin_total = np;
    KL(tmb_launch_hop(a, C.s_main));
  } else if (C.p2p && o.nsites < 0 && !o.boundary_only && (C.nranks == 1 || (in_arena(in) && (o.nfl == 1 || in_arena(o.in1))))) {
    /* peer mode: one launch; boundary slices read the neighbours' copies of `in` over NVLink */
    const size_t off = C.nranks == 1 ? 0 : (size_t)((const char *)in - C.arena);
    if (o.nfl == 2) {
      const size_t off1 = C.nranks == 1 ? 0 : (size_t)((const char *)o.in1 - C.arena);
      const bool loc = C.nranks == 1 || (C.p2p_diag & 1);
      a.in_up1 = loc ? o.in1 : (const void *)(C.up_base + off1);
      a.in_dn1 = loc ? o.in1 : (const void *)(C.dn_base + off1);
    }
    a.dist = 2; a.site0 = 0; a.nsites = C.g.Vh; a.split = a.nsites; a.gap = 0; a.variant = 0; a.xblock = 0;
    const bool local = C.nranks == 1 || (C.p2p_diag & 1);
    a.in_up = local ? in : (const void *)(C.up_base + off);
    a.in_dn = local ? in : (const void *)(C.dn_base + off);
    a.p2p_nohandshake = (C.p2p_diag & 2) ? 1 : 0; a.p2p_diag = C.p2p_diag;
    a.seq_base = C.seq_dev; a.seq_off = ++C.hop_off; a.flags = C.flags; a.up_flags = C.up_flags; a.dn_flags = C.dn_flags;
    a.p2p_err = C.p2p_err; a.p2p_copied = C.p2p_ticket;
    a.halo_up_w = C.halo_up; a.halo_dn_w = C.halo_dn;
    a.p2p_copy_ctas = C.p2p_copy_ctas;
    a.tile = hop_tile_ok(a, true);
    np = tmb_hop_grid(a);
    if (np > C.npartial) return fail(-10, "partial buffer too small (%d > %d)", np, C.npartial);
    a.fin_total = np;
    KL(tmb_launch_hop(a, C.s_main));
  } else {
    if (o.nsites >= 0) return fail(-12, "site sub-ranges are not supported with a distributed T direction");
    /* halo exchange on the comm stream, overlapped with the interior kernel:
     * replaces xchange_field(k, ieo) at operator/Hopping_Matrix.c:141-143 */
    const int S = C.g.S, Vh = C.g.Vh;
    CU(cudaEventRecord(C.ev_in, C.s_main));
    CU(cudaStreamWaitEvent(C.s_comm, C.ev_in, 0));
    const size_t face = (size_t)6 * S * (o.prec ? sizeof(float2) : sizeof(double2));
    KL(tmb_launch_pack_halo(o.prec, C.send_up, C.send_dn, in, C.g, C.s_comm));
    if (o.nfl == 2) /* the second flavour's faces behind the first's: one exchange for both */
      KL(tmb_launch_pack_halo(o.prec, (char *)C.send_up + face, (char *)C.send_dn + face, o.in1, C.g, C.s_comm));
    TRY(exchange_faces(C.send_up, C.send_dn, C.halo_up, C.halo_dn, o.nfl * face, C.s_comm));
    a.variant = 0; a.xblock = 0;
    /* the interior grid size fixes where the boundary launch puts its fused-dot partials */
    int nb_int = 0;
    tmb_hop_launch ai = a;
    if (Vh > 2 * S) { /* interior time-slices t in [1, T-2]: no halo data needed */
      ai.dist = 0; ai.site0 = S; ai.nsites = Vh - 2 * S; ai.split = ai.nsites; ai.gap = 0;
      nb_int = tmb_hop_grid(ai);
    }
    /* boundary slices t = 0 and t = T-1: launched on the HIGH-PRIORITY comm stream right behind the
     * exchange, so their CTAs interleave with the interior kernel's instead of forming a second,
     * poorly filled launch after it (outputs are disjoint sites) */
    a.dist = 1; a.site0 = 0; a.nsites = 2 * S; a.split = S; a.gap = Vh - 2 * S;
    a.partial = C.partial + nb_int;
    a.fin_total = ai.fin_total = nb_int + tmb_hop_grid(a);
    if (a.fin_total > C.npartial) return fail(-10, "partial buffer too small (%d > %d)", a.fin_total, C.npartial);
    KL(tmb_launch_hop(a, C.s_comm));
    CU(cudaEventRecord(C.ev_halo, C.s_comm));
    if (nb_int > 0 && !o.boundary_only) KL(tmb_launch_hop(ai, C.s_main));
    CU(cudaStreamWaitEvent(C.s_main, C.ev_halo, 0));
    np = nb_int + tmb_hop_grid(a);
  }
  if (zs) {
    CU(cudaStreamWaitEvent(C.s_main, C.ev_z, 0));
    tmb_zpeer zw; memset(&zw, 0, sizeof(zw));
    zw.own_up = C.zsend_up; zw.own_dn = C.zsend_dn; zw.Uzl = o.prec ? (const void *)C.Uzl32 : (const void *)C.Uzl;
    if (C.zpeer) { zw.hz_up1 = C.zloc_up[1]; zw.hz_dn1 = C.zloc_dn[1]; zw.flags = C.zflags; zw.seq_base = C.seq_dev + 2; zw.seq_off = zseq; zw.err = C.p2p_err; }
    KL(tmb_launch_zfix(o.prec, o.mode, out, in, o.prec ? (const void *)C.U32 : (const void *)C.U,
                       C.zpeer ? (const void *)C.zloc_up[0] : (const void *)C.zhalo_up, C.zpeer ? (const void *)C.zloc_dn[0] : (const void *)C.zhalo_dn,
                       o.prec ? (const void *)C.Uzh32 : (const void *)C.Uzh, C.g, a.par, C.ka[3], o.cf, o.st, &zw, C.s_main));
  }
  if (o.npartial) *o.npartial = np;
  return 0;
}

/* ------------------------------------------------------------------ host-pointer hopping, pipelined
 * Hopping_Matrix(ieo, l, k) with caller-owned HOST buffers is PCIe-bound (192 B/site each way for
 * 1536 B/site of HBM traffic).  Instead of upload -> kernel -> download, the field is cut into
 * chunks of time-slices: chunk c is computed as soon as chunks c-1, c, c+1 have arrived, and its
 * result goes back while later chunks are still coming in, so the H2D and D2H copy engines run
 * concurrently (full duplex) and the kernel time disappears behind them.
 *
 * What is left on top of the two transfers is the fill (the first chunk up before the first kernel) and the drain (what is
 * still to come down when the upload ends); the chunk schedule (tmb_host_chunk_schedule, tmb_geom.h) trades them against the
 * per-copy cost.  A dozen chunks cost more host time to ENQUEUE (copy, events, pack, hop, unpack, copy: ~9 driver calls each)
 * than some of them take to run, so the whole pipeline of one call is captured once as a CUDA graph, keyed on everything baked into it (host pointers, ieo, mode,
 * coefficient, parameter generation), and replayed: the reference calls its operators with the same few field slabs
 * over and over (init/init_spinor_field.c:38-66).
 *
 * With a split T the same pipeline runs on the interior time-slices 1 .. T-2 (no halo needed); the two boundary slices
 * follow once the whole input is on the device and the projected faces have been exchanged. */
struct HostHopGraph {
  const void *k = nullptr; void *l = nullptr; int ieo = 0, mode = 0; double cre = 0., cim = 0.; unsigned long long gen = 0;
  cudaGraphExec_t exec = nullptr; long long launches = 0; unsigned long long used = 0;
};
static std::vector<HostHopGraph> g_hgraphs;
static unsigned long long g_hgraph_clock = 0;
static void hgraphs_clear() {
  for (auto &h : g_hgraphs) if (h.exec) cudaGraphExecDestroy(h.exec);
  g_hgraphs.clear();
}

/* Pageable host memory (the reference's fields are calloc slabs, ALIGN empty unless SSE): cudaMemcpyAsync bounces it through
 * the driver's staging buffers at a fraction of the link rate and without overlap.  The recipe is to page-lock a slab ONCE:
 * explicitly with tmb_host_register(slab, bytes) right after the reference allocated it, or - with
 * TMB_AUTO_PIN=1 in the environment - automatically when a range is seen for the first time (cudaHostRegister, remembered
 * until tmb_finalize).  The automatic form is opt-in because it is only safe when the caller never frees a field while the
 * library is initialised (true for tmLQCD's init_spinor_field / init_solver_field slabs, not for a test harness that
 * allocates and drops buffers: a stale registration of a recycled address range would be used for DMA). */
struct PinnedRange { char *lo, *hi; };
static std::vector<PinnedRange> g_pinned;
static void pinned_clear() {
  for (auto &r : g_pinned) cudaHostUnregister(r.lo);
  g_pinned.clear();
}
static void ensure_pinned(const void *p, size_t bytes) {
  static int enabled = -1;
  if (enabled < 0) { const char *e = getenv("TMB_AUTO_PIN"); enabled = (e && atoi(e) == 1) ? 1 : 0; }
  if (!enabled || !p) return;
  char *lo = (char *)p, *hi = lo + bytes;
  for (auto &r : g_pinned) if (lo >= r.lo && hi <= r.hi) return;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) return; /* already pinned / managed */
  cudaGetLastError();
  const size_t page = 4096;
  lo = (char *)((size_t)lo & ~(page - 1)); hi = (char *)(((size_t)hi + page - 1) & ~(page - 1));
  if (cudaHostRegister(lo, (size_t)(hi - lo), cudaHostRegisterDefault) == cudaSuccess) {
    g_pinned.push_back({lo, hi});
    hgraphs_clear(); /* graphs captured on the pageable pointer hold pageable copy nodes */
  } else {
    cudaGetLastError(); /* overlapping an existing registration, or not allowed: the copies still work, just slower */
  }
}

struct HopTrace { int kind, idx; cudaEvent_t ev; };
static std::vector<HopTrace> g_trace; /* tmb_host_hop_timeline: timing events of one un-captured pipeline run */
static bool g_tracing = false;
static void trace_mark(int kind, int idx, cudaStream_t s) {
  if (!g_tracing) return;
  HopTrace t; t.kind = kind; t.idx = idx; t.ev = nullptr;
  if (cudaEventCreate(&t.ev) != cudaSuccess) return;
  cudaEventRecord(t.ev, s);
  g_trace.push_back(t);
}
static int host_hop_enqueue(int ieo, double *l_host, const double *k_host, int mode, double cre, double cim, double2 *din, double2 *dout) {
  const int T = C.g.T, S = C.g.S, Vh = C.g.Vh;
  /* interior slices [t_lo, t_hi): all of them on one rank, 1 .. T-2 with a split T (slices 0 and T-1 wait for the faces) */
  const int t_lo = C.dist ? 1 : 0, t_hi = C.dist ? T - 1 : T;
  const int nt = t_hi - t_lo;
  /* INPUT chunks cb[0..n] in time-slices; OUTPUT pieces are the same ranges shifted DOWN by one slice:
   *   piece p = [cb[p] - 1, cb[p+1] - 1)   (piece 0 starts at t_lo, one more piece [t_hi - 1, t_hi) closes the lattice)
   * so that piece p needs the inputs of chunks p-1 and p only - its download starts as soon as its own chunk is up, not one
   * chunk later - plus, on a periodic lattice, the wrap-around slice (last chunk, sent second).  What the pipeline adds to
   * the two transfers is then one chunk of upload before the first download and what is left to download when the upload
   * ends; copies get slower the smaller they are once both directions are busy (tmb_host_hop_timeline, 24^3x48: a 10.6 MB
   * upload takes ~255 us = 42 GB/s next to a running download, a 2.6 MB one ~105 us = 25 GB/s, against 55 GB/s alone).
   * Hence few LARGE chunks (a sixth of the field), the first one split 1/4 + 3/4, halving towards the end.
   * tmb_set_host_chunks(n > 0) forces n equal chunks, tmb_set_host_chunk_sizes an explicit schedule.  The upload stream is
   * the critical path: it is busy from the start to ~0.1 ms before the end.  Measured schedules and timelines:
   * profiles/r02_pipe_diag.log, profiles/r02_summary.md (section host-pointer calls). */
  int cb[MAXCHUNK + 1], n = 0;
  cb[0] = t_lo;
  if (nt > 0) {
    int sched_sum = 0;
    for (int i = 0; i < C.host_sched_n; i++) sched_sum += C.host_sched[i];
    if (C.host_sched_n > 0 && sched_sum == nt) {
      int t = t_lo;
      for (int i = 0; i < C.host_sched_n; i++) { t += C.host_sched[i]; cb[++n] = t; }
    } else if (C.host_chunks > 0) {
      const int spc = (nt + C.host_chunks - 1) / C.host_chunks > 0 ? (nt + C.host_chunks - 1) / C.host_chunks : 1;
      for (int t = t_lo; t < t_hi; t += spc) cb[++n] = t + spc < t_hi ? t + spc : t_hi;
    } else {
      int small = 1;
      while ((size_t)small * S * 192 < ((size_t)1 << 20) && small < nt) small++;
      int sizes[MAXCHUNK];
      const int nc = tmb_host_chunk_schedule(nt, small, sizes); /* tmb_geom.h */
      int t = t_lo;
      for (int c = 0; c < nc && n < MAXCHUNK; c++) { t += sizes[c]; cb[++n] = t; }
    }
  }
  if (n > MAXCHUNK - 2) return fail(-13, "too many chunks");
  double2 *in_aos = C.stage, *out_aos = C.stage + (size_t)12 * Vh;
  const double2 *hk = (const double2 *)k_host; double2 *hl = (double2 *)l_host;
  auto up = [&](int t0, int t1, cudaEvent_t ev) -> int { /* input slices [t0, t1) to the device, event when they are there.
    * (Splitting a chunk into two concurrent copies on two streams was tried: 1.99 against 1.66 ms per call.) */
    CU(cudaMemcpyAsync(in_aos + (size_t)t0 * S * 12, hk + (size_t)t0 * S * 12, (size_t)(t1 - t0) * S * 192, cudaMemcpyHostToDevice, C.s_h2d));
    CU(cudaEventRecord(ev, C.s_h2d));
    trace_mark(0, t0, C.s_h2d);
    return 0;
  };
  auto pack = [&](int t0, int t1, cudaEvent_t ev) -> int { /* AoS -> device layout on the compute stream, behind the copy */
    CU(cudaStreamWaitEvent(C.s_main, ev, 0));
    KL(tmb_launch_pack_eo_range(din, in_aos, Vh, t0 * S, (t1 - t0) * S, C.s_main));
    return 0;
  };
  auto down = [&](int t0, int t1, cudaEvent_t ev) -> int { /* device layout -> AoS, then output slices [t0, t1) to the host */
    KL(tmb_launch_unpack_eo_range(out_aos, dout, Vh, t0 * S, (t1 - t0) * S, C.s_main));
    CU(cudaEventRecord(ev, C.s_main));
    trace_mark(1, t0, C.s_main);
    CU(cudaStreamWaitEvent(C.s_d2h, ev, 0));
    CU(cudaMemcpyAsync(hl + (size_t)t0 * S * 12, out_aos + (size_t)t0 * S * 12, (size_t)(t1 - t0) * S * 192, cudaMemcpyDeviceToHost, C.s_d2h));
    trace_mark(2, t0, C.s_d2h);
    return 0;
  };
  /* fork: the copy streams join the work of s_main (also what makes them part of a graph capture) */
  trace_mark(-1, 0, C.s_main);
  CU(cudaEventRecord(C.ev_in, C.s_main));
  CU(cudaStreamWaitEvent(C.s_h2d, C.ev_in, 0));
  CU(cudaStreamWaitEvent(C.s_d2h, C.ev_in, 0));
  /* upload order: with a split T the two boundary slices first of all (their faces have the longest way to go), then chunk 0,
   * the wrap-around chunk (periodic lattice), then the rest in order */
  if (C.dist) {
    TRY(up(0, 1, C.ev_halo));
    CU(cudaMemcpyAsync(in_aos + (size_t)(T - 1) * S * 12, hk + (size_t)(T - 1) * S * 12, (size_t)S * 192, cudaMemcpyHostToDevice, C.s_h2d));
    CU(cudaEventRecord(C.ev_halo, C.s_h2d));
  }
  if (n > 0) TRY(up(cb[0], cb[1], C.ev_up[0]));
  if (n > 1 && !C.dist) TRY(up(cb[n - 1], cb[n], C.ev_up[n - 1]));
  for (int c = 1; c < n - (C.dist ? 0 : 1); c++) TRY(up(cb[c], cb[c + 1], C.ev_up[c]));
  if (C.dist) {
    CU(cudaStreamWaitEvent(C.s_main, C.ev_halo, 0));
    KL(tmb_launch_pack_eo_range(din, in_aos, Vh, 0, S, C.s_main));
    KL(tmb_launch_pack_eo_range(din, in_aos, Vh, (T - 1) * S, S, C.s_main));
  }
  bool packed[MAXCHUNK] = {false};
  for (int p = 0; p <= n && n > 0; p++) {
    if (p < n && !packed[p]) { TRY(pack(cb[p], cb[p + 1], C.ev_up[p])); packed[p] = true; }
    if (p == 0 && n > 1 && !C.dist) { TRY(pack(cb[n - 1], cb[n], C.ev_up[n - 1])); packed[n - 1] = true; } /* the wrap-around slice lives in the last chunk */
    const int lo = p == 0 ? t_lo : cb[p] - 1, hi = p == n ? t_hi : cb[p + 1] - 1;
    if (hi <= lo) continue;
    HopOpt o; o.mode = mode; o.cf = make_double2(cre, cim); o.site0 = lo * S; o.nsites = (hi - lo) * S; o.nocom = true;
    TRY(hop(ieo, dout, din, o));
    TRY(down(lo, hi, C.ev_done[p]));
  }
  if (C.dist) {
    /* boundary slices 0 and T-1: faces projected and exchanged (half-spinors, xchange_halffield's idea), then the halo
     * kernel on those two slices only */
    HopOpt o; o.mode = mode; o.cf = make_double2(cre, cim); o.boundary_only = true;
    TRY(hop(ieo, dout, din, o));
    TRY(down(0, 1, C.ev_done[MAXCHUNK - 2]));
    TRY(down(T - 1, T, C.ev_done[MAXCHUNK - 1]));
  }
  /* join */
  CU(cudaEventRecord(C.ev_chk[0], C.s_d2h));
  CU(cudaEventRecord(C.ev_chk[1], C.s_h2d));
  CU(cudaStreamWaitEvent(C.s_main, C.ev_chk[0], 0));
  CU(cudaStreamWaitEvent(C.s_main, C.ev_chk[1], 0));
  return 0;
}

extern "C" int tmb_Hopping_Matrix_host(int ieo, double *l_host, const double *k_host, int mode, double cre, double cim) {
  NEED_INIT();
  if (mode != 0 && mode != 1) return fail(-13, "tmb_Hopping_Matrix_host: mode must be 0 or 1");
  if (!C.gauge_loaded) return fail(-9, "no gauge field on the device: call tmb_gauge_upload first");
  SCR(din, 12); SCR(dout, 13);
  if (C.zsplit) { /* split Z: plain upload / compute / download */
    TRY(tmb_field_upload(din, k_host));
    HopOpt o; o.mode = mode; o.cf = make_double2(cre, cim);
    TRY(hop(ieo, dout, din, o));
    return tmb_field_download(l_host, dout);
  }
  if (C.compression == 12) TRY(ensure_gauge12(0)); /* nothing may allocate or synchronise inside a capture */
  ensure_pinned(k_host, FIELD_BYTES()); ensure_pinned(l_host, FIELD_BYTES());
  CU(cudaStreamSynchronize(C.s_main));
  /* a split T exchanges faces through NCCL inside the pipeline: not captured (NCCL calls stay out of graphs here) */
  const bool use_graph = C.cg_graph && !C.dist;
  auto enqueue = [&]() { return host_hop_enqueue(ieo, l_host, k_host, mode, cre, cim, din, dout); };
  if (!use_graph) {
    TRY(enqueue());
    CU(cudaStreamSynchronize(C.s_main));
    return 0;
  }
  HostHopGraph *hit = nullptr;
  for (auto &h : g_hgraphs)
    if (h.k == k_host && h.l == l_host && h.ieo == ieo && h.mode == mode && h.cre == cre && h.cim == cim && h.gen == C.param_gen) { hit = &h; break; }
  if (!hit) {
    if (g_hgraphs.size() >= 32) { /* forget the least recently used one */
      size_t lru = 0;
      for (size_t i = 1; i < g_hgraphs.size(); i++) if (g_hgraphs[i].used < g_hgraphs[lru].used) lru = i;
      cudaGraphExecDestroy(g_hgraphs[lru].exec);
      g_hgraphs.erase(g_hgraphs.begin() + lru);
    }
    const long long l0 = C.launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(C.s_main, cudaStreamCaptureModeRelaxed) != cudaSuccess) return fail(-100, "cudaStreamBeginCapture failed");
    const int rc = enqueue();
    cudaError_t e = cudaStreamEndCapture(C.s_main, &graph);
    HostHopGraph h;
    h.launches = C.launches - l0; C.launches = l0;
    if (rc < 0) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(-100, "capture of the host-pointer hop failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(-100, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    h.k = k_host; h.l = l_host; h.ieo = ieo; h.mode = mode; h.cre = cre; h.cim = cim; h.gen = C.param_gen;
    g_hgraphs.push_back(h);
    hit = &g_hgraphs.back();
  }
  hit->used = ++g_hgraph_clock;
  if (cudaGraphLaunch(hit->exec, C.s_main) != cudaSuccess) return fail(-101, "cudaGraphLaunch failed");
  C.launches += hit->launches;
  CU(cudaStreamSynchronize(C.s_main));
  return 0;
}

/* Diagnostic: one un-captured run of the pipeline with a timing event behind every upload (kind 0), every piece's kernels
 * (kind 1) and every download (kind 2).  out[3 i ..] = (kind, first time-slice, microseconds since the start); returns the
 * number of rows (at most max_rows), the last one (kind 3) being the end of the call. */
extern "C" int tmb_host_hop_timeline(int ieo, double *l_host, const double *k_host, double *out, int max_rows) {
  NEED_INIT();
  if (!C.gauge_loaded) return fail(-9, "no gauge field on the device: call tmb_gauge_upload first");
  if (C.zsplit || C.dist) return fail(-12, "timeline: one rank only");
  SCR(din, 12); SCR(dout, 13);
  ensure_pinned(k_host, FIELD_BYTES()); ensure_pinned(l_host, FIELD_BYTES());
  CU(cudaStreamSynchronize(C.s_main));
  g_trace.clear(); g_tracing = true;
  const int rc = host_hop_enqueue(ieo, l_host, k_host, 0, 0., 0., din, dout);
  trace_mark(3, 0, C.s_main);
  g_tracing = false;
  cudaStreamSynchronize(C.s_main);
  int rows = 0;
  cudaEvent_t start = nullptr;
  for (auto &t : g_trace) if (t.kind == -1) start = t.ev;
  for (auto &t : g_trace) {
    if (t.kind >= 0 && start && rows < max_rows) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, start, t.ev);
      out[3 * rows] = t.kind; out[3 * rows + 1] = t.idx; out[3 * rows + 2] = 1e3 * ms; rows++;
    }
  }
  for (auto &t : g_trace) cudaEventDestroy(t.ev);
  g_trace.clear();
  return rc < 0 ? rc : rows;
}

/* ------------------------------------------------------------------ operators on device fields */
extern "C" int tmb_Hopping_Matrix(int ieo, void *l, const void *k) {
  NEED_INIT(); HopOpt o; return hop(ieo, F(l), F(k), o);
}
/* Hopping_Matrix_nocom (operator/Hopping_Matrix_nocom.c): the same arithmetic without the exchange, which the reference
 * times against Hopping_Matrix to isolate the communication (benchmark.c:337-373).  Without xchange_field the reference
 * reads whatever its halo region holds; here the +-t neighbours of the boundary slices are the slab's own opposite
 * boundary (plain single-GPU kernel on the local slab).  On one rank it IS Hopping_Matrix. */
extern "C" int tmb_Hopping_Matrix_nocom(int ieo, void *l, const void *k) {
  NEED_INIT(); HopOpt o; o.nocom = true; return hop(ieo, F(l), F(k), o);
}
extern "C" int tmb_tm_times_Hopping_Matrix(int ieo, void *l, const void *k, double cre, double cim) {
  NEED_INIT(); HopOpt o; o.mode = 1; o.cf = make_double2(cre, cim); return hop(ieo, F(l), F(k), o);
}
extern "C" int tmb_tm_sub_Hopping_Matrix(int ieo, void *l, const void *p, const void *k, double cre, double cim) {
  NEED_INIT(); HopOpt o; o.mode = 2; o.cf = make_double2(cre, cim); o.p = F(p); return hop(ieo, F(l), F(k), o);
}
/* z of H_eo_tm_inv_psi (tm_operators.c:514-521): (1 -+ i mu)/(1+mu^2) */
static double2 z_inv(double sign) {
  const double nrm = 1. / (1. + C.mu * C.mu), sg = sign < 0. ? 1. : -1.;
  return make_double2(nrm, sg * nrm * C.mu);
}
/* z of tm_sub_H_eo_gamma5 / mul_one_pm_imu (tm_operators.c:533-541): 1 +- i mu */
static double2 z_fwd(double sign) { return make_double2(1., (sign < 0. ? -1. : 1.) * C.mu); }

extern "C" int tmb_H_eo_tm_inv_psi(void *l, const void *k, int ieo, double sign) {
  NEED_INIT(); HopOpt o; o.mode = 1; o.cf = z_inv(sign); return hop(ieo, F(l), F(k), o);
}
extern "C" int tmb_tm_sub_H_eo_gamma5(void *l, const void *p, const void *k, int ieo, double sign) {
  NEED_INIT(); HopOpt o; o.mode = 2; o.cf = z_fwd(sign); o.p = F(p); return hop(ieo, F(l), F(k), o);
}

/* Qtm_pm_psi (tm_operators.c:338-345): 4 hops, each with its diagonal fused in the epilogue.
 * dotw/st: the CG asks for <dotw, l> and the early exit.  For dotw == k (the CG's <p, A p>, cg_her.c:93) the
 * reduction moves to the SECOND hop as |Q- k|^2: Q+ is the adjoint of Q-, so <k, Q+ Q- k> = |Q- k|^2, and the
 * squared norm of a hop's own output needs no operand - 192 B/site and the widest epilogue of the CG saved. */
static int qtm_pm(double2 *l, const double2 *k, const double2 *dotw, const tmb_cg_state *st, int *np,
                  int fin_op = -1, int fin_slot = 0) {
  SCR(w0, 0); SCR(w1, 1);
  if (C.zsplit && dotw != nullptr) return fail(-12, "no fused reduction with a split Z direction");
  const bool self = dotw != nullptr && dotw == k && C.cg_selfnorm;
  HopOpt a; a.mode = 1; a.cf = z_inv(-1.); a.st = st;
  TRY(hop(0, w1, k, a));
  HopOpt b; b.mode = 2; b.cf = z_fwd(-1.); b.p = k; b.st = st;
  if (self) { b.selfnorm = true; b.npartial = np; b.fin_op = fin_op; b.fin_slot = fin_slot; }
  TRY(hop(1, w0, w1, b));
  HopOpt c; c.mode = 1; c.cf = z_inv(+1.); c.st = st;
  TRY(hop(0, w1, w0, c));
  HopOpt d; d.mode = 2; d.cf = z_fwd(+1.); d.p = w0; d.st = st;
  if (!self) { d.dotw = dotw; d.npartial = np; d.fin_op = fin_op; d.fin_slot = fin_slot; }
  TRY(hop(1, l, w1, d));
  return 0;
}
extern "C" int tmb_Qtm_pm_psi(void *l, const void *k) { NEED_INIT(); return qtm_pm(F(l), F(k), nullptr, nullptr, nullptr); }

/* One CG iteration's operator part with the vector updates folded in (cg_her.c:92-101), both precisions:
 *   hop 1, hop 2 (+ |Q- p|^2 = <p, A p>, finish: alpha), hop 3, hop 4 whose epilogue does x += alpha p, r -= alpha A p, |r|^2
 * A p itself is never written.  4 launches; the fourth leaves the |r|^2 partials (finished in place when fuse_fin()). */
static float2 *scratch32(int k);
static int reduce_to(int npart, int slot, int op);
static int qtm_pm_cg(int prec, void *x, void *r, const void *p, int err_op, int *np_err) {
  void *w0, *w1;
  if (prec) { w0 = scratch32(0); w1 = scratch32(1); } else { w0 = scratch(0); w1 = scratch(1); }
  if (!w0 || !w1) return -100;
  const bool fuse = fuse_fin();
  int np = 0;
  HopOpt a; a.prec = prec; a.mode = 1; a.cf = z_inv(-1.); a.st = C.st;
  TRY(hop(0, w1, p, a));
  HopOpt b; b.prec = prec; b.mode = 2; b.cf = z_fwd(-1.); b.p = p; b.st = C.st;
  /* <p, A p> = |Q- p|^2 comes out of the second hop as per-CTA partials; alpha is needed by the FOURTH hop only.  Default:
   * finished by the last CTA of the second hop (fence + ticket per CTA; ncu: 103.3 against 89.4 us for that launch at
   * 24^3x48).  Experiment (tmb_set_overlap bit 6): the finish as a one-CTA kernel on the side stream WHILE the third hop
   * runs.  Same sums, bit-identical solutions, but the two cross-stream dependencies cost more than the tail they remove:
   * 0.475 against 0.459 ms per iteration at 24^3x48, 0.134 against 0.118 at 16^3x32 (profiles/r02_cg_side_ab.log). */
  const bool side = fuse && C.cg_side;
  b.selfnorm = true; b.npartial = &np; b.fin_op = (fuse && !side) ? TMB_FIN_CG_PRO : -1; b.fin_slot = 1;
  TRY(hop(1, w0, w1, b));
  if (!fuse) TRY(reduce_to(np, 1, TMB_FIN_CG_PRO)); /* alpha must exist before the fourth hop starts */
  if (side) {
    CU(cudaEventRecord(C.ev_side[0], C.s_main));
    CU(cudaStreamWaitEvent(C.s_comm, C.ev_side[0], 0));
    KL(tmb_launch_final_hop(C.partial, np, C.st, 1, TMB_FIN_CG_PRO, 1, xr_tab(), C.s_comm));
    CU(cudaEventRecord(C.ev_side[1], C.s_comm));
  }
  HopOpt c; c.prec = prec; c.mode = 1; c.cf = z_inv(+1.); c.st = C.st;
  TRY(hop(0, w1, w0, c));
  if (side) CU(cudaStreamWaitEvent(C.s_main, C.ev_side[1], 0));
  HopOpt d; d.prec = prec; d.mode = 4; d.cf = z_fwd(+1.); d.p = w0; d.st = C.st;
  d.cg_x = x; d.cg_r = r; d.cg_p = p;
  d.selfnorm = true; d.npartial = np_err; d.fin_op = fuse ? err_op : -1; d.fin_slot = 2;
  TRY(hop(1, w1 /* unused: mode 4 writes x and r */, w1, d));
  return 0;
}

/* Q_+- = g5[(1 +- i mu g5) k - H_oe (1 +- i mu g5)^-1 H_eo k]  (tm_operators.c:172-177, :216-221);
 * l may alias k: k is only read site-locally by the last kernel */
static int qtm_pm_single(double2 *l, const double2 *k, double sign, int g5) {
  SCR(w1, 1);
  HopOpt a; a.mode = 1; a.cf = z_inv(sign);
  TRY(hop(0, w1, k, a));
  HopOpt b; b.mode = g5 ? 2 : 3; b.cf = z_fwd(sign); b.p = k;
  TRY(hop(1, l, w1, b));
  return 0;
}
extern "C" int tmb_Qtm_plus_psi(void *l, const void *k) { NEED_INIT(); return qtm_pm_single(F(l), F(k), +1., 1); }
extern "C" int tmb_Qtm_minus_psi(void *l, const void *k) { NEED_INIT(); return qtm_pm_single(F(l), F(k), -1., 1); }
extern "C" int tmb_Mtm_plus_psi(void *l, const void *k) { NEED_INIT(); return qtm_pm_single(F(l), F(k), +1., 0); }
extern "C" int tmb_Mtm_minus_psi(void *l, const void *k) { NEED_INIT(); return qtm_pm_single(F(l), F(k), -1., 0); }

/* M_full (tm_operators.c:117-128): En = (1+i mu g5)E - H_eo O ; On = (1+i mu g5)O - H_oe E */
static int m_full(double2 *en, double2 *on, const double2 *e, const double2 *o, int g5) {
  HopOpt a; a.mode = g5 ? 2 : 3; a.cf = z_fwd(+1.); a.p = e;
  TRY(hop(0, en, o, a));
  HopOpt b; b.mode = g5 ? 2 : 3; b.cf = z_fwd(+1.); b.p = o;
  TRY(hop(1, on, e, b));
  return 0;
}
extern "C" int tmb_M_full(void *en, void *on, const void *e, const void *o) {
  NEED_INIT();
  if (en == e || en == o || on == e || on == o) return fail(-11, "tmb_M_full: output must not alias input");
  return m_full(F(en), F(on), F(e), F(o), 0);
}
extern "C" int tmb_Q_full(void *en, void *on, const void *e, const void *o) {
  NEED_INIT();
  if (en == e || en == o || on == e || on == o) return fail(-11, "tmb_Q_full: output must not alias input");
  return m_full(F(en), F(on), F(e), F(o), 1);
}
/* D_psi (D_psi_body.c:266-375) = M_full on the eo-split field: phase_mu = -ka_mu gives the minus sign */
extern "C" int tmb_D_psi_eo(void *en, void *on, const void *e, const void *o) { return tmb_M_full(en, on, e, o); }

extern "C" int tmb_assign_mul_one_pm_imu_inv(void *l, const void *k, double sign) {
  NEED_INIT(); KL(tmb_launch_diag(F(l), F(k), z_inv(sign), N2(), HALF(), C.s_main)); return 0;
}
extern "C" int tmb_assign_mul_one_pm_imu(void *l, const void *k, double sign) {
  NEED_INIT(); KL(tmb_launch_diag(F(l), F(k), z_fwd(sign), N2(), HALF(), C.s_main)); return 0;
}
extern "C" int tmb_mul_one_pm_imu_sub_mul_gamma5(void *l, const void *k, const void *j, double sign) {
  NEED_INIT(); KL(tmb_launch_diag_sub(F(l), F(k), F(j), z_fwd(sign), 1, N2(), HALF(), C.s_main)); return 0;
}
extern "C" int tmb_mul_one_pm_imu_sub_mul(void *l, const void *k, const void *j, double sign) {
  NEED_INIT(); KL(tmb_launch_diag_sub(F(l), F(k), F(j), z_fwd(sign), 0, N2(), HALF(), C.s_main)); return 0;
}
/* generic forms: l = (z on s0,s1 | conj z on s2,s3) k   and   l = [g5]( (z | conj z) k - j ) */
extern "C" int tmb_diag(void *l, const void *k, double zre, double zim) {
  NEED_INIT(); KL(tmb_launch_diag(F(l), F(k), make_double2(zre, zim), N2(), HALF(), C.s_main)); return 0;
}
extern "C" int tmb_diag_sub(void *l, const void *k, const void *j, double zre, double zim, int g5) {
  NEED_INIT(); KL(tmb_launch_diag_sub(F(l), F(k), F(j), make_double2(zre, zim), g5 ? 1 : 0, N2(), HALF(), C.s_main)); return 0;
}
extern "C" int tmb_gamma5(void *l, const void *k) { NEED_INIT(); KL(tmb_launch_gamma5(F(l), F(k), N2(), HALF(), C.s_main)); return 0; }

/* ------------------------------------------------------------------ BLAS-1 */
static int finish_reduction(int npart, double *result) {
  KL(tmb_launch_final(C.partial, npart, C.st, 0, TMB_FIN_STORE, 0, xr_tab(), C.s_main));
  if (!C.xred) TRY(allreduce_slot(0));
  CU(cudaMemcpyAsync(result, &C.st->tmp[0], sizeof(double), cudaMemcpyDeviceToHost, C.s_main));
  CU(cudaStreamSynchronize(C.s_main));
  return 0;
}
extern "C" int tmb_square_norm(const void *p, double *result) {
  NEED_INIT(); KL(tmb_launch_norm2(0, F(p), N2(), C.partial, C.s_main)); return finish_reduction(tmb_red_grid(N2()), result);
}
extern "C" int tmb_scalar_prod_r(const void *s, const void *r, double *result) {
  NEED_INIT(); KL(tmb_launch_dot(0, F(s), F(r), N2(), C.partial, C.s_main)); return finish_reduction(tmb_red_grid(N2()), result);
}
extern "C" int tmb_assign_mul_add_r_and_square(void *r, double c, const void *s, double *result) {
  NEED_INIT(); KL(tmb_launch_xpay_norm(F(r), c, F(s), N2(), C.partial, C.s_main)); return finish_reduction(tmb_red_grid(N2()), result);
}
extern "C" int tmb_assign_add_mul_r(void *p, const void *q, double c) { NEED_INIT(); KL(tmb_launch_axpy(F(p), F(q), c, N2(), C.s_main)); return 0; }
extern "C" int tmb_assign_mul_add_r(void *r, double c, const void *s) { NEED_INIT(); KL(tmb_launch_xpay(F(r), c, F(s), N2(), C.s_main)); return 0; }
extern "C" int tmb_diff(void *q, const void *r, const void *s) { NEED_INIT(); KL(tmb_launch_lincomb(F(q), 1., F(r), -1., F(s), N2(), C.s_main)); return 0; }
extern "C" int tmb_add(void *q, const void *r, const void *s) { NEED_INIT(); KL(tmb_launch_lincomb(F(q), 1., F(r), 1., F(s), N2(), C.s_main)); return 0; }
extern "C" int tmb_assign(void *r, const void *s) {
  NEED_INIT();
  if (r != s) CU(cudaMemcpyAsync(r, s, FIELD_BYTES(), cudaMemcpyDeviceToDevice, C.s_main));
  return 0;
}
extern "C" int tmb_mul_r(void *r, double c, const void *s) { NEED_INIT(); KL(tmb_launch_scale(F(r), c, F(s), N2(), C.s_main)); return 0; }

/* ------------------------------------------------------------------ CG, solver/cg_her.c:62-143
 * Same recurrence and stopping rule as the reference.  What changes is where things live:
 * p, r, Ap, x and the scalars normsq/pro/alpha/err/beta stay in HBM; <p,Ap> is accumulated in the
 * epilogue of the 4th hop of Qtm_pm_psi; x += alpha p, r -= alpha Ap and |r|^2 are one sweep;
 * the stop test runs on the device and turns every later kernel into a no-op, so the host
 * enqueues iterations in chunks without a synchronisation per iteration. */
#define CG_CHUNK 8
static int reduce_to(int npart, int slot, int op) {
  if (!fuse_fin()) { /* NCCL fallback: partial sum, all-reduce of one double, bookkeeping */
    KL(tmb_launch_final(C.partial, npart, C.st, slot, op, 0, nullptr, C.s_main));
    TRY(allreduce_slot(slot));
    KL(tmb_launch_apply(C.st, slot, op, C.s_main));
  } else {
    KL(tmb_launch_final(C.partial, npart, C.st, slot, op, 1, xr_tab(), C.s_main));
  }
  return 0;
}

static int cg_drive(int prec, void *x, void *r, void *p, void *ap, int err_op, int trips, tmb_cg_state *hout);
extern "C" int tmb_cg_her(void *P, const void *Q, int max_iter, double eps_sq, int rel_prec) {
  NEED_INIT();
  SCR(ap, 2); SCR(r, 3); SCR(p, 4);
  double2 *x = F(P); const double2 *q = F(Q);
  const size_t n2 = N2();
  auto t0 = std::chrono::steady_clock::now();
  /* squarenorm = |Q|^2 (cg_her.c:82) */
  double sqn = 0.;
  TRY(tmb_square_norm(Q, &sqn));
  tmb_cg_state h;
  memset(&h, 0, sizeof(h));
  h.sqnorm_q = sqn; h.eps_sq = eps_sq; h.rel_prec = rel_prec;
  CU(cudaMemcpyAsync(C.st, &h, sizeof(h), cudaMemcpyHostToDevice, C.s_main));
  CU(cudaStreamSynchronize(C.s_main)); /* h is on the stack */
  /* r = Q - A x ; p = r ; normsq = |r|^2 (cg_her.c:84-88) */
  TRY(qtm_pm(ap, x, nullptr, nullptr, nullptr));
  KL(tmb_launch_lincomb(r, 1., q, -1., ap, n2, C.s_main));
  CU(cudaMemcpyAsync(p, r, FIELD_BYTES(), cudaMemcpyDeviceToDevice, C.s_main));
  KL(tmb_launch_norm2(0, r, n2, C.partial, C.s_main));
  TRY(reduce_to(tmb_red_grid(n2), 0, TMB_FIN_CG_INIT));

  /* main loop (cg_her.c:91-127), enqueued in chunks / as a replayed CUDA graph: see cg_drive() */
  TRY(cg_drive(0, x, r, p, ap, TMB_FIN_CG_ERR, max_iter, &h));
  auto t1 = std::chrono::steady_clock::now();
  C.last_iters = h.iter; C.last_err = h.err;
  C.last_seconds = std::chrono::duration<double>(t1 - t0).count();
  if (!h.converged) return -1; /* cg_her.c:141 */
  return h.iter;
}

extern "C" int tmb_solver_stats(int *iterations, double *final_err, double *seconds) {
  if (iterations) *iterations = C.last_iters;
  if (final_err) *final_err = C.last_err;
  if (seconds) *seconds = C.last_seconds;
  return 0;
}

/* invert_eo, CG branch (invert_eo.c:152-157, :252, :268-270, :306-310) */
extern "C" int tmb_invert_eo(void *even_new, void *odd_new, const void *even, const void *odd, double precision,
                             int max_iter, int rel_prec) {
  NEED_INIT();
  SCR(d, 5);
  double2 *En = F(even_new), *On = F(odd_new);
  const size_t n2 = N2();
  /* Even_new = Mee^-1 Even */
  KL(tmb_launch_diag(En, F(even), z_inv(+1.), n2, HALF(), C.s_main));
  /* DUM_DERI = g5 (Odd + H_oe Even_new): MODE 2 with cf = -1 gives g5(-Odd - H En); so do it in two steps */
  HopOpt a; TRY(hop(1, d, En, a));
  KL(tmb_launch_xpay(d, 1., F(odd), n2, C.s_main));
  KL(tmb_launch_gamma5(d, d, n2, HALF(), C.s_main));
  const int iter = tmb_cg_her(odd_new, d, max_iter, precision, rel_prec);
  if (iter < -1) return iter;
  TRY(qtm_pm_single(On, On, -1., 1));            /* Qtm_minus_psi(Odd_new, Odd_new) */
  /* Even_new += Mee^-1 H_eo Odd_new */
  HopOpt b; b.mode = 1; b.cf = z_inv(+1.);
  TRY(hop(0, d, On, b));
  KL(tmb_launch_axpy(En, d, 1., n2, C.s_main));
  CU(cudaStreamSynchronize(C.s_main));
  return iter;
}

/* ------------------------------------------------------------------ ND doublet, tm_operators_nd.c */
extern "C" int tmb_M_ee_inv_ndpsi(void *ls, void *lc, const void *ks, const void *kc, double mu, double eps) {
  NEED_INIT(); KL(tmb_launch_nd_mee_inv(F(ls), F(lc), F(ks), F(kc), mu, eps, N2(), HALF(), C.s_main)); return 0;
}
static int nd_moo(double2 *ls, double2 *lc, const double2 *ks, const double2 *kc, const double2 *js, const double2 *jc,
                  double mu, double eps) {
  KL(tmb_launch_nd_moo_sub_g5(ls, lc, ks, kc, js, jc, mu, eps, N2(), HALF(), C.s_main)); return 0;
}
extern "C" int tmb_M_oo_sub_g5_ndpsi(void *ls, void *lc, const void *ks, const void *kc, const void *js, const void *jc,
                                     double mu, double eps) {
  NEED_INIT(); return nd_moo(F(ls), F(lc), F(ks), F(kc), F(js), F(jc), mu, eps);
}
static int hop0(int ieo, double2 *l, const double2 *k) { HopOpt o; return hop(ieo, l, k, o); }

struct NdDot;
static int hop2_legacy(int ieo, void *o0, void *o1, const void *i0, const void *i1, int mode, const void *p0,
                       const void *p1, double mu, double eps, double scale, int prec, const tmb_cg_state *st, const NdDot *dot);
/* Two-flavour fused path: every Hopping_Matrix pair of tm_operators_nd.c is ONE
 * launch that streams the links once for both flavours, with M_ee_inv_ndpsi / M_oo_sub_g5_ndpsi and the
 * phmc_invmaxev scaling in its epilogue.  mode 1: out = M_ee_inv_nd(H in0, H in1); mode 2: out = scale g5(M_oo(p) - H in). */
/* Which two-flavour kernel: the NFL = 2 instantiation of hop_kernel serves every precision, compression and communication
 * mode; round 1's one-thread-two-flavours kernel (tmb_force.cu) is 13 % faster where it applies - one rank, 18-real links,
 * double (32^3x64: 1.60 against 1.84 ms per Qtm_pm_ndpsi, profiles/r02_section_nd_first.json) - and is taken there. */
static bool nd_legacy_ok() { return !C.dist && C.compression == 18 && !C.zsplit; }
static bool nd_nfl2(int prec) {
  if (C.zsplit) return false; /* split Z: single hops (each with its z fix-up) + sweeps */
  if (C.hop2_variant == 2) return true;
  if (C.hop2_variant == 0 || C.hop2_variant == 3) return !nd_legacy_ok() && prec; /* forced: double falls back to single hops where the kernel does not apply */
  if (C.hop2_variant == 1) return prec != 0;               /* the lane-paired kernel is double only */
  return !nd_legacy_ok();
}
static bool nd_fused() { return nd_nfl2(0) || nd_legacy_ok(); }
struct NdDot { const tmb_cg_state *st = nullptr; int fin_op = -1, fin_slot = 0; double scale = 1.; int *np = nullptr; bool on = false; };
static int hop2(int ieo, void *o0, void *o1, const void *i0, const void *i1, int mode, const void *p0,
                const void *p1, double mu, double eps, double scale, int prec = 0, const tmb_cg_state *st = nullptr,
                const NdDot *dot = nullptr) {
  if (nd_nfl2(prec)) { /* hop_kernel with NFL = 2: every precision, compression and communication mode */
    HopOpt o; o.nfl = 2; o.mode = mode; o.prec = prec; o.in1 = i1; o.out1 = o1; o.p = p0; o.p1 = p1; o.st = st;
    o.nd_mu = mu; o.nd_eps = eps; o.nd_scale = scale;
    if (dot && dot->on) { o.selfnorm = true; o.fin_op = dot->fin_op; o.fin_slot = dot->fin_slot; o.dot_scale = dot->scale; o.npartial = dot->np; }
    return hop(ieo, o0, i0, o);
  }
  return hop2_legacy(ieo, o0, o1, i0, i1, mode, p0, p1, mu, eps, scale, prec, st, dot);
}
/* the round-1 kernels (tmb_force.cu: one thread carries both flavours, variant 0; lane-paired flavours, variant 1): one
 * rank, 18-real links, double precision only */
static int hop2_legacy(int ieo, void *o0, void *o1, const void *i0, const void *i1, int mode, const void *p0,
                       const void *p1, double mu, double eps, double scale, int prec, const tmb_cg_state *st, const NdDot *dot) {
  if (!C.gauge_loaded) return fail(-9, "no gauge field on the device: call tmb_gauge_upload first");
  if (C.kappa == 0.) return fail(-9, "hopping parameter not set: call tmb_set_boundary first");
  if (!nd_legacy_ok()) return fail(-7, "this two-flavour kernel runs on one rank with 18-real links only");
  if (prec) TRY(ensure_gauge32());
  tmb_hop2_launch a;
  memset(&a, 0, sizeof(a));
  a.in0 = i0; a.in1 = i1; a.out0 = o0; a.out1 = o1; a.p0 = p0; a.p1 = p1; a.g = C.g; a.par = ieo ? 1 : 0;
  a.prec = prec; a.U = prec ? (const void *)C.U32 : (const void *)C.U;
  for (int m = 0; m < 4; m++) a.ka[m] = C.ka[m];
  a.mode = mode; a.mu = mu; a.eps = eps; a.scale = scale; a.hints = eff_hints(); a.variant = (C.hop2_variant == 1 && !prec) ? 1 : ((C.hop2_variant == 3 && !prec) ? 3 : 0);
  a.prefetch = C.prefetch & 1; a.prefetch_dist = C.prefetch_dist;
  a.tile = C.tile && a.variant != 1 && tmb_tile_ok(a.g);
  if (a.tile) a.prefetch = 0;
  a.st = st; a.fin_op = -1;
  if (dot && dot->on) {
    if (a.variant == 1) return fail(-7, "the lane-paired two-flavour kernel has no fused norm");
    a.dot = 2; a.dot_scale = dot->scale; a.partial = C.partial; a.st_fin = C.st;
    a.fin_op = fuse_fin() ? dot->fin_op : -1; a.fin_slot = dot->fin_slot; a.xr = a.fin_op >= 0 ? xr_tab() : nullptr;
    const int np = tmb_hop2_grid(a);
    if (np > C.npartial) return fail(-10, "partial buffer too small (%d > %d)", np, C.npartial);
    if (dot->np) *dot->np = np;
  }
  KL(tmb_launch_hop2(a, C.s_main));
  return 0;
}

/* tm_operators_nd.c:68-89 */
extern "C" int tmb_Qtm_ndpsi(void *ls_, void *lc_, const void *ks_, const void *kc_) {
  NEED_INIT();
  double2 *ls = F(ls_), *lc = F(lc_); const double2 *ks = F(ks_), *kc = F(kc_);
  SCR(s0, 6); SCR(s1, 7); SCR(s2, 8); SCR(s3, 9);
  if (nd_fused()) {
    TRY(hop2(0, s3, s2, ks, kc, 1, nullptr, nullptr, C.mubar, C.epsbar, 1.));
    return hop2(1, ls, lc, s3, s2, 2, ks, kc, -C.mubar, -C.epsbar, C.invmaxev);
  }
  TRY(hop0(0, s0, ks)); TRY(hop0(0, s1, kc));
  KL(tmb_launch_nd_mee_inv(s3, s2, s0, s1, C.mubar, C.epsbar, N2(), HALF(), C.s_main));
  TRY(hop0(1, ls, s3)); TRY(hop0(1, lc, s2));
  TRY(nd_moo(s0, s1, ks, kc, ls, lc, -C.mubar, -C.epsbar));
  KL(tmb_launch_scale(ls, C.invmaxev, s0, N2(), C.s_main));
  KL(tmb_launch_scale(lc, C.invmaxev, s1, N2(), C.s_main));
  return 0;
}
/* tm_operators_nd.c:130-152; l may alias k */
extern "C" int tmb_Qtm_dagger_ndpsi(void *ls_, void *lc_, const void *ks_, const void *kc_) {
  NEED_INIT();
  double2 *ls = F(ls_), *lc = F(lc_); const double2 *ks = F(ks_), *kc = F(kc_);
  SCR(s0, 6); SCR(s1, 7); SCR(s2, 8); SCR(s3, 9);
  if (nd_fused()) {
    TRY(hop2(0, s2, s3, kc, ks, 1, nullptr, nullptr, C.mubar, C.epsbar, 1.));
    return hop2(1, ls, lc, s3, s2, 2, ks, kc, C.mubar, -C.epsbar, C.invmaxev);
  }
  TRY(hop0(0, s0, kc)); TRY(hop0(0, s1, ks));
  KL(tmb_launch_nd_mee_inv(s2, s3, s0, s1, C.mubar, C.epsbar, N2(), HALF(), C.s_main));
  TRY(hop0(1, s0, s2)); TRY(hop0(1, s1, s3));
  TRY(nd_moo(ls, lc, ks, kc, s1, s0, C.mubar, -C.epsbar));
  KL(tmb_launch_scale(lc, C.invmaxev, lc, N2(), C.s_main));
  KL(tmb_launch_scale(ls, C.invmaxev, ls, N2(), C.s_main));
  return 0;
}
/* tm_operators_nd.c:195-238 (prec 0) and tm_operators_nd_32.c:215-262 (prec 1, float fields) as 4 two-flavour launches,
 * 8448 B/site in double instead of 8 hops + 5 sweeps, 17.7 kB/site.  Qtm_pm_ndpsi = Qhat Qhat^dagger: its first half B =
 * Qhat^dagger k is the output of the second launch, so the CG's <k, Qhat Qhat^dagger k> = invmaxev^2 |B|^2 is the squared norm
 * of that launch's own output (`dot`: fused reduction and finish, no operand load, no separate dot kernel). */
static float2 *scratch32(int k);
static int qtm_pm_nd_x(int prec, void *ls, void *lc, const void *ks, const void *kc, const tmb_cg_state *st, NdDot *dot) {
  void *a0, *a1, *b0, *b1;
  if (prec) { a0 = scratch32(6); a1 = scratch32(7); b0 = scratch32(8); b1 = scratch32(9); }
  else { a0 = scratch(6); a1 = scratch(7); b0 = scratch(8); b1 = scratch(9); }
  if (!a0 || !a1 || !b0 || !b1) return -100;
  if (dot) dot->scale = C.invmaxev * C.invmaxev;
  TRY(hop2(0, a0, a1, kc, ks, 1, nullptr, nullptr, C.mubar, C.epsbar, 1., prec, st));            /* A = Mee^-1 H_eo (kc, ks) */
  TRY(hop2(1, b0, b1, a0, a1, 2, kc, ks, -C.mubar, -C.epsbar, 1., prec, st, dot));              /* B = g5(Moo(kc,ks) - H_oe A): tau1 Qhat tau1 */
  TRY(hop2(0, a0, a1, b0, b1, 1, nullptr, nullptr, -C.mubar, C.epsbar, 1., prec, st));          /* A = (s5, s4) */
  return hop2(1, ls, lc, a1, a0, 2, b1, b0, -C.mubar, -C.epsbar, C.invmaxev * C.invmaxev, prec, st);
}
static int qtm_pm_nd(double2 *ls, double2 *lc, const double2 *ks, const double2 *kc) {
  SCR(s0, 6); SCR(s1, 7); SCR(s2, 8); SCR(s3, 9); SCR(s4, 10); SCR(s5, 11);
  if (nd_fused()) return qtm_pm_nd_x(0, ls, lc, ks, kc, nullptr, nullptr);
  TRY(hop0(0, s0, kc)); TRY(hop0(0, s1, ks));
  KL(tmb_launch_nd_mee_inv(s2, s3, s0, s1, C.mubar, C.epsbar, N2(), HALF(), C.s_main));
  TRY(hop0(1, s0, s2)); TRY(hop0(1, s1, s3));
  TRY(nd_moo(s2, s3, kc, ks, s0, s1, -C.mubar, -C.epsbar));
  TRY(hop0(0, s0, s3)); TRY(hop0(0, s1, s2));
  KL(tmb_launch_nd_mee_inv(s5, s4, s1, s0, -C.mubar, C.epsbar, N2(), HALF(), C.s_main));
  TRY(hop0(1, ls, s4)); TRY(hop0(1, lc, s5));
  TRY(nd_moo(ls, lc, s3, s2, ls, lc, -C.mubar, -C.epsbar));
  const double f = C.invmaxev * C.invmaxev;
  if (f != 1.) {
    KL(tmb_launch_scale(lc, f, lc, N2(), C.s_main));
    KL(tmb_launch_scale(ls, f, ls, N2(), C.s_main));
  }
  return 0;
}
extern "C" int tmb_Qtm_pm_ndpsi(void *ls, void *lc, const void *ks, const void *kc) {
  NEED_INIT(); return qtm_pm_nd(F(ls), F(lc), F(ks), F(kc));
}

/* tmb_cg_her_nd (solver/cg_her_nd.c:57-170): device-resident, see tmb_capi_mixed.inc */
extern "C" int tmb_cg_her_nd(void *Pup, void *Pdn, const void *Qup, const void *Qdn, int max_iter, double eps_sq, int rel_prec);

/* invert_doublet_eo.c:102-178 (NO_EXT_INV): solver_flag RGMIXEDCG (14) -> rg_mixed_cg_her_nd (:145-149), anything else -> cg_her_nd */
static int nd_pair(int k, double2 **p);
extern "C" int tmb_rg_mixed_cg_her_nd(void *Pup, void *Pdn, const void *Qup, const void *Qdn, int max_iter, double eps_sq, int rel_prec);
extern "C" int tmb_invert_doublet_eo_solver(void *ens, void *ons, void *enc, void *onc, const void *es, const void *os,
                                            const void *ec, const void *oc, double precision, int max_iter, int rel_prec, int solver_flag);
extern "C" int tmb_invert_doublet_eo(void *ens, void *ons, void *enc, void *onc, const void *es, const void *os,
                                     const void *ec, const void *oc, double precision, int max_iter, int rel_prec) {
  return tmb_invert_doublet_eo_solver(ens, ons, enc, onc, es, os, ec, oc, precision, max_iter, rel_prec, TMB_SOLVER_CG);
}
extern "C" int tmb_invert_doublet_eo_solver(void *ens, void *ons, void *enc, void *onc, const void *es, const void *os,
                                            const void *ec, const void *oc, double precision, int max_iter, int rel_prec, int solver_flag) {
  NEED_INIT();
  const size_t n2 = N2();
  /* temporaries live in the context: allocating and freeing four fields per call cost more than the rest of the
   * wrapper (0.1 - 0.6 s at 32^3x64, cudaFree synchronises and unmaps) */
  double2 *d[4], *t0 = nullptr, *t1 = nullptr;
  TRY(nd_pair(5, &t0)); TRY(nd_pair(6, &t1));
  d[0] = t0; d[1] = t0 + n2; d[2] = t1; d[3] = t1 + n2;
  int rc = 0, iter = -1;
  do {
    if ((rc = tmb_M_ee_inv_ndpsi(ens, enc, es, ec, C.mubar, C.epsbar)) < 0) break;
    if ((rc = hop0(1, d[0], F(ens))) < 0) break;
    if ((rc = hop0(1, d[1], F(enc))) < 0) break;
    if ((rc = tmb_assign_mul_add_r(d[0], 1., os)) < 0) break;
    if ((rc = tmb_assign_mul_add_r(d[1], 1., oc)) < 0) break;
    if ((rc = tmb_gamma5(d[0], d[0])) < 0) break;
    if ((rc = tmb_gamma5(d[1], d[1])) < 0) break;
    if (solver_flag == TMB_SOLVER_RGMIXEDCG) iter = tmb_rg_mixed_cg_her_nd(ons, onc, d[0], d[1], max_iter, precision, rel_prec);
    else iter = tmb_cg_her_nd(ons, onc, d[0], d[1], max_iter, precision, rel_prec);
    if (iter < -1) { rc = iter; break; }
    if ((rc = tmb_Qtm_dagger_ndpsi(ons, onc, ons, onc)) < 0) break;
    if ((rc = hop0(0, d[0], F(ons))) < 0) break;
    if ((rc = hop0(0, d[1], F(onc))) < 0) break;
    if ((rc = tmb_M_ee_inv_ndpsi(d[2], d[3], d[0], d[1], C.mubar, C.epsbar)) < 0) break;
    if ((rc = tmb_assign_add_mul_r(ens, d[2], 1.)) < 0) break;
    if ((rc = tmb_assign_add_mul_r(enc, d[3], 1.)) < 0) break;
    cudaStreamSynchronize(C.s_main);
  } while (0);
  return rc < 0 ? rc : iter;
}

#include "tmb_capi_mixed.inc"
#include "tmb_capi_hmc.inc"

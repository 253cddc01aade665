#!/usr/bin/env python
"""bench.py - Hopping_Matrix throughput (+ eo-CG time-to-solution) on B200, BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (oracle/_ref)

A step is one EO+OE pair of Hopping_Matrix calls over the whole local lattice, exactly the loop body of the reference's
benchmark.c:293-299.  Workload: N=1 -> BASELINE configs[1] lattice (24^3 x 48, kappa=0.16, mu=0.01); N>1 -> configs[2]:
48^3 x (12 N) split along T, 48^3 x 12 per GPU (weak scaling), T-neighbour fields read over NVLink (peer mode; NCCL halos as
fallback).  Inputs at EVERY N are the reference's own: start_ranlux(1, 123456); random_gauge_field; random_spinor_field_eo
(benchmark.c:247-259), drawn by the unmodified reference (oracle/_ref) on the GLOBAL lattice on rank 0 and scattered as T slabs.
One JSON line on stdout (rank 0).

Keys beyond the driver's contract:
  roofline      achieved = 1536 B x sites / mean launch time over the timed region, peak = MEASURED_PEAKS.json; `frac_sustained`
                = the same over the >= 1.5 s of identical steps that follow the timed region (clock record beside it); `traffic` =
                DRAM bytes per launch from the committed ncu capture; `copy_gbs_sustained_this_run`
  parity        N = 1: Hopping_Matrix / Qtm_pm_psi against the unmodified reference on the same inputs, invert_eo iteration count
                against the reference's cg_her.  N > 1: the N-rank results (gathered) against the reference on the GLOBAL lattice
                (hop, Qtm_pm_psi) and against a ONE-GPU device solve of the same global problem (CG count, residual, solution),
                once in peer mode and once with TMB_P2P=0 (NCCL halos + all-reduce); recipe of test/check_xchange.c:88-150 and
                hopping_test.c:305-354 (N-rank result vs trusted result)
  weak_anchor   the same Hopping_Matrix pairs on ONE GPU at the per-GPU volume of the N > 1 runs (48^3 x 12)
  comm          N > 1: Hopping_Matrix against Hopping_Matrix_nocom (benchmark.c:337-373)
  e2e           the same pairs through the reference-named Hopping_Matrix() with HOST buffers, at every N
  cg            invert_eo time to solution: device-resident, through the host-pointer drop-in, mixed precision, 12-real links
  cpu_baseline  the unmodified reference on all host cores (Hopping_Matrix pairs; its cg_her on the same solve at N = 1)
and one section per remaining BASELINE config, each in its own process (scripts/bench_sections.py): `benchmark_8x8x8x8`
(configs[0]), `nd` (configs[3]), `hmc` (configs[4]).  `--lattice TxLXxLYxLZ --global-chunk-t 12` gives the strong-scaling series
of configs[2] on one global problem (scripts/gpu_strong.sh).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_SITE = 1320.0        # north_star convention (phases folded); the reference prints 1608 (benchmark.c:327)
FLOP_SITE_REF = 1608.0
BYTES_SITE = 1536.0       # compulsory bytes per output site: 8 links x 144 + 192 in + 192 out
KAPPA, MU = 0.16, 0.01
GMU = 2 * KAPPA * MU      # g_mu = 2 kappa mu (invert_eo.c:255)
CG_EPS_SQ, CG_MAXITER = 1e-14, 5000
ANCHOR_DIMS = (12, 48, 48, 48)   # per-GPU volume of the N > 1 runs
TOL_HOP = 1e-13                  # north_star: relative L2 difference of the double-precision operator


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, pw, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1]); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "sm_mhz_min": min(sm) if sm else None, "power_w_max": max(pw) if pw else None}


def numpy_gauge(dims, seed):
    V = int(np.prod(dims))
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(V * 4, 3, 3)) + 1j * rng.normal(size=(V * 4, 3, 3))
    q, r = np.linalg.qr(a)
    dg = np.diagonal(r, axis1=1, axis2=2)
    q = q * (dg / np.abs(dg))[:, None, :]
    q = q / np.linalg.det(q)[:, None, None] ** (1.0 / 3.0)
    return np.ascontiguousarray(q.reshape(V * 4, 9)).view(np.float64).reshape(V, 4, 18)


def reference_inputs(gdims, nsrc=3):
    """Gauge field and Gaussian eo sources of the GLOBAL lattice drawn by the unmodified reference (oracle/_ref):
    start_ranlux(1, 123456); random_gauge_field(repro); random_spinor_field_eo(.., RN_GAUSS) as benchmark.c:247-259.
    Returns (gauge, [sources], Reference object or None, description).  Fallback without oracle/_ref: numpy."""
    V = int(np.prod(gdims))
    try:
        from oracle import refclient
        hs = refclient.available(halfspinor=True)
        if hs or refclient.available():
            ref = refclient.Reference(*gdims, nthreads=os.cpu_count() or 1, halfspinor=hs)
            g = ref.random_gauge(123456)
            srcs = [ref.random_spinor_eo() for _ in range(nsrc)]
            return g, srcs, ref, REFERENCE_INPUTS_HOW
    except Exception as e:  # pragma: no cover
        log("reference generator unavailable:", e)
    rng = np.random.default_rng(99)
    return (numpy_gauge(gdims, 123456), [rng.normal(scale=np.sqrt(0.5), size=(V // 2, 24)) for _ in range(nsrc)], None,
            "numpy QR random SU(3) + numpy Gaussian sources (oracle/_ref not built)")


def chunked_fields(dims, rank, chunk_t):
    """Decomposition-independent synthetic inputs for the strong-scaling series: the GLOBAL lattice is cut into chunks
    of `chunk_t` time-slices, chunk c of the global lattice is drawn from numpy seeds tied to c, and a rank assembles
    the chunks of its slab (t is the slowest index of the lexicographic and of the even/odd orderings, so chunks
    concatenate).  Every N that divides the number of chunks sees the same global gauge field and sources."""
    T, LX, LY, LZ = dims
    assert T % chunk_t == 0 and chunk_t % 2 == 0
    per = T // chunk_t
    gs, srcs = [], [[], [], []]
    for c in range(per):
        gc = rank * per + c
        gs.append(numpy_gauge((chunk_t, LX, LY, LZ), 7000 + gc))
        rng = np.random.default_rng(8000 + gc)
        for k in range(3):
            srcs[k].append(rng.normal(scale=np.sqrt(0.5), size=(chunk_t * LX * LY * LZ // 2, 24)))
    return np.concatenate(gs), [np.concatenate(x) for x in srcs], f"numpy QR random SU(3), global chunks of {chunk_t} time-slices (seed 7000 + chunk)"


def pinned(dev, shape):
    n = int(np.prod(shape))
    p = dev.lib.tmb_host_alloc(n * 8)
    if not p:
        raise RuntimeError("tmb_host_alloc failed")
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,)).reshape(shape), p


# ----------------------------------------------------------------------------------- reference arm
def run_reference(args, dims, real_stdout=sys.stdout):
    """the reference's own CPU Hopping_Matrix (half-spinor OpenMP build, fastest generic-C variant,
    SURVEY 6) on the host cores: bench loop of benchmark.c:262-327, one pair per step."""
    from oracle import refclient
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    hs = refclient.available(halfspinor=True)
    if not (hs or refclient.available()):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}), file=real_stdout, flush=True)
        return
    ref = refclient.Reference(*dims, nthreads=ncores, halfspinor=hs)
    ref.set_params(KAPPA, GMU)
    ref.random_gauge(123456)
    V = int(np.prod(dims))
    for _ in range(max(1, args.warmup)):
        ref.bench_hopping(1)
    t = ref.bench_hopping(args.steps)
    gf = V * FLOP_SITE * args.steps / t / 1e9   # V sites per pair (V/2 per call)
    out = {
        "impl": "reference", "metric": "Hopping_Matrix GFLOP/s (eo, double, 1320 flop/site)", "value": gf,
        "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, dims, REFERENCE_INPUTS_HOW, nz=(args.nz if args.gpus > 1 else 1)),
        "gflops_1608": V * FLOP_SITE_REF * args.steps / t / 1e9,
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": ref.nthreads,
                         "kind": "reference", "sample": f"{args.steps} EO+OE Hopping_Matrix pairs on {dims} "
                         f"({'half-spinor' if hs else 'full-spinor'} OpenMP build of the unmodified reference)"},
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), file=real_stdout, flush=True)


def workload_name(ngpus, dims):
    T, LX, LY, LZ = dims
    if ngpus == 1:
        return f"BASELINE configs[1] lattice {LX}^3x{T} (eo Hopping_Matrix pairs + invert_eo CG), kappa=0.16 mu=0.01, random SU(3)"
    return (f"BASELINE configs[2]: {LX}^3x{T * ngpus} split along T over {ngpus} GPUs, {LX}^3x{T} per GPU (weak), "
            "T-neighbour fields read in place over NVLink (peer mode; NCCL half-spinor halos as fallback)")


REFERENCE_INPUTS_HOW = ("unmodified reference on the global lattice: start_ranlux(1,123456); random_gauge_field; "
                        "random_spinor_field_eo(RN_GAUSS) (benchmark.c:247-259)")


def workload_config(ngpus, dims, gauge_how, nt=None, nz=1):
    """the `config` object of the JSON line - the same for both arms (`--impl reference` times the per-GPU volume on the host)"""
    V = int(np.prod(dims))
    nt = ngpus // nz if nt is None else nt
    return {"workload": workload_name(ngpus, dims), "lattice_TxLXxLYxLZ": list(dims), "kappa": KAPPA, "mu": MU,
            "gauge": gauge_how, "l2": "inputs larger than L2: gauge field %.0f MB + spinors per call" % (V * 4 * 144 / 1e6),
            "step": "one EO+OE Hopping_Matrix pair (benchmark.c:293-299)",
            "rank_grid_TxZ": [nt, nz], "global_lattice_TxLXxLYxLZ": [dims[0] * nt, dims[1], dims[2], dims[3] * nz]}


def cpu_baseline(ref, dims, target_s=10.0):
    """oracle/_ref timed on the host cores, bounded sample of the same workload"""
    V = int(np.prod(dims))
    ref.set_params(KAPPA, GMU)
    t1 = ref.bench_hopping(1)
    n = int(max(2, min(200, target_s / max(t1, 1e-6))))
    t = ref.bench_hopping(n)
    tq1 = ref.bench_Qtm_pm(1)
    nq = int(max(2, min(100, 0.3 * target_s / max(tq1, 1e-6))))
    tq = ref.bench_Qtm_pm(nq) / nq
    return {"value": V * FLOP_SITE * n / t / 1e9, "unit": "GFLOP/s", "cores": ref.nthreads, "kind": "reference",
            "sample": f"{n} EO+OE Hopping_Matrix pairs on the same lattice, "
                      f"{'half-spinor' if ref.is_halfspinor() else 'full-spinor'} OpenMP build of the unmodified "
                      f"reference (oracle/_ref), {ref.nthreads} threads",
            "gflops_1608": V * FLOP_SITE_REF * n / t / 1e9, "ms_per_pair": 1e3 * t / n,
            "qtm_pm_psi_s": tq}


# ----------------------------------------------------------------------------------- device sessions
class Session:
    """one library context on this rank's GPU: tmb_init (+ tmb_comm_init over `dist`), parameters, gauge"""

    def __init__(self, dims, world, rank, local_rank, dist, gauge, p2p=True, nz=1):
        import tmlqcd_b200 as tm
        self.tm, self.world, self.rank, self.dist = tm, world, rank, dist
        self.nz, self.nt, self.dims = nz, world // nz, dims
        os.environ["TMB_P2P"] = "1" if p2p else "0"
        self.dev = dev = tm.Device(*dims, device=local_rank)
        self.lib = lib = dev.lib
        if world > 1:
            import torch
            idbuf = (C.c_ubyte * 128)()
            if rank == 0:
                dev.ck(lib.tmb_comm_unique_id(C.cast(idbuf, C.c_void_p)))
            t_id = torch.tensor(list(idbuf), dtype=torch.uint8, device="cuda")
            dist.broadcast(t_id, 0)
            idbuf = (C.c_ubyte * 128)(*t_id.cpu().tolist())
            dev.ck(lib.tmb_comm_init_grid(C.cast(idbuf, C.c_void_p), world // nz, nz, rank))
        dev.set_params(KAPPA, GMU)
        dev.gauge_upload(gauge)
        self.path = "peer" if lib.tmb_comm_peer_mode() else ("nccl" if world > 1 else "single")
        if nz > 1:
            self.path += f"+zsplit{nz}" + ("(peer)" if lib.tmb_comm_zpeer_mode() else "(nccl)")

    def barrier(self):
        self.dev.ck(self.lib.tmb_sync())
        if self.dist is not None:
            import torch
            self.dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_pairs(self, nsteps, f0, f1, f2, nocom=False):
        """`nsteps` EO+OE pairs, CUDA events on the library's compute stream, barrier + sync on both sides, max over ranks"""
        hop = self.lib.tmb_Hopping_Matrix_nocom if nocom else self.lib.tmb_Hopping_Matrix
        self.barrier()
        n0 = self.lib.tmb_launch_count()
        self.dev.timer_start()
        for _ in range(nsteps):
            hop(0, f1, f0)
            hop(1, f2, f1)
        ms = self.dev.timer_stop()
        self.barrier()
        return self.max_over_ranks(ms), self.lib.tmb_launch_count() - n0

    def gather(self, field):
        """the global field on rank 0 (T slabs concatenate: t is the slowest index of the eo ordering)"""
        loc = self.dev.download(field)
        if self.dist is None:
            return loc
        import torch
        t = torch.from_numpy(loc).cuda()
        lst = [torch.empty_like(t) for _ in range(self.world)] if self.rank == 0 else None
        self.dist.gather(t, lst, dst=0)
        if self.rank != 0:
            return None
        if self.nz == 1:
            return torch.cat(lst).cpu().numpy()
        return unslab([x.cpu().numpy() for x in lst], self.dims, self.nt, self.nz, 2, 24)

    def operators_and_solve(self, src, E, O, keep=True):
        """the parity workload: both hops and Qtm_pm_psi on `src`, invert_eo on (E, O); fields gathered on rank 0"""
        dev = self.dev
        dk, dl = dev.field(src), dev.field()
        res = {}
        for ieo in (0, 1):
            dev.call("Hopping_Matrix", ieo, dl, dk)
            res[f"hop{ieo}"] = self.gather(dl)
        dev.call("Qtm_pm_psi", dl, dk)
        res["qtm_pm"] = self.gather(dl)
        dE, dO, dEn, dOn = dev.field(E), dev.field(O), dev.field(), dev.field()
        res["cg_iters"] = dev.call("invert_eo", dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
        _, res["cg_rr"], res["cg_loop_s"] = dev.solver_stats()
        res["inv_e"], res["inv_o"] = self.gather(dEn), self.gather(dOn)
        dev.free(dk, dl, dE, dO, dEn, dOn)
        return res

    def close(self):
        self.dev.close()


def slab(arr, dims, nt, nz, r, zdiv, cols):
    """the part of a GLOBAL field that rank r = ct * nz + cz of an (nt x nz) grid holds: t is the slowest index of the
    lexicographic and of the even/odd orderings, z the fastest (z / 2 for eo fields: zdiv = 2).  `dims` are the local extents."""
    T, LX, LY, LZ = dims
    ct, cz = r // nz, r % nz
    a = arr.reshape(T * nt, LX * LY, LZ * nz // zdiv, cols)
    return np.ascontiguousarray(a[ct * T:(ct + 1) * T, :, cz * (LZ // zdiv):(cz + 1) * (LZ // zdiv)]).reshape(-1, cols)


def unslab(parts, dims, nt, nz, zdiv, cols):
    T, LX, LY, LZ = dims
    out = np.zeros((T * nt, LX * LY, LZ * nz // zdiv, cols))
    for r, part in enumerate(parts):
        ct, cz = r // nz, r % nz
        out[ct * T:(ct + 1) * T, :, cz * (LZ // zdiv):(cz + 1) * (LZ // zdiv)] = part.reshape(T, LX * LY, LZ // zdiv, cols)
    return out.reshape(-1, cols)


def scatter_rows(dist, world, rank, arr, rows, cols, grid=None):
    """rank 0 holds the global field; every rank returns its slab of `rows` rows (numpy).  grid = (dims, nt, nz, zdiv) for a
    T x Z grid of ranks, None for T slabs (contiguous row blocks)"""
    import torch
    out = torch.empty((rows, cols), dtype=torch.float64, device="cuda")
    lst = None
    if rank == 0:
        if grid is not None and grid[2] > 1:
            dims_, nt_, nz_, zdiv_ = grid
            lst = [torch.from_numpy(slab(np.asarray(arr), dims_, nt_, nz_, r, zdiv_, cols)).cuda() for r in range(world)]
        else:
            src = torch.from_numpy(np.ascontiguousarray(arr).reshape(world * rows, cols)).cuda()
            lst = list(src.chunk(world))
    dist.scatter(out, lst, src=0)
    res = out.cpu().numpy()
    del out, lst
    torch.cuda.empty_cache()
    return res


def compare(res, trusted, what):
    """relative L2 differences of the gathered N-rank fields against trusted fields of the same names"""
    return {k: rel_l2(res[k], trusted[k]) for k in what}


def run_anchor(steps, warmup):
    """the weak-scaling anchor: this bench on ONE GPU at the per-GPU volume of the N > 1 runs, in its own process"""
    try:
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "TMB_P2P")}
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--lattice", "x".join(str(x) for x in ANCHOR_DIMS),
                            "--steps", str(steps), "--warmup", str(warmup), "--skip-cpu", "--skip-cg", "--skip-e2e", "--skip-sections",
                            "--skip-anchor", "--skip-parity"], capture_output=True, text=True, timeout=600, env=env)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        return {"lattice_TxLXxLYxLZ": list(ANCHOR_DIMS), "n_gpus": 1, "value": d["value"], "unit": "GFLOP/s",
                "us_per_hop": d["roofline"]["avg_launch_us"], "frac": d["roofline"]["frac"],
                "frac_sustained": d["roofline"].get("frac_sustained"), "steps": steps,
                "use": "weak-scaling efficiency at equal volume per GPU = value_N / (N x this value)"}
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)[:300]}


# ----------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lattice", default=None, help="local TxLXxLYxLZ, e.g. 48x24x24x24")
    ap.add_argument("--global-chunk-t", type=int, default=0,
                    help="strong-scaling series: draw gauge field and sources per chunk of this many GLOBAL time-slices, so every N sees the same global problem")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-cg", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--skip-anchor", action="store_true")
    ap.add_argument("--skip-sections", action="store_true", help="skip the configs[0], configs[3] (ND doublet) and configs[4] (HMC monomials) sections")
    ap.add_argument("--loopback", action="store_true", help="1 GPU: run the T-split halo/boundary path against itself")
    ap.add_argument("--loopback2", action="store_true", help="1 GPU: run the T-split peer-mode path against itself")
    ap.add_argument("--sweep", action="store_true", help="time every kernel variant (tuning aid, prints to stderr)")
    ap.add_argument("--sweep-sustained", action="store_true", help="time the residency variants for >= 0.6 s each (power-capped regime; stderr)")
    ap.add_argument("--overlap", type=int, default=0, help="tmb_set_overlap flags: 1 PDL, 2 L2 gauge prefetch")
    ap.add_argument("--p2p-diag", type=int, default=0, help="tmb_set_p2p_diag bits (timing diagnostics, results invalid; needs TMB_P2P_DIAG=1)")
    ap.add_argument("--nz", type=int, default=1, help="split Z over this many ranks as well (rank grid (gpus / nz) x nz); default: T only")
    ap.add_argument("--variant", type=int, default=None)
    ap.add_argument("--hints", type=int, default=None)
    ap.add_argument("--xblock", type=int, default=None)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    # rank 0 prints exactly ONE line on stdout: libraries that write there (NCCL's "NCCL version ..."
    # banner) are moved to stderr by pointing fd 1 at fd 2 and keeping a private copy of the real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.lattice:
        dims = tuple(int(x) for x in args.lattice.lower().split("x"))
    else:
        dims = (48, 24, 24, 24) if args.gpus == 1 else ANCHOR_DIMS
    nz = args.nz if world > 1 else 1
    assert world % nz == 0
    nt = world // nz
    gdims = (dims[0] * nt, dims[1], dims[2], dims[3] * nz)

    if args.impl == "reference":
        run_reference(args, dims, real_stdout)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local_rank)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_

    # ---- inputs: the reference's own hot start and Gaussian sources on the global lattice, T slabs per rank ----
    V, Vh = int(np.prod(dims)), int(np.prod(dims)) // 2
    ref = None
    G = {}  # rank 0: global gauge field and sources (parity phase)
    t_in = time.perf_counter()
    if args.global_chunk_t:
        g, srcs, gauge_how = chunked_fields(dims, rank, args.global_chunk_t)
        args.skip_parity = True
    elif world == 1:
        g, srcs, ref, gauge_how = reference_inputs(dims)
    else:
        if rank == 0:
            G["g"], G["srcs"], ref, gauge_how = reference_inputs(gdims)
        g = scatter_rows(dist, world, rank, G.get("g"), V, 72, (dims, nt, nz, 1))
        srcs = [scatter_rows(dist, world, rank, G["srcs"][k] if rank == 0 else None, Vh, 24, (dims, nt, nz, 2)) for k in range(3)]
        gauge_how = gauge_how if rank == 0 else ""
    src, E, O = srcs
    log(f"rank {rank}: inputs ready in {time.perf_counter() - t_in:.1f} s")

    S = Session(dims, world, rank, local_rank, dist, g, p2p=True, nz=nz)
    dev, lib = S.dev, S.lib
    if args.variant is not None or args.hints is not None or args.xblock is not None:
        dev.ck(lib.tmb_set_tuning(-1 if args.variant is None else args.variant, -1 if args.hints is None else args.hints, args.xblock or 0))
    dev.ck(lib.tmb_set_overlap(args.overlap))
    if args.p2p_diag:
        dev.ck(lib.tmb_set_p2p_diag(args.p2p_diag))
    if (args.loopback or args.loopback2) and world == 1:
        dev.ck(lib.tmb_comm_loopback(2 if args.loopback2 else 1))
        dev.gauge_upload(g)
    f0, f1, f2 = dev.field(src), dev.field(), dev.field()

    if args.sweep_sustained and rank == 0:
        # every residency variant of the plain kernel for >= 0.6 s each: the power-capped regime the CG lives in
        log("sustained sweep (>= 0.6 s per variant, cache policies on):")
        for variant in list(range(0, 11)):
            dev.ck(lib.tmb_set_tuning(variant, 1, 0))
            S.time_pairs(20, f0, f1, f2)
            ms1, _ = S.time_pairs(100, f0, f1, f2)
            n = int(max(200, 600.0 / (ms1 / 100)))
            ms_v, _ = S.time_pairs(n, f0, f1, f2)
            per = ms_v / (2 * n)
            log(f"  variant={variant:2d}: burst {1e3 * ms1 / 200:7.2f} us/hop, sustained {1e3 * per:7.2f} us/hop "
                f"{Vh * BYTES_SITE / per / 1e6:8.1f} GB/s = {Vh * BYTES_SITE / per / 1e6 / measured_peaks()[0]:.3f} of the measured peak")
        dev.ck(lib.tmb_set_tuning(-1 if args.variant is None else args.variant, -1 if args.hints is None else args.hints, args.xblock or 0))

    if args.sweep and rank == 0:
        results = []
        xbs = [0] + [x for x in (2, 4, 8) if dims[1] % x == 0]
        for hints in (1, 0):
            for xb in xbs:
                for variant in range(0, 10):
                    dev.ck(lib.tmb_set_tuning(variant, hints, xb))
                    S.time_pairs(5, f0, f1, f2)
                    ms, _ = S.time_pairs(50, f0, f1, f2)
                    per = ms / 100.0
                    results.append((per, variant, hints, xb))
                    log(f"sweep variant={variant} hints={hints} xblock={xb}: {per * 1e3:8.2f} us/hop "
                        f"{Vh * BYTES_SITE / per / 1e6:8.1f} GB/s {Vh * FLOP_SITE / per / 1e6:8.1f} GFLOP/s")
        results.sort()
        log("best:", results[:5])
        dev.ck(lib.tmb_set_tuning(-1 if args.variant is None else args.variant, -1 if args.hints is None else args.hints, args.xblock or 0))

    # ---- timed region: K pairs, device resident, inputs larger than L2 (gauge alone is 1152 B/site) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # let nvidia-smi come up so that its samples fall inside the loaded window
    S.time_pairs(args.warmup, f0, f1, f2)
    ms, launches = S.time_pairs(args.steps, f0, f1, f2)
    # The same step keeps running behind the timed region until ~1.5 s of this load have been seen: enough nvidia-smi samples
    # (100 ms period) for the clock record, and the SUSTAINED figure next to the short timed region's burst figure.
    extra = int(max(0, (1.5e3 - ms) / max(ms / args.steps, 1e-3)))
    ms_sus = None
    if extra > 0:
        ms_sus, _ = S.time_pairs(extra, f0, f1, f2)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = f"warm-up + {args.steps} timed steps + {extra} identical steps (timed separately: roofline.frac_sustained)"
    sites = V * world  # output sites per pair over all ranks (V/2 per call)
    gflops = sites * FLOP_SITE * args.steps / (ms * 1e-3) / 1e9
    peak, peak_how = measured_peaks()
    per_launch_ms = ms / (2 * args.steps)
    achieved = Vh * BYTES_SITE / (per_launch_ms * 1e-3) / 1e9  # per GPU

    out = {
        "metric": "Hopping_Matrix GFLOP/s (eo, double, 1320 flop/site)", "value": gflops, "unit": "GFLOP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world, dims, gauge_how, nt=nt, nz=nz),
        "gflops_1608": sites * FLOP_SITE_REF * args.steps / (ms * 1e-3) / 1e9,
        "hbm_gbs_effective_per_gpu": achieved,
        "peer_mode": bool(lib.tmb_comm_peer_mode()),
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_how, "frac_of_8TBs_nominal": achieved / 8000.0,
                     "kernel": "hop_kernel<MODE 0> (Hopping_Matrix)", "algorithmic_bytes_per_launch": Vh * BYTES_SITE,
                     "avg_launch_us": per_launch_ms * 1e3},
    }
    if ms_sus is not None:
        sus = Vh * BYTES_SITE / (ms_sus / (2 * extra) * 1e-3) / 1e9
        out["roofline"].update({"achieved_sustained": sus, "frac_sustained": sus / peak, "sustained_window_s": ms_sus * 1e-3,
                                "sustained_avg_launch_us": 1e3 * ms_sus / (2 * extra),
                                "sustained_how": f"{extra} more identical steps right behind the timed region, CUDA events, max over ranks; "
                                                 "clocks.sm_mhz is the median over the same window"})
    # the same process, the same thermal state: device-to-device copy bandwidth sustained over ~0.5 s (read + write bytes),
    # next to the burst figure of MEASURED_PEAKS.json that `peak` quotes
    try:
        gbs = C.c_double(0.)
        dev.ck(lib.tmb_measure_copy_gbs(1 << 30, 1500, C.byref(gbs)))
        out["roofline"]["copy_gbs_sustained_this_run"] = gbs.value
        out["roofline"]["frac_of_sustained_copy"] = achieved / gbs.value
        if ms_sus is not None:
            out["roofline"]["frac_sustained_of_sustained_copy"] = out["roofline"]["achieved_sustained"] / gbs.value
    except Exception as e:  # pragma: no cover
        out["roofline"]["copy_gbs_sustained_this_run"] = None
        log("copy bandwidth measurement failed:", e)
    tr = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tr):
        try:
            t = json.load(open(tr))
            if t.get("lattice") == list(dims):
                out["roofline"]["traffic"] = t["dram_bytes_per_launch"]
        except Exception:
            pass

    # ---- N > 1: communication on / off, the reference's own pair of measurements (benchmark.c:293-299 vs :337-373) ----
    if world > 1:
        nc = max(args.steps, 100)
        S.time_pairs(args.warmup, f0, f1, f2, nocom=True)
        ms_off, _ = S.time_pairs(nc, f0, f1, f2, nocom=True)
        ms_on, _ = S.time_pairs(nc, f0, f1, f2)
        out["comm"] = {"on_us_per_hop": 1e3 * ms_on / (2 * nc), "off_us_per_hop": 1e3 * ms_off / (2 * nc), "steps": nc,
                       "exchange_cost_us_per_hop": 1e3 * (ms_on - ms_off) / (2 * nc),
                       "how": "Hopping_Matrix against Hopping_Matrix_nocom (no halo exchange, the slab wraps onto itself in T), "
                              "the reference's comm-on / comm-off pair (benchmark.c:337-373)"}

    # ---- optional variant: 12-real compressed links (1152 algorithmic B/site), reported separately ----
    dev.ck(lib.tmb_set_compression(12))
    S.time_pairs(args.warmup, f0, f1, f2)
    n12 = max(args.steps // 4, 10)
    ms12, _ = S.time_pairs(n12, f0, f1, f2)
    dev.ck(lib.tmb_set_compression(18))
    per12 = ms12 / (2 * n12)
    out["compression12"] = {"us_per_hop": per12 * 1e3, "gflops_1320": Vh * FLOP_SITE / (per12 * 1e-3) / 1e9 * world,
                            "algorithmic_bytes_per_site": 1152, "hbm_gbs_effective_per_gpu": Vh * 1152.0 / (per12 * 1e-3) / 1e9,
                            "note": "tmb_set_compression(12): two link rows streamed, third rebuilt in registers; "
                                    "not the headline (value uses the reference's 18-real links)"}

    # ---- e2e: the reference-named Hopping_Matrix(ieo, l, k) with HOST buffers, copies inside the timing ----
    # (N > 1: every rank moves its own slab across its own PCIe link; the hop runs the T-split path)
    tm = S.tm
    D = None
    if not args.skip_e2e:
        D = tm.DropIn(*dims, device=local_rank)
        D.glob("g_nproc", C.c_int).value = world; D.glob("g_nproc_t", C.c_int).value = world; D.glob("g_proc_id", C.c_int).value = rank
        D.set_params(KAPPA, GMU)
        D.set_gauge(g)
        hk, _ = pinned(dev, (Vh, 24)); h1, _ = pinned(dev, (Vh, 24)); h2, _ = pinned(dev, (Vh, 24))
        hk[:] = src
        D.Hopping_Matrix(0, h1, hk)  # uploads the gauge (dirty flag) outside the timing
        ne = max(3, min(args.steps, 20))
        for _ in range(2):
            D.Hopping_Matrix(0, h1, hk); D.Hopping_Matrix(1, h2, h1)
        S.barrier()
        n0 = lib.tmb_launch_count()
        t0 = time.perf_counter()
        for _ in range(ne):
            D.Hopping_Matrix(0, h1, hk); D.Hopping_Matrix(1, h2, h1)
        S.barrier()
        dt = S.max_over_ranks(time.perf_counter() - t0)
        out["e2e"] = {"value": V * world * FLOP_SITE * ne / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": 2 * Vh * 192 * world,
                      "d2h_bytes_per_step": 2 * Vh * 192 * world, "steps": ne, "ms_per_step": 1e3 * dt / ne,
                      "gbs_per_direction_per_gpu": 2 * Vh * 192 * ne / dt / 1e9,
                      "api": "Hopping_Matrix(ieo, spinor* l, spinor* k) drop-in, pinned host buffers, upload+kernel+download per call"}
        out["gpu_launches"] += int(lib.tmb_launch_count() - n0)
        # what the host link of this box gives (pinned memory, copy engines), and the same pairs on PAGEABLE buffers - what the
        # reference's calloc'ed fields are (init/init_spinor_field.c:38-66): the library page-locks a slab when it first sees it
        try:
            la, lb, lc = C.c_double(), C.c_double(), C.c_double()
            dev.ck(lib.tmb_measure_pcie_gbs(64 << 20, 10, C.byref(la), C.byref(lb), C.byref(lc)))
            out["e2e"]["link_gbs"] = {"h2d": la.value, "d2h": lb.value, "duplex_per_direction": lc.value,
                                      "how": "64 MB pinned copies on the library's two copy streams, this process, this box"}
            out["e2e"]["frac_of_duplex_link"] = out["e2e"]["gbs_per_direction_per_gpu"] / lc.value
            pk, p1, p2 = np.array(src), np.zeros((Vh, 24)), np.zeros((Vh, 24))

            def pairs_on_pageable(n):
                S.barrier()
                t0 = time.perf_counter()
                for _ in range(n):
                    D.Hopping_Matrix(0, p1, pk); D.Hopping_Matrix(1, p2, p1)
                S.barrier()
                return S.max_over_ranks(time.perf_counter() - t0)
            pairs_on_pageable(1)
            npg = max(3, ne // 2)
            dt_plain = pairs_on_pageable(npg)   # as they are: the driver's bounce buffers
            t0 = time.perf_counter()
            for a_ in (pk, p1, p2):             # the INTEGRATION.md recipe: page-lock each slab once
                dev.ck(lib.tmb_host_register(a_.ctypes.data_as(C.c_void_p), a_.nbytes))
            t_reg = time.perf_counter() - t0
            pairs_on_pageable(1)
            dt_reg = pairs_on_pageable(ne)
            for a_ in (pk, p1, p2):
                dev.ck(lib.tmb_host_unregister(a_.ctypes.data_as(C.c_void_p)))
            out["e2e"]["pageable"] = {
                "value": V * world * FLOP_SITE * ne / dt_reg / 1e9, "unit": "GFLOP/s", "ms_per_step": 1e3 * dt_reg / ne,
                "register_once_ms": 1e3 * t_reg, "unregistered_value": V * world * FLOP_SITE * npg / dt_plain / 1e9,
                "unregistered_ms_per_step": 1e3 * dt_plain / npg,
                "how": "numpy (malloc) buffers like the reference's calloc slabs: `value` after tmb_host_register(slab) once per slab "
                       "(register_once_ms for the three 64 MB slabs), `unregistered_value` with the buffers left pageable"}
        except Exception as e:  # pragma: no cover
            out["e2e"]["link_gbs"] = {"error": repr(e)[:200]}

    # ---- eo-CG time-to-solution (the configs[1] solve): device-resident and through invert_eo with host buffers ----
    if not args.skip_cg:
        dE, dO, dEn, dOn = dev.field(E), dev.field(O), dev.field(), dev.field()
        dev.call("invert_eo", dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)  # warm-up solve
        dev.call("field_zero", dOn)
        S.barrier()
        t0 = time.perf_counter()
        it = dev.call("invert_eo", dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
        S.barrier()
        t_dev = time.perf_counter() - t0
        its, err, t_cg = dev.solver_stats()
        cg = {"iterations": it, "time_to_solution_s": t_dev, "cg_loop_s": t_cg, "final_rr": err, "eps_sq": CG_EPS_SQ,
              "rel_prec": 1, "ms_per_iteration": 1e3 * t_cg / max(it, 1),
              "gflops_cg_1608_convention": ((2 * (2 * 1608.0 + 24) + 24 + max(it, 0) * (2 * (2 * 1608.0 + 24) + 120))
                                            * Vh * world / max(t_cg, 1e-9) / 1e9)}
        # mixed-precision CG (float inner solve, double defect correction; solver/mixed_cg_her.c:65)
        dev.call("field_zero", dOn)
        dev.call("invert_eo_mixed", dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
        S.barrier()
        t0 = time.perf_counter()
        itm = dev.call("invert_eo_mixed", dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
        S.barrier()
        cg["mixed_time_to_solution_s"] = time.perf_counter() - t0
        cg["mixed_count"] = itm
        cg["mixed_true_rr"] = dev.solver_stats()[1]
        if world == 1:
            # invert_eo's RGMIXEDCG branch (invert_eo.c:242-249): reliable-update CG, float inner loops, delta = operator.c:125's default
            try:
                dev.ck(lib.tmb_set_mcg_delta(5.0e-5))
                dev.call("invert_eo_rgmixed", dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
                S.barrier()
                t0 = time.perf_counter()
                itg = dev.call("invert_eo_rgmixed", dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
                S.barrier()
                cg["rgmixed_time_to_solution_s"] = time.perf_counter() - t0
                cg["rgmixed_count"] = itg
                cg["rgmixed_true_rr"] = dev.solver_stats()[1]
            except Exception as e:  # pragma: no cover
                cg["rgmixed_error"] = repr(e)[:200]
        # the same two solves with 12-real gauge compression (CompressionType COMPRESSION_12 of invert_eo)
        dev.ck(lib.tmb_set_compression(12))
        for name, fn in (("c12", "invert_eo"), ("c12_mixed", "invert_eo_mixed")):
            dev.call("field_zero", dOn)
            dev.call(fn, dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
            dev.call("field_zero", dOn)
            S.barrier()
            t0 = time.perf_counter()
            itc = dev.call(fn, dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
            S.barrier()
            cg[name + "_time_to_solution_s"] = time.perf_counter() - t0
            cg[name + "_count"] = itc
        dev.ck(lib.tmb_set_compression(18))
        if world == 1 and D is not None:
            hE, _ = pinned(dev, (Vh, 24)); hO, _ = pinned(dev, (Vh, 24)); hEn, _ = pinned(dev, (Vh, 24)); hOn, _ = pinned(dev, (Vh, 24))
            hE[:] = E; hO[:] = O; hOn[:] = 0
            # warm-up solve like the device-resident leg (first call allocates the drop-in's device fields)
            D.invert_eo(hEn, hOn, hE, hO, CG_EPS_SQ, CG_MAXITER, 1, 1, 0, 1, 0, None, tm.capi.SolverParams(), 0, 0, 0, 18)
            hOn[:] = 0
            t0 = time.perf_counter()
            it2 = D.invert_eo(hEn, hOn, hE, hO, CG_EPS_SQ, CG_MAXITER, 1, 1, 0, 1, 0, None, tm.capi.SolverParams(), 0, 0, 0, 18)
            cg["e2e_time_to_solution_s"] = time.perf_counter() - t0
            cg["e2e_iterations"] = it2
            cg["e2e_api"] = "invert_eo(...) drop-in, pinned host buffers (4 fields across PCIe), gauge already resident"
            cg["_e2e_solution"] = (np.array(hEn), np.array(hOn))
        out["cg"] = cg

    # ---- parity (outside every timed region) ----
    parity = None
    res_paths = {}
    if not args.skip_parity and not (args.loopback or args.loopback2):
        res_paths[S.path] = S.operators_and_solve(src, E, O)
    S.barrier()
    if D is not None:
        D.close()  # tmb_dropin_finalize: the drop-in layer's host state and the library context
    else:
        S.close()
    if world > 1 and not args.skip_parity:
        # the same global problem once more with NCCL halos and NCCL all-reduces (TMB_P2P=0), a short timing beside it
        S2 = Session(dims, world, rank, local_rank, dist, g, p2p=False, nz=nz)
        a0, a1, a2 = S2.dev.field(src), S2.dev.field(), S2.dev.field()
        S2.time_pairs(args.warmup, a0, a1, a2)
        ms_n, _ = S2.time_pairs(max(args.steps, 50), a0, a1, a2)
        res_paths[S2.path] = S2.operators_and_solve(src, E, O)
        res_paths[S2.path]["us_per_hop"] = 1e3 * ms_n / (2 * max(args.steps, 50))
        S2.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- rank 0 alone from here: trusted results and the comparison ----
    if res_paths and ref is not None:
        parity = {"tolerance_hop_rel_l2": TOL_HOP, "inputs": gauge_how}
        ref.set_params(KAPPA, GMU)
        gsrc, gE, gO = (G["srcs"] if world > 1 else srcs)
        t0 = time.perf_counter()
        trusted = {}
        for ieo in (0, 1):
            trusted[f"hop{ieo}"] = ref.spinor(); ref.Hopping_Matrix(ieo, trusted[f"hop{ieo}"], gsrc)
        trusted["qtm_pm"] = ref.spinor(); ref.Qtm_pm_psi(trusted["qtm_pm"], gsrc)
        parity["cpu_reference_operators_s"] = time.perf_counter() - t0
        n1 = None
        if world > 1:
            # the SAME global problem on ONE GPU (this rank's, the other ranks have gone): trusted CG count, residual, solution
            S1 = Session(gdims, 1, 0, local_rank, None, G["g"], p2p=True)
            n1 = S1.operators_and_solve(gsrc, gE, gO)
            S1.close()
            parity["n1_device"] = {"lattice_TxLXxLYxLZ": list(gdims), "cg_iters": n1["cg_iters"], "cg_rr": n1["cg_rr"],
                                   "cg_loop_s": n1["cg_loop_s"], "vs_cpu_reference": compare(n1, trusted, ("hop0", "hop1", "qtm_pm"))}
        ok_all = True
        for path, res in res_paths.items():
            rec = {"path": path, "vs_cpu_reference": compare(res, trusted, ("hop0", "hop1", "qtm_pm"))}
            rec["hop_rel_l2"] = max(rec["vs_cpu_reference"]["hop0"], rec["vs_cpu_reference"]["hop1"])
            rec["qtm_pm_rel_l2"] = rec["vs_cpu_reference"]["qtm_pm"]
            rec["cg_iters"], rec["cg_rr"], rec["cg_loop_s"] = res["cg_iters"], res["cg_rr"], res["cg_loop_s"]
            ok = rec["hop_rel_l2"] <= TOL_HOP and rec["qtm_pm_rel_l2"] <= TOL_HOP and res["cg_iters"] > 0
            if "us_per_hop" in res:
                rec["us_per_hop"] = res["us_per_hop"]
            if n1 is not None:
                rec["vs_n1_device"] = compare(res, n1, ("hop0", "hop1", "qtm_pm", "inv_e", "inv_o"))
                rec["cg_iters_n1_device"] = n1["cg_iters"]
                rec["cg_rr_rel_diff_vs_n1_device"] = abs(res["cg_rr"] - n1["cg_rr"]) / abs(n1["cg_rr"])
                ok = ok and abs(res["cg_iters"] - n1["cg_iters"]) <= 1 and max(rec["vs_n1_device"]["inv_e"], rec["vs_n1_device"]["inv_o"]) <= 1e-9
            rec["ok"] = bool(ok)
            ok_all = ok_all and ok
            parity[path] = rec
        first = parity[next(iter(res_paths))]
        parity.update({"hop_rel_l2": first["hop_rel_l2"], "cg_iters": first["cg_iters"], "path": first["path"]})
        if world == 1 and not args.skip_cpu and not args.skip_cg:
            # the reference's own invert_eo CG branch (cg_her on Qtm_pm_psi) on the same sources: its count and wall time
            en, on = ref.spinor(), ref.spinor()
            t0 = time.perf_counter()
            itr = ref.invert_eo_cg(en, on, gE, gO, CG_EPS_SQ, CG_MAXITER, 1)
            t_ref = time.perf_counter() - t0
            parity["cg_iters_cpu_reference"] = itr
            sol = res_paths[S.path]
            parity["cg_solution_rel_l2_vs_cpu_reference"] = max(rel_l2(sol["inv_e"], en), rel_l2(sol["inv_o"], on))
            ok_all = ok_all and abs(first["cg_iters"] - itr) <= 1 and parity["cg_solution_rel_l2_vs_cpu_reference"] <= 1e-9
            if "cg" in out:
                out["cg"]["cpu_reference_iterations"] = itr
                out["cg"]["cpu_reference_time_to_solution_s"] = t_ref
                out["cg"]["cpu_reference_how"] = (f"the unmodified reference's cg_her(&Qtm_pm_psi) inside the CG branch of invert_eo on the same "
                                                  f"sources, {ref.nthreads} threads, one run")
        elif n1 is not None:
            parity["cg_iters_ref"] = n1["cg_iters"]
            parity["cg_iters_ref_how"] = "one-GPU device solve of the same global problem (itself checked against the reference's cg_her at N = 1)"
        parity["ok"] = bool(ok_all)
        out["parity"] = parity
    elif not args.skip_parity:
        out["parity"] = {"ok": None, "why": "oracle/_ref not built: no trusted result on this box"}
    if "cg" in out and "_e2e_solution" in out["cg"]:
        hEn, hOn = out["cg"].pop("_e2e_solution")
        if ref is not None and not args.skip_cpu:
            # the reference's own end-to-end check |M x - b|^2 with the CPU operator (operator.c:358-384)
            r1, r2 = ref.spinor(), ref.spinor()
            ref.M_full(r1, r2, hEn, hOn)
            out["cg"]["true_residual_sq_cpu_M_full"] = float(np.sum((r1 - E) ** 2) + np.sum((r2 - O) ** 2))
            out["cg"]["source_norm_sq"] = float(np.sum(E ** 2) + np.sum(O ** 2))

    if not args.skip_cpu and ref is not None and world == 1:
        out["cpu_baseline"] = cpu_baseline(ref, dims)
    if not args.skip_anchor:
        if tuple(dims) == ANCHOR_DIMS and world == 1:
            out["weak_anchor"] = {"lattice_TxLXxLYxLZ": list(dims), "n_gpus": 1, "value": gflops, "unit": "GFLOP/s", "note": "this run"}
        else:
            out["weak_anchor"] = run_anchor(max(args.steps, 200), args.warmup)
            if world > 1 and "value" in out["weak_anchor"]:
                out["weak_anchor"]["efficiency_at_equal_volume"] = gflops / (world * out["weak_anchor"]["value"])
    # ---- BASELINE configs[0], configs[3] and configs[4] on their own lattices, each in its own process (scripts/bench_sections.py) ----
    if world == 1 and not args.skip_sections:
        for name in ("small", "nd", "hmc"):
            key = "benchmark_8x8x8x8" if name == "small" else name
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "bench_sections.py"), name],
                                   capture_output=True, text=True, timeout=600)
                out[key] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-400:]}
            except Exception as e:  # pragma: no cover
                out[key] = {"error": repr(e)}
    print(json.dumps(out), file=real_stdout, flush=True)


if __name__ == "__main__":
    main()

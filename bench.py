#!/usr/bin/env python
"""bench.py - Hopping_Matrix throughput (+ eo-CG time-to-solution) on B200, BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (oracle/_ref)

A step is one EO+OE pair of Hopping_Matrix calls over the whole local lattice, exactly the loop
body of the reference's benchmark.c:293-299.  Workload: N=1 -> BASELINE configs[1] lattice
(24^3 x 48, kappa=0.16, mu=0.01, random SU(3) gauge); N>1 -> configs[2]: 48^3 x (12 N) split
along T, 48^3 x 12 per GPU (weak scaling), T-neighbour fields read over NVLink (peer mode; NCCL halos as fallback).
One JSON line on stdout (rank 0).

Keys beyond the driver's contract: `roofline` (achieved = 1536 B x sites / mean launch time, peak = MEASURED_PEAKS.json,
`traffic` = DRAM bytes per launch from the committed ncu capture, `copy_gbs_sustained_this_run` = device-to-device copy
bandwidth sustained in this process), `e2e` (the same pairs through the reference-named Hopping_Matrix() with pinned HOST
buffers, at every N), `cg` (invert_eo time to solution: device-resident, through the host-pointer drop-in, mixed precision,
12-real links, true residual by the CPU M_full), `cpu_baseline` (the unmodified reference on all host cores), `compression12`,
and one section per remaining BASELINE config, each in its own process (scripts/bench_sections.py): `benchmark_8x8x8x8`
(configs[0]), `nd` (configs[3]), `hmc` (configs[4]).  `--lattice TxLXxLYxLZ --global-chunk-t 12` gives the strong-scaling
series of configs[2] on one global problem (scripts/gpu_strong.sh).
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
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_gauge(dims, seed):
    """random SU(3) gauge (hot start).  With oracle/_ref present this is the reference's own
    start_ranlux(1,123456); random_gauge_field (benchmark.c:247-248), else numpy QR."""
    V = int(np.prod(dims))
    if seed == 123456:
        try:
            from oracle import refclient
            if refclient.available():
                ref = refclient.Reference(*dims, nthreads=os.cpu_count() or 1)
                return ref.random_gauge(123456), ref, "reference ranlux random_gauge_field(seed 123456)"
        except Exception as e:  # pragma: no cover
            log("reference generator unavailable:", e)
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(V * 4, 3, 3)) + 1j * rng.normal(size=(V * 4, 3, 3))
    q, r = np.linalg.qr(a)
    dg = np.diagonal(r, axis1=1, axis2=2)
    q = q * (dg / np.abs(dg))[:, None, :]
    q = q / np.linalg.det(q)[:, None, None] ** (1.0 / 3.0)
    g = np.ascontiguousarray(q.reshape(V * 4, 9)).view(np.float64).reshape(V, 4, 18)
    return g, None, f"numpy QR random SU(3) (seed {seed})"


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
        g, _, _ = make_gauge((chunk_t, LX, LY, LZ), 7000 + gc)
        gs.append(g)
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
        "config": {"workload": workload_name(args.gpus, dims), "lattice_TxLXxLYxLZ": list(dims), "kappa": KAPPA, "mu": MU},
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


def cpu_baseline(dims, gauge_ref, target_s=12.0):
    """oracle/_ref timed on the host cores, bounded sample of the same workload"""
    from oracle import refclient
    ncores = os.cpu_count() or 1
    if not refclient.available(halfspinor=True):
        return None
    # a second library instance (half-spinor build) in the same process: separate C globals
    ref = refclient.Reference(*dims, nthreads=ncores, halfspinor=True)
    ref.set_params(KAPPA, GMU)
    if gauge_ref is not None:
        ref.set_gauge(gauge_ref)
    else:
        ref.random_gauge(123456)
    V = int(np.prod(dims))
    t1 = ref.bench_hopping(1)
    n = int(max(2, min(200, target_s / max(t1, 1e-6))))
    t = ref.bench_hopping(n)
    # CG: per-application cost of Qtm_pm_psi, the reference's cg_her is 1 application + BLAS-1 per iteration
    tq1 = ref.bench_Qtm_pm(1)
    nq = int(max(2, min(100, 0.5 * target_s / max(tq1, 1e-6))))
    tq = ref.bench_Qtm_pm(nq) / nq
    return {"value": V * FLOP_SITE * n / t / 1e9, "unit": "GFLOP/s", "cores": ref.nthreads, "kind": "reference",
            "sample": f"{n} EO+OE Hopping_Matrix pairs on the same lattice, half-spinor OpenMP build of the unmodified "
                      f"reference (oracle/_ref), {ref.nthreads} threads",
            "gflops_1608": V * FLOP_SITE_REF * n / t / 1e9, "ms_per_pair": 1e3 * t / n,
            "qtm_pm_psi_s": tq}


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
    ap.add_argument("--skip-sections", action="store_true", help="skip the configs[3] (ND doublet) and configs[4] (HMC monomials) sections")
    ap.add_argument("--loopback", action="store_true", help="1 GPU: run the T-split halo/boundary path against itself")
    ap.add_argument("--loopback2", action="store_true", help="1 GPU: run the T-split peer-mode path against itself")
    ap.add_argument("--sweep", action="store_true", help="time every kernel variant (tuning aid, prints to stderr)")
    ap.add_argument("--sweep-overlap", action="store_true", help="time PDL / L2-prefetch combinations (stderr)")
    ap.add_argument("--overlap", type=int, default=0, help="tmb_set_overlap flags: 1 PDL, 2 L2 gauge prefetch")
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
        dims = (48, 24, 24, 24) if args.gpus == 1 else (12, 48, 48, 48)

    if args.impl == "reference":
        run_reference(args, dims, real_stdout)
        return

    import tmlqcd_b200 as tm
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local_rank)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_

    dev = tm.Device(*dims, device=local_rank)
    lib = dev.lib
    if world > 1:
        import torch
        idbuf = (C.c_ubyte * 128)()
        if rank == 0:
            dev.ck(lib.tmb_comm_unique_id(C.cast(idbuf, C.c_void_p)))
        t_id = torch.tensor(list(idbuf), dtype=torch.uint8, device="cuda")
        dist.broadcast(t_id, 0)
        idbuf = (C.c_ubyte * 128)(*t_id.cpu().tolist())
        dev.ck(lib.tmb_comm_init(C.cast(idbuf, C.c_void_p), world, rank))

    V, Vh = dev.V, dev.Vh
    chunk_src = None
    if args.global_chunk_t:
        g, chunk_src, gauge_how = chunked_fields(dims, rank, args.global_chunk_t)
        ref = None
    else:
        g, ref, gauge_how = make_gauge(dims, 123456 if world == 1 else 1000 + rank)
    dev.set_params(KAPPA, GMU)
    if args.variant is not None or args.hints is not None or args.xblock is not None:
        dev.ck(lib.tmb_set_tuning(-1 if args.variant is None else args.variant, -1 if args.hints is None else args.hints, args.xblock or 0))
    dev.ck(lib.tmb_set_overlap(args.overlap))
    if (args.loopback or args.loopback2) and world == 1:
        dev.ck(lib.tmb_comm_loopback(2 if args.loopback2 else 1))
    dev.gauge_upload(g)
    rng = np.random.default_rng(99 + rank)
    src = chunk_src[0] if chunk_src else rng.normal(scale=np.sqrt(0.5), size=(Vh, 24))
    f0, f1, f2 = dev.field(src), dev.field(), dev.field()

    def barrier():
        dev.ck(lib.tmb_sync())
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def time_pairs(nsteps):
        barrier()
        n0 = lib.tmb_launch_count()
        dev.timer_start()
        for _ in range(nsteps):
            lib.tmb_Hopping_Matrix(0, f1, f0)
            lib.tmb_Hopping_Matrix(1, f2, f1)
        ms = dev.timer_stop()
        barrier()
        if dist is not None:
            import torch
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, lib.tmb_launch_count() - n0

    if args.sweep and rank == 0:
        results = []
        xbs = [0] + [x for x in (2, 4, 8) if dims[1] % x == 0]
        for hints in (1, 0):
            for xb in xbs:
                for variant in range(0, 10):
                    dev.ck(lib.tmb_set_tuning(variant, hints, xb))
                    time_pairs(5)
                    ms, _ = time_pairs(50)
                    per = ms / 100.0
                    results.append((per, variant, hints, xb))
                    log(f"sweep variant={variant} hints={hints} xblock={xb}: {per * 1e3:8.2f} us/hop "
                        f"{Vh * BYTES_SITE / per / 1e6:8.1f} GB/s {Vh * FLOP_SITE / per / 1e6:8.1f} GFLOP/s")
        results.sort()
        log("best:", results[:5])
        dev.ck(lib.tmb_set_tuning(-1 if args.variant is None else args.variant, -1 if args.hints is None else args.hints, args.xblock or 0))

    if args.sweep_overlap and rank == 0:
        for flags in (0, 1, 2, 3):
            for variant in (0, 2):
                dev.ck(lib.tmb_set_tuning(variant, 1, 0)); dev.ck(lib.tmb_set_overlap(flags))
                time_pairs(10)
                ms, _ = time_pairs(200)
                per = ms / 400.0
                log(f"overlap flags={flags} (pdl={flags & 1} prefetch={flags >> 1}) variant={variant}: {per * 1e3:8.2f} us/hop "
                    f"{Vh * BYTES_SITE / per / 1e6:8.1f} GB/s")
        dev.ck(lib.tmb_set_tuning(-1 if args.variant is None else args.variant, -1 if args.hints is None else args.hints, args.xblock or 0))
        dev.ck(lib.tmb_set_overlap(args.overlap))
    # ---- timed region: K pairs, device resident, inputs larger than L2 (gauge alone is 1152 B/site) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # let nvidia-smi come up so that its samples fall inside the loaded window
    time_pairs(args.warmup)
    ms, launches = time_pairs(args.steps)
    # nvidia-smi samples every 100 ms; a short timed region yields too few samples, so the same step
    # keeps running (untimed) until the sampler has seen ~1.5 s of this load
    extra = int(max(0, (1.5e3 - ms) / max(ms / args.steps, 1e-3)))
    if extra > 0:
        time_pairs(extra)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = f"warm-up + {args.steps} timed steps + {extra} identical untimed steps"
    sites = V * world  # output sites per pair over all ranks (V/2 per call)
    gflops = sites * FLOP_SITE * args.steps / (ms * 1e-3) / 1e9
    peak, peak_how = measured_peaks()
    per_launch_ms = ms / (2 * args.steps)
    achieved = Vh * BYTES_SITE / (per_launch_ms * 1e-3) / 1e9  # per GPU

    out = {
        "metric": "Hopping_Matrix GFLOP/s (eo, double, 1320 flop/site)", "value": gflops, "unit": "GFLOP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(world, dims), "lattice_TxLXxLYxLZ": list(dims), "kappa": KAPPA, "mu": MU,
                   "gauge": gauge_how, "l2": "inputs larger than L2: gauge field %.0f MB + spinors per call" % (V * 4 * 144 / 1e6),
                   "step": "one EO+OE Hopping_Matrix pair (benchmark.c:293-299)"},
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
    # the same process, the same thermal state: device-to-device copy bandwidth sustained over ~0.5 s (read + write bytes),
    # next to the burst figure of MEASURED_PEAKS.json that `peak` quotes
    try:
        gbs = C.c_double(0.)
        dev.ck(lib.tmb_measure_copy_gbs(1 << 30, 1500, C.byref(gbs)))
        out["roofline"]["copy_gbs_sustained_this_run"] = gbs.value
        out["roofline"]["frac_of_sustained_copy"] = achieved / gbs.value
    except Exception as e:  # pragma: no cover
        out["roofline"]["copy_gbs_sustained_this_run"] = None
        print("copy bandwidth measurement failed:", e, file=sys.stderr)
    tr = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tr):
        try:
            t = json.load(open(tr))
            if t.get("lattice") == list(dims):
                out["roofline"]["traffic"] = t["dram_bytes_per_launch"]
        except Exception:
            pass

    # ---- optional variant: 12-real compressed links (1152 algorithmic B/site), reported separately ----
    dev.ck(lib.tmb_set_compression(12))
    time_pairs(args.warmup)
    ms12, _ = time_pairs(max(args.steps // 4, 10))
    dev.ck(lib.tmb_set_compression(18))
    per12 = ms12 / (2 * max(args.steps // 4, 10))
    out["compression12"] = {"us_per_hop": per12 * 1e3, "gflops_1320": Vh * FLOP_SITE / (per12 * 1e-3) / 1e9 * world,
                            "algorithmic_bytes_per_site": 1152, "hbm_gbs_effective_per_gpu": Vh * 1152.0 / (per12 * 1e-3) / 1e9,
                            "note": "tmb_set_compression(12): two link rows streamed, third rebuilt in registers; "
                                    "not the headline (value uses the reference's 18-real links)"}

    # ---- e2e: the reference-named Hopping_Matrix(ieo, l, k) with HOST buffers, copies inside the timing ----
    # (N > 1: every rank moves its own slab across its own PCIe link; the hop runs the T-split path)
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
        barrier()
        n0 = lib.tmb_launch_count()
        t0 = time.perf_counter()
        for _ in range(ne):
            D.Hopping_Matrix(0, h1, hk); D.Hopping_Matrix(1, h2, h1)
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            import torch
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        out["e2e"] = {"value": V * world * FLOP_SITE * ne / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": 2 * Vh * 192 * world,
                      "d2h_bytes_per_step": 2 * Vh * 192 * world, "steps": ne, "ms_per_step": 1e3 * dt / ne,
                      "api": "Hopping_Matrix(ieo, spinor* l, spinor* k) drop-in, pinned host buffers, upload+kernel+download per call"}
        out["gpu_launches"] += int(lib.tmb_launch_count() - n0)

    # ---- eo-CG time-to-solution (the configs[1] solve): device-resident and through invert_eo with host buffers ----
    if not args.skip_cg:
        E = rng.normal(scale=np.sqrt(0.5), size=(Vh, 24)); O = rng.normal(scale=np.sqrt(0.5), size=(Vh, 24))
        if chunk_src:
            E, O = chunk_src[1], chunk_src[2]
        dE, dO, dEn, dOn = dev.field(E), dev.field(O), dev.field(), dev.field()
        dev.call("invert_eo", dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)  # warm-up solve
        dev.call("field_zero", dOn)
        barrier()
        t0 = time.perf_counter()
        it = dev.call("invert_eo", dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
        barrier()
        t_dev = time.perf_counter() - t0
        its, err, t_cg = dev.solver_stats()
        cg = {"iterations": it, "time_to_solution_s": t_dev, "cg_loop_s": t_cg, "final_rr": err, "eps_sq": CG_EPS_SQ,
              "rel_prec": 1, "gflops_cg_1608_convention": ((2 * (2 * 1608.0 + 24) + 24 + max(it, 0) * (2 * (2 * 1608.0 + 24) + 120))
                                                             * Vh * world / max(t_cg, 1e-9) / 1e9)}
        # mixed-precision CG (float inner solve, double defect correction; solver/mixed_cg_her.c:65)
        dev.call("field_zero", dOn)
        dev.call("invert_eo_mixed", dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
        barrier()
        t0 = time.perf_counter()
        itm = dev.call("invert_eo_mixed", dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
        barrier()
        cg["mixed_time_to_solution_s"] = time.perf_counter() - t0
        cg["mixed_count"] = itm
        cg["mixed_true_rr"] = dev.solver_stats()[1]
        # the same two solves with 12-real gauge compression (CompressionType COMPRESSION_12 of invert_eo)
        dev.ck(lib.tmb_set_compression(12))
        for name, fn in (("c12", "invert_eo"), ("c12_mixed", "invert_eo_mixed")):
            dev.call("field_zero", dOn)
            dev.call(fn, dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
            dev.call("field_zero", dOn)
            barrier()
            t0 = time.perf_counter()
            itc = dev.call(fn, dEn, dOn, dE, dO, CG_EPS_SQ, CG_MAXITER, 1)
            barrier()
            cg[name + "_time_to_solution_s"] = time.perf_counter() - t0
            cg[name + "_count"] = itc
        dev.ck(lib.tmb_set_compression(18))
        if world == 1 and not args.skip_e2e:
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
            if not args.skip_cpu:
                # the reference's own end-to-end check |M x - b|^2 with the CPU operator (operator.c:358-384)
                from oracle.oracleclient import Oracle
                subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], capture_output=True)
                o = Oracle(*dims); o.set_gauge(g); o.set_params(KAPPA, GMU)
                r1, r2 = o.spinor(), o.spinor()
                o.M_full(r1, r2, np.array(hEn), np.array(hOn))
                cg["true_residual_sq_cpu_M_full"] = float(np.sum((r1 - E) ** 2) + np.sum((r2 - O) ** 2))
                cg["source_norm_sq"] = float(np.sum(E ** 2) + np.sum(O ** 2))
        out["cg"] = cg

    if not args.skip_cpu and rank == 0 and world == 1:
        cb = cpu_baseline(dims, g)
        if cb:
            out["cpu_baseline"] = cb
            if "cg" in out and out["cg"]["iterations"] > 0:
                out["cg"]["cpu_reference_time_to_solution_s_est"] = cb["qtm_pm_psi_s"] * (out["cg"]["iterations"] + 1)
                out["cg"]["cpu_reference_est_how"] = ("reference Qtm_pm_psi time per application x (iterations+1); "
                                                      "lower bound, BLAS-1 of cg_her not included")
    dev.close()
    # ---- BASELINE configs[0], configs[3] and configs[4] on their own lattices, each in its own process (scripts/bench_sections.py) ----
    if rank == 0 and world == 1 and not args.skip_sections:
        for name in ("small", "nd", "hmc"):
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "bench_sections.py"), name],
                                   capture_output=True, text=True, timeout=600)
                key = "benchmark_8x8x8x8" if name == "small" else name
                out[key] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-400:]}
            except Exception as e:  # pragma: no cover
                out[key] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(out), file=real_stdout, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

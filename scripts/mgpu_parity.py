#!/usr/bin/env python
"""Multi-GPU parity: torchrun --nproc-per-node N scripts/mgpu_parity.py
The GLOBAL lattice (T = N*T_loc) is generated identically on every rank; rank r takes its T-slab
(the reference's PARALLELT decomposition, mpi_init.c:321), runs the distributed operators / solvers (peer mode by
default; TMB_P2P=0: NCCL half-spinor halos and all-reduces; TMB_XRED=0: peer hops with NCCL all-reduces), and rank 0 compares the gathered result with the CPU oracle on the global
lattice (the decomposition-independence recipe of SURVEY 4: reproducible global fields)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tmlqcd_b200 as tm  # noqa: E402
from conftest import random_gauge, random_spinor, rel_l2  # noqa: E402
from oracle.oracleclient import Oracle  # noqa: E402

KAPPA, GMU, THETA = 0.16, 0.0032, (1., 0., 0.3, 0.)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    args = [a for a in sys.argv[1:] if not a.startswith("--grid")]
    grid = [a for a in sys.argv[1:] if a.startswith("--grid")]
    nt, nz = (int(x) for x in grid[0].split("=")[1].split("x")) if grid else (world, 1)
    assert nt * nz == world
    ct, cz = rank // nz, rank % nz
    Tl, LX, LY, LZl = (int(x) for x in (args[0] if args else "8x8x8x8").split("x"))
    T, LZ = Tl * nt, LZl * nz
    rng = np.random.default_rng(2024)
    V, Vh = T * LX * LY * LZ, T * LX * LY * LZ // 2
    g = random_gauge(rng, V)
    k, p = random_spinor(rng, Vh), random_spinor(rng, Vh)

    # the slab of rank (ct, cz): t is the slowest index of both orderings, z the fastest (eo fields: z/2 with LZl even)
    def slab_eo(f):
        return np.ascontiguousarray(f.reshape(T, LX * LY, LZ // 2, -1)[ct * Tl:(ct + 1) * Tl, :, cz * (LZl // 2):(cz + 1) * (LZl // 2)]).reshape(-1, f.shape[-1])

    def slab_lex(f):
        return np.ascontiguousarray(f.reshape(T, LX * LY, LZ, -1)[ct * Tl:(ct + 1) * Tl, :, cz * LZl:(cz + 1) * LZl]).reshape((-1,) + f.shape[1:])

    def unslab(parts, zdiv, tail):
        """inverse of slab_*: parts[rank] -> global array; zdiv = 2 for eo fields, 1 for lexicographic ones"""
        out = np.zeros((T, LX * LY, LZ // zdiv) + tail)
        for r, part in enumerate(parts):
            a, b = r // nz, r % nz
            out[a * Tl:(a + 1) * Tl, :, b * (LZl // zdiv):(b + 1) * (LZl // zdiv)] = part.reshape((Tl, LX * LY, LZl // zdiv) + tail)
        return out
    d = tm.Device(Tl, LX, LY, LZl, device=lr)
    idbuf = (C.c_ubyte * 128)()
    if rank == 0:
        d.ck(d.lib.tmb_comm_unique_id(C.cast(idbuf, C.c_void_p)))
    t_id = torch.tensor(list(idbuf), dtype=torch.uint8, device="cuda")
    dist.broadcast(t_id, 0)
    idbuf = (C.c_ubyte * 128)(*t_id.cpu().tolist())
    d.ck(d.lib.tmb_comm_init_grid(C.cast(idbuf, C.c_void_p), nt, nz, rank))
    d.set_params(KAPPA, GMU, THETA)
    d.gauge_upload(slab_lex(g))
    if rank == 0:
        print(f"grid {nt} x {nz} (T x Z), peer mode:", bool(d.lib.tmb_comm_peer_mode()), "z faces through peer memory:", bool(d.lib.tmb_comm_zpeer_mode()), flush=True)

    def gather(field):
        loc = torch.from_numpy(d.download(field)).cuda()
        out = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(out, loc)
        return unslab([o_.cpu().numpy() for o_ in out], 2, (24,)).reshape(-1, 24)

    dk, dp, dl, dx = d.field(slab_eo(k)), d.field(slab_eo(p)), d.field(), d.field()
    res = {}
    plaq = C.c_double(0.)
    if nz == 1:
        d.ck(d.lib.tmb_measure_plaquette(C.byref(plaq)))  # all-reduced: the same value on every rank
    for ieo in (0, 1):
        d.call("Hopping_Matrix", ieo, dl, dk); res[f"hop{ieo}"] = gather(dl)
        d.call("tm_sub_Hopping_Matrix", ieo, dl, dp, dk, 1.0, 0.3); res[f"tm_sub{ieo}"] = gather(dl)
    d.call("Qtm_pm_psi", dl, dk); res["Qtm_pm"] = gather(dl)
    sq = d.reduce("square_norm", dk)
    it = d.call("cg_her", dx, dk, 2000, 1e-22, 1); res["cg_x"] = gather(dx)
    dE, dO, dEn, dOn = d.field(slab_eo(k)), d.field(slab_eo(p)), d.field(), d.field()
    it2 = d.call("invert_eo", dEn, dOn, dE, dO, 1e-22, 2000, 1)
    res["inv_e"], res["inv_o"] = gather(dEn), gather(dOn)
    # mixed-precision solvers (float hops through the same halo machinery)
    itm = d.call("mixed_cg_her", dx, dk, 2000, 1e-22, 1); res["mcg_x"] = gather(dx)
    itg = d.call("rg_mixed_cg_her", dx, dk, 2000, 1e-22, 1); res["rg_x"] = gather(dx)
    # non-degenerate doublet
    d.ck(d.lib.tmb_set_nd(0.139, 0.15, 0.9))
    dls, dlc = d.field(), d.field()
    d.call("Qtm_pm_ndpsi", dls, dlc, dk, dp); res["nd_s"], res["nd_c"] = gather(dls), gather(dlc)
    d.call("field_zero", dls); d.call("field_zero", dlc)
    itn = d.call("cg_her_nd", dls, dlc, dk, dp, 2000, 1e-20, 1); res["cgnd_s"], res["cgnd_c"] = gather(dls), gather(dlc)
    d.ck(d.lib.tmb_set_mcg_delta(0.1))
    hmc = nz == 1  # the float two-flavour solver and the plaquette are T-split only; the force and the monomials run on T x Z grids too
    itr_nd = None
    if hmc:
        itr_nd = d.call("rg_mixed_cg_her_nd", dls, dlc, dk, dp, 2000, 1e-20, 1); res["rgnd_s"], res["rgnd_c"] = gather(dls), gather(dlc)
    # fermion force and a det monomial with chronological guess (deriv_Sb exchanges the projected first slices)

    def gather_df():
        loc = torch.from_numpy(d.derivative_download()).cuda()
        out = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(out, loc)
        return unslab([o_.cpu().numpy() for o_ in out], 1, (4, 8)).reshape(-1, 4, 8)
    d.call("derivative_zero")
    d.call("deriv_Sb", 0, dk, dp, 0.7); d.call("deriv_Sb", 1, dp, dk, -0.4)
    res["df"] = gather_df()
    margs = (0, 0.15, 0.01, 0.15, 0.05, 1, 3000, 1e-20, 1e-22, 2)
    assert d.lib.tmb_monomial_add(*margs) == 0
    e0 = C.c_double()
    d.ck(d.lib.tmb_monomial_heatbath(0, dk, C.byref(e0)))
    d.call("derivative_zero")
    for _ in range(2):
        d.ck(d.lib.tmb_monomial_derivative(0))
    res["mnl_df"] = gather_df()
    dH = C.c_double(); d.ck(d.lib.tmb_monomial_acc(0, C.byref(dH)))
    minfo = d.monomial_info(0)
    return finish(d, rank, world, T, LX, LY, LZ, g, k, p, Vh, res, sq, plaq, it, it2, itm, itg, itn, itr_nd, e0, dH, minfo, margs, nz)


def finish(d, rank, world, T, LX, LY, LZ, g, k, p, Vh, res, sq, plaq, it, it2, itm, itg, itn, itr_nd, e0, dH, minfo, margs, nz):
    hmc = nz == 1
    d.set_params(KAPPA, GMU, THETA)
    ok = True
    if rank == 0:
        o = Oracle(T, LX, LY, LZ)
        o.set_gauge(g); o.set_params(KAPPA, GMU, THETA)
        e = o.spinor()
        for ieo in (0, 1):
            o.Hopping_Matrix(ieo, e, k); r1 = rel_l2(res[f"hop{ieo}"], e)
            o.tm_sub_Hopping_Matrix(ieo, e, p, k, 1.0, 0.3); r2 = rel_l2(res[f"tm_sub{ieo}"], e)
            print(f"hop{ieo} rel {r1:.2e}  tm_sub{ieo} rel {r2:.2e}"); ok &= r1 <= 1e-13 and r2 <= 1e-13
        o.Qtm_pm_psi(e, k); r = rel_l2(res["Qtm_pm"], e); print(f"Qtm_pm rel {r:.2e}"); ok &= r <= 1e-13
        r = abs(sq / o.square_norm(k, Vh) - 1); print(f"global square_norm rel {r:.2e}"); ok &= r < 1e-14
        if hmc:
            r = abs(plaq.value / o.measure_plaquette() - 1); print(f"measure_plaquette rel {r:.2e}"); ok &= r < 1e-13
        x = o.spinor(); itr = o.cg_her(x, k, 2000, 1e-22, 1); r = rel_l2(res["cg_x"], x)
        print(f"cg_her iters {it} (oracle {itr}) x rel {r:.2e}"); ok &= abs(it - itr) <= 1 and r <= 1e-10
        en, on = o.spinor(), o.spinor(); itr = o.invert_eo_cg(en, on, k, p, 1e-22, 2000, 1)
        r1, r2 = rel_l2(res["inv_e"], en), rel_l2(res["inv_o"], on)
        print(f"invert_eo iters {it2} (oracle {itr}) rel {r1:.2e} {r2:.2e}"); ok &= abs(it2 - itr) <= 1 and max(r1, r2) <= 1e-10
        r1, r2 = rel_l2(res["mcg_x"], x), rel_l2(res["rg_x"], x)
        print(f"mixed_cg_her count {itm}, rg_mixed_cg_her count {itg}: x rel {r1:.2e} {r2:.2e}"); ok &= itm > 0 and itg > 0 and max(r1, r2) <= 1e-8
        o.set_nd_params(0.139, 0.15, 0.9)
        es, ec = o.spinor(), o.spinor(); o.Qtm_pm_ndpsi(es, ec, k, p)
        r1, r2 = rel_l2(res["nd_s"], es), rel_l2(res["nd_c"], ec); print(f"Qtm_pm_ndpsi rel {r1:.2e} {r2:.2e}"); ok &= max(r1, r2) <= 1e-13
        es[:] = 0; ec[:] = 0; itr = o.cg_her_nd(es, ec, k, p, 2000, 1e-20, 1)
        r1, r2 = rel_l2(res["cgnd_s"], es), rel_l2(res["cgnd_c"], ec)
        print(f"cg_her_nd iters {itn} (oracle {itr}) rel {r1:.2e} {r2:.2e}"); ok &= abs(itn - itr) <= 1 and max(r1, r2) <= 1e-9
        if hmc:
            r1, r2 = rel_l2(res["rgnd_s"], es), rel_l2(res["rgnd_c"], ec)
            print(f"rg_mixed_cg_her_nd count {itr_nd}: x rel {r1:.2e} {r2:.2e}"); ok &= itr_nd > 0 and max(r1, r2) <= 1e-8
        df = o.derivative(); o.deriv_Sb(0, k, p, df, 0.7); o.deriv_Sb(1, p, k, df, -0.4)
        r = rel_l2(res["df"], df); print(f"deriv_Sb rel {r:.2e}"); ok &= r <= 1e-13
        o.mnl_clear(); assert o.mnl_add(*margs) == 0
        e0r = o.mnl_heatbath(0, k); dfo = o.derivative()
        for _ in range(2):
            o.mnl_derivative(0, dfo)
        dHr = o.mnl_acc(0); oinfo = o.mnl_info(0)
        r = rel_l2(res["mnl_df"], dfo)
        print(f"det monomial: energy0 rel {abs(e0.value / e0r - 1):.2e}, derivative rel {r:.2e}, iter1 {minfo['iter1']} (oracle {oinfo['iter1']}), "
              f"dH {dH.value:.2e} (oracle {dHr:.2e})")
        ok &= abs(e0.value / e0r - 1) <= 1e-13 and r <= 1e-8 and abs(minfo["iter1"] - oinfo["iter1"]) <= 3 and abs(dH.value - dHr) <= 1e-7
        print("MGPU PARITY", "OK" if ok else "FAILED", f"world={world} global={T}x{LX}x{LY}x{LZ}" + ("" if hmc else " (Z split: operators, solvers, force, det monomial)"))
    return close(d, ok)


def close(d, ok):
    d.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()

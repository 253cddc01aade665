"""world_size-2 CPU test (gloo) of the N>1 path's host-side logic: T-slab decomposition of the global
fields, half-spinor face packing, the send/recv pattern of exchange_faces() (send_up -> rank+1's
halo_dn, send_dn -> rank-1's halo_up, tmb_capi.cu), the one-off gauge halo, the interior/boundary
site ranges, the first-slice link halo of the plaquette, and the global sum of reductions.  Compute on each rank is the PRODUCT's site code
compiled for the host (tests/emul); the check is the oracle on the global lattice."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, dims_loc, theta, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import random_gauge, random_spinor
    from emul_client import Emul
    Tl, LX, LY, LZ = dims_loc
    T = Tl * world
    rng = np.random.default_rng(77)  # identical global fields on every rank
    V, Vh = T * LX * LY * LZ, T * LX * LY * LZ // 2
    g, k, p = random_gauge(rng, V), random_spinor(rng, Vh), random_spinor(rng, Vh)
    Vl, Vhl = V // world, Vh // world
    e = Emul(Tl, LX, LY, LZ)
    U = e.pack_gauge(g[rank * Vl:(rank + 1) * Vl])
    sk, sp = e.pack(k[rank * Vhl:(rank + 1) * Vhl]), e.pack(p[rank * Vhl:(rank + 1) * Vhl])
    up_r, dn_r = (rank + 1) % world, (rank - 1) % world

    def exchange(send_up, send_dn):
        """same pairing as exchange_faces(): what I send up is my upper neighbour's halo_dn"""
        halo_dn, halo_up = torch.empty(send_up.size, dtype=torch.float64), torch.empty(send_dn.size, dtype=torch.float64)
        ops = [dist.P2POp(dist.isend, torch.from_numpy(send_up), up_r, tag=1), dist.P2POp(dist.irecv, halo_dn, dn_r, tag=1),
               dist.P2POp(dist.isend, torch.from_numpy(send_dn), dn_r, tag=2), dist.P2POp(dist.irecv, halo_up, up_r, tag=2)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return halo_up.numpy(), halo_dn.numpy()

    # one-off gauge halo: U_0 of rank-1's last slice
    gsend = e.pack_gauge_halo(U)
    Uhalo = torch.empty(gsend.size, dtype=torch.float64)
    for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, torch.from_numpy(gsend), up_r), dist.P2POp(dist.irecv, Uhalo, dn_r)]):
        w.wait()
    ka = np.stack([0.16 * np.cos(np.array(theta) * 3.14159265358979 / np.array([T, LX, LY, LZ])),
                   0.16 * np.sin(np.array(theta) * 3.14159265358979 / np.array([T, LX, LY, LZ]))], axis=1).reshape(-1)
    outs = {}
    for par in (0, 1):
        su, sd = e.pack_halo(sk)
        hu, hd = exchange(su, sd)
        outs[f"hop{par}"] = e.unpack(e.hop(par, sk, U, ka, 0, halo=(hu, hd, Uhalo.numpy())))
        outs[f"tm_sub{par}"] = e.unpack(e.hop(par, sk, U, ka, 2, (1.0, 0.3), sp, halo=(hu, hd, Uhalo.numpy())))
    # plaquette: the spatial links of my FIRST slice go to rank-1, I receive rank+1's (tmb_measure_plaquette)
    psend = e.pack_gauge_first_slice(U)
    pup = torch.empty(psend.size, dtype=torch.float64)
    for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, torch.from_numpy(psend), dn_r, tag=3), dist.P2POp(dist.irecv, pup, up_r, tag=3)]):
        w.wait()
    plaq = torch.tensor([e.plaquette(U, pup.numpy())], dtype=torch.float64)
    dist.all_reduce(plaq)  # the ncclAllReduce behind finish_reduction
    nrm = torch.tensor([float(np.sum(k[rank * Vhl:(rank + 1) * Vhl] ** 2))], dtype=torch.float64)
    dist.all_reduce(nrm)  # the ncclAllReduce of tmb_square_norm
    gathered = {}
    for name, loc in outs.items():
        lst = [torch.empty(loc.shape, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(lst, torch.from_numpy(np.ascontiguousarray(loc)))
        gathered[name] = torch.cat(lst).numpy()
    if rank == 0:
        q.put((gathered, float(nrm.item()), g, k, p, float(plaq.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dims_loc,theta", [((4, 4, 4, 4), (1., 0., 0., 0.)), ((2, 4, 6, 4), (0., 0.5, 0., 1.))])
def test_two_rank_T_split_matches_global_oracle(oracle_lib, dims_loc, theta):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, dims_loc, theta, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    gathered, nrm, g, k, p, plaq = q.get(timeout=180)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    Tl, LX, LY, LZ = dims_loc
    o = oracle_lib.Oracle(Tl * world, LX, LY, LZ)
    o.set_gauge(g); o.set_params(0.16, 0.0, theta)
    exp = o.spinor()
    for par in (0, 1):
        o.Hopping_Matrix(par, exp, k)
        assert np.linalg.norm(gathered[f"hop{par}"] - exp) / np.linalg.norm(exp) < 1e-14
        o.tm_sub_Hopping_Matrix(par, exp, p, k, 1.0, 0.3)
        assert np.linalg.norm(gathered[f"tm_sub{par}"] - exp) / np.linalg.norm(exp) < 1e-14
    assert abs(nrm / o.square_norm(k, o.Vh) - 1) < 1e-14
    assert abs(plaq / o.measure_plaquette() - 1) < 1e-13


def _worker_grid(rank, nt, nz, port, dims_loc, theta, q):
    """rank = ct * nz + cz of an (nt x nz) grid (tmb_comm_init_grid): bench.py's slab() decomposition of the global fields, the
    T-face exchange between ranks (ct +- 1, cz) and the z-face exchange between ranks (ct, cz +- 1) with the pairings of
    exchange_faces() / exchange_zfaces() (tmb_capi.cu), the one-off U_0 and U_z halos, the hop on the slab as if it were
    periodic in z, the fix-up of the z-face sites"""
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    world = nt * nz
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import random_gauge, random_spinor
    from emul_client import Emul
    from bench import slab
    Tl, LX, LY, LZl = dims_loc
    gd = (Tl * nt, LX, LY, LZl * nz)
    rng = np.random.default_rng(78)  # identical global fields on every rank
    V, Vh = int(np.prod(gd)), int(np.prod(gd)) // 2
    g, k, p = random_gauge(rng, V), random_spinor(rng, Vh), random_spinor(rng, Vh)
    e = Emul(*dims_loc)
    U = e.pack_gauge(slab(g.reshape(V, 72), dims_loc, nt, nz, rank, 1, 72).reshape(-1, 4, 18))
    sk = e.pack(slab(k, dims_loc, nt, nz, rank, 2, 24)); sp = e.pack(slab(p, dims_loc, nt, nz, rank, 2, 24))
    ct, cz = rank // nz, rank % nz
    t_up, t_dn = ((ct + 1) % nt) * nz + cz, ((ct - 1) % nt) * nz + cz
    z_up, z_dn = ct * nz + (cz + 1) % nz, ct * nz + (cz - 1) % nz
    Sz = Tl * LX * LY // 2

    def exchange(send_up, send_dn, up_r, dn_r, tag):
        if up_r == rank:  # one rank in this direction: it is its own neighbour (the product copies, exchange_faces())
            return send_dn.copy(), send_up.copy()
        halo_dn, halo_up = torch.empty(send_up.size, dtype=torch.float64), torch.empty(send_dn.size, dtype=torch.float64)
        ops = [dist.P2POp(dist.isend, torch.from_numpy(send_up), up_r, tag=tag), dist.P2POp(dist.irecv, halo_dn, dn_r, tag=tag),
               dist.P2POp(dist.isend, torch.from_numpy(send_dn), dn_r, tag=tag + 1), dist.P2POp(dist.irecv, halo_up, up_r, tag=tag + 1)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return halo_up.numpy(), halo_dn.numpy()

    def shift_up(buf, up_r, dn_r, tag):  # what I send up is what my lower neighbour's counterpart sends me
        if up_r == rank:
            return buf.copy()
        got = torch.empty(buf.size, dtype=torch.float64)
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, torch.from_numpy(buf), up_r, tag=tag), dist.P2POp(dist.irecv, got, dn_r, tag=tag)]):
            w.wait()
        return got.numpy()

    Uhalo = shift_up(e.pack_gauge_halo(U), t_up, t_dn, 10)
    uz = np.zeros(36 * Sz); e.E.emul_pack_gauge_zhalo(uz, U, *dims_loc)
    Uzh = shift_up(uz, z_up, z_dn, 11)
    ka = np.stack([0.16 * np.cos(np.array(theta) * 3.14159265358979 / np.array(gd)),
                   0.16 * np.sin(np.array(theta) * 3.14159265358979 / np.array(gd))], axis=1).reshape(-1)
    outs = {}
    for par in (0, 1):
        su, sd = e.pack_halo(sk)
        hu, hd = exchange(su, sd, t_up, t_dn, 20)
        zu, zd = np.zeros(12 * Sz), np.zeros(12 * Sz)
        e.E.emul_pack_zfaces(zu, zd, sk, *dims_loc, 1 - par)
        hz_up, hz_dn = exchange(zu, zd, z_up, z_dn, 30)
        for name, mode, cf, pp in ((f"hop{par}", 0, (1., 0.), None), (f"tm_sub{par}", 2, (1.0, 0.3), sp)):
            out = e.hop(par, sk, U, ka, mode, cf, pp, halo=(hu, hd, Uhalo))
            e.E.emul_zfix(mode, out, sk, U, hz_up, hz_dn, Uzh, *dims_loc, par, np.asarray(ka, dtype=np.float64), cf[0], cf[1])
            outs[name] = e.unpack(out)
    gathered = {}
    for name, loc in outs.items():
        lst = [torch.empty(loc.shape, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(lst, torch.from_numpy(np.ascontiguousarray(loc)))
        gathered[name] = [x.numpy() for x in lst]
    if rank == 0:
        q.put((gathered, g, k, p))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nt,nz,dims_loc,theta", [(2, 2, (2, 4, 4, 4), (1., 0., 0.3, 0.7)), (1, 2, (4, 2, 4, 6), (0., 0.5, 0., 1.))])
def test_rank_grid_T_x_Z_matches_global_oracle(oracle_lib, nt, nz, dims_loc, theta):
    """four (two) CPU ranks as a 2 x 2 (1 x 2) grid: the decomposition and gathering bench.py uses at N > 1 (slab / unslab), both
    exchange pairings, and the product's site code for hop + z fix-up, against the oracle on the global lattice"""
    sys.path.insert(0, ROOT)
    from bench import unslab
    world = nt * nz
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_grid, args=(r, nt, nz, port, dims_loc, theta, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    gathered, g, k, p = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    Tl, LX, LY, LZl = dims_loc
    o = oracle_lib.Oracle(Tl * nt, LX, LY, LZl * nz)
    o.set_gauge(g); o.set_params(0.16, 0.0, theta)
    exp = o.spinor()
    for par in (0, 1):
        o.Hopping_Matrix(par, exp, k)
        got = unslab(gathered[f"hop{par}"], dims_loc, nt, nz, 2, 24)
        assert np.linalg.norm(got - exp) / np.linalg.norm(exp) < 1e-14
        o.tm_sub_Hopping_Matrix(par, exp, p, k, 1.0, 0.3)
        got = unslab(gathered[f"tm_sub{par}"], dims_loc, nt, nz, 2, 24)
        assert np.linalg.norm(got - exp) / np.linalg.norm(exp) < 1e-14

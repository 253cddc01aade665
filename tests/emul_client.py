"""ctypes access to tests/emul/libtmb_emul.so: the product's __host__ __device__ site functions and
layout functors compiled for the host (TEST ONLY, see tests/emul/tmb_emul.cu)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "emul", "libtmb_emul.so")
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i, d = C.c_int, C.c_double


def load():
    src = os.path.join(HERE, "emul", "tmb_emul.cu")
    csrc = os.path.join(os.path.dirname(HERE), "tmlqcd_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("tmb_kernels.cu", "tmb_force.cu", "tmb_site.cuh", "tmb_geom.h", "tmb_kernels.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(p) > os.path.getmtime(LIB) for p in deps):
        r = subprocess.run(["bash", os.path.join(HERE, "emul", "build.sh")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    E = C.CDLL(LIB)
    sig = {
        "emul_pack_eo": [dp, dp, i], "emul_unpack_eo": [dp, dp, i],
        "emul_pack_lexic": [dp, dp, dp, i, i, i, i], "emul_unpack_lexic": [dp, dp, dp, i, i, i, i],
        "emul_pack_gauge": [dp, dp, i, i, i, i], "emul_pack_halo": [dp, dp, dp, i, i, i, i],
        "emul_pack_gauge_halo": [dp, dp, i, i, i, i], "emul_neighbours": [ip, i, i, i, i, i],
        "emul_eo2lexic": [ip, i, i, i, i], "emul_pull_halo": [dp, dp, dp, dp, i, i, i, i], "emul_xblock_perm": [ip, i, i, i, i, i], "emul_tile_perm": [ip, ip, i, i, i, i, i], "emul_host_chunk_schedule": [i, i, ip],
        "emul_hop": [i, dp, dp, dp, dp, dp, dp, dp, i, i, i, i, dp, d, d, i, i],
        "emul_hop12": [i, dp, dp, dp, dp, dp, dp, i, i, i, i, dp, i], "emul_compress12": [dp, dp, C.c_long, i],
        "emul_hop_f": [i, fp, fp, fp, i, i, i, i, dp],
        "emul_diag": [dp, dp, d, d, i], "emul_diag_sub": [dp, dp, dp, d, d, i, i], "emul_gamma5": [dp, dp, i],
        "emul_pack_deriv": [dp, dp, i, i, i, i], "emul_unpack_deriv": [dp, dp, i, i, i, i],
        "emul_pack_deriv_halo": [dp, dp, dp, i, i, i, i],
        "emul_deriv": [i, dp, dp, dp, dp, dp, i, i, i, i, dp, d, i],
        "emul_deriv_zfix": [i, dp, dp, dp, dp, dp, dp, i, i, i, i, dp, d],
        "emul_blas32": [i, fp, fp, fp, C.c_float, C.c_float, C.c_long],
        "emul_hop2": [i, dp, dp, dp, dp, dp, dp, dp, i, i, i, i, dp, i, d, d, d],
        "emul_pack_gauge_first_slice": [dp, dp, i, i, i, i], "emul_plaquette": [dp, dp, i, i, i, i, i],
        "emul_nd_mee_inv": [dp, dp, dp, dp, d, d, i], "emul_nd_moo_sub_g5": [dp, dp, dp, dp, dp, dp, d, d, i],
        "emul_pack_zfaces": [dp, dp, dp, i, i, i, i, i], "emul_pack_gauge_zhalo": [dp, dp, i, i, i, i],
        "emul_zfix": [i, dp, dp, dp, dp, dp, dp, i, i, i, i, i, dp, d, d],
    }
    for n, a in sig.items():
        getattr(E, n).argtypes = a
        getattr(E, n).restype = i if n in ("emul_hop", "emul_hop12", "emul_hop_f", "emul_tile_perm", "emul_host_chunk_schedule") else (d if n == "emul_plaquette" else None)
    return E


class Emul:
    """device-layout helper for one local lattice"""

    def __init__(self, T, LX, LY, LZ):
        self.E = load()
        self.dims = (T, LX, LY, LZ)
        self.V = T * LX * LY * LZ
        self.Vh = self.V // 2
        self.S = LX * LY * LZ // 2

    def pack(self, aos):
        out = np.zeros(24 * self.Vh); self.E.emul_pack_eo(out, np.ascontiguousarray(aos).reshape(-1), self.Vh); return out

    def unpack(self, soa):
        out = np.zeros(24 * self.Vh); self.E.emul_unpack_eo(out, soa, self.Vh); return out.reshape(self.Vh, 24)

    def pack_gauge(self, g):
        U = np.zeros(144 * self.Vh); self.E.emul_pack_gauge(U, np.ascontiguousarray(g).reshape(-1), *self.dims); return U

    def pack_halo(self, soa):
        up, dn = np.zeros(12 * self.S), np.zeros(12 * self.S)
        self.E.emul_pack_halo(up, dn, soa, *self.dims); return up, dn

    def pack_gauge_halo(self, U):
        out = np.zeros(36 * self.S); self.E.emul_pack_gauge_halo(out, U, *self.dims); return out

    def deriv(self, ieo, soa_l, soa_k, U, ka, df_lex, factor, halo=None):
        """deriv_Sb on device-layout fields; df_lex in the reference's [V][4][8] layout, returned likewise"""
        dev = np.zeros(64 * self.Vh); self.E.emul_pack_deriv(dev, np.ascontiguousarray(df_lex).reshape(-1), *self.dims)
        h = halo if halo is not None else np.zeros(2)
        self.E.emul_deriv(ieo, soa_l, soa_k, U, dev, h, *self.dims, np.asarray(ka, dtype=np.float64), factor,
                          0 if halo is None else 1)
        out = np.zeros(32 * self.V); self.E.emul_unpack_deriv(out, dev, *self.dims)
        return out.reshape(self.V, 4, 8)

    def hop2(self, par, in0, in1, U, ka, mode=0, p0=None, p1=None, mu=0., eps=0., scale=1.):
        """two-flavour hop with the epilogues of hop2_kernel; fields in device layout"""
        o0, o1 = np.zeros(24 * self.Vh), np.zeros(24 * self.Vh)
        z = np.zeros(2)
        self.E.emul_hop2(par, o0, o1, in0, in1, p0 if p0 is not None else z, p1 if p1 is not None else z, U, *self.dims,
                         np.asarray(ka, dtype=np.float64), mode, mu, eps, scale)
        return o0, o1

    def plaquette(self, U, up=None):
        """measure_plaquette on the device-layout gauge field; up = first-slice spatial links of the rank above"""
        return self.E.emul_plaquette(U, up if up is not None else np.zeros(2), *self.dims, 0 if up is None else 1)

    def pack_gauge_first_slice(self, U):
        out = np.zeros(108 * self.S); self.E.emul_pack_gauge_first_slice(out, U, *self.dims); return out

    def pack_deriv_halo(self, soa_k, soa_l):
        out = np.zeros(24 * self.S); self.E.emul_pack_deriv_halo(out, soa_k, soa_l, *self.dims); return out

    def pull_halo(self, soa_up, soa_dn):
        hu, hd = np.zeros(12 * self.S), np.zeros(12 * self.S)
        self.E.emul_pull_halo(hu, hd, soa_up, soa_dn, *self.dims); return hu, hd

    def hop(self, par, soa_in, U, ka, mode=0, cf=(1., 0.), soa_p=None, halo=None):
        out = np.zeros(24 * self.Vh)
        z = np.zeros(2)
        p = soa_p if soa_p is not None else z
        hu, hd, Uh = halo if halo is not None else (z, z, z)
        rc = self.E.emul_hop(par, out, soa_in, p, U, hu, hd, Uh, *self.dims, np.asarray(ka, dtype=np.float64),
                             cf[0], cf[1], mode, 0 if halo is None else 1)
        assert rc == 0
        return out

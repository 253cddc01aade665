"""ctypes client for oracle/_ref/libtmlqcd_ref*.so (the compiled, unmodified reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The reference keeps its state in C globals
(global.h), so one process can hold ONE lattice per loaded variant; use `run_isolated`
for anything that needs another lattice size.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REFDIR = os.path.join(_HERE, "_ref")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def lib_path(halfspinor=False):
    return os.path.join(_REFDIR, "libtmlqcd_ref_hs.so" if halfspinor else "libtmlqcd_ref.so")


def available(halfspinor=False):
    return os.path.exists(lib_path(halfspinor))


class Reference:
    """One loaded reference library bound to one lattice (T, LX, LY, LZ)."""

    def __init__(self, T, LX, LY, LZ, nthreads=1, halfspinor=False):
        path = lib_path(halfspinor)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle/ref_build` (needs /root/reference)")
        self.lib = L = C.CDLL(path)
        self.T, self.LX, self.LY, self.LZ = T, LX, LY, LZ
        self.V = T * LX * LY * LZ
        self.Vh = self.V // 2
        d, i = C.c_double, C.c_int
        sig = {
            "ref_init": (i, [i, i, i, i, i]),
            "ref_num_threads": (i, []),
            "ref_is_halfspinor": (i, []),
            "ref_set_params": (None, [d] * 6),
            "ref_set_nd_params": (None, [d] * 3),
            "ref_set_debug_level": (None, [i]),
            "ref_start_ranlux": (None, [i, i]),
            "ref_random_gauge": (None, [i]),
            "ref_get_gauge": (None, [_dp]),
            "ref_set_gauge": (None, [_dp]),
            "ref_random_spinor_eo": (None, [_dp]),
            "ref_random_spinor_lexic": (None, [_dp]),
            "ref_get_eo2lexic": (None, [_ip]),
            "ref_get_lexic2eosub": (None, [_ip]),
            "ref_get_hi": (None, [_ip]),
            "ref_get_iup": (None, [_ip]),
            "ref_get_idn": (None, [_ip]),
            "ref_get_ka": (None, [_dp]),
            "ref_Hopping_Matrix": (None, [i, _dp, _dp]),
            "ref_tm_times_Hopping_Matrix": (None, [i, _dp, _dp, d, d]),
            "ref_tm_sub_Hopping_Matrix": (None, [i, _dp, _dp, _dp, d, d]),
            "ref_H_eo_tm_inv_psi": (None, [_dp, _dp, i, d]),
            "ref_Qtm_pm_psi": (None, [_dp, _dp]),
            "ref_Qtm_plus_psi": (None, [_dp, _dp]),
            "ref_Qtm_minus_psi": (None, [_dp, _dp]),
            "ref_Mtm_plus_psi": (None, [_dp, _dp]),
            "ref_Mtm_minus_psi": (None, [_dp, _dp]),
            "ref_M_full": (None, [_dp] * 4),
            "ref_Q_full": (None, [_dp] * 4),
            "ref_D_psi": (None, [_dp, _dp]),
            "ref_Q_pm_psi": (None, [_dp, _dp]),
            "ref_gamma5": (None, [_dp, _dp, i]),
            "ref_mul_one_pm_imu_inv": (None, [_dp, d, i]),
            "ref_assign_mul_one_pm_imu_inv": (None, [_dp, _dp, d, i]),
            "ref_assign_mul_one_pm_imu": (None, [_dp, _dp, d, i]),
            "ref_mul_one_pm_imu_sub_mul_gamma5": (None, [_dp, _dp, _dp, d]),
            "ref_convert_eo_to_lexic": (None, [_dp, _dp, _dp]),
            "ref_convert_lexic_to_eo": (None, [_dp, _dp, _dp]),
            "ref_square_norm": (d, [_dp, i]),
            "ref_scalar_prod_r": (d, [_dp, _dp, i]),
            "ref_assign_add_mul_r": (None, [_dp, _dp, d, i]),
            "ref_assign_mul_add_r": (None, [_dp, d, _dp, i]),
            "ref_assign_mul_add_r_and_square": (d, [_dp, d, _dp, i]),
            "ref_diff": (None, [_dp, _dp, _dp, i]),
            "ref_assign": (None, [_dp, _dp, i]),
            "ref_mul_r": (None, [_dp, d, _dp, i]),
            "ref_cg_her": (i, [_dp, _dp, i, d, i]),
            "ref_invert_eo_cg": (i, [_dp, _dp, _dp, _dp, d, i, i]),
            "ref_invert_eo_flags": (i, [_dp, _dp, _dp, _dp, d, i, i, i, i, d]),
            "ref_Qtm_pm_ndpsi": (None, [_dp] * 4),
            "ref_Qtm_ndpsi": (None, [_dp] * 4),
            "ref_Qtm_dagger_ndpsi": (None, [_dp] * 4),
            "ref_M_ee_inv_ndpsi": (None, [_dp] * 4 + [d, d]),
            "ref_cg_her_nd": (i, [_dp] * 4 + [i, d, i]),
            "ref_invert_doublet_eo_cg": (i, [_dp] * 8 + [d, i, i]),
            "ref_init32": (i, []), "ref_update_gauge32": (None, []), "ref_set_mixcg": (None, [d, i]),
            "ref_Hopping_Matrix_32": (None, [i, _fp, _fp]), "ref_Qtm_pm_psi_32": (None, [_fp, _fp]),
            "ref_mixed_cg_her": (i, [_dp, _dp, i, d, i]),
            "ref_hmc_init": (i, []), "ref_set_relative_precision_flag": (None, [i]),
            "ref_deriv_Sb": (None, [i, _dp, _dp, _dp, d]),
            "ref_mnl_add": (i, [i, d, d, d, d, i, i, d, d, i]), "ref_mnl_init": (i, []),
            "ref_mnl_heatbath": (None, [i]), "ref_mnl_acc": (d, [i]), "ref_mnl_derivative": (None, [i, _dp]),
            "ref_mnl_get_pf": (None, [i, _dp]), "ref_mnl_set_pf": (None, [i, _dp]),
            "ref_mnl_info": (None, [i, C.POINTER(d), C.POINTER(d), C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
            "ref_solve_degenerate": (i, [_dp, _dp, i, d, i, i]),
            "ref_bench_mnl_derivative": (d, [i, i]),
            **{"ref_" + n: (None, [_dp, _dp]) for n in ("Qtm_plus_sym_psi", "Qtm_minus_sym_psi", "Mtm_plus_sym_psi",
               "Mtm_minus_sym_psi", "Mtm_plus_sym_dagg_psi", "Qtm_pm_sym_psi", "M_minus_psi", "D_dagg_psi", "Q_plus_psi", "Q_minus_psi")},
            "ref_Mee_psi": (None, [_dp, _dp, d]), "ref_Mee_inv_psi": (None, [_dp, _dp, d]),
            "ref_mul_one_sub_mul_gamma5": (None, [_dp] * 3), "ref_mul_one_pm_imu_sub_mul": (None, [_dp] * 3 + [d, i]),
            "ref_M_minus_1_timesC": (None, [_dp] * 4), "ref_H_eo_tm_ndpsi": (None, [_dp] * 4 + [i]),
            "ref_M_oo_sub_g5_ndpsi": (None, [_dp] * 6 + [d, d]), "ref_mul_one_pm_iconst": (None, [_dp, _dp, d, i]),
            "ref_rg_mixed_cg_her": (i, [_dp, _dp, i, d, i, d]),
            "ref_measure_plaquette": (d, []),
            "ref_write_gauge": (i, [C.c_char_p, i, d, i]), "ref_read_gauge": (i, [C.c_char_p, i]),
            "ref_write_propagator": (i, [C.c_char_p, _dp, _dp, i, d, i]), "ref_read_spinor": (i, [_dp, _dp, C.c_char_p, i]),
            "ref_bench_hopping": (d, [i]),
            "ref_bench_D_psi": (d, [i]),
            "ref_bench_Qtm_pm": (d, [i]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        rc = L.ref_init(T, LX, LY, LZ, nthreads)
        if rc != 0:
            raise RuntimeError(f"ref_init failed rc={rc}")
        self.nthreads = L.ref_num_threads()

    # ---- field helpers (reference AoS layouts) ----
    def spinor(self, n=None):
        return np.zeros((self.Vh if n is None else n, 24), dtype=np.float64)

    def random_gauge(self, seed=123456):
        self.lib.ref_random_gauge(seed)
        return self.get_gauge()

    def get_gauge(self):
        g = np.zeros((self.V, 4, 18), dtype=np.float64)
        self.lib.ref_get_gauge(g)
        return g

    def set_gauge(self, g):
        self.lib.ref_set_gauge(np.ascontiguousarray(g, dtype=np.float64))

    def random_spinor_eo(self):
        s = self.spinor()
        self.lib.ref_random_spinor_eo(s)
        return s

    def random_spinor_lexic(self):
        s = self.spinor(self.V)
        self.lib.ref_random_spinor_lexic(s)
        return s

    def set_params(self, kappa, gmu, theta=(0., 0., 0., 0.)):
        self.lib.ref_set_params(kappa, gmu, *[float(t) for t in theta])

    def derivative(self):
        """hf->derivative as [V][4][8] doubles"""
        return np.zeros((self.V, 4, 8), dtype=np.float64)

    def mnl_info(self, id):
        e0, e1, i0, i1, n = C.c_double(), C.c_double(), C.c_int(), C.c_int(), C.c_int()
        self.lib.ref_mnl_info(id, C.byref(e0), C.byref(e1), C.byref(i0), C.byref(i1), C.byref(n))
        return {"energy0": e0.value, "energy1": e1.value, "iter0": i0.value, "iter1": i1.value, "csg_n": n.value}

    def table(self, name):
        n = {"eo2lexic": 1, "lexic2eosub": 1, "hi": 16, "iup": 4, "idn": 4}[name]
        out = np.zeros(self.V * n, dtype=np.int32)
        getattr(self.lib, "ref_get_" + name)(out)
        return out.reshape(self.V, n) if n > 1 else out

    def __getattr__(self, name):
        # forward e.g. ref.Hopping_Matrix(...) -> lib.ref_Hopping_Matrix(...)
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.lib, "ref_" + name)

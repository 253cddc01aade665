"""ctypes client for oracle/libtmoracle.so, the plain-C restatement of the reference path.

TEST INFRASTRUCTURE ONLY (see the header of tmoracle.c).  Function names mirror the
reference: Oracle.Hopping_Matrix(ieo, l, k), Oracle.Qtm_pm_psi(l, k), Oracle.cg_her(...).
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libtmoracle.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class Oracle:
    def __init__(self, T, LX, LY, LZ):
        if not os.path.exists(LIB):
            raise FileNotFoundError(f"{LIB} missing: run `make -C oracle`")
        self.lib = L = C.CDLL(LIB)
        d, i = C.c_double, C.c_int
        sig = {
            "orc_init": (i, [i] * 4),
            "orc_finalize": (None, []),
            "orc_set_params": (None, [d] * 6),
            "orc_set_nd_params": (None, [d] * 3),
            "orc_set_gauge": (None, [_dp]),
            "orc_get_eo2lexic": (None, [_ip]),
            "orc_get_lexic2eosub": (None, [_ip]),
            "orc_Hopping_Matrix": (None, [i, _dp, _dp]),
            "orc_tm_times_Hopping_Matrix": (None, [i, _dp, _dp, d, d]),
            "orc_tm_sub_Hopping_Matrix": (None, [i, _dp, _dp, _dp, d, d]),
            "orc_mul_one_pm_imu_inv": (None, [_dp, d, i]),
            "orc_assign_mul_one_pm_imu_inv": (None, [_dp, _dp, d, i]),
            "orc_assign_mul_one_pm_imu": (None, [_dp, _dp, d, i]),
            "orc_mul_one_pm_imu": (None, [_dp, d]),
            "orc_mul_one_pm_imu_sub_mul_gamma5": (None, [_dp, _dp, _dp, d]),
            "orc_mul_one_pm_imu_sub_mul": (None, [_dp, _dp, _dp, d, i]),
            "orc_gamma5": (None, [_dp, _dp, i]),
            "orc_square_norm": (d, [_dp, i]),
            "orc_scalar_prod_r": (d, [_dp, _dp, i]),
            "orc_assign_add_mul_r": (None, [_dp, _dp, d, i]),
            "orc_assign_mul_add_r": (None, [_dp, d, _dp, i]),
            "orc_assign_mul_add_r_and_square": (d, [_dp, d, _dp, i]),
            "orc_diff": (None, [_dp, _dp, _dp, i]),
            "orc_assign": (None, [_dp, _dp, i]),
            "orc_mul_r": (None, [_dp, d, _dp, i]),
            "orc_add": (None, [_dp, _dp, _dp, i]),
            "orc_convert_eo_to_lexic": (None, [_dp, _dp, _dp]),
            "orc_convert_lexic_to_eo": (None, [_dp, _dp, _dp]),
            "orc_H_eo_tm_inv_psi": (None, [_dp, _dp, i, d]),
            "orc_tm_sub_H_eo_gamma5": (None, [_dp, _dp, _dp, i, d]),
            "orc_Qtm_pm_psi": (None, [_dp, _dp]),
            "orc_Qtm_plus_psi": (None, [_dp, _dp]),
            "orc_Qtm_minus_psi": (None, [_dp, _dp]),
            "orc_Mtm_plus_psi": (None, [_dp, _dp]),
            "orc_Mtm_minus_psi": (None, [_dp, _dp]),
            "orc_M_full": (None, [_dp] * 4),
            "orc_Q_full": (None, [_dp] * 4),
            "orc_D_psi": (None, [_dp, _dp]),
            "orc_Q_pm_psi": (None, [_dp, _dp]),
            "orc_cg_her": (i, [_dp, _dp, i, d, i]),
            "orc_cg_her_full": (i, [_dp, _dp, i, d, i]),
            "orc_invert_eo_cg": (i, [_dp] * 4 + [d, i, i]),
            "orc_M_ee_inv_ndpsi": (None, [_dp] * 4 + [d, d]),
            "orc_M_oo_sub_g5_ndpsi": (None, [_dp] * 6 + [d, d]),
            "orc_Qtm_ndpsi": (None, [_dp] * 4),
            "orc_Qtm_dagger_ndpsi": (None, [_dp] * 4),
            "orc_Qtm_pm_ndpsi": (None, [_dp] * 4),
            "orc_cg_her_nd": (i, [_dp] * 4 + [i, d, i]),
            "orc_invert_doublet_eo_cg": (i, [_dp] * 8 + [d, i, i]),
            "orc_deriv_Sb": (None, [i, _dp, _dp, _dp, d]),
            "orc_measure_plaquette": (d, []),
            "orc_set_relative_precision_flag": (None, [i]),
            "orc_mnl_add": (i, [i, d, d, d, d, i, i, d, d, i]),
            "orc_mnl_clear": (None, []),
            "orc_mnl_heatbath": (d, [i, _dp]),
            "orc_mnl_derivative": (None, [i, _dp]),
            "orc_mnl_acc": (d, [i]),
            "orc_mnl_get_pf": (None, [i, _dp]),
            "orc_mnl_set_pf": (None, [i, _dp]),
            "orc_mnl_info": (None, [i, C.POINTER(d), C.POINTER(d), C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        if L.orc_init(T, LX, LY, LZ) != 0:
            raise RuntimeError("orc_init failed (odd volume?)")
        self.T, self.LX, self.LY, self.LZ = T, LX, LY, LZ
        self.V = T * LX * LY * LZ
        self.Vh = self.V // 2
        self._gauge = None

    def spinor(self, n=None):
        return np.zeros((self.Vh if n is None else n, 24), dtype=np.float64)

    def set_gauge(self, g):
        self._gauge = np.ascontiguousarray(g, dtype=np.float64)  # keep alive: oracle stores the pointer
        self.lib.orc_set_gauge(self._gauge)

    def set_params(self, kappa, gmu, theta=(0., 0., 0., 0.)):
        self.lib.orc_set_params(kappa, gmu, *[float(t) for t in theta])

    def derivative(self):
        """hf->derivative as [V][4][8] doubles"""
        return np.zeros((self.V, 4, 8), dtype=np.float64)

    def mnl_info(self, id):
        e0, e1, i0, i1, n = C.c_double(), C.c_double(), C.c_int(), C.c_int(), C.c_int()
        self.lib.orc_mnl_info(id, C.byref(e0), C.byref(e1), C.byref(i0), C.byref(i1), C.byref(n))
        return {"energy0": e0.value, "energy1": e1.value, "iter0": i0.value, "iter1": i1.value, "csg_n": n.value}

    def lexic2eosub(self):
        out = np.zeros(self.V, dtype=np.int32)
        self.lib.orc_get_lexic2eosub(out)
        return out

    def eo2lexic(self):
        out = np.zeros(self.V, dtype=np.int32)
        self.lib.orc_get_eo2lexic(out)
        return out

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.lib, "orc_" + name)

"""ctypes bindings for libtmlqcd_b200.so.

`Device`  : the device-level C ABI of include/tmlqcd_b200.h (tmb_*), fields resident in HBM.
`DropIn`  : the reference-named entry points of include/tmlqcd_b200_dropin.h with host
            (numpy) buffers, mirroring tmLQCD's own C interface for this path.
There is NO fallback: if the shared library is missing or no GPU is present the calls raise.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "lib", "libtmlqcd_b200.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_vp, _d, _i = C.c_void_p, C.c_double, C.c_int


def lib_path():
    return _LIB


def build(force=False):
    """Compile every CUDA source for sm_100a into tmlqcd_b200/lib (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libtmlqcd_b200.so failed:\n" + r.stdout + r.stderr)
    return _LIB


# name -> (restype, argtypes); every symbol declared in include/tmlqcd_b200.h
DEVICE_API = {
    "tmb_init": (_i, [_i] * 5), "tmb_finalize": (_i, []), "tmb_is_initialized": (_i, []),
    "tmb_last_error": (C.c_char_p, []), "tmb_volume_half": (_i, []),
    "tmb_comm_unique_id": (_i, [_vp]), "tmb_comm_init": (_i, [_vp, _i, _i]), "tmb_comm_loopback": (_i, [_i]), "tmb_comm_loopback_z": (_i, [_i]),
    "tmb_comm_init_grid": (_i, [_vp, _i, _i, _i]), "tmb_comm_grid": (_i, [C.POINTER(_i), C.POINTER(_i)]),
    "tmb_comm_nranks": (_i, []), "tmb_comm_peer_mode": (_i, []), "tmb_comm_zpeer_mode": (_i, []), "tmb_comm_sequence_counts": (_i, [_vp, _vp]),
    "tmb_set_boundary": (_i, [_d, _dp]), "tmb_set_hopping_phases": (_i, [_dp]), "tmb_set_mu": (_i, [_d]),
    "tmb_set_nd": (_i, [_d] * 3), "tmb_set_tuning": (_i, [_i] * 3), "tmb_set_hop2_variant": (_i, [_i]), "tmb_set_tile": (_i, [_i]), "tmb_set_prefetch_distance": (_i, [_i]), "tmb_set_host_chunk_sizes": (_i, [_vp, _i]), "tmb_host_hop_timeline": (_i, [_i, _vp, _vp, _vp, _i]), "tmb_set_overlap": (_i, [_i]), "tmb_set_p2p_diag": (_i, [_i]), "tmb_set_host_chunks": (_i, [_i]), "tmb_set_compression": (_i, [_i]),
    "tmb_field_alloc": (_vp, []), "tmb_field_free": (_i, [_vp]), "tmb_field_zero": (_i, [_vp]),
    "tmb_host_alloc": (_vp, [C.c_size_t]), "tmb_host_free": (_i, [_vp]),
    "tmb_host_register": (_i, [_vp, C.c_size_t]), "tmb_host_unregister": (_i, [_vp]),
    "tmb_field_upload": (_i, [_vp, _vp]), "tmb_field_download": (_i, [_vp, _vp]),
    "tmb_field_upload_lexic": (_i, [_vp, _vp, _vp]), "tmb_field_download_lexic": (_i, [_vp, _vp, _vp]),
    "tmb_gauge_upload": (_i, [_vp]), "tmb_sync": (_i, []),
    "tmb_timer_start": (_i, []), "tmb_timer_stop": (_i, [C.POINTER(C.c_float)]),
    "tmb_Hopping_Matrix": (_i, [_i, _vp, _vp]), "tmb_Hopping_Matrix_nocom": (_i, [_i, _vp, _vp]), "tmb_Hopping_Matrix_host": (_i, [_i, _vp, _vp, _i, _d, _d]),
    "tmb_tm_times_Hopping_Matrix": (_i, [_i, _vp, _vp, _d, _d]),
    "tmb_tm_sub_Hopping_Matrix": (_i, [_i, _vp, _vp, _vp, _d, _d]),
    "tmb_H_eo_tm_inv_psi": (_i, [_vp, _vp, _i, _d]), "tmb_tm_sub_H_eo_gamma5": (_i, [_vp, _vp, _vp, _i, _d]),
    "tmb_Qtm_pm_psi": (_i, [_vp, _vp]), "tmb_Qtm_plus_psi": (_i, [_vp, _vp]), "tmb_Qtm_minus_psi": (_i, [_vp, _vp]),
    "tmb_Mtm_plus_psi": (_i, [_vp, _vp]), "tmb_Mtm_minus_psi": (_i, [_vp, _vp]),
    "tmb_M_full": (_i, [_vp] * 4), "tmb_Q_full": (_i, [_vp] * 4), "tmb_D_psi_eo": (_i, [_vp] * 4),
    "tmb_assign_mul_one_pm_imu_inv": (_i, [_vp, _vp, _d]), "tmb_assign_mul_one_pm_imu": (_i, [_vp, _vp, _d]),
    "tmb_mul_one_pm_imu_sub_mul_gamma5": (_i, [_vp, _vp, _vp, _d]), "tmb_mul_one_pm_imu_sub_mul": (_i, [_vp, _vp, _vp, _d]),
    "tmb_gamma5": (_i, [_vp, _vp]), "tmb_diag": (_i, [_vp, _vp, _d, _d]), "tmb_diag_sub": (_i, [_vp, _vp, _vp, _d, _d, _i]),
    "tmb_M_oo_sub_g5_ndpsi": (_i, [_vp] * 6 + [_d, _d]),
    "tmb_square_norm": (_i, [_vp, C.POINTER(_d)]), "tmb_scalar_prod_r": (_i, [_vp, _vp, C.POINTER(_d)]),
    "tmb_assign_add_mul_r": (_i, [_vp, _vp, _d]), "tmb_assign_mul_add_r": (_i, [_vp, _d, _vp]),
    "tmb_assign_mul_add_r_and_square": (_i, [_vp, _d, _vp, C.POINTER(_d)]),
    "tmb_diff": (_i, [_vp] * 3), "tmb_add": (_i, [_vp] * 3), "tmb_assign": (_i, [_vp] * 2), "tmb_mul_r": (_i, [_vp, _d, _vp]),
    "tmb_cg_her": (_i, [_vp, _vp, _i, _d, _i]), "tmb_invert_eo": (_i, [_vp] * 4 + [_d, _i, _i]),
    "tmb_solver_stats": (_i, [C.POINTER(_i), C.POINTER(_d), C.POINTER(_d)]),
    "tmb_M_ee_inv_ndpsi": (_i, [_vp] * 4 + [_d, _d]), "tmb_Qtm_ndpsi": (_i, [_vp] * 4),
    "tmb_Qtm_dagger_ndpsi": (_i, [_vp] * 4), "tmb_Qtm_pm_ndpsi": (_i, [_vp] * 4),
    "tmb_cg_her_nd": (_i, [_vp] * 4 + [_i, _d, _i]), "tmb_invert_doublet_eo": (_i, [_vp] * 8 + [_d, _i, _i]),
    "tmb_invert_doublet_eo_solver": (_i, [_vp] * 8 + [_d, _i, _i, _i]), "tmb_Qtm_pm_ndpsi_32": (_i, [_vp] * 4),
    "tmb_rg_mixed_cg_her_nd": (_i, [_vp] * 4 + [_i, _d, _i]), "tmb_solver_stats_rg": (_i, [C.POINTER(_i)] * 3),
    "tmb_field32_alloc": (_vp, []), "tmb_field32_upload": (_i, [_vp, _vp]), "tmb_field32_download": (_i, [_vp, _vp]),
    "tmb_assign_to_32": (_i, [_vp, _vp]), "tmb_assign_to_64": (_i, [_vp, _vp]),
    "tmb_Hopping_Matrix_32": (_i, [_i, _vp, _vp]), "tmb_Qtm_pm_psi_32": (_i, [_vp, _vp]),
    "tmb_D_psi_eo_32": (_i, [_vp] * 4), "tmb_M_full_32": (_i, [_vp] * 4 + [_i]),
    "tmb_field32_upload_lexic": (_i, [_vp, _vp, _vp]), "tmb_field32_download_lexic": (_i, [_vp, _vp, _vp]),
    "tmb_set_mixcg": (_i, [_d, _i]), "tmb_mixed_cg_her": (_i, [_vp, _vp, _i, _d, _i]),
    "tmb_invert_eo_mixed": (_i, [_vp] * 4 + [_d, _i, _i]), "tmb_invert_eo_rgmixed": (_i, [_vp] * 4 + [_d, _i, _i]),
    "tmb_set_mcg_delta": (_i, [_d]), "tmb_rg_mixed_cg_her": (_i, [_vp, _vp, _i, _d, _i]),
    "tmb_derivative_zero": (_i, []), "tmb_derivative_upload": (_i, [_vp]), "tmb_derivative_download": (_i, [_vp]),
    "tmb_deriv_Sb": (_i, [_i, _vp, _vp, _d]),
    "tmb_scalar_prod": (_i, [_vp, _vp, C.POINTER(_d), C.POINTER(_d)]),
    "tmb_assign_add_mul": (_i, [_vp, _vp, _d, _d]), "tmb_assign_diff_mul": (_i, [_vp, _vp, _d, _d]),
    "tmb_mul": (_i, [_vp, _d, _d, _vp]),
    "tmb_chrono_add_solution": (_i, [_vp, C.POINTER(_vp), C.POINTER(_i), _i, C.POINTER(_i)]),
    "tmb_chrono_guess": (_i, [_vp, _vp, C.POINTER(_vp), C.POINTER(_i), _i, _i, _i]),
    "tmb_solve_degenerate": (_i, [_vp, _vp, _i, _d, _i, _i]),
    "tmb_monomial_add": (_i, [_i, _d, _d, _d, _d, _i, _i, _d, _d, _i]), "tmb_monomial_clear": (_i, []),
    "tmb_set_relative_precision_flag": (_i, [_i]),
    "tmb_monomial_heatbath": (_i, [_i, _vp, C.POINTER(_d)]), "tmb_monomial_derivative": (_i, [_i]),
    "tmb_monomial_acc": (_i, [_i, C.POINTER(_d)]),
    "tmb_monomial_info": (_i, [_i, C.POINTER(_d), C.POINTER(_d), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "tmb_monomial_pf": (_vp, [_i]), "tmb_monomial_wfield": (_vp, [_i]),
    "tmb_blas32": (_i, [_i, _vp, _vp, _vp, _d, _d]), "tmb_square_norm_32": (_i, [_vp, C.POINTER(_d)]), "tmb_scalar_prod_r_32": (_i, [_vp, _vp, C.POINTER(_d)]),
    "tmb_measure_plaquette": (_i, [C.POINTER(_d)]), "tmb_launch_count": (C.c_longlong, []), "tmb_measure_copy_gbs": (_i, [C.c_size_t, _i, C.POINTER(_d)]),
    "tmb_measure_pcie_gbs": (_i, [C.c_size_t, _i, C.POINTER(_d), C.POINTER(_d), C.POINTER(_d)]),
}

_sp = _dp  # host spinor buffers (reference AoS layout) as float64 arrays


class SolverParams(C.Structure):
    """solver_params_t (solver/solver_params.h:39-101), passed by value to invert_eo."""
    _fields_ = [("eigcg_nrhs", _i), ("eigcg_nrhs1", _i), ("eigcg_nev", _i), ("eigcg_vmax", _i), ("eigcg_ldh", _i),
                ("eigcg_tolsq1", _d), ("eigcg_tolsq", _d), ("eigcg_restolsq", _d), ("eigcg_rand_guess_opt", _i),
                ("mcg_delta", C.c_float), ("type", _i), ("max_iter", _i), ("rel_prec", _i), ("no_shifts", _i),
                ("sdim", _i), ("squared_solver_prec", _d), ("M_psi", _vp), ("M_psi32", _vp), ("M_ndpsi", _vp),
                ("M_ndpsi32", _vp), ("shifts", _vp), ("solution_type", _i), ("compression_type", _i),
                ("sloppy_precision", _i), ("external_inverter", _i)]


class HamiltonianField(C.Structure):
    """hamiltonian_field_t (hamiltonian_field.h:28-34)"""
    _fields_ = [("gaugefield", _vp), ("momenta", _vp), ("derivative", C.POINTER(C.POINTER(_d))),
                ("update_gauge_copy", _i), ("traj_counter", _i)]


RANDOM_SPINOR_FN = C.CFUNCTYPE(None, C.POINTER(_d), _i, _i)

# reference-named entry points (include/tmlqcd_b200_dropin.h).  `_Complex double` by value is
# two doubles in SSE registers under the x86-64 SysV ABI, hence the (_d, _d) pairs.
DROPIN_API = {
    "tmb_dropin_init": (_i, [_i] * 5), "tmb_dropin_finalize": (_i, []),
    "boundary": (None, [_d]),
    "Hopping_Matrix": (None, [_i, _sp, _sp]), "Hopping_Matrix_nocom": (None, [_i, _sp, _sp]),
    "tm_times_Hopping_Matrix": (None, [_i, _sp, _sp, _d, _d]),
    "tm_sub_Hopping_Matrix": (None, [_i, _sp, _sp, _sp, _d, _d]),
    "Qtm_pm_psi": (None, [_sp, _sp]), "Qtm_plus_psi": (None, [_sp, _sp]), "Qtm_minus_psi": (None, [_sp, _sp]),
    "Mtm_plus_psi": (None, [_sp, _sp]), "Mtm_minus_psi": (None, [_sp, _sp]),
    "H_eo_tm_inv_psi": (None, [_sp, _sp, _i, _d]), "tm_sub_H_eo_gamma5": (None, [_sp, _sp, _sp, _i, _d]),
    "M_full": (None, [_sp] * 4), "Q_full": (None, [_sp] * 4),
    "mul_one_pm_imu_inv": (None, [_sp, _d, _i]), "assign_mul_one_pm_imu_inv": (None, [_sp, _sp, _d, _i]),
    "mul_one_pm_imu": (None, [_sp, _d]), "assign_mul_one_pm_imu": (None, [_sp, _sp, _d, _i]),
    "mul_one_pm_imu_sub_mul_gamma5": (None, [_sp, _sp, _sp, _d]),
    "D_psi": (None, [_sp, _sp]), "Q_pm_psi": (None, [_sp, _sp]), "Q_plus_psi": (None, [_sp, _sp]),
    "Q_minus_psi": (None, [_sp, _sp]), "gamma5": (None, [_sp, _sp, _i]),
    "square_norm": (_d, [_sp, _i, _i]), "scalar_prod_r": (_d, [_sp, _sp, _i, _i]),
    "assign_add_mul_r": (None, [_sp, _sp, _d, _i]), "assign_mul_add_r": (None, [_sp, _d, _sp, _i]),
    "assign_mul_add_r_and_square": (_d, [_sp, _d, _sp, _i, _i]),
    "diff": (None, [_sp, _sp, _sp, _i]), "add": (None, [_sp, _sp, _sp, _i]), "assign": (None, [_sp, _sp, _i]),
    "mul_r": (None, [_sp, _d, _sp, _i]),
    "convert_eo_to_lexic": (None, [_sp, _sp, _sp]), "convert_lexic_to_eo": (None, [_sp, _sp, _sp]),
    "cg_her": (_i, [_sp, _sp, _i, _d, _i, _i, _vp]),
    "invert_eo": (_i, [_sp] * 4 + [_d, _i, _i, _i, _i, _i, _i, _vp, SolverParams, _i, _i, _i, _i]),
    "Hopping_Matrix_32": (None, [_i, _fp, _fp]), "Qtm_pm_psi_32": (None, [_fp, _fp]),
    "D_psi_32": (None, [_fp, _fp]), "Q_pm_psi_32": (None, [_fp, _fp]),
    "mixed_cg_her": (_i, [_sp, _sp, SolverParams, _i, _d, _i, _i, _vp, _vp]),
    "M_ee_inv_ndpsi": (None, [_sp] * 4 + [_d, _d]), "Qtm_ndpsi": (None, [_sp] * 4),
    "Qtm_dagger_ndpsi": (None, [_sp] * 4), "Qtm_pm_ndpsi": (None, [_sp] * 4),
    "cg_her_nd": (_i, [_sp] * 4 + [_i, _d, _i, _i, _vp]),
    "Qtm_pm_ndpsi_32": (None, [_fp] * 4), "rg_mixed_cg_her_nd": (_i, [_sp] * 4 + [SolverParams, _i, _d, _i, _i, _vp, _vp]),
    "invert_doublet_eo": (_i, [_sp] * 8 + [_d, _i, _i, _i, SolverParams, _i, _i, _i]),
    "Mee_inv_psi": (None, [_sp, _sp, _d]), "Mee_psi": (None, [_sp, _sp, _d]),
    "mul_one_pm_imu_sub_mul": (None, [_sp, _sp, _sp, _d, _i]), "mul_one_sub_mul_gamma5": (None, [_sp, _sp, _sp]),
    "M_minus_1_timesC": (None, [_sp] * 4),
    "Qtm_plus_sym_psi": (None, [_sp, _sp]), "Qtm_minus_sym_psi": (None, [_sp, _sp]), "Mtm_plus_sym_psi": (None, [_sp, _sp]),
    "Mtm_minus_sym_psi": (None, [_sp, _sp]), "Mtm_plus_sym_dagg_psi": (None, [_sp, _sp]), "Qtm_pm_sym_psi": (None, [_sp, _sp]),
    "Qtm_plus_sym_psi_nocom": (None, [_sp, _sp]), "Mtm_plus_sym_psi_nocom": (None, [_sp, _sp]),
    "Mtm_minus_sym_psi_nocom": (None, [_sp, _sp]), "Qtm_plus_psi_nocom": (None, [_sp, _sp]),
    "Mtm_plus_psi_nocom": (None, [_sp, _sp]), "Qtm_pm_psi_nocom": (None, [_sp, _sp]),
    "M_minus_psi": (None, [_sp, _sp]), "D_dagg_psi": (None, [_sp, _sp]),
    "zero_spinor_field": (None, [_sp, _i]), "assign_to_32": (None, [_fp, _sp, _i]), "assign_to_64": (None, [_sp, _fp, _i]),
    "addto_32": (None, [_sp, _fp, _i]),
    "init_solver_field": (_i, [C.POINTER(C.POINTER(_vp)), _i, _i]), "finalize_solver": (None, [C.POINTER(_vp), _i]),
    "H_eo_tm_ndpsi": (None, [_sp] * 4 + [_i]), "M_oo_sub_g5_ndpsi": (None, [_sp] * 6 + [_d, _d]),
    "mul_one_pm_iconst": (None, [_sp, _sp, _d, _i]),
    "rg_mixed_cg_her": (_i, [_sp, _sp, SolverParams, _i, _d, _i, _i, _vp, _vp]),
    "deriv_Sb": (None, [_i, _sp, _sp, C.POINTER(HamiltonianField), _d]),
    "chrono_add_solution": (None, [_sp, C.POINTER(_vp), C.POINTER(_i), _i, C.POINTER(_i), _i]),
    "chrono_guess": (_i, [_sp, _sp, C.POINTER(_vp), C.POINTER(_i), _i, _i, _i, _vp]),
    "solve_degenerate": (_i, [_sp, _sp, SolverParams, _i, _d, _i, _i, _vp, _i]),
    "tmb_dropin_register_monomial": (_i, [_i, _i, _d, _d, _d, _d, _i, _i, _d, _d, _i]),
    "tmb_dropin_set_random_spinor_field_eo": (None, [RANDOM_SPINOR_FN]),
    "tmb_dropin_monomial_info": (_i, [_i, C.POINTER(_d), C.POINTER(_d), C.POINTER(_i), C.POINTER(_i)]),
    "det_heatbath": (None, [_i, C.POINTER(HamiltonianField)]), "det_acc": (_d, [_i, C.POINTER(HamiltonianField)]),
    "det_derivative": (None, [_i, C.POINTER(HamiltonianField)]),
    "detratio_heatbath": (None, [_i, C.POINTER(HamiltonianField)]), "detratio_acc": (_d, [_i, C.POINTER(HamiltonianField)]),
    "detratio_derivative": (None, [_i, C.POINTER(HamiltonianField)]),
    "square_norm_32": (C.c_float, [_fp, _i, _i]), "scalar_prod_r_32": (C.c_float, [_fp, _fp, _i, _i]),
    "assign_add_mul_r_32": (None, [_fp, _fp, C.c_float, _i]), "assign_mul_add_r_32": (None, [_fp, C.c_float, _fp, _i]),
    "diff_32": (None, [_fp, _fp, _fp, _i]), "mul_r_32": (None, [_fp, C.c_float, _fp, _i]),
    "assign_mul_add_mul_r_32": (None, [_fp, _fp, C.c_float, C.c_float, _i]), "gamma5_32": (None, [_fp, _fp, _i]),
    "measure_plaquette": (_d, [_vp]),
    "construct_paramsXlfInfo": (_vp, [_d, _i]), "read_gauge_field": (_i, [C.c_char_p, _vp]),
    "write_gauge_field": (_i, [C.c_char_p, _i, _vp]), "read_spinor": (_i, [_sp, _sp, C.c_char_p, _i]),
    "tmb_write_propagator": (_i, [C.c_char_p, _sp, _sp, _i, _d, _i, C.c_char_p, _i]),
    "tmLQCD_b200_set_io": (_i, [C.c_char_p, C.c_char_p, _i]),
    "tmLQCD_invert_init": (_i, [_i, _vp, _i, _i]), "tmLQCD_read_gauge": (_i, [_i]),
    "tmLQCD_invert": (_i, [_sp, _sp, _i, _i]), "tmLQCD_finalise": (_i, []),
    "tmLQCD_get_gauge_field_pointer": (_i, [C.POINTER(C.POINTER(_d))]),
    "tmLQCD_get_mpi_params": (_i, [_vp]), "tmLQCD_get_lat_params": (_i, [_vp]),
    "tmLQCD_b200_set_lattice": (_i, [_i] * 4), "tmLQCD_b200_add_operator": (_i, [_d, _d, _d, _i, _i]),
    "tmLQCD_b200_set_theta": (_i, [_d] * 4), "tmLQCD_b200_get_solver_info": (_i, [_i, C.POINTER(_i), C.POINTER(_d)]),
    "tmLQCD_b200_set_operator_solver": (_i, [_i, _i, _i, _d]),
}
DROPIN_GLOBALS = ["T", "L", "LX", "LY", "LZ", "VOLUME", "RAND", "VOLUMEPLUSRAND", "g_update_gauge_copy", "g_proc_id",
                  "g_debug_level", "g_nproc", "g_nproc_t", "g_nproc_x", "g_nproc_y", "g_nproc_z", "g_kappa", "g_mu", "g_mubar", "g_epsbar", "phmc_invmaxev",
                  "X0", "X1", "X2", "X3", "ka0", "ka1", "ka2", "ka3", "phase_0", "phase_1", "phase_2", "phase_3",
                  "g_gauge_field", "mixcg_innereps", "mixcg_maxinnersolverit", "g_relative_precision_flag",
                  "GaugeInfo", "gauge_precision_read_flag", "g_disable_IO_checks", "g_beta", "g_rgi_C1"]

_SOLVERS = {"cg_her", "invert_eo", "cg_her_nd", "invert_doublet_eo", "mixed_cg_her", "invert_eo_mixed", "invert_eo_rgmixed",
            "rg_mixed_cg_her", "solve_degenerate", "invert_doublet_eo_solver", "rg_mixed_cg_her_nd"}
_lib = None


def load():
    """dlopen the product library; raises if it was not built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise RuntimeError(f"{_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU implementation to fall back to)")
    lib = C.CDLL(_LIB)
    for table in (DEVICE_API, DROPIN_API):
        for name, (res, args) in table.items():
            f = getattr(lib, name)
            f.restype = res
            f.argtypes = args
    _lib = lib
    return lib


class TmbError(RuntimeError):
    pass


def _addr(a):
    """device pointer (int) or host numpy array -> void*"""
    if isinstance(a, np.ndarray):
        assert a.dtype in (np.float64, np.float32) and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(_vp)
    return _vp(a)


class Device:
    """Device-level API: one GPU context, device-resident eo spinor fields."""

    def __init__(self, T, LX, LY, LZ, device=0):
        self.lib = load()
        self.dims = (T, LX, LY, LZ)
        self.V = T * LX * LY * LZ
        self.Vh = self.V // 2
        self.ck(self.lib.tmb_init(T, LX, LY, LZ, device))

    def ck(self, rc):
        if rc is not None and rc < 0:
            raise TmbError(f"rc={rc}: {self.lib.tmb_last_error().decode()}")
        return rc

    def call(self, name, *args):
        """tmb_<name>(*args) with error check; device fields are ints, host arrays numpy."""
        f = getattr(self.lib, "tmb_" + name)
        conv = [(_addr(a) if (isinstance(a, np.ndarray) or (t is _vp and not isinstance(a, _vp))) else a)
                for a, t in zip(args, f.argtypes)]
        rc = f(*conv)
        if name in _SOLVERS:  # iteration count, -1 = not converged (solver/cg_her.c:141), < -1 = error
            if rc < -1:
                self.ck(rc)
            return rc
        return self.ck(rc)

    # ---- fields ----
    def field(self, host=None):
        p = self.lib.tmb_field_alloc()
        if not p:
            raise TmbError(self.lib.tmb_last_error().decode())
        if host is not None:
            self.upload(p, host)
        return p

    def free(self, *fields):
        for p in fields:
            self.ck(self.lib.tmb_field_free(p))

    def upload(self, field, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.size == self.Vh * 24
        self.ck(self.lib.tmb_field_upload(field, host.ctypes.data_as(_vp)))

    def download(self, field):
        out = np.zeros((self.Vh, 24), dtype=np.float64)
        self.ck(self.lib.tmb_field_download(out.ctypes.data_as(_vp), field))
        return out

    def field32(self, host=None):
        p = self.lib.tmb_field32_alloc()
        if not p:
            raise TmbError(self.lib.tmb_last_error().decode())
        if host is not None:
            host = np.ascontiguousarray(host, dtype=np.float32)
            assert host.size == self.Vh * 24
            self.ck(self.lib.tmb_field32_upload(p, host.ctypes.data_as(_vp)))
        return p

    def download32(self, field):
        out = np.zeros((self.Vh, 24), dtype=np.float32)
        self.ck(self.lib.tmb_field32_download(out.ctypes.data_as(_vp), field))
        return out

    def upload_lexic(self, even, odd, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.size == self.V * 24
        self.ck(self.lib.tmb_field_upload_lexic(even, odd, host.ctypes.data_as(_vp)))

    def download_lexic(self, even, odd):
        out = np.zeros((self.V, 24), dtype=np.float64)
        self.ck(self.lib.tmb_field_download_lexic(out.ctypes.data_as(_vp), even, odd))
        return out

    def gauge_upload(self, g):
        g = np.ascontiguousarray(g, dtype=np.float64)
        assert g.size == self.V * 4 * 18
        self.ck(self.lib.tmb_gauge_upload(g.ctypes.data_as(_vp)))

    def set_params(self, kappa, gmu, theta=(0., 0., 0., 0.)):
        self.ck(self.lib.tmb_set_boundary(kappa, np.asarray(theta, dtype=np.float64)))
        self.ck(self.lib.tmb_set_mu(gmu))

    def reduce(self, name, *args):
        out = _d(0.)
        self.ck(getattr(self.lib, "tmb_" + name)(*args, C.byref(out)))
        return out.value

    def timer_start(self):
        self.ck(self.lib.tmb_timer_start())

    def timer_stop(self):
        ms = C.c_float(0.)
        self.ck(self.lib.tmb_timer_stop(C.byref(ms)))
        return ms.value

    # ---- HMC side ----
    def derivative_upload(self, df):
        df = np.ascontiguousarray(df, dtype=np.float64)
        assert df.size == self.V * 32
        self.ck(self.lib.tmb_derivative_upload(df.ctypes.data_as(_vp)))

    def derivative_download(self):
        out = np.zeros((self.V, 4, 8), dtype=np.float64)
        self.ck(self.lib.tmb_derivative_download(out.ctypes.data_as(_vp)))
        return out

    def monomial_info(self, id):
        e0, e1, i0, i1, n = _d(0.), _d(0.), _i(0), _i(0), _i(0)
        self.ck(self.lib.tmb_monomial_info(id, C.byref(e0), C.byref(e1), C.byref(i0), C.byref(i1), C.byref(n)))
        return {"energy0": e0.value, "energy1": e1.value, "iter0": i0.value, "iter1": i1.value, "csg_n": n.value}

    def solver_stats(self):
        it, err, sec = _i(0), _d(0.), _d(0.)
        self.lib.tmb_solver_stats(C.byref(it), C.byref(err), C.byref(sec))
        return it.value, err.value, sec.value

    def close(self):
        self.ck(self.lib.tmb_finalize())


class DropIn:
    """The reference-named entry points with numpy host buffers (reference AoS layouts)."""

    def __init__(self, T, LX, LY, LZ, device=0):
        self.lib = load()
        self.dims = (T, LX, LY, LZ)
        self.V = T * LX * LY * LZ
        self.Vh = self.V // 2
        if self.lib.tmb_dropin_init(T, LX, LY, LZ, device) != 0:
            raise TmbError(self.lib.tmb_last_error().decode())

    def glob(self, name, ctype=_d):
        return ctype.in_dll(self.lib, name)

    def set_params(self, kappa, gmu, theta=(0., 0., 0., 0.)):
        for n, v in zip(("X0", "X1", "X2", "X3"), theta):
            self.glob(n).value = float(v)
        self.glob("g_kappa").value = kappa
        self.glob("g_mu").value = gmu
        self.lib.boundary(kappa)

    def set_nd_params(self, mubar, epsbar, invmaxev):
        self.glob("g_mubar").value = mubar
        self.glob("g_epsbar").value = epsbar
        self.glob("phmc_invmaxev").value = invmaxev

    def set_gauge(self, g):
        """write g_gauge_field[ix][mu] and raise the reference's dirty flag"""
        g = np.ascontiguousarray(g, dtype=np.float64)
        gf = C.POINTER(C.POINTER(_d)).in_dll(self.lib, "g_gauge_field")
        C.memmove(gf[0], g.ctypes.data, g.nbytes)
        self.glob("g_update_gauge_copy", _i).value = 1

    def spinor(self, n=None):
        return np.zeros((self.Vh if n is None else n, 24), dtype=np.float64)

    def hamiltonian_field(self, df):
        """hamiltonian_field_t over g_gauge_field and a caller-owned derivative array df[V][4][8]"""
        assert df.dtype == np.float64 and df.flags["C_CONTIGUOUS"] and df.size == self.V * 32
        rows = (C.POINTER(_d) * self.V)()
        base = df.ctypes.data
        for ix in range(self.V):
            rows[ix] = C.cast(base + ix * 4 * 8 * 8, C.POINTER(_d))
        hf = HamiltonianField()
        hf.gaugefield = C.cast(C.POINTER(C.POINTER(_d)).in_dll(self.lib, "g_gauge_field"), _vp)
        hf.derivative = C.cast(rows, C.POINTER(C.POINTER(_d)))
        hf._keep = (rows, df)
        return hf

    def fptr(self, name):
        return C.cast(getattr(self.lib, name), _vp)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.lib, name)

    def close(self):
        self.lib.tmb_dropin_finalize()

"""ctypes binding of libafmg.so (C ABI: include/afmg.h).

The library is the product; this module only loads it.  There is no CPU fallback: if the
shared library is missing the import of the solver fails loudly, and without a CUDA device
``afmg_create`` returns AFMG_ERR_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libafmg.so")

AFMG_OK = 0
AFMG_MAX_RANKS = 8
AFMG_COMM_BLOB_BYTES = 192
ERR_NAMES = {-1: "AFMG_ERR_ARG", -2: "AFMG_ERR_CUDA", -3: "AFMG_ERR_UNSUPPORTED", -4: "AFMG_ERR_STATE",
             -5: "AFMG_ERR_SINGULAR", -6: "AFMG_ERR_COMM", -7: "AFMG_ERR_NOT_CONVERGED"}


class AfmgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Opts(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("n_cell", C.c_int32), ("coord_t", C.c_int32), ("n_cycle_down", C.c_int32),
        ("n_cycle_up", C.c_int32), ("use_corners", C.c_int32), ("subtract_mean", C.c_int32),
        ("prolongation_type", C.c_int32), ("operator_mask", C.c_int32), ("has_eps", C.c_int32),
        ("device", C.c_int32), ("n_gpus", C.c_int32), ("helmholtz_lambda", C.c_double),
        ("lsf_boundary_value", C.c_double), ("coarse_grid_size", C.c_int32 * 3), ("periodic", C.c_int32 * 3),
        ("dr_base", C.c_double * 3), ("r_base", C.c_double * 3),
    ]


class StencilDesc(C.Structure):
    _fields_ = [
        ("box_id", C.c_int32), ("op_stype", C.c_int32), ("prolong_shape", C.c_int32), ("prolong_stype", C.c_int32),
        ("tag", C.c_int32), ("cylindrical_gradient", C.c_int32), ("op_offset", C.c_int64), ("f_offset", C.c_int64),
        ("prolong_offset", C.c_int64),
    ]


class LsfOpts(C.Structure):
    _fields_ = [("dist_method", C.c_int32), ("gradient_safety_factor", C.c_double), ("length_scale", C.c_double),
                ("tol", C.c_double), ("min_rel_distance", C.c_double)]


LSF_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_void_p)  # afmg_lsf_fn


class Electrode(C.Structure):
    """afmg_electrode: the built-in electrode shapes of src/m_field.f90:254-362."""
    _fields_ = [("type", C.c_int32), ("ndim", C.c_int32), ("electrode_grounded", C.c_int32),
                ("electrode2_grounded", C.c_int32), ("current_voltage", C.c_double),
                ("rod_r0", C.c_double * 3), ("rod_r1", C.c_double * 3), ("rod_radius", C.c_double),
                ("cone_tip_radius", C.c_double), ("cone_length_frac", C.c_double),
                ("rod2_r0", C.c_double * 3), ("rod2_r1", C.c_double * 3), ("rod2_radius", C.c_double),
                ("cone2_tip_radius", C.c_double), ("cone2_length_frac", C.c_double), ("domain_center", C.c_double * 3),
                ("cone_tip_center", C.c_double * 3), ("cone_tip_r_curvature", C.c_double),
                ("cone2_tip_center", C.c_double * 3), ("cone2_tip_r_curvature", C.c_double)]


class TreeDesc(C.Structure):
    _fields_ = [
        ("highest_lvl", C.c_int32), ("highest_id", C.c_int32), ("lvl_counts", C.POINTER(C.c_int32)),
        ("lvl_ids", C.POINTER(C.c_int32)), ("lvl", C.POINTER(C.c_int32)), ("ix", C.POINTER(C.c_int32)),
        ("parent", C.POINTER(C.c_int32)), ("children", C.POINTER(C.c_int32)),
        ("neighbors", C.POINTER(C.c_int32)), ("neighbor_mat", C.POINTER(C.c_int32)),
        ("r_min", C.POINTER(C.c_double)),
    ]


# every symbol include/afmg.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_I = C.c_int32
_IP = C.POINTER(C.c_int32)
_DP = C.POINTER(C.c_double)
SYMBOLS = {
    "afmg_create": (C.c_int, [C.POINTER(_H), C.POINTER(Opts)]),
    "afmg_destroy": (C.c_int, [_H]),
    "afmg_last_error": (C.c_char_p, [_H]),
    "afmg_set_tree": (C.c_int, [_H, C.POINTER(TreeDesc)]),
    "afmg_set_bc": (C.c_int, [_H, _I, _IP, _IP, _IP, _DP]),
    "afmg_set_helmholtz_lambda": (C.c_int, [_H, C.c_double]),
    "afmg_set_lsf_boundary_value": (C.c_int, [_H, C.c_double]),
    "afmg_set_lsf_boundary_values": (C.c_int, [_H, _I, _IP, _DP]),
    "afmg_set_stencils": (C.c_int, [_H, _I, C.c_void_p, _DP, C.c_int64]),
    "afmg_update_operator_stencil": (C.c_int, [_H]),
    "afmg_upload": (C.c_int, [_H, _I, _I, _IP, C.c_void_p]),
    "afmg_download": (C.c_int, [_H, _I, _I, _IP, C.c_void_p]),
    "afmg_upload_interior": (C.c_int, [_H, _I, _I, _IP, C.c_void_p]),
    "afmg_download_interior": (C.c_int, [_H, _I, _I, _IP, C.c_void_p]),
    "afmg_host_alloc": (C.c_void_p, [C.c_size_t]),
    "afmg_host_free": (None, [C.c_void_p]),
    "afmg_upload_device": (C.c_int, [_H, _I, _I, _IP, C.c_void_p]),
    "afmg_download_device": (C.c_int, [_H, _I, _I, _IP, C.c_void_p]),
    "afmg_field_set_rhs": (C.c_int, [_H, _I, _IP, _I, _DP, C.POINTER(C.c_void_p), _I]),
    "afmg_clear": (C.c_int, [_H, _I]),
    "afmg_fas_fmg": (C.c_int, [_H, _I, _I]),
    "afmg_fas_vcycle": (C.c_int, [_H, _I, _I, _I]),
    "afmg_fas_fmg_async": (C.c_int, [_H, _I, _I, _I]),
    "afmg_fas_vcycle_async": (C.c_int, [_H, _I, _I, _I]),
    "afmg_sync": (C.c_int, [_H]),
    "afmg_field_solve": (C.c_int, [_H, _I, C.c_double, C.c_double, _I, _I, _DP, _IP, _IP]),
    "afmg_compute_phi_gradient": (C.c_int, [_H, C.c_double, _I]),
    "afmg_compute_field_norm": (C.c_int, [_H]),
    "afmg_gc_tree": (C.c_int, [_H, _I, _I]),
    "afmg_field_from_potential": (C.c_int, [_H, C.c_double]),
    "afmg_set_fld_bc": (C.c_int, [_H, _I, _IP, _IP, _IP, _DP]),
    "afmg_set_lsf_distances": (C.c_int, [_H, _I, _IP, _IP, _IP, _DP, _DP]),
    "afmg_upload_fc": (C.c_int, [_H, _I, _IP, C.c_void_p]),
    "afmg_download_fc": (C.c_int, [_H, _I, _IP, C.c_void_p]),
    "afmg_helmholtz_compute": (C.c_int, [C.POINTER(_H), _I, _DP, _I, C.c_double, _IP, _DP]),
    "afmg_gsrb_boxes": (C.c_int, [_H, _I, _I]),
    "afmg_gsrb_halfsweep": (C.c_int, [_H, _I, _I]),
    "afmg_gc_lvl": (C.c_int, [_H, _I, _I, _I]),
    "afmg_update_coarse": (C.c_int, [_H, _I, _I]),
    "afmg_correct_children": (C.c_int, [_H, _I]),
    "afmg_correct_children_gc": (C.c_int, [_H, _I]),
    "afmg_residual_lvl": (C.c_int, [_H, _I]),
    "afmg_solve_coarse_grid": (C.c_int, [_H]),
    "afmg_init_phi_rhs": (C.c_int, [_H]),
    "afmg_max_abs": (C.c_int, [_H, _I, _DP]),
    "afmg_tree_sum": (C.c_int, [_H, _I, _DP]),
    "afmg_checksum": (C.c_int, [_H, _I, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "afmg_slab_bytes": (C.c_int, [_H, _I, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _IP]),
    "afmg_set_mega": (C.c_int, [_H, _I, _I]),
    "afmg_mega_active": (C.c_int32, [_H]),
    "afmg_kernel_launches": (C.c_int64, [_H]),
    "afmg_last_cycle_ms": (C.c_int, [_H, _DP]),
    "afmg_set_profiling": (C.c_int, [_H, _I]),
    "afmg_profile": (C.c_int, [_H, _I, C.c_void_p, _DP, C.POINTER(C.c_int64), _IP]),
    "afmg_cell_updates": (C.c_int, [_H, _I, _I, _DP]),
    "afmg_layout_offset": (C.c_int32, [_I, _I, _I, _I, _I]),
    "afmg_layout_box_len": (C.c_int32, [_I, _I]),
    "afmg_morton_key": (C.c_int64, [_I, _I, _I, _I]),
    "afmg_slot_of_box": (C.c_int32, [_H, _I]),
    "afmg_comm_init": (C.c_int, [_H, _I, _I]),
    "afmg_comm_export": (C.c_int, [_H, C.c_void_p]),
    "afmg_comm_connect": (C.c_int, [_H, C.c_void_p]),
    "afmg_owner_of_box": (C.c_int32, [_H, _I]),
    "afmg_partition": (C.c_int, [_I, _I, _IP, _IP]),
    "afmg_partition_min": (C.c_int, [_I, _I, _IP, _I, _IP]),
    "afmg_lsf_opts_default": (None, [C.POINTER(LsfOpts)]),
    "afmg_electrode_prepare": (C.c_int, [C.POINTER(Electrode)]),
    "afmg_electrode_lsf": (C.c_double, [_DP, C.c_void_p]),
    "afmg_electrode_potential": (C.c_double, [_DP, C.c_void_p]),
    "afmg_build_stencils_device": (C.c_int, [_H, C.POINTER(Electrode), C.POINTER(LsfOpts)]),
    "afmg_built_stencils": (C.c_int, [_H, _IP, _IP, _IP, _IP, _DP]),
    "afmg_build_box_tag": (C.c_int32, [_I, _I, _DP, _I]),
    "afmg_build_box_operator": (C.c_int, [_I, _I, _I, _I, _DP, _DP, _DP, _DP, _DP, _DP, _IP, _IP, _IP]),
    "afmg_build_box_prolongation": (C.c_int, [_I, _I, _I, _IP, _DP, _DP, _DP, _IP, _IP]),
    "afmg_build_box_lsf_distances": (C.c_int, [_I, _I, _DP, _DP, LSF_FN, C.c_void_p, C.POINTER(LsfOpts), _DP,
                                               C.POINTER(C.c_uint8), _DP, _IP]),
    "afmg_build_box_lsf_prolong_distances": (C.c_int, [_I, _I, _DP, _DP, _IP, _DP, _DP, LSF_FN, C.c_void_p,
                                                       C.POINTER(LsfOpts), C.POINTER(C.c_uint8), _DP]),
}

_lib = None


def lib():
    """Load libafmg.so; raises if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build the CUDA library first (make -C afivo_streamer_b200/csrc). "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib

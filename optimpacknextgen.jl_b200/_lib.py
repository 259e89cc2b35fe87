"""ctypes binding of libopkb200.so (the C ABI declared in include/opkb200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
present, every entry point raises `OpkbError`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libopkb200.so")

FLOAT, DOUBLE = 0, 1
METHOD_LBFGS, METHOD_BLMVM, METHOD_VMLMB = 0, 1, 2
OBJ_USER, OBJ_ROSENBROCK, OBJ_QUADRATIC = 0, 1, 2
FORWARD, BACKWARD = 1, -1
MEM_MAX = 20        # pairs of the Gram-form kernels
SLOTS_MAX = 256     # pairs a workspace can store (sequential two-loop beyond MEM_MAX)
FG_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double))

i64 = C.c_int64
f64 = C.c_double
vp = C.c_void_p
Pd = C.POINTER(f64)
Pi = C.POINTER(C.c_int)


class OpkbError(RuntimeError):
    """Raised for any non-zero status of the library (text of opkb_last_error)."""

    def __init__(self, code, message):
        super().__init__(f"libopkb200 error {code}: {message}")
        self.code = code


class VmlmbOptions(C.Structure):
    """opkb_vmlmb_options of include/opkb200.h"""
    _fields_ = [("blmvm", C.c_int), ("fmin", f64), ("maxiter", i64), ("maxeval", i64),
                ("xatol", f64), ("xrtol", f64), ("fatol", f64), ("frtol", f64), ("gatol", f64),
                ("grtol", f64), ("epsilon", f64), ("lnsrch", C.c_int), ("ls_ftol", f64),
                ("ls_gtol", f64), ("ls_xtol", f64), ("ls_gamma1", f64), ("ls_gamma2", f64),
                ("ls_strict", C.c_int), ("twoloop", C.c_int), ("speculate", C.c_int)]


TASK_ERROR, TASK_START, TASK_COMPUTE_FG, TASK_NEW_X, TASK_FINAL_X, TASK_WARNING = -1, 0, 1, 2, 3, 4

# name -> (restype, argtypes); the list mirrors include/opkb200.h one for one
_SIGNATURES = {
    "opkb_last_error": (C.c_char_p, []),
    "opkb_version": (C.c_int, []),
    "opkb_context_create": (C.c_int, [C.POINTER(vp), C.c_int]),
    "opkb_context_destroy": (C.c_int, [vp]),
    "opkb_context_synchronize": (C.c_int, [vp]),
    "opkb_context_device": (C.c_int, [vp]),
    "opkb_context_sm_count": (C.c_int, [vp]),
    "opkb_context_stream": (vp, [vp]),
    "opkb_context_wait_stream": (C.c_int, [vp, vp]),
    "opkb_context_trim": (C.c_int, [vp]),
    "opkb_context_cached_bytes": (i64, [vp]),
    "opkb_context_launch_count": (i64, [vp]),
    "opkb_timer_start": (C.c_int, [vp]),
    "opkb_timer_stop": (C.c_int, [vp, Pd]),
    "opkb_profile_enable": (C.c_int, [vp, C.c_int]),
    "opkb_profile_reset": (C.c_int, [vp]),
    "opkb_profile_get": (C.c_int, [vp, C.c_int, Pd, C.POINTER(i64)]),
    "opkb_flush_l2": (C.c_int, [vp]),
    "opkb_comm_unique_id": (C.c_int, [vp]),
    "opkb_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    "opkb_comm_uses_mailbox": (C.c_int, [vp]),
    "opkb_comm_rank": (C.c_int, [vp]),
    "opkb_comm_size": (C.c_int, [vp]),
    "opkb_group_create": (C.c_int, [C.POINTER(vp), C.c_int, Pi]),
    "opkb_group_destroy": (C.c_int, [vp]),
    "opkb_group_size": (C.c_int, [vp]),
    "opkb_group_context": (vp, [vp, C.c_int]),
    "opkb_group_solver_run": (C.c_int, [vp, C.POINTER(vp), i64, Pi]),
    "opkb_group_spg_solver_run": (C.c_int, [vp, C.POINTER(vp), i64, Pi]),
    "opkb_group_run": (C.c_int, [vp, vp, vp]),
    "opkb_vector_create": (C.c_int, [vp, C.c_int, i64, C.POINTER(vp)]),
    "opkb_vector_destroy": (C.c_int, [vp]),
    "opkb_vector_length": (i64, [vp]),
    "opkb_vector_eltype": (C.c_int, [vp]),
    "opkb_vector_data": (vp, [vp]),
    "opkb_vector_upload": (C.c_int, [vp, vp, i64, i64]),
    "opkb_vector_download": (C.c_int, [vp, vp, i64, i64]),
    "opkb_vector_fill": (C.c_int, [vp, f64]),
    "opkb_vector_fill_random": (C.c_int, [vp, C.c_uint64, i64, f64, f64, C.c_int]),
    "opkb_vdot": (C.c_int, [vp, vp, vp, Pd]),
    "opkb_vdot_masked": (C.c_int, [vp, vp, vp, vp, Pd]),
    "opkb_vnorm2": (C.c_int, [vp, vp, Pd]),
    "opkb_vnorminf": (C.c_int, [vp, vp, Pd]),
    "opkb_vcopy": (C.c_int, [vp, vp, vp]),
    "opkb_vscale": (C.c_int, [vp, vp, f64]),
    "opkb_vupdate": (C.c_int, [vp, vp, f64, vp]),
    "opkb_vupdate_masked": (C.c_int, [vp, vp, vp, f64, vp]),
    "opkb_vcombine": (C.c_int, [vp, vp, f64, vp, f64, vp]),
    "opkb_vproduct": (C.c_int, [vp, vp, vp, vp]),
    "opkb_vswap": (C.c_int, [vp, vp, vp]),
    "opkb_vnorm1": (C.c_int, [vp, vp, Pd]),
    "opkb_vcombine3": (C.c_int, [vp, vp, f64, vp, f64, vp, f64, vp]),
    "opkb_vproduct_masked": (C.c_int, [vp, vp, vp, vp, vp]),
    "opkb_project_variables": (C.c_int, [vp, vp, vp, vp, f64, vp, f64]),
    "opkb_project_direction": (C.c_int, [vp, vp, vp, vp, f64, vp, f64, C.c_int, vp]),
    "opkb_step_limits": (C.c_int, [vp, vp, vp, f64, vp, f64, C.c_int, vp, Pd, Pd]),
    "opkb_free_variables": (C.c_int, [vp, vp, C.POINTER(i64), vp]),
    "opkb_rosenbrock_fg": (C.c_int, [vp, vp, vp, Pd]),
    "opkb_quadratic_fg": (C.c_int, [vp, vp, vp, vp, vp, Pd]),
    "opkb_smooth2d_fg": (C.c_int, [vp, vp, vp, f64, i64, vp, vp, vp, vp, Pd]),
    "opkb_smooth2d_apply": (C.c_int, [vp, vp, f64, i64, vp, vp, vp, vp, Pd]),
    "opkb_cg_smooth2d_step": (C.c_int, [vp, vp, f64, i64, vp, vp, vp, vp, vp, vp, vp, vp, f64, C.c_int, Pd]),
    "opkb_cg_update": (C.c_int, [vp, vp, vp, vp, vp, f64, Pd]),
    "opkb_cg_curvature": (C.c_int, [vp, vp, vp, Pd]),
    "opkb_cg_diag_step": (C.c_int, [vp, vp, vp, vp, vp, f64, C.c_int, Pd]),
    "opkb_vmlmb_create": (C.c_int, [vp, C.c_int, i64, C.c_int, C.c_int, vp, f64, vp, f64,
                                    C.POINTER(vp)]),
    "opkb_vmlmb_destroy": (C.c_int, [vp]),
    "opkb_vmlmb_set_objective": (C.c_int, [vp, C.c_int, vp, vp]),
    "opkb_vmlmb_vector": (vp, [vp, C.c_int, C.c_int]),
    "opkb_vmlmb_trial": (C.c_int, [vp, C.c_int, f64, C.c_int, Pd]),
    "opkb_vmlmb_trial_begin": (C.c_int, [vp, C.c_int, f64, C.c_int]),
    "opkb_vmlmb_trial_end": (C.c_int, [vp, C.c_int, f64, C.c_int, Pd]),
    "opkb_vmlmb_gram": (C.c_int, [vp, C.c_int, C.c_int, Pi, Pd]),
    "opkb_vmlmb_direction": (C.c_int, [vp, C.c_int, Pi, Pd, Pd, f64, C.c_int, Pd]),
    "opkb_vmlmb_direction_trial": (C.c_int, [vp, C.c_int, Pi, Pd, Pd, f64, C.c_int, Pd, Pd]),
    "opkb_vmlmb_discard_trial": (C.c_int, [vp, C.c_int]),
    "opkb_vmlmb_history": (C.c_int, [vp]),
    "opkb_vmlmb_carry_stats": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]),
    "opkb_vmlmb_chain_stats": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
    "opkb_vmlmb_chain_check_stats": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]),
    "opkb_vmlmb_last_gram": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(f64)]),
    "opkb_vmlmb_pair_update": (C.c_int, [vp, C.c_int, Pd]),
    "opkb_vmlmb_twoloop_step": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, f64, C.c_int, f64, C.c_int,
                                          C.c_int, C.c_int, Pd]),
    "opkb_vmlmb_finish_direction": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, f64, Pd]),
    "opkb_vmlmb_steepest": (C.c_int, [vp, C.c_int, Pd]),
    "opkb_vmlmb_revert": (C.c_int, [vp, C.c_int, f64]),
    "opkb_vmlmb_default_options": (None, [C.POINTER(VmlmbOptions)]),
    "opkb_solver_create": (C.c_int, [vp, C.POINTER(VmlmbOptions), C.POINTER(vp)]),
    "opkb_solver_destroy": (C.c_int, [vp]),
    "opkb_solver_start": (C.c_int, [vp, Pi]),
    "opkb_solver_iterate": (C.c_int, [vp, f64, Pi]),
    "opkb_solver_run": (C.c_int, [vp, i64, Pi]),
    "opkb_solver_set_fg_callback": (C.c_int, [vp, vp, vp, C.c_int]),
    "opkb_solver_set_chain": (C.c_int, [vp, C.c_int]),
    "opkb_solver_chain_enabled": (C.c_int, [vp]),
    "opkb_solver_status": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), Pd, Pd, Pd,
                                     Pi, Pi]),
    "opkb_solver_reason": (C.c_char_p, [vp]),
    "opkb_solver_mark": (C.c_int, [vp]),
    "opkb_spg_create": (C.c_int, [vp, C.c_int, i64, vp, f64, vp, f64, C.POINTER(vp)]),
    "opkb_spg_destroy": (C.c_int, [vp]),
    "opkb_spg_set_objective": (C.c_int, [vp, C.c_int, vp, vp]),
    "opkb_spg_vector": (vp, [vp, C.c_int]),
    "opkb_spg_start": (C.c_int, [vp, f64, Pd]),
    "opkb_spg_step": (C.c_int, [vp, f64, f64, Pd]),
    "opkb_spg_backtrack": (C.c_int, [vp, f64, f64, f64, Pd]),
    "opkb_spg_mark_best": (C.c_int, [vp]),
    "opkb_spg_solver_create": (C.c_int, [vp, C.c_int, i64, i64, f64, f64, f64, C.POINTER(vp)]),
    "opkb_spg_solver_destroy": (C.c_int, [vp]),
    "opkb_spg_solver_run": (C.c_int, [vp, i64, Pi]),
    "opkb_spg_solver_info": (C.c_int, [vp, Pd, Pd, Pd, Pd, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64),
                                      Pi, Pi]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_LIB = None


def lib():
    """Load libopkb200.so (once).  Fails loudly when the extension is absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise OpkbError(
                -2, f"{SO_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
        L = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _LIB = L
    return _LIB


def check(rc: int) -> None:
    if rc != 0:
        raise OpkbError(rc, lib().opkb_last_error().decode(errors="replace"))

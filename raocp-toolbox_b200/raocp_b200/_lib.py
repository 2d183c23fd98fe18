"""ctypes binding of libraocp_b200.so (C-ABI: include/raocp_b200.h).

There is NO CPU fallback: if the library has not been built, or no CUDA device is usable, every compute entry point
raises.  Build with `python __graft_entry__.py` (or `__graft_entry__.build()`), which runs nvcc for sm_100a.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libraocp_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int32)


class RbProblem(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("m", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32),
        ("num_stages", C.c_int32), ("batch", C.c_int32),
        ("stage_off", c_int_p), ("parent", c_int_p), ("child_first", c_int_p), ("child_count", c_int_p),
        ("num_dyn", C.c_int32), ("dyn_idx", c_int_p), ("A", c_double_p), ("B", c_double_p),
        ("num_cost", C.c_int32), ("cost_idx", c_int_p), ("sqrtQ", c_double_p), ("sqrtR", c_double_p),
        ("num_leafcost", C.c_int32), ("leafcost_idx", c_int_p), ("sqrtQf", c_double_p),
        ("num_nl_rect", C.c_int32), ("nl_rect_idx", c_int_p), ("nl_lo", c_double_p), ("nl_hi", c_double_p),
        ("num_leaf_rect", C.c_int32), ("leaf_rect_idx", c_int_p), ("leaf_lo", c_double_p), ("leaf_hi", c_double_p),
        ("risk_alpha", c_double_p), ("cond_prob", c_double_p),
        ("num_cls", C.c_int32), ("cls", c_int_p),
        ("device", C.c_int32),
        ("shard_rank", C.c_int32), ("shard_world", C.c_int32),
        ("sweep_cut1_min", C.c_int32), ("sweep_cut2_min", C.c_int32),
    ]


# every symbol include/raocp_b200.h declares: (name, restype, argtypes)
_H = C.c_void_p
SYMBOLS = [
    ("rb_create", C.c_int, [C.POINTER(RbProblem), C.POINTER(_H)]),
    ("rb_destroy", None, [_H]),
    ("rb_last_error", C.c_char_p, [_H]),
    ("rb_set_stream", C.c_int, [_H, C.c_void_p]),
    ("rb_sizes", C.c_int, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("rb_synchronize", C.c_int, [_H]),
    ("rb_offline", C.c_int, [_H]),
    ("rb_get_offline", C.c_int, [_H, c_double_p, c_double_p, c_double_p]),
    ("rb_set_primal", C.c_int, [_H, C.c_int, c_double_p]),
    ("rb_get_primal", C.c_int, [_H, C.c_int, c_double_p]),
    ("rb_set_dual", C.c_int, [_H, C.c_int, c_double_p]),
    ("rb_get_dual", C.c_int, [_H, C.c_int, c_double_p]),
    ("rb_set_initial_state", C.c_int, [_H, c_double_p]),
    ("rb_update_cache", C.c_int, [_H]),
    ("rb_apply_L", C.c_int, [_H, c_double_p, c_double_p]),
    ("rb_apply_Lt", C.c_int, [_H, c_double_p, c_double_p]),
    ("rb_lambda_max", C.c_int, [_H, c_double_p]),
    ("rb_primal_half", C.c_int, [_H, C.c_double]),
    ("rb_prox_f", C.c_int, [_H, C.c_double]),
    ("rb_s0_shift", C.c_int, [_H, C.c_double]),
    ("rb_project_dynamics", C.c_int, [_H]),
    ("rb_project_kernel", C.c_int, [_H]),
    ("rb_dual_half", C.c_int, [_H, C.c_double]),
    ("rb_prox_g_conj", C.c_int, [_H, C.c_double]),
    ("rb_modify_dual", C.c_int, [_H, C.c_double]),
    ("rb_add_halves", C.c_int, [_H]),
    ("rb_project_nonleaf", C.c_int, [_H]),
    ("rb_project_leaf", C.c_int, [_H]),
    ("rb_modify_projection", C.c_int, [_H, C.c_double, c_double_p]),
    ("rb_residuals", C.c_int, [_H, C.c_double, c_double_p, c_double_p]),
    ("rb_iterate", C.c_int, [_H, C.c_double, C.c_int32, C.c_double, C.c_int32, c_double_p, c_double_p, C.c_int32,
                             c_int_p, c_int_p]),
    ("rb_iterate_fixed", C.c_int, [_H, C.c_double, C.c_int32, c_double_p]),
    ("rb_loop_begin", C.c_int, [_H, C.c_double, C.c_int32, C.c_double, C.c_int32]),
    ("rb_loop_enqueue", C.c_int, [_H, C.c_int32]),
    ("rb_loop_poll", C.c_int, [_H, c_int_p, c_int_p, c_double_p]),
    ("rb_step", C.c_int, [_H, c_double_p, c_double_p]),
    ("rb_loop_end", C.c_int, [_H, c_double_p, c_double_p, c_int_p, c_int_p]),
    ("rb_profile_iteration", C.c_int, [_H, C.POINTER(C.c_float)]),
    ("rb_use_graphs", C.c_int, [_H, C.c_int32]),
    ("rb_use_batch_panels", C.c_int, [_H, C.c_int32]),
    ("rb_use_launch_overlap", C.c_int, [_H, C.c_int32]),
    ("rb_use_table_prefetch", C.c_int, [_H, C.c_int32]),
    ("rb_use_fused_check", C.c_int, [_H, C.c_int32]),
    ("rb_use_pipeline", C.c_int, [_H, C.c_int32]),
    ("rb_pipeline_info", C.c_int, [_H, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("rb_force_dense_costs", C.c_int, [_H, C.c_int32]),
    ("rb_use_lane_kernels", C.c_int, [_H, C.c_int32]),
    ("rb_use_mma_sweeps", C.c_int, [_H, C.c_int32]),
    ("rb_use_tree_kernels", C.c_int, [_H, C.c_int32]),
    ("rb_shard_unique_id", C.c_int, [C.c_char_p]),
    ("rb_shard_init", C.c_int, [_H, C.c_char_p]),
    ("rb_shard_p2p_export", C.c_int, [_H, C.c_char_p]),
    ("rb_shard_p2p_open", C.c_int, [_H, C.c_char_p]),
    ("rb_shard_info", C.c_int, [_H, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("rb_launch_count", C.c_int, [_H, C.POINTER(C.c_int64)]),
    ("rb_cone_project", C.c_int, [C.c_int32, C.c_int32, c_double_p, c_double_p]),
    ("rb_box_project", C.c_int, [C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p]),
]

_lib = None


def load():
    """dlopen the library and set the prototypes; raises if it has not been built"""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build the CUDA library first (python __graft_entry__.py). "
                               "raocp_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


ERR_NUMERIC = -4


def dptr(arr):
    return arr.ctypes.data_as(c_double_p)


def iptr(arr):
    return arr.ctypes.data_as(c_int_p)


def check(rc, handle=None):
    if rc == 0:
        return
    msg = load().rb_last_error(handle).decode()
    if rc == ERR_NUMERIC and "Rectangle" in msg:
        raise ValueError(msg)      # the reference raises ValueError here (rectangle.py:58-59)
    raise Exception(f"raocp_b200 error {rc}: {msg}")


def _vec(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))


def cone_project(code, vector):
    """device projection of one vector (cones.py:30-132)"""
    flat = _vec(vector)
    out = np.empty_like(flat)
    check(load().rb_cone_project(code, flat.size, dptr(flat), dptr(out)))
    return out.reshape(np.asarray(vector).shape)


def box_project(vector, lo, hi):
    """device clip of one vector (rectangle.py:29-35)"""
    flat, lo, hi = _vec(vector), _vec(lo), _vec(hi)
    out = np.empty_like(flat)
    check(load().rb_box_project(flat.size, dptr(flat), dptr(lo), dptr(hi), dptr(out)))
    return out.reshape(np.asarray(vector).shape)

"""ctypes binding of libcodes_b200.so (see include/codes_b200.h).

The product path has NO fallback: if the shared library is missing or a call
fails, a ``RuntimeError`` is raised (it surfaces through the reference's per-task
``try/except`` in ``codes/train/train_fcts.py:241-248`` like any other error).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

CB_MAX_LAYERS = 16

ACT_IDS = {
    "identity": 0, "relu": 1, "leakyrelu": 2, "tanh": 3, "gelu": 4, "elu": 5,
    "silu": 6, "softplus": 7, "mish": 8, "sigmoid": 9, "prelu": 10,
}
LOSS_IDS = {"mse": 0, "smoothl1": 1, "l1": 2}


class MlpDesc(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32),
        ("dims", C.c_int32 * (CB_MAX_LAYERS + 1)),
        ("act", C.c_int32),
        ("final_act", C.c_int32),
        ("act_param", C.c_float),
    ]


class LnodeDesc(C.Structure):
    _fields_ = [
        ("ode", MlpDesc),
        ("tanh_reg", C.c_int32),
        ("reg_factor", C.c_float),
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("max_steps", C.c_int32),
        ("n_err", C.c_int32),
    ]


class LnodeTape(C.Structure):
    _fields_ = [("d_n", C.c_void_p), ("d_t", C.c_void_p), ("d_dt", C.c_void_p), ("d_y", C.c_void_p),
                ("d_k", C.c_void_p), ("cap", C.c_int32)]


_LIB = None
_LOCK = threading.Lock()
LIB_PATH = os.environ.get("CODES_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                           "libcodes_b200.so")   # (override: kernel-variant experiments)

_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_I32, _I64, _F = C.c_int32, C.c_int64, C.c_float

_SIGNATURES = {
    "cb_last_error": (C.c_char_p, []),
    "cb_version": (_I32, []),
    "cb_launch_count": (_I64, []),
    "cb_reset_launch_count": (None, []),
    "cb_mlp_saved_floats": (_I64, [C.POINTER(MlpDesc), _I64]),
    "cb_mlp_workspace_floats": (_I64, [C.POINTER(MlpDesc), _I64, _I32]),
    "cb_mlp_forward_f32": (_I32, [C.POINTER(MlpDesc), _PP, _PP, _P, _I64, _P, _P, _P, _I64, _P]),
    "cb_mlp_backward_f32": (_I32, [C.POINTER(MlpDesc), _PP, _P, _P, _P, _I64, _PP, _PP, _P, _P, _I64, _P]),
    "cb_mlp_forward_f32_p": (_I32, [C.POINTER(MlpDesc), _PP, _PP, _P, _I64, _P, _P, _P, _I64, _P, _P]),
    "cb_mlp_backward_f32_p": (_I32, [C.POINTER(MlpDesc), _PP, _P, _P, _P, _I64, _PP, _PP, _P, _P, _I64, _P, _P, _P]),
    "cb_deeponet_combine_fwd": (_I32, [_P, _P, _I64, _I32, _I32, _P, _P]),
    "cb_deeponet_combine_bwd": (_I32, [_P, _P, _P, _I64, _I32, _I32, _P, _P, _P]),
    "cb_deeponet_grid_combine_f32": (_I32, [_P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I32, _P, _P]),
    "cb_poly_eval_fwd": (_I32, [_P, _P, _P, _I64, _I32, _I32, _I32, _P, _P]),
    "cb_poly_eval_bwd": (_I32, [_P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _I64, _P]),
    "cb_poly_eval_batched_fwd": (_I32, [_P, _P, _P, _I64, _I32, _I32, _I32, _P, _P]),
    "cb_poly_eval_batched_bwd": (_I32, [_P, _P, _I64, _I32, _I32, _I32, _P, _P, _P]),
    "cb_pointwise_loss": (_I32, [_P, _P, _I64, _I32, _F, _P, _P, _P, _I64, _P]),
    "cb_error_sums_workspace_doubles": (_I64, []),
    "cb_error_sums": (_I32, [_P, _P, _I64, _P, _P, _I64, _P]),
    "cb_latent_loss_fwd_bwd": (_I32, [_P, _P, _I64, _I32, _I32, _I32, _I32, _F, _F, _F, _P, _P, _P, _I64, _P]),
    "cb_adamw_step": (_I32, [_PP, _PP, _PP, _PP, C.POINTER(_I64), _I32, _F, _F, _F, _F, _F, _I64, _P]),
    "cb_sgd_step": (_I32, [_PP, _PP, _PP, C.POINTER(_I64), _I32, _F, _F, _F, _I64, _P]),
    "cb_schedulefree_adamw_step": (_I32, [_PP, _PP, _PP, _PP, C.POINTER(_I64), _I32, _F, _F, _F, _F, _F, _F, _F, _P]),
    "cb_schedulefree_sgd_step": (_I32, [_PP, _PP, _PP, C.POINTER(_I64), _I32, _F, _F, _F, _F, _P]),
    "cb_set_floats": (_I32, [_P, _I32, C.POINTER(_F), _P]),
    "cb_optimizer_step_dev": (_I32, [_I32, _PP, _PP, _PP, _PP, C.POINTER(_I64), _I32, _P, _P]),
    "cb_lerp_tensors": (_I32, [_PP, _PP, C.POINTER(_I64), _I32, _F, _P]),
    "cb_gather_rows": (_I32, [_P, _P, _I64, _I64, _I32, _P, _P]),
    "cb_denormalize": (_I32, [_P, _I64, _I32, _I32, _P, _P, _I32, _P, _P]),
    "cb_lnode_workspace_floats": (_I64, [C.POINTER(LnodeDesc)]),
    "cb_lnode_solve_fwd": (_I32, [C.POINTER(LnodeDesc), _PP, _PP, _P, _P, _I64, _I32, _P, _P, _P, _I64,
                                  C.POINTER(LnodeTape), _P]),
    "cb_lnode_bwd_workspace_floats": (_I64, [C.POINTER(LnodeDesc), _I64]),
    "cb_lnode_bwd_supported": (_I32, [C.POINTER(LnodeDesc), _I32]),
    "cb_lnode_solve_bwd": (_I32, [C.POINTER(LnodeDesc), _PP, _PP, _P, _I64, _I32, C.POINTER(LnodeTape), _P, _P, _PP,
                                  _PP, _P, _P, _I64, _P]),
    "cb_tc_supported": (_I32, []),
    "cb_tc_chain_supported": (_I32, [C.POINTER(MlpDesc)]),
    "cb_tc_chain_create": (_I32, [C.POINTER(MlpDesc), _I32, C.POINTER(_P)]),
    "cb_tc_chain_destroy": (None, [_P]),
    "cb_tc_chain_set_weights": (_I32, [_P, _PP, _PP, _P]),
    "cb_tc_chain_forward": (_I32, [_P, _P, _I64, _P, _P]),
    "cb_tc_chain_create_grouped": (_I32, [C.POINTER(MlpDesc), _I32, _I32, C.POINTER(_P)]),
    "cb_tc_chain_set_weights_group": (_I32, [_P, _I32, _PP, _PP, _P]),
    "cb_tc_chain_forward_grouped": (_I32, [_P, _P, _I64, _I64, _P, _P]),
    "cb_tc_multionet_supported": (_I32, [C.POINTER(MlpDesc), C.POINTER(MlpDesc), _I32]),
    "cb_tc_multionet_create": (_I32, [C.POINTER(MlpDesc), C.POINTER(MlpDesc), _I32, _I32, C.POINTER(_P)]),
    "cb_tc_multionet_destroy": (None, [_P]),
    "cb_tc_multionet_workspace_bytes": (_I64, [_P, _I64]),
    "cb_tc_multionet_set_weights": (_I32, [_P, _PP, _PP, _PP, _PP, _P]),
    "cb_tc_multionet_forward": (_I32, [_P, _P, _P, _I64, _P, _P, _I64, _P]),
    "cb_tc_multionet_forward_timed": (_I32, [_P, _P, _P, _I64, _P, _P, _I64, _P, C.POINTER(C.c_float)]),
    "cb_tc_multionet_final_train": (_I32, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _I64, _P]),
    "cb_tc_multionet_grid_supported": (_I32, [C.POINTER(MlpDesc), _I32]),
    "cb_tc_multionet_grid_create": (_I32, [C.POINTER(MlpDesc), _I32, _I32, C.POINTER(_P)]),
    "cb_tc_multionet_grid_destroy": (None, [_P]),
    "cb_tc_multionet_grid_set_weights": (_I32, [_P, _PP, _PP, _P]),
    "cb_tc_multionet_grid_workspace_bytes": (_I64, [_P, _I64, _I32]),
    "cb_tc_multionet_grid_forward": (_I32, [_P, _P, _I64, _P, _I32, _P, _I32, _P, _P, _I32, _P, _I64, _P]),
    "cb_tc_multionet_grid_forward_timed": (_I32, [_P, _P, _I64, _P, _I32, _P, _I32, _P, _P, _I32, _P, _I64, _P,
                                                  C.POINTER(C.c_float)]),
    "cb_tc_train_create": (_I32, [C.POINTER(MlpDesc), _I32, _I32, C.POINTER(_P)]),
    "cb_tc_train_destroy": (None, [_P]),
    "cb_tc_train_ld": (_I32, [_P, _I32]),
    "cb_tc_train_saved_bytes": (_I64, [_P, _I64]),
    "cb_tc_train_scratch_bytes": (_I64, [_P, _I64]),
    "cb_tc_train_forward": (_I32, [_P, _PP, _PP, _P, _I64, _P, _P, _I64, _P]),
    "cb_tc_train_forward_hidden": (_I32, [_P, _PP, _PP, _P, _I64, _P, _I64, _P]),
    "cb_tc_train_saved_offset": (_I64, [_P, _I64, _I32]),
    "cb_tc_train_infer_bytes": (_I64, [_P, _I64]),
    "cb_tc_train_infer": (_I32, [_P, _PP, _PP, _P, _I64, _P, _P, _I64, _P]),
    "cb_tc_x3_create": (_I32, [C.POINTER(MlpDesc), _I32, C.POINTER(_P)]),
    "cb_tc_x3_destroy": (None, [_P]),
    "cb_tc_x3_set_weights": (_I32, [_P, _PP, _PP, _P]),
    "cb_tc_x3_workspace_bytes": (_I64, [_P, _I64]),
    "cb_tc_x3_forward": (_I32, [_P, _P, _I64, _P, _P, _I64, _P]),
    "cb_tc_train_backward": (_I32, [_P, _P, _P, _I64, _P, _P, _I64, _PP, _PP, _P, _P]),
    "cb_tc_combine_fwd": (_I32, [_P, _P, _I64, _I32, _I32, _I32, _P, _P]),
    "cb_tc_combine_bwd": (_I32, [_P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P]),
}


def exported_symbols():
    """Names every build of the library must export (checked by the CPU tests)."""
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise if it is missing."""
    global _LIB
    if _LIB is None:
        with _LOCK:
            if _LIB is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"codes_b200: native library not found at {LIB_PATH}; run "
                        "`python codes-benchmark_b200/build.py` (there is no CPU fallback)")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _LIB = handle
    return _LIB


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().cb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"codes_b200.{what} failed (status {status}): {msg}")


def make_desc(dims, act: str, final_act: str = "identity", act_param: float = 0.01) -> MlpDesc:
    n = len(dims) - 1
    if not 1 <= n <= CB_MAX_LAYERS:
        raise ValueError(f"MLP chain with {n} Linear layers unsupported (max {CB_MAX_LAYERS})")
    d = MlpDesc()
    d.n_layers = n
    for i, v in enumerate(dims):
        d.dims[i] = int(v)
    d.act = ACT_IDS[act]
    d.final_act = ACT_IDS[final_act]
    d.act_param = float(act_param)
    return d


def ptr_array(tensors):
    """Host array of device pointers (None -> NULL)."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def launch_count() -> int:
    return int(lib().cb_launch_count())


def reset_launch_count() -> None:
    lib().cb_reset_launch_count()

"""ctypes binding of libsr4d.so (C ABI declared in include/sr4d.h).

There is no CPU or PyTorch fallback: importing the engine without the compiled
library raises, and creating a handle without an sm_100 GPU raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsr4d.so")

OK, EINVAL, ENODEVICE, ECUDA, ENOMEM, ESTATE = 0, -1, -2, -3, -4, -5
OPT_CONV_IMPL, OPT_SAVE_ACTS, OPT_PROFILE, OPT_FUSED_DGRAD, OPT_DGRAD_SINGLE, OPT_WGRAD_SINGLE, OPT_NVTX = 1, 2, 3, 4, 5, 6, 7
OPT_FWD_CHAIN = 8
PROF_CLASSES = ("conv64_fwd_lr", "conv64_fwd_hr", "conv64_dgrad_lr", "conv64_dgrad_hr", "conv64_wgrad_lr",
                "conv64_wgrad_hr")
CONV_AUTO, CONV_SIMT, CONV_TCGEN05 = 0, 1, 2
METRIC_TAIL = 4096     # SR4D_METRIC_TAIL

ERRNAMES = {EINVAL: "SR4D_EINVAL", ENODEVICE: "SR4D_ENODEVICE", ECUDA: "SR4D_ECUDA",
            ENOMEM: "SR4D_ENOMEM", ESTATE: "SR4D_ESTATE"}


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("offset", C.c_int64), ("count", C.c_int64),
                ("ndim", C.c_int32), ("shape", C.c_int32 * 5), ("is_kernel", C.c_int32)]


# every symbol include/sr4d.h declares: (restype, argtypes)
_P, _F = C.c_void_p, C.c_float
SYMBOLS = {
    "sr4d_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "sr4d_destroy": (None, [_P]),
    "sr4d_last_error": (C.c_char_p, [_P]),
    "sr4d_version": (C.c_char_p, []),
    "sr4d_set_option": (C.c_int, [_P, C.c_int, C.c_int]),
    "sr4d_get_option": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    "sr4d_param_count": (C.c_int64, [_P]),
    "sr4d_flat_size": (C.c_int64, [_P]),
    "sr4d_num_tensors": (C.c_int, [_P]),
    "sr4d_param_table": (C.c_int, [_P, C.POINTER(TensorDesc), C.c_int]),
    "sr4d_params": (_P, [_P]),
    "sr4d_grads": (_P, [_P]),
    "sr4d_adam_m": (_P, [_P]),
    "sr4d_adam_v": (_P, [_P]),
    "sr4d_params_changed": (C.c_int, [_P, _P]),
    "sr4d_forward": (C.c_int, [_P] + [_P] * 6 + [_P, C.c_int, _P]),
    "sr4d_loss_metrics": (C.c_int, [_P] + [_P] * 5 + [C.c_int, _P, _P]),
    "sr4d_train_fwd_bwd": (C.c_int, [_P] + [_P] * 10 + [C.c_int, _P, _P, _P, _P]),
    "sr4d_train_forward": (C.c_int, [_P] + [_P] * 6 + [C.c_int, _P, _P]),
    "sr4d_train_backward": (C.c_int, [_P] + [_P] * 4 + [C.c_int, _P, _P, _P]),
    "sr4d_adam_step": (C.c_int, [_P, _F, _F, _F, _F, C.c_int64, _F, _P]),
    "sr4d_adam_step_counted": (C.c_int, [_P, _F, _F, _F, _F, C.c_int64, _F, C.c_int, _P]),
    "sr4d_grads_size": (C.c_int64, [_P]),
    "sr4d_stitch": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _F, C.c_int, _P, _P]),
    "sr4d_conv64_layer": (C.c_int, [_P, _P, _P, _P, _P, _F, _P, C.c_int, C.c_int, C.c_int, _P]),
    "sr4d_upsample_layer": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "sr4d_conv64_layer_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "sr4d_head_layer_bwd": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "sr4d_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "sr4d_activation_overflow": (C.c_int, [_P, C.POINTER(C.c_int), C.c_int, _P]),
    "sr4d_launch_count": (C.c_int64, [_P]),
    "sr4d_reset_launch_count": (None, [_P]),
}

_lib = None


def load():
    """Load libsr4d.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)       # AttributeError if the ABI drifted from the header
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

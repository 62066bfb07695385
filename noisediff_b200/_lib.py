"""ctypes binding of the C ABI declared in ``include/noisediff_b200.h`` (the library lives in-tree at
``noisediff_b200/csrc/libnoisediff_b200.so``; ``__graft_entry__.build()`` / ``make -C noisediff_b200/csrc`` builds it).
There is no fallback: a missing library is an ImportError-like RuntimeError at first use."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libnoisediff_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "noisediff_b200.h")

FLAG_CONV_DIRECT = 1
FLAG_NO_GRAPH = 2
FLAG_KEEP_ACTIVATIONS = 4
FLAG_INIT_SIMT = 8
FLAG_UNFUSED = 16
FLAG_PDL = 32
FLAG_HALO1 = 64
FLAG_NO_XF = 128
FLAG_INIT_WINDOWS = 256


class Config(C.Structure):
    _fields_ = [("dim", C.c_int32), ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("device", C.c_int32), ("flags", C.c_int32)]


class ConvEx(C.Structure):      # ndiff_conv_ex
    _fields_ = [("out2", C.c_void_p), ("bias2", C.c_void_p), ("xf_stats", C.c_void_p), ("xf_gamma", C.c_void_p),
                ("xf_beta", C.c_void_p), ("xf_ss", C.c_void_p), ("xf_ss_ld", C.c_int32), ("xf_groups", C.c_int32)]


class Step(C.Structure):
    _fields_ = [("t", C.c_int32), ("p", C.c_float), ("q", C.c_float), ("a", C.c_float), ("b", C.c_float),
                ("c", C.c_float), ("r1", C.c_float), ("r2", C.c_float), ("sigma", C.c_float), ("clip", C.c_int32),
                ("reserved", C.c_int32 * 2)]


_P = C.c_void_p
_SIGS = {
    "ndiff_abi_version": (C.c_int32, []),
    "ndiff_last_error": (C.c_char_p, []),
    "ndiff_engine_create": (C.c_int32, [C.POINTER(Config), C.POINTER(_P)]),
    "ndiff_engine_destroy": (None, [_P]),
    "ndiff_load_param": (C.c_int32, [_P, C.c_char_p, _P, C.c_int32, C.POINTER(C.c_int64)]),
    "ndiff_load_param_async": (C.c_int32, [_P, C.c_char_p, _P, C.c_int32, C.POINTER(C.c_int64), _P]),
    "ndiff_finalize_params": (C.c_int32, [_P, _P]),
    "ndiff_set_condition": (C.c_int32, [_P, _P, _P, _P, _P]),
    "ndiff_forward": (C.c_int32, [_P, _P, _P, _P, _P]),
    "ndiff_chain_begin": (C.c_int32, [_P, C.POINTER(Step), C.c_int32, _P, C.c_uint64, _P]),
    "ndiff_chain_run": (C.c_int32, [_P, C.c_int32, _P, _P, _P, _P]),
    "ndiff_chain_read": (C.c_int32, [_P, _P, _P]),
    "ndiff_chain_seek": (C.c_int32, [_P, C.c_int32, _P, C.c_uint64, _P]),
    "ndiff_sample_host": (C.c_int32, [_P, _P, _P, _P, C.POINTER(Step), C.c_int32, C.c_uint64, _P]),
    "ndiff_compose_noisy": (C.c_int32, [_P, _P, _P, _P, C.c_int64, _P]),
    "ndiff_debug_tensor": (C.c_int32, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _P]),
    "ndiff_launches_per_step": (C.c_int64, [_P]),
    "ndiff_conv_flops_per_step": (C.c_double, [_P]),
    "ndiff_time_layers": (C.c_int32, [_P, C.c_int32, C.POINTER(C.c_float), C.c_char_p, C.c_int32,
                                      C.POINTER(C.c_int32), _P]),
    "ndiff_op_conv": (C.c_int32, [C.c_int32] * 4 + [_P, C.c_int32, _P, C.c_int32] + [C.c_int32] * 4 +
                      [_P, C.c_int32, _P, _P, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P, C.c_int32, C.c_int32, _P]),
    "ndiff_op_conv_time": (C.c_int32, [C.c_int32] * 4 + [_P, C.c_int32, _P, C.c_int32] + [C.c_int32] * 4 +
                           [_P, C.c_int32, _P, _P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), _P]),
    "ndiff_op_gn_apply": (C.c_int32, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P] + [C.c_int32] * 4 + [_P]),
    "ndiff_op_layernorm": (C.c_int32, [_P, _P, C.c_int32, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "ndiff_op_pixel_chain": (C.c_int32, [C.c_int32] * 3 + [_P] * 6 + [C.c_int32, _P, _P, _P, _P]),
    "ndiff_op_philox_normal": (C.c_int32, [_P, C.c_int64, C.c_uint64, C.c_uint64, _P]),
    "ndiff_op_conv_ex": (C.c_int32, [C.c_int32] * 4 + [_P, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P, _P, C.c_int32, _P, _P, _P]),
    "ndiff_op_tail_chain": (C.c_int32, [C.c_int32, C.c_int32] + [_P] * 8 + [C.c_int32, _P, _P]),
    # ---- training row (noisediff_b200/training.py)
    "ndiff_trainer_create": (C.c_int32, [C.POINTER(Config), C.POINTER(_P)]),
    "ndiff_trainer_destroy": (None, [_P]),
    "ndiff_trainer_engine": (_P, [_P]),
    "ndiff_trainer_finalize": (C.c_int32, [_P, _P]),
    "ndiff_trainer_forward_backward": (C.c_int32, [_P, _P, _P, _P, _P, C.POINTER(C.c_double), _P]),
    "ndiff_trainer_flat": (C.c_int32, [_P, C.c_int32, C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ndiff_trainer_slot": (C.c_int32, [_P, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ndiff_trainer_adam_step": (C.c_int32, [_P] + [C.c_float] * 6 + [_P]),
    "ndiff_trainer_ema_update": (C.c_int32, [_P, C.c_float, _P]),
    "ndiff_trainer_time": (C.c_int32, [_P, C.c_char_p, C.c_int32, _P]),
    "ndiff_trainer_activation_bytes": (C.c_int64, [_P]),
    "ndiff_trainer_launches": (C.c_int64, [_P, C.c_int32]),
    "ndiff_op_wgrad": (C.c_int32, [C.c_int32] * 4 + [_P, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P, _P]),
    "ndiff_op_gn_backward": (C.c_int32, [_P] * 6 + [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P] + [C.c_int32] * 4 + [_P]),
    "ndiff_op_layernorm_backward": (C.c_int32, [_P, _P, C.c_int32, _P, _P, _P, _P, _P] + [C.c_int32] * 3 + [_P]),
}

_lib = None


def declared_symbols():
    """Every function name the public header declares (used by the CPU test that checks the export table)."""
    with open(HEADER_PATH) as f:
        return sorted(set(re.findall(r"NDIFF_API[^;(]*?\b(ndiff_\w+)\s*\(", f.read())))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"noisediff_b200: CUDA library not built ({LIB_PATH} missing). Run `python -c 'import __graft_entry__ "
                f"as g; g.build()'` or `make -C noisediff_b200/csrc`. There is no CPU / PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.ndiff_abi_version() != 1:
            raise RuntimeError("noisediff_b200: ABI version mismatch between the Python host and the library")
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().ndiff_last_error()
        raise RuntimeError("noisediff_b200: " + (msg.decode() if msg else f"error {rc}"))
